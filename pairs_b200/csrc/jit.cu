// User-defined kernels: the part of the reference's code generator that matters for the MD path, re-done for one GPU target.
// The reference traces a Python kernel body into IR and prints C++/CUDA for it (src/pairs/mapping/funcs.py:39-334,
// code_gen/cgen.py); here pairs_b200/kernelgen.py prints CUDA for the SAME restricted vocabulary directly (one statement per
// operation, in Python's evaluation order, like the reference's generated code), and this file compiles it at run time with
// NVRTC for sm_100a (--fmad=false: no contraction, so every operation is the IEEE operation the reference's C++ performs),
// loads the cubin through the runtime's library API and launches it on the context's stream.  Kernels that match one of the
// hand-written families (md_kernels.cu, dem_kernels.cu) never come here.
//
// Kernel ABI: extern "C" __global__ void <name>(PbJitArgs a) -- one thread per local particle; the struct below is repeated in
// the prelude the generated source starts with (pb_jit_prelude()).
#include <dlfcn.h>
#include <nvrtc.h>

#include <string>
#include <vector>

#include "ctx.cuh"

struct PbJitArgs {
    int nlocal, nslots, cap, pad;
    double cutsq;
    const double4 *pos;     // x, y, z, w = type bits (locals + ghosts)
    double4 *pos_w;         // same array, writable (particle kernels)
    double *vel;            // [3][cap]
    double *force;          // [3][cap]
    const double *mass;     // [cap]
    const int *flags;       // [cap]
    const int *numneigh;    // [cap]
    const int *neigh;       // sliced ELLPACK: ((i / 32) * nslots + k) * 32 + i % 32
    double *xdata;          // user-defined properties, rows of [cap]: component d of a property at (row0 + d) * cap + i
    const int *uid;         // [cap]
    const int *shape;       // [cap]
    double *radius;         // DEM scripts only (null otherwise): [cap]
    double *angvel;         //                                    [3][cap]
    double *torque;         //                                    [3][cap]
    double *inv_inertia;    //                                    [9][cap]
    double *rotmat;         //                                    [9][cap]
    double *quat;           //                                    [4][cap]
    const int *particle_cell;   // pair kernels over CELL lists (no neighbour lists built): flat cell of every particle,
    const int *cell_start;      // CSR over the reference's cells (cell 0 = INFINITE particles),
    const int *cell_list;       // particle indices in cell order
    int ncells, dim1, dim2, pad2;
};

static const char *PB_JIT_PRELUDE = R"PRELUDE(
struct PbJitArgs {
    int nlocal, nslots, cap, pad;
    double cutsq;
    const double4 *pos;
    double4 *pos_w;
    double *vel;
    double *force;
    const double *mass;
    const int *flags;
    const int *numneigh;
    const int *neigh;
    double *xdata;
    const int *uid;
    const int *shape;
    double *radius;
    double *angvel;
    double *torque;
    double *inv_inertia;
    double *rotmat;
    double *quat;
    const int *particle_cell;
    const int *cell_start;
    const int *cell_list;
    int ncells, dim1, dim2, pad2;
};
#ifndef PB_INFINITY
#define PB_INFINITY __longlong_as_double(0x7ff0000000000000LL)
#endif
#define PB_FLAG_FIXED 4
#define PB_SHAPE_SPHERE 0
__device__ __forceinline__ int pb_w_type(double w) { return (int) (__double_as_longlong(w) & 0xffffffffLL); }
__device__ __forceinline__ double4 pb_ld_pos(const double4 *p) {
    double4 r;
#ifdef PB_HAVE_LD256          /* one 256-bit load; needs the PTX of CUDA >= 12.9 */
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
#else
    const double2 lo = __ldg(reinterpret_cast<const double2 *>(p)), hi = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    r.x = lo.x; r.y = lo.y; r.z = hi.x; r.w = hi.y;
#endif
    return r;
}
)PRELUDE";

extern "C" const char *pb_jit_prelude(void) { return PB_JIT_PRELUDE; }

struct NvrtcApi {
    void *handle = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
    nvrtcResult (*Version)(int *, int *) = nullptr;
    const char *(*GetErrorString)(nvrtcResult) = nullptr;
    bool ld256 = false;
};

static NvrtcApi g_rtc;

static bool pb_nvrtc_load(std::string *err) {
    if(g_rtc.handle != nullptr) { return true; }
    void *h = nullptr;
    // the toolkit's copy first, bound to its own symbols: a Python process that imported PyTorch already holds an older
    // libnvrtc.so.12 (12.8) whose PTX does not know 256-bit loads
    for(const char *n : {"/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so"}) {
        h = dlopen(n, RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
        if(h != nullptr) { break; }
    }
    if(h == nullptr) { *err = std::string("cannot load libnvrtc: ") + dlerror(); return false; }
#define PB_SYM(field, name)                                                     \
    *(void **) (&g_rtc.field) = dlsym(h, name);                                 \
    if(g_rtc.field == nullptr) { *err = std::string("missing NVRTC symbol ") + name; return false; }
    PB_SYM(CreateProgram, "nvrtcCreateProgram");
    PB_SYM(CompileProgram, "nvrtcCompileProgram");
    PB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    PB_SYM(GetProgramLog, "nvrtcGetProgramLog");
    PB_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    PB_SYM(GetCUBIN, "nvrtcGetCUBIN");
    PB_SYM(DestroyProgram, "nvrtcDestroyProgram");
    PB_SYM(GetErrorString, "nvrtcGetErrorString");
    PB_SYM(Version, "nvrtcVersion");
#undef PB_SYM
    int major = 0, minor = 0;
    g_rtc.Version(&major, &minor);
    g_rtc.ld256 = major > 12 || (major == 12 && minor >= 9);
    g_rtc.handle = h;
    return true;
}

// source -> sm_100a cubin; returns false with the compiler log in *err
static bool pb_nvrtc_compile(const char *source, std::vector<char> *cubin, std::string *err, int maxreg = 0) {
    if(!pb_nvrtc_load(err)) { return false; }
    nvrtcProgram prog;
    nvrtcResult r = g_rtc.CreateProgram(&prog, source, "pairs_user_kernel.cu", 0, nullptr, nullptr);
    if(r != NVRTC_SUCCESS) { *err = std::string("nvrtcCreateProgram: ") + g_rtc.GetErrorString(r); return false; }
    const std::string reg_opt = "--maxrregcount=" + std::to_string(maxreg);
    std::vector<const char *> opts = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "-lineinfo"};
    if(g_rtc.ld256) { opts.push_back("-DPB_HAVE_LD256"); }
    if(maxreg > 0) { opts.push_back(reg_opt.c_str()); }      // tuning: register cap of a re-built kernel (option "dem_force_maxreg")
    r = g_rtc.CompileProgram(prog, (int) opts.size(), opts.data());
    if(r != NVRTC_SUCCESS) {
        size_t n = 0;
        g_rtc.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if(n > 0) { g_rtc.GetProgramLog(prog, &log[0]); }
        *err = std::string("NVRTC: ") + g_rtc.GetErrorString(r) + "\n" + log;
        g_rtc.DestroyProgram(&prog);
        return false;
    }
    size_t n = 0;
    g_rtc.GetCUBINSize(prog, &n);
    cubin->resize(n);
    g_rtc.GetCUBIN(prog, cubin->data());
    g_rtc.DestroyProgram(&prog);
    return n > 0;
}

// compile only: lets the host side (and the CPU test-suite) reject a kernel before any GPU is involved
extern "C" int pb_jit_check(const char *source, char *log, int log_cap) {
    std::vector<char> cubin;
    std::string err;
    const bool ok = pb_nvrtc_compile(source, &cubin, &err);
    if(log != nullptr && log_cap > 0) {
        snprintf(log, (size_t) log_cap, "%s", ok ? "" : err.c_str());
    }
    return ok ? (int) cubin.size() : -1;
}

struct PbJitKernel {
    cudaLibrary_t lib;
    cudaKernel_t kernel;
    std::string name;
};

static std::vector<PbJitKernel> *pb_jit_table(pb_ctx *ctx) {
    if(ctx->jit == nullptr) { ctx->jit = new std::vector<PbJitKernel>(); }
    return (std::vector<PbJitKernel> *) ctx->jit;
}

struct PbDemUserForce {
    cudaLibrary_t lib;
    cudaKernel_t staged, fused;
};

void pb_jit_destroy(pb_ctx *ctx) {
    if(ctx->dem_user_force != nullptr) {
        auto *u = (PbDemUserForce *) ctx->dem_user_force;
        cudaLibraryUnload(u->lib);
        delete u;
        ctx->dem_user_force = nullptr;
    }
    if(ctx->jit == nullptr) { return; }
    auto *tab = (std::vector<PbJitKernel> *) ctx->jit;
    for(auto &k : *tab) { cudaLibraryUnload(k.lib); }
    delete tab;
    ctx->jit = nullptr;
}

extern "C" int pb_jit_compile(pb_ctx *ctx, const char *source, const char *kernel_name, int *handle) {
    PB_CHECK(cudaSetDevice(ctx->device));
    std::vector<char> cubin;
    std::string err;
    if(!pb_nvrtc_compile(source, &cubin, &err)) { ctx->set_error(err); return -1; }
    PbJitKernel k;
    k.name = kernel_name;
    PB_CHECK(cudaLibraryLoadData(&k.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    PB_CHECK(cudaLibraryGetKernel(&k.kernel, k.lib, kernel_name));
    auto *tab = pb_jit_table(ctx);
    tab->push_back(k);
    *handle = (int) tab->size() - 1;
    return 0;
}

int pb_materialise_force_reset(pb_ctx *ctx);

// kind 0: pair kernel over the neighbour lists (needs current lists; `cutoff` is the interaction cutoff of compute());
// kind 1: per-particle kernel; kind 2: pair kernel over the cell lists (needs current cell lists);
// kind 3: pair kernel generated for compute_half() over half neighbour lists
extern "C" int pb_jit_launch(pb_ctx *ctx, int handle, int kind, double cutoff) {
    PB_CHECK(cudaSetDevice(ctx->device));
    auto *tab = pb_jit_table(ctx);
    if(handle < 0 || handle >= (int) tab->size()) { ctx->set_error("pb_jit_launch: bad kernel handle"); return -1; }
    if(kind < 0 || kind > 3) { ctx->set_error("pb_jit_launch: kind is 0 (pair / lists), 1 (particle), 2 (pair / cells) or 3 (pair / half lists)"); return -1; }
    PbJitKernel &k = (*tab)[handle];
    PbStage st(ctx, k.name.c_str());
    if(kind == 0 || kind == 3) {
        if(!pb_lists_valid(ctx)) { ctx->set_error("pb_jit_launch: neighbour lists are stale"); return -1; }
        PB_TRY(pb_require_neigh32(ctx));      // generated pair kernels walk per-particle lists
        if(ctx->lanes != 1) { ctx->set_error("pb_jit_launch: user pair kernels need one lane per particle"); return -1; }
        if(ctx->half_lists != (kind == 3)) {
            ctx->set_error("pb_jit_launch: the kernel was generated for the other kind of neighbour lists (compute_half)");
            return -1;
        }
    }
    if(kind == 2 && ctx->cells_n != ctx->nlocal + ctx->nghost) { ctx->set_error("pb_jit_launch: cell lists are stale"); return -1; }
    PB_TRY(pb_materialise_force_reset(ctx));      // a deferred reset_volatile_properties must be visible to user code
    if(ctx->nlocal == 0) { return 0; }
    PbJitArgs a;
    a.nlocal = ctx->nlocal; a.nslots = ctx->nslots; a.cap = ctx->pcap; a.pad = 0;
    a.cutsq = cutoff * cutoff;
    a.pos = ctx->pos; a.pos_w = ctx->pos; a.vel = ctx->vel; a.force = ctx->force; a.mass = ctx->mass; a.flags = ctx->flags;
    a.numneigh = ctx->numneigh; a.neigh = ctx->neigh; a.xdata = ctx->xdata;
    a.uid = ctx->uid; a.shape = ctx->shape;
    a.radius = ctx->radius; a.angvel = ctx->angvel; a.torque = ctx->torque;
    a.inv_inertia = ctx->inv_inertia; a.rotmat = ctx->rotmat; a.quat = ctx->quat;
    a.particle_cell = ctx->particle_cell; a.cell_start = ctx->cell_start; a.cell_list = ctx->cell_list;
    a.ncells = ctx->ncells; a.dim1 = ctx->dim_cells[1]; a.dim2 = ctx->dim_cells[2]; a.pad2 = 0;
    void *params[] = {&a};
    PB_CHECK(cudaLaunchKernel((const void *) k.kernel, dim3(pb_blocks(ctx->nlocal, 128)), dim3(128), params, 0, ctx->stream));
    ctx->launches++;
    return 0;
}

// ---- user-defined DEM contact models ----------------------------------------------------------------------------------------
// examples/dem.py's linear_spring_dashpot is ONE contact model; the reference generates code for whatever body the user writes
// (mapping/funcs.py:39-334, contact properties :230-263).  Here the contact KERNEL stays the hand-written one -- detection pass,
// history lookup / insert keyed by the partner's uid, usage marks, clean-up, force / torque accumulation
// (dem_force_kernel.cuh) -- and only the per-pair model is exchanged: kernelgen.py prints the user's body as a device function
//     bool <name>(xi, vi, wi, mi, ri, xj, vj, wj, mj, rj, n, cp, delta, tij, tsd, ivm, sticking, F, T)
// and this file compiles  prelude + dem_math.h + that function + dem_force_kernel.cuh  (the texts of the two headers are embedded
// in the library at build time) with NVRTC into the staged and the fused variant of the kernel.
#include "embedded_sources.inc"

static const char *PB_DEM_USER_WRAPPERS = R"WRAP(
extern "C" __global__ void __launch_bounds__(128) pb_user_dem_force_staged(PbDemForceArgs a) { pb_dem_force_body<false>(a); }
extern "C" __global__ void __launch_bounds__(128) pb_user_dem_force_fused(PbDemForceArgs a) { pb_dem_force_body<true>(a); }
)WRAP";

static std::string pb_dem_user_source(const char *model_source, const char *model_name) {
    std::string src = PB_JIT_PRELUDE;
    src += PB_SRC_DEM_MATH;
    src += "\n";
    if(model_source != nullptr) {           // null: the contact model of examples/dem.py (pb_dem_pair_force), re-built at run time
        src += model_source;
        src += "\n#define PB_DEM_USER_PAIR ";
        src += model_name;
        src += "\n";
    }
    src += PB_SRC_DEM_FORCE_KERNEL;
    src += PB_DEM_USER_WRAPPERS;
    return src;
}

// compile only (no GPU needed): cubin size, or -1 with the compiler log in `log`
extern "C" int pb_jit_check_dem_model(const char *model_source, const char *model_name, char *log, int log_cap) {
    return pb_jit_check(pb_dem_user_source(model_source, model_name).c_str(), log, log_cap);
}

// installs the contact model: pb_dem_linear_spring_dashpot and pb_dem_run use it from now on (NULL source: back to the built-in)
extern "C" int pb_jit_set_dem_model(pb_ctx *ctx, const char *model_source, const char *model_name) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(!ctx->dem) { ctx->set_error("pb_jit_set_dem_model: call pb_dem_enable first"); return -1; }
    if(ctx->dem_user_force != nullptr) {
        auto *old = (PbDemUserForce *) ctx->dem_user_force;
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
        cudaLibraryUnload(old->lib);
        delete old;
        ctx->dem_user_force = nullptr;
    }
    if(model_source == nullptr && ctx->dem_force_maxreg <= 0) { return 0; }
    std::vector<char> cubin;
    std::string err;
    if(!pb_nvrtc_compile(pb_dem_user_source(model_source, model_name).c_str(), &cubin, &err, ctx->dem_force_maxreg)) { ctx->set_error(err); return -1; }
    PbDemUserForce u;
    PB_CHECK(cudaLibraryLoadData(&u.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    cudaError_t e = cudaLibraryGetKernel(&u.staged, u.lib, "pb_user_dem_force_staged");
    if(e == cudaSuccess) { e = cudaLibraryGetKernel(&u.fused, u.lib, "pb_user_dem_force_fused"); }
    if(e != cudaSuccess) {
        cudaLibraryUnload(u.lib);
        ctx->set_error(std::string("pb_jit_set_dem_model: ") + cudaGetErrorString(e));
        return -1;
    }
    ctx->dem_user_force = new PbDemUserForce(u);
    return 0;
}

// PbDemForceArgs is defined in dem_force_kernel.cuh; the launch sites in dem_kernels.cu build it once for either kernel
struct PbDemParams;
int pb_jit_launch_dem_force_raw(pb_ctx *ctx, int fused, void *args_struct) {
    auto *u = (PbDemUserForce *) ctx->dem_user_force;
    if(u == nullptr) { ctx->set_error("no user-defined DEM contact model installed"); return -1; }
    void *params[] = {args_struct};
    PB_CHECK(cudaLaunchKernel((const void *) (fused ? u->fused : u->staged), dim3(pb_blocks(ctx->nlocal, 512)), dim3(128), params, 0,
                              ctx->stream));      // PB_DEM_CTA_PARTICLES = 512 particles per CTA (dem_force_kernel.cuh)
    ctx->launches++;
    return 0;
}
