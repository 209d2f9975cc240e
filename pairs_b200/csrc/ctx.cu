// Context lifecycle, device store, domain decomposition, upload/download.
// Replaces the reference's PairsSimulation registry + dirty-flag copy engine (runtime/pairs.hpp:27-517,
// runtime/pairs.cpp:31-327) by a device-resident SoA store: nothing but set-up data, thermo scalars and
// capacity counters ever crosses PCIe.
#include <cstring>
#include <algorithm>
#include <cmath>

#include "ctx.cuh"

static std::string g_create_error;

extern "C" const char *pb_version(void) { return "pairs_b200 0.1 (sm_100a)"; }

extern "C" const char *pb_last_error(const pb_ctx *ctx) {
    return ctx != nullptr ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" void pb_destroy(pb_ctx *ctx);

extern "C" int pb_create(pb_ctx **out, int device) {
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if(e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("pairs_b200: no CUDA device available (") + cudaGetErrorString(e) +
                         "); this backend has no CPU fallback";
        return -1;
    }
    if(device < 0 || device >= ndev) {
        g_create_error = "pairs_b200: invalid device index " + std::to_string(device);
        return -1;
    }
    e = cudaSetDevice(device);
    if(e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return -1; }
    pb_ctx *ctx = new pb_ctx();
    ctx->device = device;
    // every resource of the context, checked: a half-built context is destroyed again and the first error reported
    auto fail = [&](cudaError_t err, const char *what) {
        g_create_error = std::string("pairs_b200: ") + what + ": " + cudaGetErrorString(err);
        pb_destroy(ctx);
        return -1;
    };
#define PB_CREATE_CHECK(call)                                   \
    do {                                                        \
        e = (call);                                             \
        if(e != cudaSuccess) { return fail(e, #call); }         \
    } while(0)
    PB_CREATE_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    PB_CREATE_CHECK(cudaEventCreate(&ctx->ev0));
    PB_CREATE_CHECK(cudaEventCreate(&ctx->ev1));
    {
        int prio_lo = 0, prio_hi = 0;      // comm stream gets the highest priority: its blocks are scheduled ahead of the force kernel's
        PB_CREATE_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        PB_CREATE_CHECK(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, prio_hi));
    }
    PB_CREATE_CHECK(cudaEventCreateWithFlags(&ctx->ev_prev, cudaEventDisableTiming));
    PB_CREATE_CHECK(cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventDisableTiming));
    PB_CREATE_CHECK(cudaEventCreateWithFlags(&ctx->ev_io[0], cudaEventDisableTiming));
    PB_CREATE_CHECK(cudaEventCreateWithFlags(&ctx->ev_io[1], cudaEventDisableTiming));
    PB_CREATE_CHECK(cudaMalloc(&ctx->d_scalars, sizeof(int) * PB_NSCALARS));
    PB_CREATE_CHECK(cudaMemsetAsync(ctx->d_scalars, 0, sizeof(int) * PB_NSCALARS, ctx->stream));      // (on the stream the kernels use: the legacy stream does not order against it)
    PB_CREATE_CHECK(cudaMallocHost(&ctx->h_scalars, sizeof(int) * PB_NSCALARS));
#undef PB_CREATE_CHECK
    *out = ctx;
    return 0;
}

void pb_nccl_destroy(pb_ctx *ctx);
void pb_jit_destroy(pb_ctx *ctx);

extern "C" void pb_destroy(pb_ctx *ctx) {
    if(ctx == nullptr) { return; }
    cudaSetDevice(ctx->device);
    if(ctx->stream != nullptr) { cudaStreamSynchronize(ctx->stream); }
    pb_nccl_destroy(ctx);
    pb_jit_destroy(ctx);
    void *bufs[] = {ctx->pos, ctx->pos_alt, ctx->vel, ctx->vel_alt, ctx->force, ctx->mass, ctx->mass_alt, ctx->type,
                    ctx->type_alt, ctx->flags, ctx->flags_alt, ctx->uid, ctx->uid_alt, ctx->shape, ctx->shape_alt, ctx->tag,
                    ctx->tag_alt, ctx->particle_cell, ctx->cell_count, ctx->cell_start, ctx->cell_slot, ctx->cell_list,
                    ctx->cell_key, ctx->sub_start, ctx->scan_tmp, ctx->neigh, ctx->numneigh, ctx->d_eps, ctx->d_sig6, ctx->send_map, ctx->send_mult,
                    ctx->send_buf, ctx->recv_buf, ctx->sel_blocks, ctx->sel_flag, ctx->sel_scan, ctx->d_partial, ctx->d_scalars,
                    ctx->group_flag, ctx->group_scan, ctx->groups_interior, ctx->groups_boundary,
                    ctx->radius, ctx->angvel, ctx->torque, ctx->normal, ctx->inv_inertia, ctx->rotmat, ctx->quat, ctx->num_contacts,
                    ctx->contact_uid, ctx->contact_used, ctx->contact_stick, ctx->contact_tsd, ctx->contact_ivm, ctx->contact_x, ctx->m32, ctx->d_fric_static,
                    ctx->d_fric_dynamic, ctx->d_dem_flag, ctx->xdata, ctx->xdata_alt,
                    ctx->tiles, ctx->tile_lvl, ctx->tile_cnt, ctx->tile_off, ctx->tile_pad, ctx->tile_row, ctx->twords, ctx->tile_flag,
                    ctx->tile_scan, ctx->tiles_interior, ctx->tiles_boundary, ctx->tile_hdrs, ctx->tile_rowsrc, ctx->mxy[0], ctx->mxy[1], ctx->mz[0], ctx->mz[1],
                    ctx->mmeta, ctx->ghost_csr, ctx->io_stage};
    for(void *b : bufs) { if(b != nullptr) { cudaFree(b); } }
    if(ctx->h_scalars != nullptr) { cudaFreeHost(ctx->h_scalars); }
    for(auto &kv : ctx->timers) { for(auto &pr : kv.second.pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); } }
    for(cudaEvent_t e : ctx->event_pool) { cudaEventDestroy(e); }
    for(cudaEvent_t e : {ctx->ev0, ctx->ev1, ctx->ev_prev, ctx->ev_sync, ctx->ev_io[0], ctx->ev_io[1]}) { if(e != nullptr) { cudaEventDestroy(e); } }
    if(ctx->upload_perm != nullptr) { cudaFree(ctx->upload_perm); }
    if(ctx->comm_stream != nullptr) { cudaStreamDestroy(ctx->comm_stream); }
    if(ctx->stream != nullptr) { cudaStreamDestroy(ctx->stream); }
    delete ctx;
}

// ---- growable device arrays -------------------------------------------------------------------------------
template<typename T>
int pb_regrow(pb_ctx *ctx, T **p, size_t old_count, size_t new_count, bool keep) {
    T *q = nullptr;
    PB_CHECK(cudaMalloc(&q, sizeof(T) * new_count));
    if(keep && *p != nullptr && old_count > 0) {
        PB_CHECK(cudaMemcpyAsync(q, *p, sizeof(T) * old_count, cudaMemcpyDeviceToDevice, ctx->stream));
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    if(*p != nullptr) { PB_CHECK(cudaFree(*p)); }
    *p = q;
    return 0;
}

template int pb_regrow<int>(pb_ctx *, int **, size_t, size_t, bool);
template int pb_regrow<double>(pb_ctx *, double **, size_t, size_t, bool);

// SoA [comps][cap] arrays need a strided move when the capacity changes
int pb_regrow_soa(pb_ctx *ctx, double **p, int comps, size_t old_cap, size_t new_cap, size_t used, bool keep) {
    double *q = nullptr;
    PB_CHECK(cudaMalloc(&q, sizeof(double) * comps * new_cap));
    PB_CHECK(cudaMemsetAsync(q, 0, sizeof(double) * comps * new_cap, ctx->stream));
    if(keep && *p != nullptr && used > 0) {
        for(int d = 0; d < comps; d++) {
            PB_CHECK(cudaMemcpyAsync(q + d * new_cap, *p + d * old_cap, sizeof(double) * used, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    if(*p != nullptr) { PB_CHECK(cudaFree(*p)); }
    *p = q;
    return 0;
}

static int pb_regrow_soa3(pb_ctx *ctx, double **p, size_t old_cap, size_t new_cap, size_t used, bool keep) {
    return pb_regrow_soa(ctx, p, 3, old_cap, new_cap, used, keep);
}

int pb_ensure_particle_capacity(pb_ctx *ctx, int needed) {
    if(needed <= ctx->pcap) { return 0; }
    if(ctx->upload_pending) {      // the arrays move: the deferred half of the upload must have landed in the old ones
        PB_CHECK(cudaStreamSynchronize(ctx->comm_stream));
    }
    // the reference doubles (transformations/modules.py:159-203); 1.5x keeps the 180 GB budget for 4M-atom boxes
    size_t newcap = std::max<size_t>((size_t) needed, (size_t) ctx->pcap + (size_t) ctx->pcap / 2);
    newcap = (newcap + 255) / 256 * 256;
    const size_t used = (size_t) ctx->nlocal + (size_t) ctx->nghost;
    const size_t oldcap = (size_t) ctx->pcap;
    PB_TRY(pb_regrow(ctx, &ctx->pos, used, newcap, true));
    PB_TRY(pb_regrow(ctx, &ctx->pos_alt, 0, newcap, false));
    PB_TRY(pb_regrow_soa3(ctx, &ctx->vel, oldcap, newcap, used, true));
    PB_TRY(pb_regrow_soa3(ctx, &ctx->vel_alt, oldcap, newcap, 0, false));
    PB_TRY(pb_regrow_soa3(ctx, &ctx->force, oldcap, newcap, used, true));
    PB_TRY(pb_regrow(ctx, &ctx->mass, used, newcap, true));
    PB_TRY(pb_regrow(ctx, &ctx->mass_alt, 0, newcap, false));
    int **ints[] = {&ctx->type, &ctx->flags, &ctx->uid, &ctx->shape, &ctx->tag};
    int **alts[] = {&ctx->type_alt, &ctx->flags_alt, &ctx->uid_alt, &ctx->shape_alt, &ctx->tag_alt};
    for(int k = 0; k < 5; k++) {
        PB_TRY(pb_regrow(ctx, ints[k], used, newcap, true));
        PB_TRY(pb_regrow(ctx, alts[k], 0, newcap, false));
    }
    PB_TRY(pb_regrow(ctx, &ctx->particle_cell, used, newcap, true));
    PB_TRY(pb_regrow(ctx, &ctx->cell_slot, 0, newcap, false));
    PB_TRY(pb_regrow(ctx, &ctx->cell_list, 0, newcap, false));
    PB_TRY(pb_regrow(ctx, &ctx->cell_key, 0, newcap, false));
    PB_TRY(pb_regrow(ctx, &ctx->sel_flag, 0, newcap + 1, false));
    PB_TRY(pb_regrow(ctx, &ctx->sel_scan, 0, newcap + 1, false));
    PB_TRY(pb_regrow(ctx, &ctx->numneigh, 0, newcap, false));
    if(ctx->dem) { PB_TRY(pb_dem_grow(ctx, oldcap, newcap, used)); }
    PB_TRY(pb_xprops_grow(ctx, oldcap, newcap, used));
    ctx->pcap = (int) newcap;
    return 0;
}

int pb_ensure_send_capacity(pb_ctx *ctx, int needed) {
    needed = std::max(needed, 1);
    if(needed <= ctx->send_cap) { return 0; }
    size_t newcap = std::max<size_t>((size_t) needed, (size_t) ctx->send_cap * 2);
    newcap = (newcap + 255) / 256 * 256;
    PB_TRY(pb_regrow(ctx, &ctx->send_map, (size_t) ctx->nsend_all, newcap, true));
    PB_TRY(pb_regrow(ctx, &ctx->send_mult, (size_t) ctx->nsend_all * 3, newcap * 3, true));
    PB_TRY(pb_regrow(ctx, &ctx->send_buf, 0, newcap * pb_record_elems(ctx), false));
    ctx->send_cap = (int) newcap;
    return 0;
}

extern "C" int pb_reserve(pb_ctx *ctx, int particle_capacity, int neighbor_capacity) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(particle_capacity > 0) { PB_TRY(pb_ensure_particle_capacity(ctx, particle_capacity)); }
    if(neighbor_capacity > ctx->ncap) { ctx->ncap = neighbor_capacity; }
    return 0;
}

// ---- domain -----------------------------------------------------------------------------------------------
// Regular6DStencil::setConfig (runtime/domain/regular_6d_stencil.cpp:10-54): the (i,j,k), i*j*k = world, that
// minimises the summed face area; ties resolved by loop order (strict <).
extern "C" int pb_rank_grid(int world_size, const double grid[6], int partitioner, int nranks[3]) {
    const int part[3] = {1, 1, partitioner == PB_PARTITION_REGULAR_XY ? 0 : 1};
    double area[3];
    double best_surf = 0.0;
    int d = 0;
    for(int d1 = 0; d1 < 3; d1++) {
        nranks[d1] = 1;
        for(int d2 = d1 + 1; d2 < 3; d2++) {
            area[d] = (grid[d1 * 2 + 1] - grid[d1 * 2]) * (grid[d2 * 2 + 1] - grid[d2 * 2]);
            best_surf += 2.0 * area[d];
            d++;
        }
    }
    for(int i = 1; i <= world_size; i++) {
        if(world_size % i != 0) { continue; }
        const int rem = world_size / i;
        for(int j = 1; j <= rem; j++) {
            if(rem % j != 0) { continue; }
            const int k = rem / j;
            if((part[0] || i == 1) && (part[1] || j == 1) && (part[2] || k == 1)) {
                const double surf = (area[0] / i / j) + (area[1] / i / k) + (area[2] / j / k);
                if(surf < best_surf) {
                    nranks[0] = i; nranks[1] = j; nranks[2] = k;
                    best_surf = surf;
                }
            }
        }
    }
    return 0;
}

// setBoundingBox (:56-82) with MPI_Cart_create(reorder = 0) semantics: row-major rank numbering, periodic shifts.
extern "C" int pb_init_domain(pb_ctx *ctx, const double grid[6], const int pbc[3], int partitioner, int world_size, int rank) {
    if(world_size < 1 || rank < 0 || rank >= world_size) { ctx->set_error("pb_init_domain: bad rank/world_size"); return -1; }
    memcpy(ctx->grid, grid, sizeof(double) * 6);
    memcpy(ctx->pbc_flag, pbc, sizeof(int) * 3);
    ctx->partitioner = partitioner;
    ctx->world = world_size;
    ctx->rank = rank;
    pb_rank_grid(world_size, grid, partitioner, ctx->nranks);
    const int *n = ctx->nranks;
    int rem = rank;
    int c[3];
    c[2] = rem % n[2]; rem /= n[2];
    c[1] = rem % n[1]; rem /= n[1];
    c[0] = rem;
    for(int d = 0; d < 3; d++) {
        const double rank_length = (grid[d * 2 + 1] - grid[d * 2]) / (double) n[d];
        int cp[3] = {c[0], c[1], c[2]}, cn[3] = {c[0], c[1], c[2]};
        cp[d] = (c[d] - 1 + n[d]) % n[d];
        cn[d] = (c[d] + 1) % n[d];
        ctx->coords[d] = c[d];
        ctx->neighbor_ranks[d * 2 + 0] = (cp[0] * n[1] + cp[1]) * n[2] + cp[2];
        ctx->neighbor_ranks[d * 2 + 1] = (cn[0] * n[1] + cn[1]) * n[2] + cn[2];
        ctx->pbc[d * 2 + 0] = (c[d] == 0) ? 1 : 0;
        ctx->pbc[d * 2 + 1] = (c[d] == n[d] - 1) ? -1 : 0;
        ctx->subdom[d * 2 + 0] = grid[d * 2] + rank_length * (double) c[d];
        ctx->subdom[d * 2 + 1] = ctx->subdom[d * 2 + 0] + rank_length;
    }
    ctx->domain_set = true;
    ctx->cells_set = false;
    return 0;
}

extern "C" int pb_get_decomposition(const pb_ctx *ctx, int nranks[3], int neighbor_ranks[6], int pbc[6], double subdom[6]) {
    memcpy(nranks, ctx->nranks, sizeof(int) * 3);
    memcpy(neighbor_ranks, ctx->neighbor_ranks, sizeof(int) * 6);
    memcpy(pbc, ctx->pbc, sizeof(int) * 6);
    memcpy(subdom, ctx->subdom, sizeof(double) * 6);
    return 0;
}

// ---- upload / download ------------------------------------------------------------------------------------
__global__ void pb_k_pack_pos(int n, const double *__restrict__ aos, const int *__restrict__ type, double4 *__restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) { pos[i] = make_double4(aos[i * 3], aos[i * 3 + 1], aos[i * 3 + 2], pb_type_w(type[i])); }
}

__global__ void pb_k_aos_to_soa3(int n, int cap, const double *__restrict__ aos, double *__restrict__ soa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        soa[i] = aos[i * 3];
        soa[cap + i] = aos[i * 3 + 1];
        soa[2 * cap + i] = aos[i * 3 + 2];
    }
}

__global__ void pb_k_soa3_to_aos(int n, int cap, const double *__restrict__ soa, double *__restrict__ aos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        aos[i * 3] = soa[i];
        aos[i * 3 + 1] = soa[cap + i];
        aos[i * 3 + 2] = soa[2 * cap + i];
    }
}

__global__ void pb_k_unpack_pos(int n, const double4 *__restrict__ pos, double *__restrict__ aos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        const double4 p = pos[i];
        aos[i * 3] = p.x; aos[i * 3 + 1] = p.y; aos[i * 3 + 2] = p.z;
    }
}

__global__ void pb_k_fill_int(int n, int *p, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) { p[i] = v; }
}

__global__ void pb_k_fill_real(int n, double *p, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) { p[i] = v; }
}

__global__ void pb_k_iota(int n, int *p, int base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) { p[i] = base + i; }
}

int pb_io_stage(pb_ctx *ctx, size_t bytes, double **out) {
    if(bytes > ctx->io_stage_bytes) {
        if(ctx->io_stage != nullptr) { PB_CHECK(cudaFree(ctx->io_stage)); ctx->io_stage = nullptr; ctx->io_stage_bytes = 0; }
        const size_t want = bytes + bytes / 8;
        PB_CHECK(cudaMalloc(&ctx->io_stage, want));
        ctx->io_stage_bytes = want;
    }
    *out = ctx->io_stage;
    return 0;
}

// defer: velocities and masses travel on comm_stream behind the positions and nothing is waited for -- the caller
// (pb_md_run_from_host) joins the streams once the first list build, which needs positions only, is under way
static int pb_upload_impl(pb_ctx *ctx, int n, const double *position, const double *velocity, const double *mass,
                          const int *type, const int *flags, const int *uid, const int *shape, bool defer) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(position == nullptr && n > 0) { ctx->set_error("pb_upload_particles: position is required"); return -1; }
    PB_TRY(pb_ensure_particle_capacity(ctx, std::max(n + n / 4 + 1024, 1024)));
    ctx->nlocal = n;
    ctx->nghost = 0;
    ctx->nsend_all = 0;
    ctx->neigh_n = 0;
    ctx->tiles_n = -1;
    ctx->cells_n = 0;
    if(n == 0) { return 0; }
    const int T = 256, B = pb_blocks(n, T);
    // positions and velocities pass through the staging area (AoS -> double4 / SoA), one half each: all copies are issued back to
    // back, the kernels behind them, ONE synchronisation at the end
    double *stage = nullptr;
    PB_TRY(pb_io_stage(ctx, sizeof(double) * 6 * (size_t) n, &stage));
    double *const stage_v = stage + 3 * (size_t) n;
    auto upload_int = [&](const int *src, int *dst, int dflt) -> int {
        if(src != nullptr) {
            PB_CHECK(cudaMemcpyAsync(dst, src, sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            PB_LAUNCH(pb_k_fill_int, B, T, n, dst, dflt);
        }
        return 0;
    };
    PB_TRY(upload_int(type, ctx->type, 0));
    PB_TRY(upload_int(flags, ctx->flags, 0));
    PB_TRY(upload_int(uid, ctx->uid, 0));
    PB_TRY(upload_int(shape, ctx->shape, PB_SHAPE_POINTMASS));
    PB_LAUNCH(pb_k_iota, B, T, n, ctx->tag, ctx->tag_base);
    PB_CHECK(cudaMemcpyAsync(stage, position, sizeof(double) * 3 * (size_t) n, cudaMemcpyHostToDevice, ctx->stream));
    PB_LAUNCH(pb_k_pack_pos, B, T, n, stage, ctx->type, ctx->pos);
    if(defer) {
        PB_CHECK(cudaEventRecord(ctx->ev_io[0], ctx->stream));
        PB_CHECK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_io[0], 0));      // the link carries the positions first
        std::swap(ctx->stream, ctx->comm_stream);
    }
    if(velocity != nullptr) {
        PB_CHECK(cudaMemcpyAsync(stage_v, velocity, sizeof(double) * 3 * (size_t) n, cudaMemcpyHostToDevice, ctx->stream));
        PB_LAUNCH(pb_k_aos_to_soa3, B, T, n, ctx->pcap, stage_v, ctx->vel);
    } else {
        PB_CHECK(cudaMemsetAsync(ctx->vel, 0, sizeof(double) * 3 * (size_t) ctx->pcap, ctx->stream));
    }
    if(mass != nullptr) {
        PB_CHECK(cudaMemcpyAsync(ctx->mass, mass, sizeof(double) * (size_t) n, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        PB_LAUNCH(pb_k_fill_real, B, T, n, ctx->mass, 1.0);
    }
    if(defer) {
        const cudaError_t e = cudaEventRecord(ctx->ev_io[1], ctx->stream);
        std::swap(ctx->stream, ctx->comm_stream);
        PB_CHECK(e);
        ctx->upload_pending = true;
    }
    PB_CHECK(cudaMemsetAsync(ctx->force, 0, sizeof(double) * 3 * (size_t) ctx->pcap, ctx->stream));
    if(ctx->dem) {   // DEM extras start zeroed (defaults of add_property); contact tables are emptied
        PB_CHECK(cudaMemsetAsync(ctx->torque, 0, sizeof(double) * 3 * (size_t) ctx->pcap, ctx->stream));
        PB_CHECK(cudaMemsetAsync(ctx->angvel, 0, sizeof(double) * 3 * (size_t) ctx->pcap, ctx->stream));
        PB_CHECK(cudaMemsetAsync(ctx->normal, 0, sizeof(double) * 3 * (size_t) ctx->pcap, ctx->stream));
        PB_CHECK(cudaMemsetAsync(ctx->num_contacts, 0, sizeof(int) * (size_t) ctx->pcap, ctx->stream));
    }
    PB_TRY(pb_xprops_defaults(ctx));     // user-defined properties of the new particles: their declared defaults
    ctx->force_is_zero = false;
    if(!defer) { PB_CHECK(cudaStreamSynchronize(ctx->stream)); }
    return 0;
}

extern "C" int pb_upload_particles(pb_ctx *ctx, int n, const double *position, const double *velocity, const double *mass,
                                   const int *type, const int *flags, const int *uid, const int *shape) {
    return pb_upload_impl(ctx, n, position, velocity, mass, type, flags, uid, shape, false);
}

// main stream waits for the deferred half of an upload (no-op otherwise)
int pb_upload_join(pb_ctx *ctx) {
    if(ctx->upload_pending) {
        ctx->upload_pending = false;
        PB_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_io[1], 0));
    }
    return 0;
}

// pb_upload_particles + pb_md_run in one call, for callers whose state lives in HOST memory: the copies of velocities and masses
// (more than half of the bytes) overlap the first neighbour-list build, which works on positions alone.  The loop must start with
// a reneighbouring iteration (ts_begin == 0) on a single-rank, non-DEM context without user-defined properties; otherwise the two
// calls simply run one after the other.  Results are those of the two separate calls, bit for bit.  Returns with both streams idle:
// the host arrays are free again.
extern "C" int pb_md_run_from_host(pb_ctx *ctx, const pb_md_params *p, int n, const double *position, const double *velocity, const double *mass,
                                   const int *type, const int *flags, const int *uid, const int *shape, int ts_begin, int ts_end,
                                   double *thermo_out, int thermo_cap, int *n_thermo) {
    const bool defer = ctx->world == 1 && !ctx->dem && ctx->xprops.empty() && velocity != nullptr && n > 0 && ts_begin == 0 && ts_end > 0;
    PB_TRY(pb_upload_impl(ctx, n, position, velocity, mass, type, flags, uid, shape, defer));
    const int rc = pb_md_run(ctx, p, ts_begin, ts_end, thermo_out, thermo_cap, n_thermo);
    if(defer) {
        ctx->upload_pending = false;
        const cudaError_t e1 = cudaStreamSynchronize(ctx->comm_stream), e2 = cudaStreamSynchronize(ctx->stream);
        if(rc >= 0) { PB_CHECK(e1); PB_CHECK(e2); }
    }
    return rc;
}

extern "C" int pb_counts(const pb_ctx *ctx, int *nlocal, int *nghost) {
    if(nlocal != nullptr) { *nlocal = ctx->nlocal; }
    if(nghost != nullptr) { *nghost = ctx->nghost; }
    return 0;
}

int pb_materialise_force_reset(pb_ctx *ctx);

extern "C" int pb_download_real(pb_ctx *ctx, const char *name, double *out, int with_ghosts) {
    PB_CHECK(cudaSetDevice(ctx->device));
    const int n = ctx->nlocal + (with_ghosts ? ctx->nghost : 0);
    if(n == 0) { return 0; }
    const std::string nm(name);
    const int T = 256, B = pb_blocks(n, T);
    if(nm == "mass") {
        PB_CHECK(cudaMemcpyAsync(out, ctx->mass, sizeof(double) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    double *stage = nullptr;
    PB_TRY(pb_io_stage(ctx, sizeof(double) * 3 * (size_t) n, &stage));
    if(nm == "position") {
        PB_LAUNCH(pb_k_unpack_pos, B, T, n, ctx->pos, stage);
    } else if(nm == "linear_velocity") {
        PB_LAUNCH(pb_k_soa3_to_aos, B, T, n, ctx->pcap, ctx->vel, stage);
    } else if(nm == "force") {
        PB_TRY(pb_materialise_force_reset(ctx));
        PB_LAUNCH(pb_k_soa3_to_aos, B, T, n, ctx->pcap, ctx->force, stage);
    } else {
        ctx->set_error("pb_download_real: unknown property " + nm);
        return -1;
    }
    PB_CHECK(cudaMemcpyAsync(out, stage, sizeof(double) * 3 * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int pb_download_int(pb_ctx *ctx, const char *name, int *out, int with_ghosts) {
    PB_CHECK(cudaSetDevice(ctx->device));
    int n = ctx->nlocal + (with_ghosts ? ctx->nghost : 0);
    const std::string nm(name);
    const int *src = nullptr;
    if(nm == "type") { src = ctx->type; }
    else if(nm == "flags") { src = ctx->flags; }
    else if(nm == "uid") { src = ctx->uid; }
    else if(nm == "shape") { src = ctx->shape; }
    else if(nm == "tag") { src = ctx->tag; }
    else if(nm == "particle_cell") { src = ctx->particle_cell; n = std::min(n, ctx->cells_n); }
    else if(nm == "numneighs") { src = ctx->numneigh; n = std::min(ctx->nlocal, std::max(ctx->neigh_n, ctx->tiles_n)); }
    else { ctx->set_error("pb_download_int: unknown property " + nm); return -1; }
    if(n == 0) { return 0; }
    PB_CHECK(cudaMemcpyAsync(out, src, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int pb_download_ghost_map(pb_ctx *ctx, int *src, int *mult) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(ctx->nsend_all == 0) { return 0; }
    PB_CHECK(cudaMemcpyAsync(src, ctx->send_map, sizeof(int) * (size_t) ctx->nsend_all, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(mult, ctx->send_mult, sizeof(int) * 3 * (size_t) ctx->nsend_all, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int pb_synchronize_device(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int pb_timers_enable(pb_ctx *ctx, int on) { ctx->timers_on = on != 0; return 0; }

static void pb_timer_collect(pb_ctx *ctx, PbTimer &t) {
    for(auto &pr : t.pending) {
        cudaEventSynchronize(pr.second);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, pr.first, pr.second);
        t.ms += ms;
        t.calls += 1;
        ctx->event_pool.push_back(pr.first);
        ctx->event_pool.push_back(pr.second);
    }
    t.pending.clear();
}

extern "C" int pb_timers_reset(pb_ctx *ctx) {
    for(auto &kv : ctx->timers) { pb_timer_collect(ctx, kv.second); }
    ctx->timers.clear();
    return 0;
}

extern "C" int pb_timers_get(pb_ctx *ctx, const char *name, double *ms, long *calls) {
    auto it = ctx->timers.find(name);
    if(it == ctx->timers.end()) { *ms = 0.0; *calls = 0; return 0; }
    pb_timer_collect(ctx, it->second);
    *ms = it->second.ms;
    *calls = it->second.calls;
    return 0;
}

// whole-region device timing on the launching stream
extern "C" int pb_stream_timer_start(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PB_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    return 0;
}

extern "C" int pb_stream_timer_stop(pb_ctx *ctx, double *ms) {
    PB_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    PB_CHECK(cudaEventSynchronize(ctx->ev1));
    float f = 0.f;
    PB_CHECK(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
    *ms = (double) f;
    return 0;
}

extern "C" long pb_kernel_launches(const pb_ctx *ctx) { return ctx->launches; }

// Page-lock / unlock a caller-owned host buffer so that uploads and downloads through the C-ABI are true DMA
// transfers (bench.py's end-to-end leg).
extern "C" int pb_host_register(pb_ctx *ctx, void *ptr, size_t bytes) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PB_CHECK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return 0;
}

extern "C" int pb_host_unregister(pb_ctx *ctx, void *ptr) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PB_CHECK(cudaHostUnregister(ptr));
    return 0;
}
