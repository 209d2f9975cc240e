// DEM path of examples/dem.py on the GPU: spheres + half-spaces, linear spring-dashpot contacts with tangential history.
//
// Replaces the generated modules update_mass_and_inertia / gravity / linear_spring_dashpot / euler /
// reset_contact_history_usage_status / clear_unused_contact_history (examples/dem.py:6-91, sim/contact_history.py:75-127,
// mapping/funcs.py:230-263) and the set-up function pairs::dem_sc_grid (runtime/dem_sc_grid.hpp:62-172).
// The per-pair and per-particle arithmetic lives in dem_math.h (bit-identical to the reference's generated code).
//
// Traversal = the reference's cell-list traversal (no Verlet lists in dem.py): for every local, non-FIXED particle i the
// sphere sweep over cell 0 + the 27 stencil cells, then the half-space sweep (sim/interaction.py:91-92: shape loop
// outermost).  Contacts are keyed by the partner's uid in a per-particle table of `ccontacts` slots, stored slot-major
// ([slot][particle]) so the 32 lanes of a warp touch consecutive addresses.  Slot order and the order in which a particle's
// pair forces are summed follow OUR cell-list order, not the reference's: contact history is compared as per-uid sets,
// forces to 1e-12.
//
// Particles are not re-sorted at every reneighbouring in DEM (contact rows stay with their particle index; every `dem_sort_every`
// iterations pb_dem_sort_locals puts them into cell order, rows and all).  Between GPUs a migrating
// particle takes its whole state along -- DEM properties and the contact table -- in one fixed-size record (migrate.cu);
// unlike the reference (whose pack_contact_history runs after the leaver's slot was overwritten)
// the history that arrives is the particle's own.
#include <algorithm>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include <type_traits>

#include "ctx.cuh"
#include "dem_math.h"

int pb_materialise_force_reset(pb_ctx *ctx);

static PbDemParams pb_dem_params(const pb_ctx *ctx) {
    PbDemParams P;
    memcpy(&P, ctx->dem_params, sizeof(PbDemParams));
    return P;
}

// ---- allocation -------------------------------------------------------------------------------------------------
int pb_dem_grow(pb_ctx *ctx, size_t oldcap, size_t newcap, size_t used) {
    const int C = ctx->ccontacts;
    PB_TRY(pb_regrow(ctx, &ctx->radius, used, newcap, true));
    PB_TRY(pb_regrow_soa(ctx, &ctx->angvel, 3, oldcap, newcap, used, true));
    PB_TRY(pb_regrow_soa(ctx, &ctx->torque, 3, oldcap, newcap, used, true));
    PB_TRY(pb_regrow_soa(ctx, &ctx->normal, 3, oldcap, newcap, used, true));
    PB_TRY(pb_regrow_soa(ctx, &ctx->inv_inertia, 9, oldcap, newcap, used, true));
    PB_TRY(pb_regrow_soa(ctx, &ctx->rotmat, 9, oldcap, newcap, used, true));
    PB_TRY(pb_regrow_soa(ctx, &ctx->quat, 4, oldcap, newcap, used, true));
    // contact tables: [slot][particle] -> strided move
    auto grow_int = [&](int **p, int comps) -> int {
        int *q = nullptr;
        PB_CHECK(cudaMalloc(&q, sizeof(int) * comps * newcap));
        PB_CHECK(cudaMemsetAsync(q, 0, sizeof(int) * comps * newcap, ctx->stream));
        if(*p != nullptr && used > 0) {
            for(int c = 0; c < comps; c++) {
                PB_CHECK(cudaMemcpyAsync(q + c * newcap, *p + c * oldcap, sizeof(int) * used, cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
        if(*p != nullptr) { PB_CHECK(cudaFree(*p)); }
        *p = q;
        return 0;
    };
    PB_TRY(grow_int(&ctx->num_contacts, 1));
    PB_TRY(grow_int(&ctx->contact_uid, C));
    PB_TRY(grow_int(&ctx->contact_used, C));
    PB_TRY(grow_int(&ctx->contact_stick, C));
    PB_TRY(pb_regrow_soa(ctx, &ctx->contact_tsd, 3 * C, oldcap, newcap, used, true));
    PB_TRY(pb_regrow_soa(ctx, &ctx->contact_ivm, C, oldcap, newcap, used, true));
    if(ctx->cx > 0) { PB_TRY(pb_regrow_soa(ctx, &ctx->contact_x, ctx->cx * C, oldcap, newcap, used, true)); }
    return 0;
}

// use_contact_history=True path of pairs.simulation(): allocates the DEM property set.  contact_capacity = neighbor_capacity.
extern "C" int pb_dem_enable(pb_ctx *ctx, int contact_capacity) { return pb_dem_enable_ex(ctx, contact_capacity, 0, nullptr); }

// ... with `extra_lanes` further double lanes per contact (contact properties beyond dem.py's three, used by contact models
// compiled at run time) and the values a fresh contact starts from
extern "C" int pb_dem_enable_ex(pb_ctx *ctx, int contact_capacity, int extra_lanes, const double *extra_defaults) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(ctx->dem) { return 0; }
    if(contact_capacity < 1 || contact_capacity > 64) { ctx->set_error("pb_dem_enable: 1 <= contact_capacity <= 64"); return -1; }
    if(extra_lanes < 0 || extra_lanes > 16) { ctx->set_error("pb_dem_enable_ex: 0 <= extra_lanes <= 16"); return -1; }
    ctx->cx = extra_lanes;
    for(int k = 0; k < extra_lanes; k++) { ctx->cx_default[k] = (extra_defaults != nullptr) ? extra_defaults[k] : 0.0; }
    ctx->ccontacts = contact_capacity;
    ctx->dem = true;
    if(ctx->send_cap > 0) {        // wire records grow (contact history travels with migrating particles): re-size the buffers
        const int keep = ctx->send_cap;
        ctx->send_cap = 0;
        PB_TRY(pb_ensure_send_capacity(ctx, keep));
    }
    if(ctx->recv_buf != nullptr) { PB_CHECK(cudaFree(ctx->recv_buf)); ctx->recv_buf = nullptr; ctx->recv_cap = 0; }
    ctx->cells_set = false;       // DEM bins without z slabs: cell arrays are re-sized by the next pb_setup_cells
    PB_CHECK(cudaMalloc(&ctx->d_dem_flag, sizeof(int) * 4));
    PB_CHECK(cudaMemsetAsync(ctx->d_dem_flag, 0, sizeof(int) * 4, ctx->stream));
    if(ctx->pcap > 0) { PB_TRY(pb_dem_grow(ctx, (size_t) ctx->pcap, (size_t) ctx->pcap, (size_t) ctx->nlocal + ctx->nghost)); }
    return 0;
}

// symbols of the dem.py kernels + feature properties friction_static / friction_dynamic [ntypes*ntypes]
extern "C" int pb_dem_set_params(pb_ctx *ctx, double dt, double pi, double kappa, double ln_dry_res_coeff, double collision_time,
                                 double density_particle, double density_fluid, double gravity, int ntypes,
                                 const double *friction_static, const double *friction_dynamic) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(!ctx->dem) { ctx->set_error("pb_dem_set_params: call pb_dem_enable first"); return -1; }
    if(ntypes < 1 || ntypes > 8) { ctx->set_error("pb_dem_set_params: 1 <= ntypes <= 8"); return -1; }
    PbDemParams P;
    P.dt = dt;
    P.c_sum = pi * pi + ln_dry_res_coeff * ln_dry_res_coeff;
    P.ct2 = collision_time * collision_time;
    P.ct = collision_time;
    P.ln_coeff = ln_dry_res_coeff;
    P.kappa = kappa;
    P.sqrt_kappa = sqrt(kappa);
    P.grav_coeff = density_particle - density_fluid;
    P.gravity = gravity;
    P.pi = pi;
    static_assert(sizeof(PbDemParams) <= sizeof(double) * 16, "PbDemParams too large");
    memcpy(ctx->dem_params, &P, sizeof(PbDemParams));
    ctx->dem_ntypes = ntypes;
    if(ctx->d_fric_static == nullptr) {
        PB_CHECK(cudaMalloc(&ctx->d_fric_static, sizeof(double) * 64));
        PB_CHECK(cudaMalloc(&ctx->d_fric_dynamic, sizeof(double) * 64));
    }
    PB_CHECK(cudaMemcpyAsync(ctx->d_fric_static, friction_static, sizeof(double) * ntypes * ntypes, cudaMemcpyHostToDevice, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(ctx->d_fric_dynamic, friction_dynamic, sizeof(double) * ntypes * ntypes, cudaMemcpyHostToDevice, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));      // (stream-ordered with the kernels that read them; the caller's arrays are free again)
    return 0;
}

// ---- set-up: pairs::dem_sc_grid (runtime/dem_sc_grid.hpp:62-172), host side -------------------------------------------
// Same libstdc++ generator (std::mt19937, default seed) and distribution as the reference, hence the same stream.  Particle
// types use the private rand() stream of setup.cu's convention (seed 1).  Two-call pattern: out pointers may be NULL to
// count.  Arrays are the reference's AoS host layout.
extern "C" int pb_dem_sc_grid(pb_ctx *ctx, double xmax, double ymax, double zmax, double spacing, double diameter, double min_diameter,
                              double max_diameter, double initial_velocity, double particle_density, int ntypes, int capacity,
                              int *uid, int *type, double *mass, double *radius, double *position, double *velocity, int *count) {
    if(!ctx->domain_set) { ctx->set_error("pb_dem_sc_grid: domain not initialised"); return -1; }
    std::mt19937 generator;
    struct random_data rd;
    char statebuf[128];
    memset(&rd, 0, sizeof(rd));
    memset(statebuf, 0, sizeof(statebuf));
    initstate_r(1, statebuf, sizeof(statebuf), &rd);
    auto real_random = [&](double lo, double hi) {
        std::uniform_real_distribution<double> distribution(lo, hi);
        return distribution(generator);
    };
    auto within = [&](const double *p, const double *aabb) {
        return p[0] >= aabb[0] && p[0] < aabb[3] && p[1] >= aabb[1] && p[1] < aabb[4] && p[2] >= aabb[2] && p[2] < aabb[5];
    };
    auto in_subdomain = [&](double x, double y, double z) {
        return x >= ctx->subdom[0] && x < ctx->subdom[1] - 0.00001 && y >= ctx->subdom[2] && y < ctx->subdom[3] - 0.00001 &&
               z >= ctx->subdom[4] && z < ctx->subdom[5] - 0.00001;
    };
    int last_uid = 1, n = 0;
    const double xmin = 0.0, ymin = 0.0, zmin = diameter;
    double gen_domain[] = {xmin, ymin, zmin, xmax, ymax, zmax};
    double ref_point[] = {spacing * 0.5, spacing * 0.5, spacing * 0.5};
    const int iret = (int) (ceil((xmin - ref_point[0]) / spacing));
    const int jret = (int) (ceil((ymin - ref_point[1]) / spacing));
    const int kret = (int) (ceil((zmin - ref_point[2]) / spacing));
    int i = iret, j = jret, k = kret;
    double point[3] = {ref_point[0] + i * spacing, ref_point[1] + j * spacing, ref_point[2] + k * spacing};
    while(within(point, gen_domain)) {
        const double diam = real_random(min_diameter, max_diameter);
        if(in_subdomain(point[0], point[1], point[2])) {
            const double rad = diam * 0.5;
            const double vx = 0.1 * real_random(-initial_velocity, initial_velocity);
            const double vy = 0.1 * real_random(-initial_velocity, initial_velocity);
            int32_t r = 0;
            random_r(&rd, &r);
            if(uid != nullptr) {
                if(n >= capacity) { ctx->set_error("pb_dem_sc_grid: capacity too small"); return -1; }
                uid[n] = last_uid;
                radius[n] = rad;
                mass[n] = ((4.0 / 3.0) * M_PI) * rad * rad * rad * particle_density;
                position[n * 3] = point[0]; position[n * 3 + 1] = point[1]; position[n * 3 + 2] = point[2];
                velocity[n * 3] = vx; velocity[n * 3 + 1] = vy; velocity[n * 3 + 2] = -initial_velocity;
                type[n] = (int) (r % ntypes);
            }
            n++;
        }
        ++i;
        point[0] = ref_point[0] + i * spacing; point[1] = ref_point[1] + j * spacing; point[2] = ref_point[2] + k * spacing;
        if(!within(point, gen_domain)) {
            i = iret; j++;
            point[0] = ref_point[0] + i * spacing; point[1] = ref_point[1] + j * spacing; point[2] = ref_point[2] + k * spacing;
            if(!within(point, gen_domain)) {
                j = jret; k++;
                point[0] = ref_point[0] + i * spacing; point[1] = ref_point[1] + j * spacing; point[2] = ref_point[2] + k * spacing;
                if(!within(point, gen_domain)) { break; }
            }
        }
        last_uid++;
    }
    *count = n;
    return 0;
}

// ---- generic upload / download of the DEM properties (reference AoS host layout) ------------------------------------
struct PbDemProp {
    double *ptr;
    int comps;
};

static bool pb_dem_prop(pb_ctx *ctx, const std::string &nm, PbDemProp *out) {
    if(nm == "radius") { *out = {ctx->radius, 1}; return true; }
    if(nm == "angular_velocity") { *out = {ctx->angvel, 3}; return true; }
    if(nm == "torque") { *out = {ctx->torque, 3}; return true; }
    if(nm == "normal") { *out = {ctx->normal, 3}; return true; }
    if(nm == "inv_inertia") { *out = {ctx->inv_inertia, 9}; return true; }
    if(nm == "rotation_matrix") { *out = {ctx->rotmat, 9}; return true; }
    if(nm == "rotation_quat") { *out = {ctx->quat, 4}; return true; }
    if(nm == "force") { *out = {ctx->force, 3}; return true; }
    if(nm == "mass") { *out = {ctx->mass, 1}; return true; }
    if(nm == "linear_velocity") { *out = {ctx->vel, 3}; return true; }
    return false;
}

__global__ void pb_k_aos_to_soa(int n, int cap, int comps, const double *__restrict__ aos, double *__restrict__ soa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) { for(int c = 0; c < comps; c++) { soa[(size_t) c * cap + i] = aos[(size_t) i * comps + c]; } }
}

__global__ void pb_k_soa_to_aos(int n, int cap, int comps, const double *__restrict__ soa, double *__restrict__ aos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) { for(int c = 0; c < comps; c++) { aos[(size_t) i * comps + c] = soa[(size_t) c * cap + i]; } }
}

// n particles starting at index `first` (so ghosts can be placed explicitly by module-level tests)
extern "C" int pb_dem_upload_real(pb_ctx *ctx, const char *name, int first, int n, const double *data) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbDemProp pr;
    if(!ctx->dem || !pb_dem_prop(ctx, name, &pr)) { ctx->set_error(std::string("pb_dem_upload_real: unknown property ") + name); return -1; }
    if(first + n > ctx->pcap) { ctx->set_error("pb_dem_upload_real: beyond capacity"); return -1; }
    if(n == 0) { return 0; }
    PbScratch stage_buf;
    PB_CHECK(stage_buf.alloc(sizeof(double) * (size_t) n * pr.comps));
    double *const stage = stage_buf.as<double>();
    PB_CHECK(cudaMemcpyAsync(stage, data, sizeof(double) * (size_t) n * pr.comps, cudaMemcpyHostToDevice, ctx->stream));
    PB_LAUNCH(pb_k_aos_to_soa, pb_blocks(n, 256), 256, n, ctx->pcap, pr.comps, stage, pr.ptr + first);
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    if(std::string(name) == "force") { ctx->force_is_zero = false; }
    return 0;
}

extern "C" int pb_dem_download_real(pb_ctx *ctx, const char *name, int first, int n, double *out) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbDemProp pr;
    if(!ctx->dem || !pb_dem_prop(ctx, name, &pr)) { ctx->set_error(std::string("pb_dem_download_real: unknown property ") + name); return -1; }
    if(n == 0) { return 0; }
    if(std::string(name) == "force" || std::string(name) == "torque") { PB_TRY(pb_materialise_force_reset(ctx)); }
    PbScratch stage_buf;
    PB_CHECK(stage_buf.alloc(sizeof(double) * (size_t) n * pr.comps));
    double *const stage = stage_buf.as<double>();
    PB_LAUNCH(pb_k_soa_to_aos, pb_blocks(n, 256), 256, n, ctx->pcap, pr.comps, pr.ptr + first, stage);
    PB_CHECK(cudaMemcpyAsync(out, stage, sizeof(double) * (size_t) n * pr.comps, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// explicit particle counts (module-level tests upload locals AND ghosts of a reference snapshot)
extern "C" int pb_set_counts(pb_ctx *ctx, int nlocal, int nghost) {
    if(nlocal + nghost > ctx->pcap) { ctx->set_error("pb_set_counts: beyond capacity"); return -1; }
    ctx->nlocal = nlocal;
    ctx->nghost = nghost;
    ctx->cells_n = 0;
    return 0;
}

// contact history in the reference's host layout: num[n], uid/sticking [n][C], tsd [n][C][3], ivm [n][C]
__global__ void pb_k_contacts_in(int n, int cap, int C, const int *__restrict__ num, const int *__restrict__ uid,
                                 const int *__restrict__ stick, const double *__restrict__ tsd, const double *__restrict__ ivm,
                                 int *num_o, int *uid_o, int *used_o, int *stick_o, double *tsd_o, double *ivm_o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    num_o[i] = num[i];
    for(int c = 0; c < C; c++) {
        uid_o[(size_t) c * cap + i] = uid[(size_t) i * C + c];
        used_o[(size_t) c * cap + i] = 0;
        stick_o[(size_t) c * cap + i] = stick[(size_t) i * C + c];
        ivm_o[(size_t) c * cap + i] = ivm[(size_t) i * C + c];
        for(int d = 0; d < 3; d++) { tsd_o[((size_t) d * C + c) * cap + i] = tsd[((size_t) i * C + c) * 3 + d]; }
    }
}

__global__ void pb_k_contacts_out(int n, int cap, int C, const int *num_i, const int *uid_i, const int *used_i, const int *stick_i,
                                  const double *tsd_i, const double *ivm_i, int *__restrict__ num, int *__restrict__ uid,
                                  int *__restrict__ used, int *__restrict__ stick, double *__restrict__ tsd, double *__restrict__ ivm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    num[i] = num_i[i];
    for(int c = 0; c < C; c++) {
        uid[(size_t) i * C + c] = uid_i[(size_t) c * cap + i];
        used[(size_t) i * C + c] = used_i[(size_t) c * cap + i];
        stick[(size_t) i * C + c] = stick_i[(size_t) c * cap + i];
        ivm[(size_t) i * C + c] = ivm_i[(size_t) c * cap + i];
        for(int d = 0; d < 3; d++) { tsd[((size_t) i * C + c) * 3 + d] = tsd_i[((size_t) d * C + c) * cap + i]; }
    }
}

extern "C" int pb_dem_upload_contacts(pb_ctx *ctx, int n, const int *num, const int *uid, const int *sticking, const double *tsd,
                                      const double *ivm) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(!ctx->dem || n > ctx->pcap) { ctx->set_error("pb_dem_upload_contacts: DEM not enabled / beyond capacity"); return -1; }
    if(n == 0) { return 0; }
    const int C = ctx->ccontacts;
    PbScratch b_num, b_uid, b_st, b_tsd, b_ivm;
    PB_CHECK(b_num.alloc(sizeof(int) * n));
    PB_CHECK(b_uid.alloc(sizeof(int) * (size_t) n * C));
    PB_CHECK(b_st.alloc(sizeof(int) * (size_t) n * C));
    PB_CHECK(b_tsd.alloc(sizeof(double) * (size_t) n * C * 3));
    PB_CHECK(b_ivm.alloc(sizeof(double) * (size_t) n * C));
    int *d_num = b_num.as<int>(), *d_uid = b_uid.as<int>(), *d_st = b_st.as<int>();
    double *d_tsd = b_tsd.as<double>(), *d_ivm = b_ivm.as<double>();
    // on the context's stream: a synchronous cudaMemcpy from pageable memory may return before its DMA has landed, and the legacy
    // stream it runs on does not order against this (non-blocking) stream -- the kernel below was seen reading the last array
    // before it had arrived (one contact of 115 with a zero impact velocity, once in many runs)
    PB_CHECK(cudaMemcpyAsync(d_num, num, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(d_uid, uid, sizeof(int) * (size_t) n * C, cudaMemcpyHostToDevice, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(d_st, sticking, sizeof(int) * (size_t) n * C, cudaMemcpyHostToDevice, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(d_tsd, tsd, sizeof(double) * (size_t) n * C * 3, cudaMemcpyHostToDevice, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(d_ivm, ivm, sizeof(double) * (size_t) n * C, cudaMemcpyHostToDevice, ctx->stream));
    PB_LAUNCH(pb_k_contacts_in, pb_blocks(n, 128), 128, n, ctx->pcap, C, d_num, d_uid, d_st, d_tsd, d_ivm, ctx->num_contacts,
              ctx->contact_uid, ctx->contact_used, ctx->contact_stick, ctx->contact_tsd, ctx->contact_ivm);
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// the extra lanes of the contact rows (pb_dem_enable_ex): host layout [n][contact_capacity][extra_lanes]
__global__ void __launch_bounds__(128) pb_k_contact_extras(int n, int cap, int C, int nx, int to_device, double *__restrict__ host_layout,
                                                           double *__restrict__ cx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    for(int c = 0; c < C; c++) {
        for(int x = 0; x < nx; x++) {
            double *h = host_layout + ((size_t) i * C + c) * nx + x, *d = cx + ((size_t) x * C + c) * cap + i;
            if(to_device) { *d = *h; } else { *h = *d; }
        }
    }
}

extern "C" int pb_dem_contact_extras(pb_ctx *ctx, int n, double *values, int upload) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(!ctx->dem || n > ctx->pcap) { ctx->set_error("pb_dem_contact_extras: DEM not enabled / beyond capacity"); return -1; }
    if(n == 0 || ctx->cx == 0) { return 0; }
    const size_t bytes = sizeof(double) * (size_t) n * ctx->ccontacts * ctx->cx;
    PbScratch buf;
    PB_CHECK(buf.alloc(bytes));
    if(upload) { PB_CHECK(cudaMemcpyAsync(buf.as<double>(), values, bytes, cudaMemcpyHostToDevice, ctx->stream)); }
    PB_LAUNCH(pb_k_contact_extras, pb_blocks(n, 128), 128, n, ctx->pcap, ctx->ccontacts, ctx->cx, upload, buf.as<double>(), ctx->contact_x);
    if(!upload) { PB_CHECK(cudaMemcpyAsync(values, buf.as<double>(), bytes, cudaMemcpyDeviceToHost, ctx->stream)); }
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int pb_dem_contact_extra_lanes(const pb_ctx *ctx) { return ctx->cx; }

extern "C" int pb_dem_download_contacts(pb_ctx *ctx, int n, int *num, int *uid, int *used, int *sticking, double *tsd, double *ivm) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(!ctx->dem || n > ctx->pcap) { ctx->set_error("pb_dem_download_contacts: DEM not enabled / beyond capacity"); return -1; }
    if(n == 0) { return 0; }
    const int C = ctx->ccontacts;
    PbScratch b_num, b_uid, b_us, b_st, b_tsd, b_ivm;
    PB_CHECK(b_num.alloc(sizeof(int) * n));
    PB_CHECK(b_uid.alloc(sizeof(int) * (size_t) n * C));
    PB_CHECK(b_us.alloc(sizeof(int) * (size_t) n * C));
    PB_CHECK(b_st.alloc(sizeof(int) * (size_t) n * C));
    PB_CHECK(b_tsd.alloc(sizeof(double) * (size_t) n * C * 3));
    PB_CHECK(b_ivm.alloc(sizeof(double) * (size_t) n * C));
    int *d_num = b_num.as<int>(), *d_uid = b_uid.as<int>(), *d_us = b_us.as<int>(), *d_st = b_st.as<int>();
    double *d_tsd = b_tsd.as<double>(), *d_ivm = b_ivm.as<double>();
    PB_LAUNCH(pb_k_contacts_out, pb_blocks(n, 128), 128, n, ctx->pcap, C, ctx->num_contacts, ctx->contact_uid, ctx->contact_used,
              ctx->contact_stick, ctx->contact_tsd, ctx->contact_ivm, d_num, d_uid, d_us, d_st, d_tsd, d_ivm);
    PB_CHECK(cudaMemcpyAsync(num, d_num, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(uid, d_uid, sizeof(int) * (size_t) n * C, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(used, d_us, sizeof(int) * (size_t) n * C, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(sticking, d_st, sizeof(int) * (size_t) n * C, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(tsd, d_tsd, sizeof(double) * (size_t) n * C * 3, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(ivm, d_ivm, sizeof(double) * (size_t) n * C, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- migration support (called from migrate.cu) -------------------------------------------------------------------------
// DEM part of an exchange record, appended after the 12 base elements: radius, angvel[3], normal[3], inv_inertia[9],
// rotmat[9], quat[4], num_contacts, then per slot: uid, sticking, tsd[3], ivm.
struct PbDemArrays {
    double *radius, *angvel, *normal, *inv_inertia, *rotmat, *quat, *c_tsd, *c_ivm, *c_x;
    int *num_contacts, *c_uid, *c_used, *c_stick;
    int cap, C, nx;
};

static PbDemArrays pb_dem_arrays(const pb_ctx *ctx) {
    PbDemArrays a;
    a.radius = ctx->radius; a.angvel = ctx->angvel; a.normal = ctx->normal; a.inv_inertia = ctx->inv_inertia; a.rotmat = ctx->rotmat;
    a.quat = ctx->quat; a.c_tsd = ctx->contact_tsd; a.c_ivm = ctx->contact_ivm; a.num_contacts = ctx->num_contacts;
    a.c_uid = ctx->contact_uid; a.c_used = ctx->contact_used; a.c_stick = ctx->contact_stick; a.cap = ctx->pcap; a.C = ctx->ccontacts;
    a.c_x = ctx->contact_x; a.nx = ctx->cx;
    return a;
}

__device__ __forceinline__ void pb_dem_pack_one(const PbDemArrays &a, int p, double *b) {
    const size_t cap = a.cap;
    int k = 0;
    b[k++] = a.radius[p];
    for(int d = 0; d < 3; d++) { b[k++] = a.angvel[d * cap + p]; }
    for(int d = 0; d < 3; d++) { b[k++] = a.normal[d * cap + p]; }
    for(int d = 0; d < 9; d++) { b[k++] = a.inv_inertia[d * cap + p]; }
    for(int d = 0; d < 9; d++) { b[k++] = a.rotmat[d * cap + p]; }
    for(int d = 0; d < 4; d++) { b[k++] = a.quat[d * cap + p]; }
    const int nc = a.num_contacts[p];
    b[k++] = (double) nc;
    for(int c = 0; c < a.C; c++) {
        const bool live = c < nc;
        b[k++] = live ? (double) a.c_uid[c * cap + p] : 0.0;
        b[k++] = live ? (double) a.c_stick[c * cap + p] : 0.0;
        for(int d = 0; d < 3; d++) { b[k++] = live ? a.c_tsd[((size_t) d * a.C + c) * cap + p] : 0.0; }
        b[k++] = live ? a.c_ivm[c * cap + p] : 0.0;
        for(int x = 0; x < a.nx; x++) { b[k++] = live ? a.c_x[((size_t) x * a.C + c) * cap + p] : 0.0; }
    }
}

__device__ __forceinline__ void pb_dem_unpack_one(const PbDemArrays &a, int p, const double *b) {
    const size_t cap = a.cap;
    int k = 0;
    a.radius[p] = b[k++];
    for(int d = 0; d < 3; d++) { a.angvel[d * cap + p] = b[k++]; }
    for(int d = 0; d < 3; d++) { a.normal[d * cap + p] = b[k++]; }
    for(int d = 0; d < 9; d++) { a.inv_inertia[d * cap + p] = b[k++]; }
    for(int d = 0; d < 9; d++) { a.rotmat[d * cap + p] = b[k++]; }
    for(int d = 0; d < 4; d++) { a.quat[d * cap + p] = b[k++]; }
    const int nc = (int) b[k++];
    a.num_contacts[p] = nc;
    for(int c = 0; c < a.C; c++) {
        a.c_uid[c * cap + p] = (int) b[k++];
        a.c_stick[c * cap + p] = (int) b[k++];
        for(int d = 0; d < 3; d++) { a.c_tsd[((size_t) d * a.C + c) * cap + p] = b[k++]; }
        a.c_ivm[c * cap + p] = b[k++];
        for(int x = 0; x < a.nx; x++) { a.c_x[((size_t) x * a.C + c) * cap + p] = b[k++]; }
        a.c_used[c * cap + p] = (c < nc) ? 1 : 0;      // as the reference's unpack marks transferred contacts used
    }
}

// records of the leavers: e = record index (same mapping as the base pack kernel of migrate.cu)
__global__ void __launch_bounds__(128) pb_k_dem_pack_exchange(int n, int stride, const int *__restrict__ rec, PbDemArrays a,
                                                              double *__restrict__ buf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    const int e = rec[i];
    if(e < 0) { return; }
    pb_dem_pack_one(a, i, buf + (size_t) e * stride + 12);
}

__global__ void __launch_bounds__(128) pb_k_dem_unpack_exchange(int count, int dst0, int stride, PbDemArrays a, const double *__restrict__ buf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= count) { return; }
    pb_dem_unpack_one(a, dst0 + k, buf + (size_t) k * stride + 12);
}

// hole filling: particle `src` (a stayer from the tail) moves into slot `dst` (a leaver's slot below the new nlocal)
__global__ void __launch_bounds__(128) pb_k_dem_move(const int *__restrict__ count, const int *__restrict__ src_idx,
                                                     const int *__restrict__ dst_idx, PbDemArrays a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= *count) { return; }
    const int s = src_idx[k], t = dst_idx[k];
    const size_t cap = a.cap;
    a.radius[t] = a.radius[s];
    for(int d = 0; d < 3; d++) { a.angvel[d * cap + t] = a.angvel[d * cap + s]; a.normal[d * cap + t] = a.normal[d * cap + s]; }
    for(int d = 0; d < 9; d++) { a.inv_inertia[d * cap + t] = a.inv_inertia[d * cap + s]; a.rotmat[d * cap + t] = a.rotmat[d * cap + s]; }
    for(int d = 0; d < 4; d++) { a.quat[d * cap + t] = a.quat[d * cap + s]; }
    a.num_contacts[t] = a.num_contacts[s];
    for(int c = 0; c < a.C; c++) {
        a.c_uid[c * cap + t] = a.c_uid[c * cap + s];
        a.c_used[c * cap + t] = a.c_used[c * cap + s];
        a.c_stick[c * cap + t] = a.c_stick[c * cap + s];
        a.c_ivm[c * cap + t] = a.c_ivm[c * cap + s];
        for(int d = 0; d < 3; d++) { a.c_tsd[((size_t) d * a.C + c) * cap + t] = a.c_tsd[((size_t) d * a.C + c) * cap + s]; }
        for(int x = 0; x < a.nx; x++) { a.c_x[((size_t) x * a.C + c) * cap + t] = a.c_x[((size_t) x * a.C + c) * cap + s]; }
    }
}

int pb_dem_pack_exchange(pb_ctx *ctx, int n, int stride, const int *rec) {
    PB_LAUNCH(pb_k_dem_pack_exchange, pb_blocks(n, 128), 128, n, stride, rec, pb_dem_arrays(ctx), ctx->send_buf);
    return 0;
}

int pb_dem_unpack_exchange(pb_ctx *ctx, int count, int dst0, int stride, const double *src) {
    PB_LAUNCH(pb_k_dem_unpack_exchange, pb_blocks(count, 128), 128, count, dst0, stride, pb_dem_arrays(ctx), src);
    return 0;
}

int pb_dem_move(pb_ctx *ctx, int max_count, const int *count, const int *src_idx, const int *dst_idx) {
    PB_LAUNCH(pb_k_dem_move, pb_blocks(max_count, 128), 128, count, src_idx, dst_idx, pb_dem_arrays(ctx));
    return 0;
}

// ---- spatial re-sort of the locals --------------------------------------------------------------------------------------
// dem.py rebuilds its cell lists every iteration but never reorders particles; after the bed has mixed, a particle's 27
// stencil cells point all over memory (ncu, settled 1 M-sphere bed: 6.6 % L1 hit rate in the detection kernel, L2 at 72 %).
// Every `dem_sort_every` iterations the locals are therefore put into cell order, each DEM property and the whole contact
// table travelling with its particle: 109 double rows and 61 int rows of [pcap] are gathered through the permutation, six rows
// at a time, into the (volatile, currently unused) force / torque arrays and copied back -- about 1 ms per million particles,
// i.e. ~1 % of the step time at the default interval.  Only the ORDER in which a particle meets its partners changes.
__global__ void __launch_bounds__(256) pb_k_gather_rows_f64(int n, size_t cap, const int *__restrict__ perm, const double *__restrict__ src,
                                                            double *__restrict__ dst) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= n) { return; }
    const size_t row = blockIdx.y;
    dst[row * cap + k] = src[row * cap + perm[k]];
}

__global__ void __launch_bounds__(256) pb_k_gather_rows_i32(int n, size_t cap, const int *__restrict__ perm, const int *__restrict__ src,
                                                            int *__restrict__ dst) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= n) { return; }
    const size_t row = blockIdx.y;
    dst[row * cap + k] = src[row * cap + perm[k]];
}

static int pb_permute_f64(pb_ctx *ctx, double *base, int rows, const int *perm, int n) {
    const size_t cap = (size_t) ctx->pcap;
    for(int r0 = 0; r0 < rows; r0 += 3) {
        const int c = std::min(3, rows - r0);
        double *scratch = ctx->force;           // [3][pcap]
        pb_k_gather_rows_f64<<<dim3(pb_blocks(n, 256), c), 256, 0, ctx->stream>>>(n, cap, perm, base + (size_t) r0 * cap, scratch);
        ctx->launches++;
        PB_CHECK(cudaGetLastError());
        PB_CHECK(cudaMemcpy2DAsync(base + (size_t) r0 * cap, cap * sizeof(double), scratch, cap * sizeof(double), (size_t) n * sizeof(double), c,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return 0;
}

static int pb_permute_i32(pb_ctx *ctx, int *base, int rows, const int *perm, int n) {
    const size_t cap = (size_t) ctx->pcap;
    int *scratch = (int *) ctx->torque;          // [3][pcap] doubles = 6 int rows
    for(int r0 = 0; r0 < rows; r0 += 6) {
        const int c = std::min(6, rows - r0);
        pb_k_gather_rows_i32<<<dim3(pb_blocks(n, 256), c), 256, 0, ctx->stream>>>(n, cap, perm, base + (size_t) r0 * cap, scratch);
        ctx->launches++;
        PB_CHECK(cudaGetLastError());
        PB_CHECK(cudaMemcpy2DAsync(base + (size_t) r0 * cap, cap * sizeof(int), scratch, cap * sizeof(int), (size_t) n * sizeof(int), c,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return 0;
}

// called between exchange and borders (no ghosts exist); force / torque hold nothing that is still needed (volatile, reset
// before the next evaluation)
int pb_dem_sort_locals(pb_ctx *ctx) {
    const int n = ctx->nlocal;
    if(n == 0) { return 0; }
    PbStage st(ctx, "dem_sort");
    PB_TRY(pb_sort_locals(ctx));                 // base arrays; leaves the permutation (new index -> old index) in cell_list
    const int *perm = ctx->cell_list;
    const int C = ctx->ccontacts;
    PB_TRY(pb_permute_f64(ctx, ctx->radius, 1, perm, n));
    PB_TRY(pb_permute_f64(ctx, ctx->angvel, 3, perm, n));
    PB_TRY(pb_permute_f64(ctx, ctx->normal, 3, perm, n));
    PB_TRY(pb_permute_f64(ctx, ctx->inv_inertia, 9, perm, n));
    PB_TRY(pb_permute_f64(ctx, ctx->rotmat, 9, perm, n));
    PB_TRY(pb_permute_f64(ctx, ctx->quat, 4, perm, n));
    PB_TRY(pb_permute_f64(ctx, ctx->contact_tsd, 3 * C, perm, n));
    PB_TRY(pb_permute_f64(ctx, ctx->contact_ivm, C, perm, n));
    if(ctx->cx > 0) { PB_TRY(pb_permute_f64(ctx, ctx->contact_x, ctx->cx * C, perm, n)); }
    PB_TRY(pb_permute_i32(ctx, ctx->num_contacts, 1, perm, n));
    PB_TRY(pb_permute_i32(ctx, ctx->contact_uid, C, perm, n));
    PB_TRY(pb_permute_i32(ctx, ctx->contact_used, C, perm, n));
    PB_TRY(pb_permute_i32(ctx, ctx->contact_stick, C, perm, n));
    ctx->force_is_zero = true;                   // the scratch rows are garbage now: make sure they are cleared before use
    ctx->cells_n = 0;
    return 0;
}

// ---- kernels ----------------------------------------------------------------------------------------------------
// update_mass_and_inertia (examples/dem.py:6-15): runs once over all locals (a setup() function: no FIXED filter)
__global__ void __launch_bounds__(256) pb_k_dem_update_mass_inertia(int n, int cap, const int *__restrict__ shape, double *__restrict__ mass,
                                                                   const double *__restrict__ radius, double *__restrict__ Iinv,
                                                                   double *__restrict__ R, double *__restrict__ q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    const double eye[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
    for(int c = 0; c < 9; c++) { R[(size_t) c * cap + i] = eye[c]; }
    q[i] = 1.0; q[(size_t) cap + i] = 0.0; q[(size_t) 2 * cap + i] = 0.0; q[(size_t) 3 * cap + i] = 0.0;
    double I[9];
    if(shape[i] == PB_SHAPE_SPHERE) {
        pb_dem_sphere_inv_inertia(mass[i], radius[i], I);
    } else {
        mass[i] = INFINITY;
        for(int c = 0; c < 9; c++) { I[c] = 0.0; }
    }
    for(int c = 0; c < 9; c++) { Iinv[(size_t) c * cap + i] = I[c]; }
}

extern "C" int pb_dem_update_mass_and_inertia(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(!ctx->dem) { ctx->set_error("DEM not enabled"); return -1; }
    if(ctx->nlocal == 0) { return 0; }
    PB_LAUNCH(pb_k_dem_update_mass_inertia, pb_blocks(ctx->nlocal, 256), 256, ctx->nlocal, ctx->pcap, ctx->shape, ctx->mass, ctx->radius,
              ctx->inv_inertia, ctx->rotmat, ctx->quat);
    return 0;
}

// gravity (examples/dem.py:88-90), compute() kernel: FIXED particles are skipped (mapping/funcs.py:305-310)
__global__ void __launch_bounds__(256) pb_k_dem_gravity(int n, int cap, PbDemParams P, const int *__restrict__ flags,
                                                       const double *__restrict__ radius, double *__restrict__ force) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n || (flags[i] & PB_FLAG_FIXED) != 0) { return; }
    force[(size_t) 2 * cap + i] = pb_dem_gravity(P, radius[i], force[(size_t) 2 * cap + i]);
}

extern "C" int pb_dem_gravity(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbStage st(ctx, "gravity");
    PB_TRY(pb_materialise_force_reset(ctx));
    if(ctx->nlocal == 0) { return 0; }
    PB_LAUNCH(pb_k_dem_gravity, pb_blocks(ctx->nlocal, 256), 256, ctx->nlocal, ctx->pcap, pb_dem_params(ctx), ctx->flags, ctx->radius,
              ctx->force);
    return 0;
}

// reset_contact_history_usage_status (sim/contact_history.py:75-87)
__global__ void __launch_bounds__(256) pb_k_dem_reset_usage(int n, int cap, const int *__restrict__ num, int *__restrict__ used) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    const int c_n = num[i];
    for(int c = 0; c < c_n; c++) { used[(size_t) c * cap + i] = 0; }
}

extern "C" int pb_dem_reset_contact_usage(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbStage st(ctx, "reset_contact_history_usage_status");
    if(ctx->nlocal == 0) { return 0; }
    PB_LAUNCH(pb_k_dem_reset_usage, pb_blocks(ctx->nlocal, 256), 256, ctx->nlocal, ctx->pcap, ctx->num_contacts, ctx->contact_used);
    return 0;
}

// clear_unused_contact_history (sim/contact_history.py:90-127, cell-list branch): unused slots are overwritten by the last slot
__global__ void __launch_bounds__(256) pb_k_dem_clear_unused(int n, int cap, int C, int *__restrict__ num, int *__restrict__ uid,
                                                            int *__restrict__ used, int *__restrict__ stick, double *__restrict__ tsd,
                                                            double *__restrict__ ivm, int nx, double *__restrict__ cx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    int c = 0, cnt = num[i];
    while(c < cnt) {
        if(used[(size_t) c * cap + i] == 0) {
            const int last = cnt - 1;
            if(last > 0) {
                stick[(size_t) c * cap + i] = stick[(size_t) last * cap + i];
                for(int d = 0; d < 3; d++) { tsd[((size_t) d * C + c) * cap + i] = tsd[((size_t) d * C + last) * cap + i]; }
                ivm[(size_t) c * cap + i] = ivm[(size_t) last * cap + i];
                for(int x = 0; x < nx; x++) { cx[((size_t) x * C + c) * cap + i] = cx[((size_t) x * C + last) * cap + i]; }
                uid[(size_t) c * cap + i] = uid[(size_t) last * cap + i];
                used[(size_t) c * cap + i] = used[(size_t) last * cap + i];
            }
            cnt--;
        } else {
            c++;
        }
    }
    num[i] = cnt;
}

extern "C" int pb_dem_clear_unused_contacts(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbStage st(ctx, "clear_unused_contact_history");
    if(ctx->nlocal == 0) { return 0; }
    PB_LAUNCH(pb_k_dem_clear_unused, pb_blocks(ctx->nlocal, 256), 256, ctx->nlocal, ctx->pcap, ctx->ccontacts, ctx->num_contacts,
              ctx->contact_uid, ctx->contact_used, ctx->contact_stick, ctx->contact_tsd, ctx->contact_ivm, ctx->cx, ctx->contact_x);
    return 0;
}

// linear_spring_dashpot (examples/dem.py:18-74) over the cell lists, in two passes (one thread per local particle each):
//   pass 1  pb_k_dem_detect   the reference's traversal (sphere sweep over cell 0 + 27 stencil cells, then the half-space sweep)
//                             with the contact geometry test only; writes the partners in contact, in traversal order.  Light
//                             (few registers, high occupancy): it is the latency-bound part -- dependent loads per candidate.
//   pass 2  pb_k_dem_force    history lookup / insert, the contact model, force and torque accumulation over the 0..~6 partners
//                             found.  Heavy arithmetic (146 registers), but no divergent search around it.
// A single fused kernel ran at 18 % occupancy and spent its time waiting on the candidate gathers (2.5 ms per step for 1M
// settled spheres); the split keeps the arithmetic and its order identical.
__global__ void __launch_bounds__(128) pb_k_dem_detect(int nlocal, int cap, int ncells, int dim1, int dim2, int maxp,
                                                       const double4 *__restrict__ pos, const double *__restrict__ radius,
                                                       const double *__restrict__ normal, const int *__restrict__ flags,
                                                       const int *__restrict__ shape, const int *__restrict__ particle_cell,
                                                       const int *__restrict__ cell_start, const int *__restrict__ cell_list,
                                                       int *__restrict__ npairs, int *__restrict__ pairs, int *__restrict__ overflow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= nlocal) { return; }
    int np = 0;
    if((flags[i] & PB_FLAG_FIXED) == 0) {
        const double4 pi4 = pb_ld_pos(pos + i);
        const double xi[3] = {pi4.x, pi4.y, pi4.z};
        const double ri = radius[i];
        const int pc = particle_cell[i];
        bool other_in_stencil = false;        // a non-sphere particle was seen in one of the 27 stencil cells
        for(int sh = 0; sh < 2; sh++) {       // shape loop outermost: spheres, then half-spaces (sim/interaction.py:91-92)
            // the half-space sweep visits the same cells as the sphere sweep: if that one saw no non-sphere particle in the
            // stencil cells, only cell 0 (where INFINITE half-spaces are binned) can contribute -- skipping is exact
            const int nruns = (sh == 0 || other_in_stencil) ? 10 : 1;
            for(int run = 0; run < nruns; run++) {
                int c_lo, c_hi;
                if(run == 0) {
                    c_lo = 0; c_hi = 0;
                } else {
                    const int r = run - 1;
                    const int mid = pc + ((r / 3 - 1) * dim1 + (r % 3 - 1)) * dim2;
                    c_lo = max(mid - 1, 1);
                    c_hi = min(mid + 1, ncells - 1);
                    if(c_lo > c_hi) { continue; }
                }
                const int b = cell_start[c_lo], e = cell_start[c_hi + 1];
                for(int k = b; k < e; k++) {
                    const int j = __ldg(cell_list + k);
                    const int sj = shape[j];
                    if(sh == 0 && run > 0 && sj != PB_SHAPE_SPHERE) { other_in_stencil = true; }
                    if(j == i || sj != sh) { continue; }
                    const double4 pj4 = pb_ld_pos(pos + j);
                    const double xj[3] = {pj4.x, pj4.y, pj4.z};
                    double n[3], cp[3], delta;
                    int hit;
                    if(sh == PB_SHAPE_SPHERE) {
                        hit = pb_dem_geom_sphere(xi, ri, xj, radius[j], n, cp, &delta);
                    } else {
                        const double nj[3] = {normal[j], normal[(size_t) cap + j], normal[(size_t) 2 * cap + j]};
                        hit = pb_dem_geom_halfspace(xi, ri, xj, nj, n, cp, &delta);
                    }
                    if(!hit) { continue; }
                    if(np >= maxp) { atomicMax(overflow, np + 1); continue; }
                    pairs[(size_t) np * cap + i] = j;
                    np++;
                }
            }
        }
    }
    npairs[i] = np;
}

#include "dem_force_kernel.cuh"

template<bool FUSED>
__global__ void __launch_bounds__(128) pb_k_dem_force(PbDemForceArgs a) { pb_dem_force_body<FUSED>(a); }

// the argument block of the contact kernel (dem_force_kernel.cuh), the same for the built-in and for a user-defined contact model
static PbDemForceArgs pb_dem_force_args(pb_ctx *ctx, int accumulate) {
    PbDemForceArgs a;
    a.nlocal = ctx->nlocal; a.cap = ctx->pcap; a.C = ctx->ccontacts; a.ntypes = ctx->dem_ntypes;
    a.P = pb_dem_params(ctx);
    a.pos = ctx->pos; a.vel = ctx->vel; a.angvel = ctx->angvel; a.mass = ctx->mass; a.radius = ctx->radius; a.normal = ctx->normal;
    a.flags = ctx->flags; a.shape = ctx->shape; a.uid = ctx->uid; a.npairs = ctx->numneigh; a.pairs = ctx->neigh;
    a.fric_s = ctx->d_fric_static; a.fric_d = ctx->d_fric_dynamic;
    a.num_contacts = ctx->num_contacts; a.c_uid = ctx->contact_uid; a.c_used = ctx->contact_used; a.c_stick = ctx->contact_stick;
    a.c_tsd = ctx->contact_tsd; a.c_ivm = ctx->contact_ivm; a.force = ctx->force; a.torque = ctx->torque;
    a.c_x = ctx->contact_x; a.nx = ctx->cx;
    for(int k = 0; k < 16; k++) { a.x_default[k] = ctx->cx_default[k]; }
    a.accumulate = accumulate;
    a.overflow = ctx->d_dem_flag;
    return a;
}

int pb_jit_launch_dem_force_raw(pb_ctx *ctx, int fused, void *args_struct);     // jit.cu: the user's contact model, if one is installed

template<bool FUSED>
static int pb_launch_dem_force(pb_ctx *ctx, int accumulate) {
    const PbDemForceArgs a = pb_dem_force_args(ctx, accumulate);
    if(ctx->dem_user_force != nullptr) { return pb_jit_launch_dem_force_raw(ctx, FUSED ? 1 : 0, (void *) &a); }
    PB_LAUNCH(pb_k_dem_force<FUSED>, pb_blocks(ctx->nlocal, PB_DEM_CTA_PARTICLES), 128, a);
    return 0;
}

extern "C" int pb_dem_linear_spring_dashpot(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbStage st(ctx, "linear_spring_dashpot");
    if(!ctx->dem) { ctx->set_error("DEM not enabled"); return -1; }
    if(ctx->cells_n != ctx->nlocal + ctx->nghost) { ctx->set_error("pb_dem_linear_spring_dashpot: cell lists are stale"); return -1; }
    PB_TRY(pb_materialise_force_reset(ctx));     // gravity precedes this kernel and already needs the zeroed force
    if(ctx->nlocal == 0) { return 0; }
    const int C = ctx->ccontacts;
    // pass-1 output: partner indices [C][pcap] + counts; neigh / numneigh are free in DEM (no Verlet lists)
    const size_t need = sizeof(int) * (size_t) C * (size_t) ctx->pcap;
    if(need > ctx->neigh_bytes) {
        if(ctx->neigh != nullptr) { PB_CHECK(cudaFree(ctx->neigh)); ctx->neigh = nullptr; }
        PB_CHECK(cudaMalloc(&ctx->neigh, need));
        ctx->neigh_bytes = need;
    }
    PB_LAUNCH(pb_k_dem_detect, pb_blocks(ctx->nlocal, 128), 128, ctx->nlocal, ctx->pcap, ctx->ncells, ctx->dim_cells[1], ctx->dim_cells[2], C,
              ctx->pos, ctx->radius, ctx->normal, ctx->flags, ctx->shape, ctx->particle_cell, ctx->cell_start, ctx->cell_list, ctx->numneigh,
              ctx->neigh, ctx->d_dem_flag);
    PB_TRY(pb_launch_dem_force<false>(ctx, 1));
    ctx->neigh_n = -1;
    return 0;
}

// reset_contact_history_usage_status + reset_volatile_properties + gravity + linear_spring_dashpot + clear_unused_contact_history
// of one iteration of the generated loop, in two launches (pb_dem_run; contact capacity <= 32)
static int pb_dem_contacts_fused(pb_ctx *ctx) {
    PbStage st(ctx, "linear_spring_dashpot");
    if(ctx->cells_n != ctx->nlocal + ctx->nghost) { ctx->set_error("pb_dem_run: cell lists are stale"); return -1; }
    ctx->force_is_zero = false;
    if(ctx->nlocal == 0) { return 0; }
    const int C = ctx->ccontacts;
    const size_t need = sizeof(int) * (size_t) C * (size_t) ctx->pcap;
    if(need > ctx->neigh_bytes) {
        if(ctx->neigh != nullptr) { PB_CHECK(cudaFree(ctx->neigh)); ctx->neigh = nullptr; }
        PB_CHECK(cudaMalloc(&ctx->neigh, need));
        ctx->neigh_bytes = need;
    }
    PB_LAUNCH(pb_k_dem_detect, pb_blocks(ctx->nlocal, 128), 128, ctx->nlocal, ctx->pcap, ctx->ncells, ctx->dim_cells[1], ctx->dim_cells[2], C,
              ctx->pos, ctx->radius, ctx->normal, ctx->flags, ctx->shape, ctx->particle_cell, ctx->cell_start, ctx->cell_list, ctx->numneigh,
              ctx->neigh, ctx->d_dem_flag);
    PB_TRY(pb_launch_dem_force<true>(ctx, 0));
    ctx->neigh_n = -1;
    return 0;
}

// Contact-capacity growth (the reference's resize protocol for `neighbor_capacity` of a contact-history simulation,
// transformations/modules.py:159-203, re-runs the module after growing; the fused contact kernel cannot be re-run, so the
// capacity grows AHEAD of need): the contact kernel records the fullest row it has seen once a row comes within
// PB_DEM_CONTACT_MARGIN slots of the capacity; the loops check that mark (pb_dem_check_contacts) and double the capacity -- all
// ranks together, the wire record of a migrating particle holds its whole row -- up to the 64 slots the kernels support.
int pb_allreduce_sum(pb_ctx *ctx, double *vals, int n);      // comm_nccl.cu

static int pb_dem_grow_contacts(pb_ctx *ctx, int newC) {
    const size_t C = (size_t) ctx->ccontacts, cap = (size_t) ctx->pcap, NC = (size_t) newC;
    auto grow = [&](auto **p, int blocks) -> int {
        using T = std::remove_pointer_t<std::remove_pointer_t<decltype(p)>>;
        T *q = nullptr;
        PB_CHECK(cudaMalloc(&q, sizeof(T) * blocks * NC * cap));
        PB_CHECK(cudaMemsetAsync(q, 0, sizeof(T) * blocks * NC * cap, ctx->stream));
        if(*p != nullptr) {
            for(int b = 0; b < blocks; b++) {      // [block][slot][particle]: the old slots are the first C of every block
                PB_CHECK(cudaMemcpyAsync(q + (size_t) b * NC * cap, *p + (size_t) b * C * cap, sizeof(T) * C * cap, cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
        if(*p != nullptr) { PB_CHECK(cudaFree(*p)); }
        *p = q;
        return 0;
    };
    if(cap > 0) {
        PB_TRY(grow(&ctx->contact_uid, 1));
        PB_TRY(grow(&ctx->contact_used, 1));
        PB_TRY(grow(&ctx->contact_stick, 1));
        PB_TRY(grow(&ctx->contact_ivm, 1));
        PB_TRY(grow(&ctx->contact_tsd, 3));
        if(ctx->cx > 0) { PB_TRY(grow(&ctx->contact_x, ctx->cx)); }
    }
    ctx->ccontacts = newC;
    if(ctx->send_cap > 0) {        // the wire records grew with the rows
        const int keep = ctx->send_cap;
        ctx->send_cap = 0;
        PB_TRY(pb_ensure_send_capacity(ctx, keep));
    }
    if(ctx->recv_buf != nullptr) { PB_CHECK(cudaFree(ctx->recv_buf)); ctx->recv_buf = nullptr; ctx->recv_cap = 0; }
    return 0;
}

extern "C" int pb_dem_contact_capacity(const pb_ctx *ctx) { return ctx->ccontacts; }

// 0: fine (possibly after growing); < 0: a contact was lost before the capacity could grow
extern "C" int pb_dem_check_contacts(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(!ctx->dem) { return 0; }
    PB_CHECK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_dem_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaMemsetAsync(ctx->d_dem_flag, 0, 2 * sizeof(int), ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    double v[2] = {(double) ctx->h_scalars[0], (ctx->h_scalars[1] + PB_DEM_CONTACT_MARGIN > ctx->ccontacts) ? 1.0 : 0.0};
    double lost = v[0];
    if(ctx->world > 1) {
        double s[2] = {v[0] > 0.0 ? 1.0 : 0.0, v[1]};
        PB_TRY(pb_allreduce_sum(ctx, s, 2));
        lost = (s[0] > 0.0) ? std::max(v[0], 1.0) : 0.0;
        v[1] = s[1];
    }
    if(lost > 0.0) {
        ctx->set_error("contact capacity exceeded: a particle needed " + std::to_string((int) lost) + " contact slots before the capacity (" +
                       std::to_string(ctx->ccontacts) + ") could grow -- raise neighbor_capacity of pairs.simulation()");
        return -1;
    }
    if(v[1] > 0.0 && ctx->ccontacts < 64) { PB_TRY(pb_dem_grow_contacts(ctx, std::min(64, ctx->ccontacts * 2))); }
    return 0;
}

// returns > 0 (needed capacity) if a particle ran out of contact slots since the last check (the mark is read and cleared)
extern "C" int pb_dem_contact_overflow(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PB_CHECK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_dem_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaMemsetAsync(ctx->d_dem_flag, 0, sizeof(int), ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return ctx->h_scalars[0];
}

// euler (examples/dem.py:77-85)
__global__ void __launch_bounds__(128) pb_k_dem_euler(int n, int cap, double dt, const int *__restrict__ flags, const double *__restrict__ mass,
                                                      const double *__restrict__ force, const double *__restrict__ torque,
                                                      const double *__restrict__ Iinv, double4 *__restrict__ pos, double *__restrict__ vel,
                                                      double *__restrict__ angvel, double *__restrict__ quat, double *__restrict__ rotmat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n || (flags[i] & PB_FLAG_FIXED) != 0) { return; }
    double4 p = pos[i];
    double x[3] = {p.x, p.y, p.z}, v[3], w[3], f[3], tau[3], q[4], R[9], I[9];
    for(int d = 0; d < 3; d++) {
        v[d] = vel[(size_t) d * cap + i]; w[d] = angvel[(size_t) d * cap + i];
        f[d] = force[(size_t) d * cap + i]; tau[d] = torque[(size_t) d * cap + i];
    }
    for(int c = 0; c < 4; c++) { q[c] = quat[(size_t) c * cap + i]; }
    for(int c = 0; c < 9; c++) { R[c] = rotmat[(size_t) c * cap + i]; I[c] = Iinv[(size_t) c * cap + i]; }
    pb_dem_euler(dt, mass[i], f, tau, I, x, v, w, q, R);
    p.x = x[0]; p.y = x[1]; p.z = x[2];
    pos[i] = p;
    for(int d = 0; d < 3; d++) { vel[(size_t) d * cap + i] = v[d]; angvel[(size_t) d * cap + i] = w[d]; }
    for(int c = 0; c < 4; c++) { quat[(size_t) c * cap + i] = q[c]; }
    for(int c = 0; c < 9; c++) { rotmat[(size_t) c * cap + i] = R[c]; }
}

extern "C" int pb_dem_euler(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbStage st(ctx, "euler");
    PB_TRY(pb_materialise_force_reset(ctx));
    if(ctx->nlocal == 0) { return 0; }
    PB_LAUNCH(pb_k_dem_euler, pb_blocks(ctx->nlocal, 128), 128, ctx->nlocal, ctx->pcap, pb_dem_params(ctx).dt, ctx->flags, ctx->mass, ctx->force,
              ctx->torque, ctx->inv_inertia, ctx->pos, ctx->vel, ctx->angvel, ctx->quat, ctx->rotmat);
    return 0;
}

// ---- the generated DEM timestep loop (sim/simulation.py:387-417 with use_contact_history, reneighbour every step) ----
extern "C" int pb_dem_run(pb_ctx *ctx, double cell_spacing, int ts_begin, int ts_end) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(!ctx->dem) { ctx->set_error("DEM not enabled"); return -1; }
    if(!ctx->cells_set || ctx->spacing != cell_spacing) { PB_TRY(pb_setup_cells(ctx, cell_spacing)); }
    for(int ts = ts_begin; ts < ts_end; ts++) {
        PB_TRY(pb_exchange(ctx));
        if(ctx->dem_sort_every > 0 && ts > 0 && ts % ctx->dem_sort_every == 0) { PB_TRY(pb_dem_sort_locals(ctx)); }
        PB_TRY(pb_borders(ctx));
        PB_TRY(pb_build_cell_lists(ctx));
        if(ctx->ccontacts <= 32 && ctx->dem_fuse) {
            if(ctx->xrows > ctx->xrows_nv) { PB_TRY(pb_xprops_reset_volatile(ctx)); }      // volatile user-defined properties
            PB_TRY(pb_dem_contacts_fused(ctx));      // usage reset + volatile reset + gravity + contacts + history clean-up
            PB_TRY(pb_dem_euler(ctx));
        } else {
            PB_TRY(pb_dem_reset_contact_usage(ctx));
            PB_TRY(pb_reset_volatile(ctx));
            PB_TRY(pb_dem_gravity(ctx));
            PB_TRY(pb_dem_linear_spring_dashpot(ctx));
            PB_TRY(pb_dem_euler(ctx));
            PB_TRY(pb_dem_clear_unused_contacts(ctx));
        }
        if(((ts + 1) & 7) == 0) { PB_TRY(pb_dem_check_contacts(ctx)); }      // contact rows nearly full: grow ahead of need
    }
    return pb_dem_check_contacts(ctx);
}
