// The DEM contact kernel (pass 2 of linear_spring_dashpot, see dem_kernels.cu): history lookup / insert, the contact model, force and
// torque accumulation.  This file is compiled twice: into libpairs_b200.so with the contact model of examples/dem.py
// (pb_dem_pair_force, dem_math.h), and -- its text embedded in the library -- by NVRTC together with a contact model that
// pairs_b200/kernelgen.py generated from a user's Python kernel body (PB_DEM_USER_PAIR, csrc/jit.cu pb_jit_compile_dem_force).  Both
// take the same argument block, so the launch sites do not care which one runs.
#pragma once

// what a fresh contact slot starts from: the defaults of add_contact_property() (sim/contact_history.py:66, mapping/funcs.py:258);
// examples/dem.py declares zeros, a generated contact model defines its script's values before this header
#ifndef PB_DEM_DEFAULT_STICK
#define PB_DEM_DEFAULT_STICK 0
#endif
#ifndef PB_DEM_DEFAULT_TSD
#define PB_DEM_DEFAULT_TSD(d) 0.0
#endif
#ifndef PB_DEM_DEFAULT_IVM
#define PB_DEM_DEFAULT_IVM 0.0
#endif

#ifndef PB_DEM_CONTACT_MARGIN
#define PB_DEM_CONTACT_MARGIN 4      // free slots below which a contact row counts as "nearly full"
#endif

struct PbDemForceArgs {
    int nlocal, cap, C, ntypes;
    PbDemParams P;
    const double4 *pos;
    const double *vel, *angvel, *mass, *radius, *normal;
    const int *flags, *shape, *uid, *npairs, *pairs;
    const double *fric_s, *fric_d;
    int *num_contacts, *c_uid, *c_used, *c_stick;
    double *c_tsd, *c_ivm, *force, *torque;
    int accumulate;
    int *overflow;
    // further contact properties (pb_dem_enable_ex): nx double lanes per contact, [lane][slot][particle], and their defaults
    double *c_x;
    int nx;
    double x_default[16];
};

// A generated contact model that declares further contact properties defines PB_DEM_NX (= their lane count, equal to a.nx) before
// this header, so that the lanes of the contact at hand live in registers around the model call.  Without it (the library's build,
// generated models without extras) the lanes are only initialised and compacted with their row, over the run-time count.
#ifdef PB_DEM_NX
#define PB_DEM_NX_LOOP(x, a) _Pragma("unroll") for(int x = 0; x < PB_DEM_NX; x++)
#else
#define PB_DEM_NX_LOOP(x, a) for(int x = 0; x < (a).nx; x++)
#endif

// FUSED folds the cheap per-particle modules around the contact evaluation of the generated loop into this kernel (the thread
// owns particle i's force, torque and contact row anyway):
//   reset_volatile_properties + gravity    f = 0; f.z = gravity(f.z)   before the contact sums are added, same operations
//   reset_contact_history_usage_status     the "used" marks live in a register bit mask (contact capacity <= 32)
//   clear_unused_contact_history           the swap-with-last compaction runs on the row at the end, driven by the mask
// (euler sits between the contact kernel and the clean-up in the reference's list; it touches no contact data, so the order
// does not matter).  Nine launches -- six memsets and three kernels -- and their passes over the arrays disappear.
// One thread per particle, but not thread t for particle t: a thread walks its particle's touching partners one after the other
// (~1800 issued instructions per partner, 455 of them fp64, in one long dependent chain), and with 0 ... 12 partners per particle
// a warp of 32 consecutive particles ran with 54 % of its lanes busy.  So a CTA takes PB_DEM_CTA_PARTICLES consecutive particles, sorts them by partner count in
// shared memory (counting sort; the order inside a bucket is arbitrary and does not matter: particles are independent), and its
// warps take groups of 32 from that order in boustrophedon fashion, so that every warp gets long and short groups.  Every particle is still evaluated
// by the same serial code: identical bits.  Measured (10^6 settled spheres, both contact kernels): 0.995 -> 0.964 ms -- the warp
// iterations halve (ncu: 140 k, 27 of 32 lanes), the time hardly moves: at 158 registers (12 warps per SM) the kernel issues 0.19
// instructions per cycle and scheduler, each warp-iteration takes ~13 000 cycles of dependent fp64 and load latency.  Also
// measured, without gain: 4 lanes per particle with ordered shuffles (same bits; 1.20 ms), prefetching the next partner's state
// (0.959 ms), register caps of 80 ... 128 through NVRTC (0.997 ... 1.016 ms).
#ifndef PB_DEM_CTA_PARTICLES
#define PB_DEM_CTA_PARTICLES 512
#endif

template<bool FUSED>
__device__ __forceinline__ void pb_dem_force_particle(const PbDemForceArgs &a, int i);

template<bool FUSED>
__device__ __forceinline__ void pb_dem_force_body(const PbDemForceArgs &a) {
    constexpr int NP = PB_DEM_CTA_PARTICLES, NG = NP / 32;
    __shared__ int s_off[34];
    __shared__ short s_perm[NP];
    const int base = blockIdx.x * NP, tid = threadIdx.x, nthreads = blockDim.x;
    if(tid < 34) { s_off[tid] = 0; }
    __syncthreads();
    // bucket = partner count (a particle beyond the end, or without partners: bucket 0); clamped to 31
    for(int k = tid; k < NP; k += nthreads) {
        const int i = base + k;
        const int key = (i < a.nlocal) ? min(a.npairs[i], 31) : 0;
        atomicAdd(&s_off[key + 2], 1);
    }
    __syncthreads();
    if(tid == 0) { for(int b = 2; b < 34; b++) { s_off[b] += s_off[b - 1]; } }      // s_off[key + 1] = first position of bucket key
    __syncthreads();
    for(int k = tid; k < NP; k += nthreads) {
        const int i = base + k;
        const int key = (i < a.nlocal) ? min(a.npairs[i], 31) : 0;
        s_perm[atomicAdd(&s_off[key + 1], 1)] = (short) k;
    }
    __syncthreads();
    // groups of 32 in sorted order; warp w of W takes them alternately from both ends, so that every warp gets long and short ones
    const int warp = tid >> 5, lane = tid & 31, W = nthreads >> 5;
    for(int p = 0; p * W < NG; p++) {
        const int g = p * W + ((p & 1) ? (W - 1 - warp) : warp);      // boustrophedon over the sorted groups
        if(g >= NG) { continue; }
        const int i = base + (int) s_perm[g * 32 + lane];
        if(i < a.nlocal) { pb_dem_force_particle<FUSED>(a, i); }
    }
}

template<bool FUSED>
__device__ __forceinline__ void pb_dem_force_particle(const PbDemForceArgs &a, const int i) {
    const int nlocal = a.nlocal, cap = a.cap, C = a.C, ntypes = a.ntypes;
    const PbDemParams &P = a.P;
    const double4 *__restrict__ pos = a.pos;
    const double *__restrict__ vel = a.vel, *__restrict__ angvel = a.angvel, *__restrict__ mass = a.mass, *__restrict__ radius = a.radius,
                 *__restrict__ normal = a.normal;
    const int *__restrict__ flags = a.flags, *__restrict__ shape = a.shape, *__restrict__ uid = a.uid, *__restrict__ npairs = a.npairs,
              *__restrict__ pairs = a.pairs;
    const double *__restrict__ fric_s = a.fric_s, *__restrict__ fric_d = a.fric_d;
    int *__restrict__ num_contacts = a.num_contacts, *__restrict__ c_uid = a.c_uid, *__restrict__ c_used = a.c_used, *__restrict__ c_stick = a.c_stick;
    double *__restrict__ c_tsd = a.c_tsd, *__restrict__ c_ivm = a.c_ivm, *__restrict__ force = a.force, *__restrict__ torque = a.torque;
    double *__restrict__ c_x = a.c_x;
    const int accumulate = a.accumulate;
    int *__restrict__ overflow = a.overflow;
    (void) P; (void) fric_s; (void) fric_d; (void) nlocal;
    const bool fixed = (flags[i] & PB_FLAG_FIXED) != 0;
    double Fs[3] = {0.0, 0.0, 0.0}, Ts[3] = {0.0, 0.0, 0.0}, Fh[3] = {0.0, 0.0, 0.0}, Th[3] = {0.0, 0.0, 0.0};
    const int np = npairs[i];
    unsigned usedmask = 0u;
    int ncont_end = FUSED ? num_contacts[i] : 0;
    if(!fixed && np > 0) {
        const double4 pi4 = pb_ld_pos(pos + i);
        const double xi[3] = {pi4.x, pi4.y, pi4.z};
        const int ti = pb_w_type(pi4.w) * ntypes;
        const double vi[3] = {vel[i], vel[(size_t) cap + i], vel[(size_t) 2 * cap + i]};
        const double wi[3] = {angvel[i], angvel[(size_t) cap + i], angvel[(size_t) 2 * cap + i]};
        const double ri = radius[i];
        const double inv_mi = 1.0 / mass[i];
        (void) inv_mi; (void) ri;
        int ncont = num_contacts[i];
        for(int q = 0; q < np; q++) {
            const int j = pairs[(size_t) q * cap + i];
            const int sh = shape[j];
            double *F = (sh == 0) ? Fs : Fh, *T = (sh == 0) ? Ts : Th;
            const double4 pj4 = pb_ld_pos(pos + j);
            const double xj[3] = {pj4.x, pj4.y, pj4.z};
            double n[3], cp[3], delta;
            if(sh == PB_SHAPE_SPHERE) {
                pb_dem_geom_sphere(xi, ri, xj, radius[j], n, cp, &delta);      // same function, same inputs as pass 1: same values
            } else {
                const double nj[3] = {normal[j], normal[(size_t) cap + j], normal[(size_t) 2 * cap + j]};
                pb_dem_geom_halfspace(xi, ri, xj, nj, n, cp, &delta);
            }
            // contact-history slot keyed by uid[j] (mapping/funcs.py:240-263): last match wins, miss -> append defaults
            const int uj = uid[j];
            int slot = -1;
            for(int c = 0; c < ncont; c++) { if(c_uid[(size_t) c * cap + i] == uj) { slot = c; } }
            if(slot == -1) {
                if(ncont >= C) { atomicMax(overflow, ncont + 1); continue; }
                slot = ncont++;
                c_uid[(size_t) slot * cap + i] = uj;
                c_stick[(size_t) slot * cap + i] = PB_DEM_DEFAULT_STICK;
                for(int d = 0; d < 3; d++) { c_tsd[((size_t) d * C + slot) * cap + i] = PB_DEM_DEFAULT_TSD(d); }
                c_ivm[(size_t) slot * cap + i] = PB_DEM_DEFAULT_IVM;
                PB_DEM_NX_LOOP(x, a) { c_x[((size_t) x * C + slot) * cap + i] = a.x_default[x]; }
            }
            if(FUSED) { usedmask |= 1u << slot; } else { c_used[(size_t) slot * cap + i] = 1; }
            double tsd[3] = {c_tsd[((size_t) 0 * C + slot) * cap + i], c_tsd[((size_t) 1 * C + slot) * cap + i],
                             c_tsd[((size_t) 2 * C + slot) * cap + i]};
            double ivm = c_ivm[(size_t) slot * cap + i];
            int stick = c_stick[(size_t) slot * cap + i];
#ifdef PB_DEM_NX
            double cx[PB_DEM_NX];
            PB_DEM_NX_LOOP(x, a) { cx[x] = c_x[((size_t) x * C + slot) * cap + i]; }
#else
            double *cx = nullptr;
            (void) cx;
#endif
            const double vj[3] = {vel[j], vel[(size_t) cap + j], vel[(size_t) 2 * cap + j]};
            const double wj[3] = {angvel[j], angvel[(size_t) cap + j], angvel[(size_t) 2 * cap + j]};
            const int tj = pb_w_type(pj4.w);
            double Fp[3], Tp[3];
#ifdef PB_DEM_USER_PAIR
            // a contact model generated from a user's kernel body; false = skip_when() left the pair (no force, no torque)
            if(!PB_DEM_USER_PAIR(xi, vi, wi, mass[i], ri, xj, vj, wj, mass[j], radius[j], n, cp, delta, ti + tj, tsd, &ivm, &stick, cx, Fp, Tp)) {
                for(int d = 0; d < 3; d++) { Fp[d] = 0.0; Tp[d] = 0.0; }
            }
#else
            pb_dem_pair_force(P, xi, vi, wi, inv_mi, xj, vj, wj, mass[j], n, cp, delta, fric_s[ti + tj], fric_d[ti + tj], tsd, &ivm, &stick,
                              Fp, Tp);
#endif
            for(int d = 0; d < 3; d++) { c_tsd[((size_t) d * C + slot) * cap + i] = tsd[d]; }
            c_ivm[(size_t) slot * cap + i] = ivm;
            c_stick[(size_t) slot * cap + i] = stick;
#ifdef PB_DEM_NX
            PB_DEM_NX_LOOP(x, a) { c_x[((size_t) x * C + slot) * cap + i] = cx[x]; }
#endif
            for(int d = 0; d < 3; d++) { F[d] = F[d] + Fp[d]; T[d] = T[d] + Tp[d]; }
        }
        if(FUSED) { ncont_end = ncont; } else { num_contacts[i] = ncont; }
        // high-water mark of the contact rows: the host grows the capacity AHEAD of need (pb_dem_check_contacts)
        if(ncont + PB_DEM_CONTACT_MARGIN > C) { atomicMax(overflow + 1, ncont); }
    }
    if(FUSED) {
        // clear_unused_contact_history (sim/contact_history.py:90-127): an unused slot is overwritten by the last one
        int c = 0, cnt = ncont_end;
        while(c < cnt) {
            if(((usedmask >> c) & 1u) == 0u) {
                const int last = cnt - 1;
                if(last > 0) {
                    c_stick[(size_t) c * cap + i] = c_stick[(size_t) last * cap + i];
                    for(int d = 0; d < 3; d++) { c_tsd[((size_t) d * C + c) * cap + i] = c_tsd[((size_t) d * C + last) * cap + i]; }
                    c_ivm[(size_t) c * cap + i] = c_ivm[(size_t) last * cap + i];
                    PB_DEM_NX_LOOP(x, a) { c_x[((size_t) x * C + c) * cap + i] = c_x[((size_t) x * C + last) * cap + i]; }
                    c_uid[(size_t) c * cap + i] = c_uid[(size_t) last * cap + i];
                    usedmask = (usedmask & ~(1u << c)) | (((usedmask >> last) & 1u) << c);
                }
                cnt--;
            } else {
                c++;
            }
        }
        for(int k = 0; k < cnt; k++) { c_used[(size_t) k * cap + i] = 1; }
        num_contacts[i] = cnt;
    }
    // prop[i] = prop[i] + (acc_sphere + acc_halfspace)  (sim/interaction.py:280-292)
    for(int d = 0; d < 3; d++) {
        double f_old = accumulate ? force[(size_t) d * cap + i] : 0.0;
        const double t_old = accumulate ? torque[(size_t) d * cap + i] : 0.0;
        if(FUSED && d == 2 && !fixed) { f_old = pb_dem_gravity(P, radius[i], f_old); }      // gravity on the freshly reset force
        if(!fixed) {
            force[(size_t) d * cap + i] = f_old + (Fs[d] + Fh[d]);
            torque[(size_t) d * cap + i] = t_old + (Ts[d] + Th[d]);
        } else if(!accumulate) {
            force[(size_t) d * cap + i] = 0.0;
            torque[(size_t) d * cap + i] = 0.0;
        }
    }
}
