// Pair lists (option "pair_lists", experimental): ONE neighbour list per pair of consecutive cell-sorted particles (2p, 2p + 1).
//
// Why: the Lennard-Jones kernel is bound by the L1TEX wavefronts of its position gathers (90 %), the fp64 pipe
// sits at 51 %.  Two consecutive particles of the cell order are ~1.8 sigma apart and share most of their partners: the union of
// their lists holds ~109 entries instead of 2 x 76.6 (counted on the oracle's lists), so gathering each partner once and testing it
// against BOTH owners needs 29 % fewer gathers for ~16 % more fp64 work.  No membership masks: a partner that is in one owner's
// list only cannot come inside the cutoff of the other while the lists are valid, so `j != i` and `rsq < rc^2` decide, as in the
// per-particle kernel.  Every pair term is computed by the same functions (md_math.h) in the same operand order as there; only the
// order in which a particle's terms are summed differs (union order), which is inside the 1e-12 budget that the cell-sorted
// lists already use.
//
// The bodies of the two kernels are host+device functions of one "thread" so that the CPU test-suite can run them on the oracle's
// lists (tests/host/pair_lists_host.cpp): pb_pairlist_merge builds the union list of a pair from the two per-particle lists,
// pb_lj_pairs_thread evaluates the forces of both owners and applies the epilogue of the per-particle kernel (accumulate / fused
// reset, fused integrator halves) to each.
#pragma once
#include "md_math.h"

#if defined(__CUDA_ARCH__)
typedef double4 PbPos4;
#define PB_PL_LD_POS(p) pb_ld_pos(p)
#define PB_PL_LDG(p) __ldg(p)
#else
#if !defined(__CUDACC__)
struct PbPos4 { double x, y, z, w; };
#else
typedef double4 PbPos4;
#endif
#define PB_PL_LD_POS(p) (*(p))
#define PB_PL_LDG(p) (*(p))
#endif

#ifndef PB_FLAG_FIXED
#define PB_FLAG_FIXED 4
#endif

PB_MD_HD int pb_pl_type(double w) {
    long long bits;
#if defined(__CUDA_ARCH__)
    bits = __double_as_longlong(w);
#else
    union { double d; long long l; } u;
    u.d = w;
    bits = u.l;
#endif
    return (int) (bits & 0xffffffffLL);
}

// k-th entry of particle i's list in the sliced-ELLPACK layout with one lane per particle (PbNeighLayout, G = 1)
PB_MD_HD size_t pb_pl_idx(int i, int k, int T) { return ((size_t) (i >> 5) * T + (size_t) k) * 32 + (size_t) (i & 31); }

// Union list of pair p: the list of i0 = 2p without i1, then the entries of i1's list that are not in i0's list.  "In i0's list"
// is decided by the list builder's own predicate -- rsq(i0, j) < cutsq_lists with i0 as the first operand -- so no search is needed
// (FIXED owners have no list: the builder leaves them out).
// Returns the union length; entries beyond the capacity T2 are counted but not stored (the caller grows and repeats).
PB_MD_HD int pb_pairlist_merge(int p, int nlocal, int T, int T2, double cutsq_lists, const PbPos4 *pos, const int *flags,
                               const int *numneigh, const int *neigh, int *pneigh) {
    const int i0 = 2 * p, i1 = 2 * p + 1;
    int n = 0;
    const int n0 = numneigh[i0];
    for(int k = 0; k < n0; k++) {
        const int j = neigh[pb_pl_idx(i0, k, T)];
        if(j == i1) { continue; }
        if(n < T2) { pneigh[pb_pl_idx(p, n, T2)] = j; }
        n++;
    }
    if(i1 < nlocal) {
        const PbPos4 x0 = pos[i0];
        const int n1 = numneigh[i1];
        // a FIXED particle has no list (the builder skips it, as the reference's does): nothing of i1's list is "already there"
        const bool has_list0 = (flags[i0] & PB_FLAG_FIXED) == 0;
        for(int k = 0; k < n1; k++) {
            const int j = neigh[pb_pl_idx(i1, k, T)];
            if(j == i0) { continue; }
            const PbPos4 xj = pos[j];
            double dx, dy, dz;
            const double rsq = pb_pair_rsq(x0.x, x0.y, x0.z, xj.x, xj.y, xj.z, &dx, &dy, &dz);
            if(has_list0 && rsq < cutsq_lists) { continue; }          // already there: it is in i0's list
            if(n < T2) { pneigh[pb_pl_idx(p, n, T2)] = j; }
            n++;
        }
    }
    return n;
}

struct PbLjPairArgs {
    int nlocal, T2, cap, ntypes;
    double cutsq, eps_u, sig6_u, dt, half_dt;
    const double *eps_t, *sig6_t;      // [ntypes * ntypes] (shared-memory copies on the device)
    const PbPos4 *pos;
    const int *flags, *pnum, *pneigh;
    double *force;
    const double *mass;
    double *vel;
    PbPos4 *pos_next;
};

// epilogue of the per-particle kernel (md_kernels.cu pb_k_lennard_jones), operation for operation
template<bool ACCUMULATE, int FUSE>
PB_MD_HD void pb_lj_pair_epilogue(const PbLjPairArgs &a, int i, bool fixed, double fx, double fy, double fz, PbPos4 pi) {
    const int cap = a.cap;
    if(ACCUMULATE) {
        if(!fixed) {
            fx = PB_ADD(a.force[i], fx);
            fy = PB_ADD(a.force[cap + i], fy);
            fz = PB_ADD(a.force[2 * cap + i], fz);
            a.force[i] = fx;
            a.force[cap + i] = fy;
            a.force[2 * cap + i] = fz;
        }
    } else {
        fx = PB_ADD(0.0, fx);
        fy = PB_ADD(0.0, fy);
        fz = PB_ADD(0.0, fz);
        a.force[i] = fx;
        a.force[cap + i] = fy;
        a.force[2 * cap + i] = fz;
    }
    if(FUSE != 0) {
        if(!fixed) {
            const double m = a.mass[i];
            double vx = a.vel[i], vy = a.vel[cap + i], vz = a.vel[2 * cap + i];
            if(FUSE & 1) {
                vx = PB_ADD(vx, PB_DIV(PB_MUL(a.half_dt, fx), m));
                vy = PB_ADD(vy, PB_DIV(PB_MUL(a.half_dt, fy), m));
                vz = PB_ADD(vz, PB_DIV(PB_MUL(a.half_dt, fz), m));
            }
            if(FUSE & 2) {
                vx = PB_ADD(vx, PB_DIV(PB_MUL(a.half_dt, fx), m));
                vy = PB_ADD(vy, PB_DIV(PB_MUL(a.half_dt, fy), m));
                vz = PB_ADD(vz, PB_DIV(PB_MUL(a.half_dt, fz), m));
                pi.x = PB_ADD(pi.x, PB_MUL(a.dt, vx));
                pi.y = PB_ADD(pi.y, PB_MUL(a.dt, vy));
                pi.z = PB_ADD(pi.z, PB_MUL(a.dt, vz));
            }
            a.vel[i] = vx;
            a.vel[cap + i] = vy;
            a.vel[2 * cap + i] = vz;
        }
        if(FUSE & 2) { a.pos_next[i] = pi; }
    }
}

// one thread = one pair of particles
template<bool UNIFORM, bool ACCUMULATE, int FUSE>
PB_MD_HD void pb_lj_pairs_thread(const PbLjPairArgs &a, int p) {
    const int i0 = 2 * p, i1 = 2 * p + 1;
    if(i0 >= a.nlocal) { return; }
    const bool has1 = i1 < a.nlocal;
    const PbPos4 p0 = PB_PL_LD_POS(a.pos + i0);
    const PbPos4 p1 = has1 ? PB_PL_LD_POS(a.pos + i1) : p0;
    const bool fixed0 = (a.flags[i0] & PB_FLAG_FIXED) != 0;
    const bool fixed1 = has1 && (a.flags[i1] & PB_FLAG_FIXED) != 0;
    const bool act0 = !fixed0, act1 = has1 && !fixed1;
    const int t0 = UNIFORM ? 0 : pb_pl_type(p0.w) * a.ntypes, t1 = UNIFORM ? 0 : pb_pl_type(p1.w) * a.ntypes;
    double f0x = 0.0, f0y = 0.0, f0z = 0.0, f1x = 0.0, f1y = 0.0, f1z = 0.0;
#define PB_PL_TERM(PI, TI, PJ, FX, FY, FZ)                                                                         \
    {                                                                                                              \
        double dx, dy, dz;                                                                                         \
        const double rsq = pb_pair_rsq((PI).x, (PI).y, (PI).z, (PJ).x, (PJ).y, (PJ).z, &dx, &dy, &dz);             \
        if(rsq < a.cutsq) {                                                                                        \
            const double sig6 = UNIFORM ? a.sig6_u : a.sig6_t[(TI) + pb_pl_type((PJ).w)];                          \
            const double eps = UNIFORM ? a.eps_u : a.eps_t[(TI) + pb_pl_type((PJ).w)];                             \
            const double f = pb_lj_fpair(rsq, sig6, eps);                                                          \
            FX = PB_ADD(FX, PB_MUL(dx, f));                                                                        \
            FY = PB_ADD(FY, PB_MUL(dy, f));                                                                        \
            FZ = PB_ADD(FZ, PB_MUL(dz, f));                                                                        \
        }                                                                                                          \
    }
    // the two owners see each other (each from its own side, as two per-particle lists would)
    if(has1) {
        if(act0) { PB_PL_TERM(p0, t0, p1, f0x, f0y, f0z) }
        if(act1) { PB_PL_TERM(p1, t1, p0, f1x, f1y, f1z) }
    }
    if(act0 || act1) {
        const int nn = a.pnum[p];
        const int *nb = a.pneigh + (size_t) (p >> 5) * a.T2 * 32 + (p & 31);
        for(int k0 = 0; k0 < nn; k0 += 4) {
            int jj[4];
            PbPos4 pp[4];
#pragma unroll
            for(int u = 0; u < 4; u++) { jj[u] = (k0 + u < nn) ? PB_PL_LDG(nb + (size_t) (k0 + u) * 32) : i0; }
#pragma unroll
            for(int u = 0; u < 4; u++) { pp[u] = PB_PL_LD_POS(a.pos + jj[u]); }
#pragma unroll
            for(int u = 0; u < 4; u++) {
                if(k0 + u >= nn) { break; }
                if(act0) { PB_PL_TERM(p0, t0, pp[u], f0x, f0y, f0z) }
                if(act1) { PB_PL_TERM(p1, t1, pp[u], f1x, f1y, f1z) }
            }
        }
    }
#undef PB_PL_TERM
    pb_lj_pair_epilogue<ACCUMULATE, FUSE>(a, i0, fixed0, f0x, f0y, f0z, p0);
    if(has1) { pb_lj_pair_epilogue<ACCUMULATE, FUSE>(a, i1, fixed1, f1x, f1y, f1z, p1); }
}
