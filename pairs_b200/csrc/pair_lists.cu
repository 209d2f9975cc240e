// Pair lists (option "pair_lists", experimental, default off): see pair_lists.h for what they are and why.  This file holds the two
// kernels as thin wrappers around the host+device thread functions, the list storage and the launch.
#include <algorithm>

#include "ctx.cuh"
#include "pair_lists.h"

__global__ void __launch_bounds__(128) pb_k_pairlist_merge(int npairs, int nlocal, int T, int T2, double cutsq_lists,
                                                           const double4 *__restrict__ pos, const int *__restrict__ flags,
                                                           const int *__restrict__ numneigh, const int *__restrict__ neigh,
                                                           int *__restrict__ pneigh, int *__restrict__ pnum) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= npairs) { return; }
    pnum[p] = pb_pairlist_merge(p, nlocal, T, T2, cutsq_lists, pos, flags, numneigh, neigh, pneigh);
}

// called at the end of pb_build_neighbor_lists (full lists, one lane per particle)
int pb_build_pair_lists(pb_ctx *ctx, double cutsq_lists) {
    const int n = ctx->nlocal;
    ctx->pairs_n = -1;
    if(n == 0) { return 0; }
    const int npairs = (n + 1) / 2;
    // the union of two lists cannot be longer than twice the longest list: no overflow protocol needed
    const int T2 = std::max(4, (2 * ctx->max_neigh + 3) / 4 * 4);
    const size_t rows = ((size_t) npairs + 31) / 32;
    const size_t need = sizeof(int) * rows * (size_t) T2 * 32;
    if(need > ctx->pneigh_bytes) {
        if(ctx->pneigh != nullptr) { PB_CHECK(cudaFree(ctx->pneigh)); ctx->pneigh = nullptr; ctx->pneigh_bytes = 0; }
        PB_CHECK(cudaMalloc(&ctx->pneigh, need + need / 8));
        ctx->pneigh_bytes = need + need / 8;
    }
    if(npairs > ctx->pnum_cap) {
        if(ctx->pnum != nullptr) { PB_CHECK(cudaFree(ctx->pnum)); ctx->pnum = nullptr; ctx->pnum_cap = 0; }
        PB_CHECK(cudaMalloc(&ctx->pnum, sizeof(int) * ((size_t) npairs + npairs / 4 + 32)));
        ctx->pnum_cap = npairs + npairs / 4 + 32;
    }
    ctx->pair_T2 = T2;
    PB_LAUNCH(pb_k_pairlist_merge, pb_blocks(npairs, 128), 128, npairs, n, ctx->nslots, T2, cutsq_lists, ctx->pos, ctx->flags, ctx->numneigh,
              ctx->neigh, ctx->pneigh, ctx->pnum);
    ctx->pairs_n = n;
    return 0;
}

template<bool UNIFORM, bool ACCUMULATE, int FUSE>
__global__ void __launch_bounds__(128) pb_k_lj_pairs(PbLjPairArgs a, int npairs) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= npairs) { return; }
    pb_lj_pairs_thread<UNIFORM, ACCUMULATE, FUSE>(a, p);
}

template<bool UNIFORM, bool ACCUMULATE>
static int pb_launch_lj_pairs(pb_ctx *ctx, const PbLjPairArgs &a, int npairs, int fuse) {
    const int B = pb_blocks(npairs, 128);
    switch(fuse) {
        case 1: PB_LAUNCH((pb_k_lj_pairs<UNIFORM, ACCUMULATE, 1>), B, 128, a, npairs); break;
        case 2: PB_LAUNCH((pb_k_lj_pairs<UNIFORM, ACCUMULATE, 2>), B, 128, a, npairs); break;
        case 3: PB_LAUNCH((pb_k_lj_pairs<UNIFORM, ACCUMULATE, 3>), B, 128, a, npairs); break;
        default: PB_LAUNCH((pb_k_lj_pairs<UNIFORM, ACCUMULATE, 0>), B, 128, a, npairs); break;
    }
    return 0;
}

// the force evaluation of pb_lennard_jones_fused (md_kernels.cu) over the pair lists; same contract: `fuse` bits, force reset folded
// in when it is pending, new positions into pos_alt when bit 1 is set (the caller swaps the buffers)
int pb_lennard_jones_pairs(pb_ctx *ctx, double cutsq, double dt, int fuse) {
    const int n = ctx->nlocal;
    const int npairs = (n + 1) / 2;
    PbLjPairArgs a;
    a.nlocal = n; a.T2 = ctx->pair_T2; a.cap = ctx->pcap; a.ntypes = ctx->ntypes;
    a.cutsq = cutsq; a.eps_u = ctx->h_eps[0]; a.sig6_u = ctx->h_sig6[0]; a.dt = dt; a.half_dt = dt * 0.5;
    a.eps_t = ctx->d_eps; a.sig6_t = ctx->d_sig6;
    a.pos = ctx->pos; a.flags = ctx->flags; a.pnum = ctx->pnum; a.pneigh = ctx->pneigh;
    a.force = ctx->force; a.mass = ctx->mass; a.vel = ctx->vel; a.pos_next = ctx->pos_alt;
    const bool acc = !ctx->force_is_zero;
    if(ctx->lj_uniform) {
        return acc ? pb_launch_lj_pairs<true, true>(ctx, a, npairs, fuse) : pb_launch_lj_pairs<true, false>(ctx, a, npairs, fuse);
    }
    return acc ? pb_launch_lj_pairs<false, true>(ctx, a, npairs, fuse) : pb_launch_lj_pairs<false, false>(ctx, a, npairs, fuse);
}
