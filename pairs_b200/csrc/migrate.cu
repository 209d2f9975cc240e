// Particle migration between ranks: the multi-rank branch of Comm.exchange (sim/comm.py:100-151) for one
// dimension.  Reference: determine_exchange_particles (atomics) -> pack -> remove_exchanged_particles pt1 (HOST loop)
// + pt2 (hole filling from the tail) -> MPI -> unpack (append).  Here: ordered select of the leavers, pack, a
// DETERMINISTIC hole filling (k-th staying particle of the tail into the k-th freed slot; device only, no host loop,
// no atomics), transport, append.  The resulting SET of local particles per rank is the reference's; their order is
// ours (MD: replaced by cell order right afterwards; DEM: kept, contact rows move with the particle).
#include <algorithm>

#include "ctx.cuh"

int pb_transport_sizes(pb_ctx *ctx, int dim);
int pb_transport_data(pb_ctx *ctx, int dim_begin, int dim_end, int elem, const double **recv_src);

static const int EXCH_ELEMS = 12;      // base record; DEM appends its properties and the contact table (dem_kernels.cu)

int pb_dem_pack_exchange(pb_ctx *ctx, int n, int stride, const int *rec);
int pb_dem_unpack_exchange(pb_ctx *ctx, int count, int dst0, int stride, const double *src);
int pb_dem_move(pb_ctx *ctx, int max_count, const int *count, const int *src_idx, const int *dst_idx);

struct PbBox3 {
    double len[3];
};

// Leavers are ranked without per-particle prefix-sum arrays: per-block counts of the two sides -> scan of the block counts
// (pb_k_sel2_scan, comm.cu) -> the pack kernel re-evaluates the predicate and ranks inside its block with ballots.  It also
// leaves behind, per particle, rec[i] = index of its wire record (-1: stays) and lb[i] = number of leavers in front of i,
// which is all the DEM pack and the hole filling need.
static const int MIG_T = 256;

__device__ __forceinline__ void pb_leave_flags(int i, int n, int dim, double lo, double hi, int do_lo, int do_hi,
                                               const double4 *__restrict__ pos, const int *__restrict__ flags, int *a, int *b) {
    *a = 0; *b = 0;
    if(i < n && (flags[i] & (PB_FLAG_INFINITE | PB_FLAG_GLOBAL)) == 0) {
        const double4 p = pos[i];
        const double x = (dim == 0) ? p.x : ((dim == 1) ? p.y : p.z);
        *a = do_lo && (x < lo);
        *b = do_hi && (x > hi);
    }
}

__global__ void __launch_bounds__(MIG_T) pb_k_leave_count(int n, int nblocks, int dim, double lo, double hi, int do_lo, int do_hi,
                                                          const double4 *__restrict__ pos, const int *__restrict__ flags,
                                                          int *__restrict__ block_counts) {
    const int i = blockIdx.x * MIG_T + threadIdx.x;
    int a, b;
    pb_leave_flags(i, n, dim, lo, hi, do_lo, do_hi, pos, flags, &a, &b);
    const int c_lo = __syncthreads_count(a), c_hi = __syncthreads_count(b);
    if(threadIdx.x == 0) { block_counts[blockIdx.x] = c_lo; block_counts[nblocks + blockIdx.x] = c_hi; }
}

__global__ void __launch_bounds__(MIG_T) pb_k_pack_exchange(int n, int nblocks, int cap, int stride, int dim, double lo, double hi, int do_lo,
                                                            int do_hi, int mult_lo, int mult_hi, double len, int buf_cap,
                                                            const int *__restrict__ block_offsets, const int *__restrict__ totals,
                                                            const double4 *__restrict__ pos, const double *__restrict__ vel,
                                                            const double *__restrict__ mass, const int *__restrict__ flags,
                                                            const int *__restrict__ uid, const int *__restrict__ shape,
                                                            const int *__restrict__ tag, double *__restrict__ buf, int *__restrict__ rec,
                                                            int *__restrict__ lb) {
    __shared__ int w_lo[MIG_T / 32], w_hi[MIG_T / 32];
    const int i = blockIdx.x * MIG_T + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int is_lo, is_hi;
    pb_leave_flags(i, n, dim, lo, hi, do_lo, do_hi, pos, flags, &is_lo, &is_hi);
    const unsigned m_lo = __ballot_sync(0xffffffffu, is_lo), m_hi = __ballot_sync(0xffffffffu, is_hi);
    if(lane == 0) { w_lo[wid] = __popc(m_lo); w_hi[wid] = __popc(m_hi); }
    __syncthreads();
    int r_lo = block_offsets[blockIdx.x], r_hi = block_offsets[nblocks + blockIdx.x];
    for(int w = 0; w < wid; w++) { r_lo += w_lo[w]; r_hi += w_hi[w]; }
    const unsigned below = (1u << lane) - 1u;
    r_lo += __popc(m_lo & below);
    r_hi += __popc(m_hi & below);
    if(i >= n) { return; }
    lb[i] = r_lo + r_hi;
    int e, mult;
    if(is_lo) { e = r_lo; mult = mult_lo; }
    else if(is_hi) { e = totals[0] + r_hi; mult = mult_hi; }
    else { rec[i] = -1; return; }
    rec[i] = e;
    if(e >= buf_cap) { return; }
    const double4 x = pos[i];
    const double sh = __dmul_rn((double) mult, len);
    double *b = buf + (size_t) e * stride;
    b[0] = (double) uid[i];
    b[1] = (double) shape[i];
    b[2] = (double) flags[i];
    b[3] = (dim == 0) ? __dadd_rn(x.x, sh) : x.x;
    b[4] = (dim == 1) ? __dadd_rn(x.y, sh) : x.y;
    b[5] = (dim == 2) ? __dadd_rn(x.z, sh) : x.z;
    b[6] = mass[i];
    b[7] = vel[i];
    b[8] = vel[cap + i];
    b[9] = vel[2 * cap + i];
    b[10] = (double) pb_w_type(x.w);
    b[11] = (double) tag[i];
}

__global__ void __launch_bounds__(256) pb_k_unpack_exchange(int count, int dst0, int cap, int stride, const double *__restrict__ buf,
                                                            double4 *__restrict__ pos, double *__restrict__ vel,
                                                            double *__restrict__ mass, int *__restrict__ type, int *__restrict__ flags,
                                                            int *__restrict__ uid, int *__restrict__ shape, int *__restrict__ tag) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= count) { return; }
    const double *b = buf + (size_t) k * stride;
    const int p = dst0 + k;
    const int t = (int) b[10];
    uid[p] = (int) b[0];
    shape[p] = (int) b[1];
    flags[p] = (int) b[2];
    pos[p] = make_double4(b[3], b[4], b[5], pb_type_w(t));
    mass[p] = b[6];
    vel[p] = b[7];
    vel[cap + p] = b[8];
    vel[2 * cap + p] = b[9];
    type[p] = t;
    tag[p] = (int) b[11];
}

// ---- hole filling (leavers are few: ~0.5 % of the particles per reneighbouring; a full compaction would move everything) ----
// new_n = n - L.  Holes = leavers below new_n, fillers = stayers at or above new_n; both in ascending order, k-th filler -> k-th
// hole (deterministic).  lb[i] = number of leavers before i, so there are lb[new_n] holes.
__global__ void __launch_bounds__(256) pb_k_hole_list(int n, int new_n, const int *__restrict__ rec, const int *__restrict__ lb,
                                                      int *__restrict__ hole_idx, int *__restrict__ fill_idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    const int before = lb[i];
    if(i < new_n) {
        if(rec[i] >= 0) { hole_idx[before] = i; }
    } else if(rec[i] < 0) {
        fill_idx[(i - new_n) - (before - lb[new_n])] = i;
    }
}

// `count` lives on the device (lb[new_n]): the launch covers the upper bound L, surplus threads leave
__global__ void __launch_bounds__(256) pb_k_move_base(const int *__restrict__ count, int cap, const int *__restrict__ src_idx,
                                                      const int *__restrict__ dst_idx, double4 *__restrict__ pos, double *__restrict__ vel,
                                                      double *__restrict__ mass, int *__restrict__ type, int *__restrict__ flags,
                                                      int *__restrict__ uid, int *__restrict__ shape, int *__restrict__ tag) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= *count) { return; }
    const int s = src_idx[k], t = dst_idx[k];
    pos[t] = pos[s];
    vel[t] = vel[s]; vel[cap + t] = vel[cap + s]; vel[2 * cap + t] = vel[2 * cap + s];
    mass[t] = mass[s]; type[t] = type[s]; flags[t] = flags[s]; uid[t] = uid[s]; shape[t] = shape[s]; tag[t] = tag[s];
}

int pb_sel2_scan(pb_ctx *ctx, int nblocks, int *block_counts, int *totals);      // comm.cu

int pb_exchange_multi(pb_ctx *ctx, int dim) {
    const int n = ctx->nlocal;
    const int j0 = dim * 2, j1 = dim * 2 + 1;
    const int do_lo = ctx->pbc_flag[dim] || ctx->pbc[j0] == 0;
    const int do_hi = ctx->pbc_flag[dim] || ctx->pbc[j1] == 0;
    // scratch ([pcap] ints, all free at this point of the reneighbouring sequence): rec / lb per particle, hole / filler lists
    int *rec = ctx->sel_flag, *lb = ctx->sel_scan, *hole_idx = ctx->cell_key, *fill_idx = ctx->particle_cell;
    for(int j = 0; j < 6; j++) { ctx->nsend[j] = 0; ctx->nrecv[j] = 0; ctx->send_offsets[j] = 0; ctx->recv_offsets[j] = 0; }
    int c_lo = 0, c_hi = 0;
    const int nblocks = pb_blocks(std::max(n, 1), MIG_T);
    if(2 * nblocks + 2 > ctx->sel_blocks_cap) {
        if(ctx->sel_blocks != nullptr) { PB_CHECK(cudaFree(ctx->sel_blocks)); }
        ctx->sel_blocks_cap = 2 * nblocks + 2 + 4096;
        PB_CHECK(cudaMalloc(&ctx->sel_blocks, sizeof(int) * (size_t) ctx->sel_blocks_cap));
    }
    int *totals = ctx->sel_blocks + 2 * (size_t) nblocks;
    if(n > 0) {
        PB_LAUNCH(pb_k_leave_count, nblocks, MIG_T, n, nblocks, dim, ctx->subdom[j0], ctx->subdom[j1], do_lo, do_hi, ctx->pos, ctx->flags,
                  ctx->sel_blocks);
        PB_TRY(pb_sel2_scan(ctx, nblocks, ctx->sel_blocks, totals));
        PB_CHECK(cudaMemcpyAsync(ctx->h_scalars, totals, sizeof(int) * 2, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
        c_lo = ctx->h_scalars[0]; c_hi = ctx->h_scalars[1];
    }
    const int c_stay = n - c_lo - c_hi;
    ctx->nsend[j0] = c_lo;
    ctx->nsend[j1] = c_hi;
    ctx->send_offsets[j0] = 0;
    ctx->send_offsets[j1] = c_lo;
    PB_TRY(pb_ensure_send_capacity(ctx, c_lo + c_hi));
    const int xoff = pb_exchange_base_elems(ctx);        // the non-volatile user-defined rows (props.cu) follow the built-in elements
    const int stride = ctx->dem ? pb_record_elems(ctx) : EXCH_ELEMS + ctx->xrows_nv;
    if(c_lo + c_hi > 0) {
        const double len = ctx->grid[dim * 2 + 1] - ctx->grid[dim * 2];
        PB_LAUNCH(pb_k_pack_exchange, nblocks, MIG_T, n, nblocks, ctx->pcap, stride, dim, ctx->subdom[j0], ctx->subdom[j1], do_lo, do_hi,
                  ctx->pbc[j0], ctx->pbc[j1], len, ctx->send_cap, ctx->sel_blocks, totals, ctx->pos, ctx->vel, ctx->mass, ctx->flags, ctx->uid,
                  ctx->shape, ctx->tag, ctx->send_buf, rec, lb);
        if(ctx->dem) { PB_TRY(pb_dem_pack_exchange(ctx, n, stride, rec)); }
        PB_TRY(pb_xprops_pack_leavers(ctx, n, stride, xoff, rec, ctx->send_buf));
    }
    PB_TRY(pb_transport_sizes(ctx, dim));
    ctx->recv_offsets[j0] = 0;
    ctx->recv_offsets[j1] = ctx->nrecv[j0];
    const int nr = ctx->nrecv[j0] + ctx->nrecv[j1];
    if(c_lo + c_hi > 0 && c_stay > 0) {
        // hole filling: k-th stayer of the tail [c_stay, n) moves into the k-th leaver slot below c_stay
        const int L = c_lo + c_hi;
        PB_LAUNCH(pb_k_hole_list, pb_blocks(n, 256), 256, n, c_stay, rec, lb, hole_idx, fill_idx);
        PB_LAUNCH(pb_k_move_base, pb_blocks(L, 256), 256, lb + c_stay, ctx->pcap, fill_idx, hole_idx, ctx->pos, ctx->vel, ctx->mass, ctx->type,
                  ctx->flags, ctx->uid, ctx->shape, ctx->tag);
        if(ctx->dem) { PB_TRY(pb_dem_move(ctx, L, lb + c_stay, fill_idx, hole_idx)); }
        PB_TRY(pb_xprops_move(ctx, L, lb + c_stay, fill_idx, hole_idx));
    }
    // grow only now: the scratch above is re-allocated (not kept) by a capacity change
    ctx->nlocal = c_stay;
    PB_TRY(pb_ensure_particle_capacity(ctx, c_stay + nr));
    const double *src = nullptr;
    PB_TRY(pb_transport_data(ctx, dim, dim + 1, stride, &src));
    if(nr > 0) {
        PB_LAUNCH(pb_k_unpack_exchange, pb_blocks(nr, 256), 256, nr, c_stay, ctx->pcap, stride, src, ctx->pos, ctx->vel, ctx->mass, ctx->type,
                  ctx->flags, ctx->uid, ctx->shape, ctx->tag);
        if(ctx->dem) { PB_TRY(pb_dem_unpack_exchange(ctx, nr, c_stay, stride, src)); }
        PB_TRY(pb_xprops_unpack(ctx, 0, nr, c_stay, stride, xoff, src));
    }
    ctx->nlocal = c_stay + nr;
    for(int j = 0; j < 6; j++) { ctx->nsend[j] = 0; ctx->nrecv[j] = 0; ctx->send_offsets[j] = 0; ctx->recv_offsets[j] = 0; }
    return 0;
}
