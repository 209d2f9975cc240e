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

int pb_dem_pack_exchange(pb_ctx *ctx, int n, int stride, int base_hi, const int *sel_lo, const int *scan_lo, const int *sel_hi, const int *scan_hi);
int pb_dem_unpack_exchange(pb_ctx *ctx, int count, int dst0, int stride, const double *src);
int pb_dem_move(pb_ctx *ctx, int count, const int *src_idx, const int *dst_idx);

struct PbBox3 {
    double len[3];
};

__global__ void __launch_bounds__(256) pb_k_sel_leave(int n, int dim, double lo, double hi, int do_lo, int do_hi,
                                                      const double4 *__restrict__ pos, const int *__restrict__ flags,
                                                      int *__restrict__ sel_lo, int *__restrict__ sel_hi, int *__restrict__ stay) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    int a = 0, b = 0;
    if((flags[i] & (PB_FLAG_INFINITE | PB_FLAG_GLOBAL)) == 0) {
        const double4 p = pos[i];
        const double x = (dim == 0) ? p.x : ((dim == 1) ? p.y : p.z);
        a = do_lo && (x < lo);
        b = do_hi && (x > hi);
    }
    sel_lo[i] = a;
    sel_hi[i] = b;
    stay[i] = !(a || b);
}

__global__ void __launch_bounds__(256) pb_k_pack_exchange(int n, int cap, int stride, int dim, int mult_lo, int mult_hi, int base_hi, double len,
                                                          const int *__restrict__ sel_lo, const int *__restrict__ scan_lo,
                                                          const int *__restrict__ sel_hi, const int *__restrict__ scan_hi,
                                                          const double4 *__restrict__ pos, const double *__restrict__ vel,
                                                          const double *__restrict__ mass, const int *__restrict__ flags,
                                                          const int *__restrict__ uid, const int *__restrict__ shape,
                                                          const int *__restrict__ tag, double *__restrict__ buf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    int e, mult;
    if(sel_lo[i]) { e = scan_lo[i]; mult = mult_lo; }
    else if(sel_hi[i]) { e = base_hi + scan_hi[i]; mult = mult_hi; }
    else { return; }
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
// hole (deterministic).  rank_leave = scan_lo + scan_hi is the number of leavers before an index.
__global__ void __launch_bounds__(256) pb_k_hole_list(int n, int new_n, const int *__restrict__ stay, const int *__restrict__ scan_lo,
                                                      const int *__restrict__ scan_hi, int *__restrict__ hole_idx, int *__restrict__ fill_idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    const int before = scan_lo[i] + scan_hi[i];
    if(i < new_n) {
        if(!stay[i]) { hole_idx[before] = i; }
    } else if(stay[i]) {
        const int before_tail = scan_lo[new_n] + scan_hi[new_n];
        fill_idx[(i - new_n) - (before - before_tail)] = i;
    }
}

__global__ void __launch_bounds__(256) pb_k_move_base(int count, int cap, const int *__restrict__ src_idx, const int *__restrict__ dst_idx,
                                                      double4 *__restrict__ pos, double *__restrict__ vel, double *__restrict__ mass,
                                                      int *__restrict__ type, int *__restrict__ flags, int *__restrict__ uid,
                                                      int *__restrict__ shape, int *__restrict__ tag) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= count) { return; }
    const int s = src_idx[k], t = dst_idx[k];
    pos[t] = pos[s];
    vel[t] = vel[s]; vel[cap + t] = vel[cap + s]; vel[2 * cap + t] = vel[2 * cap + s];
    mass[t] = mass[s]; type[t] = type[s]; flags[t] = flags[s]; uid[t] = uid[s]; shape[t] = shape[s]; tag[t] = tag[s];
}

int pb_exchange_multi(pb_ctx *ctx, int dim) {
    const int n = ctx->nlocal;
    const int j0 = dim * 2, j1 = dim * 2 + 1;
    if(ctx->mig_scan_a == nullptr) {     // capacity was reserved before the domain became multi-rank
        PB_CHECK(cudaMalloc(&ctx->mig_scan_a, sizeof(int) * ((size_t) ctx->pcap + 1)));
        PB_CHECK(cudaMalloc(&ctx->mig_scan_b, sizeof(int) * ((size_t) ctx->pcap + 1)));
    }
    const int do_lo = ctx->pbc_flag[dim] || ctx->pbc[j0] == 0;
    const int do_hi = ctx->pbc_flag[dim] || ctx->pbc[j1] == 0;
    // scratch: sel_lo = sel_flag, sel_hi / stay / scans live in cell_slot, cell_list, particle_cell, sel_scan (all [pcap])
    // (all of them are free at this point of the reneighbouring sequence; mig_scan* are [pcap+1] and persistent)
    int *sel_lo = ctx->sel_flag, *sel_hi = ctx->cell_slot, *stay = ctx->cell_list;
    int *scan_lo = ctx->sel_scan, *scan_hi = ctx->mig_scan_a, *scan_stay = ctx->mig_scan_b;
    for(int j = 0; j < 6; j++) { ctx->nsend[j] = 0; ctx->nrecv[j] = 0; ctx->send_offsets[j] = 0; ctx->recv_offsets[j] = 0; }
    int c_lo = 0, c_hi = 0, c_stay = 0;
    if(n > 0) {
        PB_LAUNCH(pb_k_sel_leave, pb_blocks(n, 256), 256, n, dim, ctx->subdom[j0], ctx->subdom[j1], do_lo, do_hi, ctx->pos, ctx->flags,
                  sel_lo, sel_hi, stay);
        PB_TRY(pb_exclusive_scan(ctx, sel_lo, scan_lo, n));
        PB_TRY(pb_exclusive_scan(ctx, sel_hi, scan_hi, n));
        PB_TRY(pb_exclusive_scan(ctx, stay, scan_stay, n));
        PB_CHECK(cudaMemcpyAsync(ctx->h_scalars + 0, scan_lo + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CHECK(cudaMemcpyAsync(ctx->h_scalars + 1, scan_hi + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CHECK(cudaMemcpyAsync(ctx->h_scalars + 2, scan_stay + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
        c_lo = ctx->h_scalars[0]; c_hi = ctx->h_scalars[1]; c_stay = ctx->h_scalars[2];
    }
    ctx->nsend[j0] = c_lo;
    ctx->nsend[j1] = c_hi;
    ctx->send_offsets[j0] = 0;
    ctx->send_offsets[j1] = c_lo;
    PB_TRY(pb_ensure_send_capacity(ctx, c_lo + c_hi));
    const int stride = ctx->dem ? pb_record_elems(ctx) : EXCH_ELEMS;
    if(c_lo + c_hi > 0) {
        const double len = ctx->grid[dim * 2 + 1] - ctx->grid[dim * 2];
        PB_LAUNCH(pb_k_pack_exchange, pb_blocks(n, 256), 256, n, ctx->pcap, stride, dim, ctx->pbc[j0], ctx->pbc[j1], c_lo, len, sel_lo, scan_lo,
                  sel_hi, scan_hi, ctx->pos, ctx->vel, ctx->mass, ctx->flags, ctx->uid, ctx->shape, ctx->tag, ctx->send_buf);
        if(ctx->dem) { PB_TRY(pb_dem_pack_exchange(ctx, n, stride, c_lo, sel_lo, scan_lo, sel_hi, scan_hi)); }
    }
    PB_TRY(pb_transport_sizes(ctx, dim));
    ctx->recv_offsets[j0] = 0;
    ctx->recv_offsets[j1] = ctx->nrecv[j0];
    const int nr = ctx->nrecv[j0] + ctx->nrecv[j1];
    if(n > 0 && c_stay < n) {
        // hole filling: k-th stayer of the tail [c_stay, n) moves into the k-th leaver slot below c_stay
        int *hole_idx = ctx->cell_key, *fill_idx = ctx->particle_cell;          // [pcap] scratch, free during exchange
        PB_LAUNCH(pb_k_hole_list, pb_blocks(n, 256), 256, n, c_stay, stay, scan_lo, scan_hi, hole_idx, fill_idx);
        PB_CHECK(cudaMemcpyAsync(ctx->h_scalars + 3, scan_stay + c_stay, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
        const int nholes = c_stay - ctx->h_scalars[3];      // leavers below c_stay = c_stay - stayers below c_stay
        if(nholes > 0) {
            PB_LAUNCH(pb_k_move_base, pb_blocks(nholes, 256), 256, nholes, ctx->pcap, fill_idx, hole_idx, ctx->pos, ctx->vel, ctx->mass, ctx->type,
                      ctx->flags, ctx->uid, ctx->shape, ctx->tag);
            if(ctx->dem) { PB_TRY(pb_dem_move(ctx, nholes, fill_idx, hole_idx)); }
        }
    }
    // grow only now: the selection / scan scratch above is re-allocated (not kept) by a capacity change
    ctx->nlocal = c_stay;
    PB_TRY(pb_ensure_particle_capacity(ctx, c_stay + nr));
    const double *src = nullptr;
    PB_TRY(pb_transport_data(ctx, dim, dim + 1, stride, &src));
    if(nr > 0) {
        PB_LAUNCH(pb_k_unpack_exchange, pb_blocks(nr, 256), 256, nr, c_stay, ctx->pcap, stride, src, ctx->pos, ctx->vel, ctx->mass, ctx->type,
                  ctx->flags, ctx->uid, ctx->shape, ctx->tag);
        if(ctx->dem) { PB_TRY(pb_dem_unpack_exchange(ctx, nr, c_stay, stride, src)); }
    }
    ctx->nlocal = c_stay + nr;
    for(int j = 0; j < 6; j++) { ctx->nsend[j] = 0; ctx->nrecv[j] = 0; ctx->send_offsets[j] = 0; ctx->recv_offsets[j] = 0; }
    return 0;
}
