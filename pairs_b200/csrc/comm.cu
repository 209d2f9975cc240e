// Communication stages of sim/comm.py as device-side select / pack / unpack kernels:
//   pb_exchange     Comm.exchange   (comm.py:100-151)  migration + periodic wrap, then cell-order sort of the locals
//   pb_borders      Comm.borders    (comm.py:56-98)    ghost creation, 3 ordered phases x -> y -> z
//   pb_synchronize  Comm.synchronize(comm.py:45-54)    per-step ghost refresh
// The reference selects with atomics (send order = whatever the atomics give, serial order on the CPU target);
// here selection is an ORDERED stream compaction (per-block counts, scan of the block counts, ballot ranks), so send lists -- and
// with them ghost numbering -- are deterministic: (dim, side, ascending source index).
//
// Transport between ranks is abstracted by pb_transport_* (comm_nccl.cu): a neighbour that is the rank itself
// (single rank in that dimension, Regular6DStencil::communicateData's copy_in_device branch,
// runtime/domain/regular_6d_stencil.cpp:156-157,175-176) is served from the send buffer directly.
//
// Wire records are rows of doubles as in the reference (ints cast to double, comm.py:328-329), one extra element
// carries the particle tag:   exchange: uid shape flags x y z mass vx vy vz type tag          (12 doubles)
//                             borders : uid type mass x y z vx vy vz shape tag              (11 doubles)
//                             borders (DEM): ... + radius wx wy wz                           (15 doubles)
//                             sync    : x y z vx vy vz                                      ( 6 doubles)
//                             sync (DEM): ... + wx wy wz                                     ( 9 doubles; comm.py:49 lists angular_velocity)
// Positions are shifted by send_mult * L on ALL three axes exactly as comm.py:320-321 does (two of the
// multipliers are zero), so ghost coordinates are bit-identical to the reference's.
#include <algorithm>

#include "ctx.cuh"

int pb_sort_locals(pb_ctx *ctx);
int pb_transport_sizes(pb_ctx *ctx, int dim);
int pb_transport_data(pb_ctx *ctx, int dim_begin, int dim_end, int elem, const double **recv_src);

static const int BORDER_ELEMS = 11, BORDER_ELEMS_DEM = 15, SYNC_ELEMS = 6, SYNC_ELEMS_DEM = 9;

struct PbBox {
    double len[3];
};

// ---- selection ---------------------------------------------------------------------------------------------
// sim/domain_partitioning.py:31-65: side 0 -> pos < min + offset, side 1 -> pos > max - offset; INFINITE/GLOBAL skipped.
// Both sides of a dimension are selected in ONE ordered stream compaction: per-block counts -> scan of the block counts
// (one block) -> scatter that re-evaluates the predicate and ranks inside the block with ballots.  Four launches and one
// read-back per dimension (the reference needs the two counts on the host too, for its MPI message sizes).
static const int SEL_T = 256;

__device__ __forceinline__ void pb_sel_flags(int i, int n, int dim, int on_lo, int on_hi, double b_lo, double b_hi,
                                             const double4 *__restrict__ pos, const int *__restrict__ flags, int *lo, int *hi) {
    *lo = 0; *hi = 0;
    if(i < n && (flags[i] & (PB_FLAG_INFINITE | PB_FLAG_GLOBAL)) == 0) {
        const double4 p = pos[i];
        const double x = (dim == 0) ? p.x : ((dim == 1) ? p.y : p.z);
        *lo = on_lo && (x < b_lo);
        *hi = on_hi && (x > b_hi);
    }
}

__global__ void __launch_bounds__(SEL_T) pb_k_sel2_count(int n, int nblocks, int dim, int on_lo, int on_hi, double b_lo, double b_hi,
                                                         const double4 *__restrict__ pos, const int *__restrict__ flags,
                                                         int *__restrict__ block_counts) {
    const int i = blockIdx.x * SEL_T + threadIdx.x;
    int lo, hi;
    pb_sel_flags(i, n, dim, on_lo, on_hi, b_lo, b_hi, pos, flags, &lo, &hi);
    const int c_lo = __syncthreads_count(lo), c_hi = __syncthreads_count(hi);
    if(threadIdx.x == 0) { block_counts[blockIdx.x] = c_lo; block_counts[nblocks + blockIdx.x] = c_hi; }
}

// one block per side: in-place exclusive scan of that side's block counts, total -> totals[side]
__global__ void __launch_bounds__(1024) pb_k_sel2_scan(int nblocks, int *__restrict__ block_counts, int *__restrict__ totals) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    int *c = block_counts + (size_t) blockIdx.x * nblocks;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if(threadIdx.x == 0) { carry_s = 0; }
    __syncthreads();
    for(int base = 0; base < nblocks; base += 1024) {
        const int k = base + threadIdx.x;
        const int v = (k < nblocks) ? c[k] : 0;
        int x = v;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if(lane >= o) { x += y; }
        }
        if(lane == 31) { warp_sums[wid] = x; }
        __syncthreads();
        if(wid == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for(int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if(lane >= o) { w += y; }
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + ((wid == 0) ? 0 : warp_sums[wid - 1]) + x - v;
        if(k < nblocks) { c[k] = excl; }
        __syncthreads();
        if(threadIdx.x == 1023) { carry_s = carry + warp_sums[31]; }
        __syncthreads();
    }
    if(threadIdx.x == 0) { totals[blockIdx.x] = carry_s; }
}

int pb_sel2_scan(pb_ctx *ctx, int nblocks, int *block_counts, int *totals) {
    PB_LAUNCH(pb_k_sel2_scan, 2, 1024, nblocks, block_counts, totals);
    return 0;
}

// entries of side 0 start at `base`, those of side 1 right behind them (base + totals[0]); nothing is written beyond `cap`
__global__ void __launch_bounds__(SEL_T) pb_k_sel2_scatter(int n, int nblocks, int dim, int on_lo, int on_hi, double b_lo, double b_hi,
                                                           int mult_lo, int mult_hi, int base, int cap,
                                                           const double4 *__restrict__ pos, const int *__restrict__ flags,
                                                           const int *__restrict__ block_offsets, const int *__restrict__ totals,
                                                           int *__restrict__ send_map, int *__restrict__ send_mult) {
    __shared__ int w_lo[SEL_T / 32], w_hi[SEL_T / 32];
    const int i = blockIdx.x * SEL_T + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int lo, hi;
    pb_sel_flags(i, n, dim, on_lo, on_hi, b_lo, b_hi, pos, flags, &lo, &hi);
    const unsigned m_lo = __ballot_sync(0xffffffffu, lo), m_hi = __ballot_sync(0xffffffffu, hi);
    if(lane == 0) { w_lo[wid] = __popc(m_lo); w_hi[wid] = __popc(m_hi); }
    __syncthreads();
    int off_lo = 0, off_hi = 0;
    for(int w = 0; w < wid; w++) { off_lo += w_lo[w]; off_hi += w_hi[w]; }
    const unsigned below = (1u << lane) - 1u;
    if(lo) {
        const int k = base + block_offsets[blockIdx.x] + off_lo + __popc(m_lo & below);
        if(k < cap) {
            send_map[k] = i;
            send_mult[k * 3 + 0] = (dim == 0) ? mult_lo : 0;
            send_mult[k * 3 + 1] = (dim == 1) ? mult_lo : 0;
            send_mult[k * 3 + 2] = (dim == 2) ? mult_lo : 0;
        }
    }
    if(hi) {
        const int k = base + totals[0] + block_offsets[nblocks + blockIdx.x] + off_hi + __popc(m_hi & below);
        if(k < cap) {
            send_map[k] = i;
            send_mult[k * 3 + 0] = (dim == 0) ? mult_hi : 0;
            send_mult[k * 3 + 1] = (dim == 1) ? mult_hi : 0;
            send_mult[k * 3 + 2] = (dim == 2) ? mult_hi : 0;
        }
    }
}

// Appends the particles of [0, n) selected for the two sides of `dim` to the send lists (side 0 first, ascending index
// inside a side); returns their numbers.
static int pb_select_dim(pb_ctx *ctx, int n, int dim, double offset, int *count_lo, int *count_hi) {
    *count_lo = 0; *count_hi = 0;
    const int j0 = dim * 2, j1 = dim * 2 + 1;
    const int on_lo = !(!ctx->pbc_flag[dim] && ctx->pbc[j0] != 0), on_hi = !(!ctx->pbc_flag[dim] && ctx->pbc[j1] != 0);
    if(n == 0 || (!on_lo && !on_hi)) { return 0; }
    const double b_lo = ctx->subdom[j0] + offset, b_hi = ctx->subdom[j1] - offset;
    const int nblocks = pb_blocks(n, SEL_T);
    if(2 * nblocks + 2 > ctx->sel_blocks_cap) {
        if(ctx->sel_blocks != nullptr) { PB_CHECK(cudaFree(ctx->sel_blocks)); }
        ctx->sel_blocks_cap = 2 * nblocks + 2 + 4096;
        PB_CHECK(cudaMalloc(&ctx->sel_blocks, sizeof(int) * (size_t) ctx->sel_blocks_cap));
    }
    int *totals = ctx->sel_blocks + 2 * (size_t) nblocks;
    PB_LAUNCH(pb_k_sel2_count, nblocks, SEL_T, n, nblocks, dim, on_lo, on_hi, b_lo, b_hi, ctx->pos, ctx->flags, ctx->sel_blocks);
    PB_LAUNCH(pb_k_sel2_scan, 2, 1024, nblocks, ctx->sel_blocks, totals);
    for(int attempt = 0; attempt < 2; attempt++) {
        // optimistic scatter into the current capacity; if the lists turn out too short, grow and scatter again
        PB_LAUNCH(pb_k_sel2_scatter, nblocks, SEL_T, n, nblocks, dim, on_lo, on_hi, b_lo, b_hi, ctx->pbc[j0], ctx->pbc[j1], ctx->nsend_all,
                  ctx->send_cap, ctx->pos, ctx->flags, ctx->sel_blocks, totals, ctx->send_map, ctx->send_mult);
        if(attempt == 0) {
            PB_CHECK(cudaMemcpyAsync(ctx->h_scalars, totals, sizeof(int) * 2, cudaMemcpyDeviceToHost, ctx->stream));
            PB_CHECK(cudaStreamSynchronize(ctx->stream));
        }
        const int need = ctx->nsend_all + ctx->h_scalars[0] + ctx->h_scalars[1];
        if(need <= ctx->send_cap) { break; }
        PB_TRY(pb_ensure_send_capacity(ctx, need));
    }
    *count_lo = ctx->h_scalars[0];
    *count_hi = ctx->h_scalars[1];
    return 0;
}

// SetCommunicationOffsets (comm.py:262-288)
static void pb_set_offsets(pb_ctx *ctx, int step) {
    int isend = 0, irecv = 0;
    for(int i = 0; i < step; i++) {
        for(int j = i * 2; j < i * 2 + 2; j++) { isend += ctx->nsend[j]; irecv += ctx->nrecv[j]; }
    }
    for(int j = step * 2; j < step * 2 + 2; j++) {
        ctx->send_offsets[j] = isend; ctx->recv_offsets[j] = irecv;
        isend += ctx->nsend[j]; irecv += ctx->nrecv[j];
    }
}

// ---- borders -----------------------------------------------------------------------------------------------
template<bool DEM>
__global__ void __launch_bounds__(256) pb_k_pack_border(int first, int count, int cap, int stride, PbBox box, const int *__restrict__ send_map,
                                                        const int *__restrict__ send_mult, const double4 *__restrict__ pos,
                                                        const double *__restrict__ vel, const double *__restrict__ mass,
                                                        const int *__restrict__ uid, const int *__restrict__ shape,
                                                        const int *__restrict__ tag, double *__restrict__ buf,
                                                        const double *__restrict__ radius, const double *__restrict__ angvel) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= count) { return; }
    const int e = first + k;
    const int p = send_map[e];
    const double4 x = pos[p];
    double *b = buf + (size_t) e * stride;
    if(DEM) {
        b[11] = radius[p];
        b[12] = angvel[p];
        b[13] = angvel[cap + p];
        b[14] = angvel[2 * cap + p];
    }
    b[0] = (double) uid[p];
    b[1] = (double) pb_w_type(x.w);
    b[2] = mass[p];
    b[3] = __dadd_rn(x.x, __dmul_rn((double) send_mult[e * 3 + 0], box.len[0]));
    b[4] = __dadd_rn(x.y, __dmul_rn((double) send_mult[e * 3 + 1], box.len[1]));
    b[5] = __dadd_rn(x.z, __dmul_rn((double) send_mult[e * 3 + 2], box.len[2]));
    b[6] = vel[p];
    b[7] = vel[cap + p];
    b[8] = vel[2 * cap + p];
    b[9] = (double) shape[p];
    b[10] = (double) tag[p];
}

template<bool DEM>
__global__ void __launch_bounds__(256) pb_k_unpack_border(int first_rec, int count, int dst0, int cap, int stride, const double *__restrict__ buf,
                                                          double4 *__restrict__ pos, double *__restrict__ vel, double *__restrict__ mass,
                                                          int *__restrict__ type, int *__restrict__ flags, int *__restrict__ uid,
                                                          int *__restrict__ shape, int *__restrict__ tag, double *__restrict__ radius,
                                                          double *__restrict__ angvel) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= count) { return; }
    const double *b = buf + (size_t) (first_rec + k) * stride;
    const int p = dst0 + k;
    if(DEM) {
        radius[p] = b[11];
        angvel[p] = b[12];
        angvel[cap + p] = b[13];
        angvel[2 * cap + p] = b[14];
    }
    const int t = (int) b[1];
    uid[p] = (int) b[0];
    type[p] = t;
    mass[p] = b[2];
    pos[p] = make_double4(b[3], b[4], b[5], pb_type_w(t));
    vel[p] = b[6];
    vel[cap + p] = b[7];
    vel[2 * cap + p] = b[8];
    shape[p] = (int) b[9];
    tag[p] = (int) b[10];
    // The reference never transmits ghost flags (the slot keeps a stale value); here they are
    // defined: GHOST only -- so ghosts are binned (not INFINITE), forwarded (not GLOBAL) and never integrated.
    flags[p] = PB_FLAG_GHOST;
}

static PbBox pb_box(const pb_ctx *ctx) {
    PbBox b;
    for(int d = 0; d < 3; d++) { b.len[d] = ctx->grid[d * 2 + 1] - ctx->grid[d * 2]; }
    return b;
}

static int pb_borders_move(pb_ctx *ctx, int step, int stride, int base_elems);

extern "C" int pb_borders(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbStage st(ctx, "borders");
    if(!ctx->cells_set) { ctx->set_error("pb_borders: cell spacing unknown (pb_setup_cells)"); return -1; }
    ctx->nsend_all = 0;
    ctx->nghost = 0;
    ctx->cells_n = 0;
    for(int j = 0; j < 6; j++) { ctx->nsend[j] = 0; ctx->nrecv[j] = 0; ctx->send_offsets[j] = 0; ctx->recv_offsets[j] = 0; }
    // record length: the built-in elements, then the non-volatile rows of the user-defined properties (props.cu)
    const int base_elems = ctx->dem ? BORDER_ELEMS_DEM : BORDER_ELEMS;
    const int stride = base_elems + ctx->xrows_nv;
    for(int step = 0; step < 3; step++) {
        const int n = ctx->nlocal + ctx->nghost;     // locals AND ghosts received so far: edges/corners are forwarded
        int c_lo = 0, c_hi = 0;
        PB_TRY(pb_select_dim(ctx, n, step, ctx->spacing, &c_lo, &c_hi));
        ctx->nsend[step * 2] = c_lo;
        ctx->nsend[step * 2 + 1] = c_hi;
        ctx->nsend_all += c_lo + c_hi;
        PB_TRY(pb_transport_sizes(ctx, step));
        pb_set_offsets(ctx, step);
        const int nr = ctx->nrecv[step * 2] + ctx->nrecv[step * 2 + 1];
        PB_TRY(pb_ensure_particle_capacity(ctx, ctx->nlocal + ctx->nghost + nr));
        PB_TRY(pb_borders_move(ctx, step, stride, base_elems));
        ctx->nghost += nr;
    }
    return 0;
}

// pb_md_run_from_host: the ghosts of the first list build were created while velocities and masses were still on their way, so
// their records carried whatever the arrays held.  Once the arrays are complete the three phases move the records again -- same
// send lists, same order (a forwarded ghost takes its values from the ghost of the earlier phase) -- and every ghost holds what
// pb_borders would have given it.
int pb_borders_refill(pb_ctx *ctx) {
    const int base_elems = ctx->dem ? BORDER_ELEMS_DEM : BORDER_ELEMS;
    const int stride = base_elems + ctx->xrows_nv;
    for(int step = 0; step < 3; step++) { PB_TRY(pb_borders_move(ctx, step, stride, base_elems)); }
    return 0;
}

// pack -> transport -> unpack of one phase of the ghost creation, over the send lists pb_select_dim left
static int pb_borders_move(pb_ctx *ctx, int step, int stride, int base_elems) {
    const int ns = ctx->nsend[step * 2] + ctx->nsend[step * 2 + 1];
    const int nr = ctx->nrecv[step * 2] + ctx->nrecv[step * 2 + 1];
    if(ns > 0) {
        if(ctx->dem) {
            PB_LAUNCH(pb_k_pack_border<true>, pb_blocks(ns, 256), 256, ctx->send_offsets[step * 2], ns, ctx->pcap, stride, pb_box(ctx),
                      ctx->send_map, ctx->send_mult, ctx->pos, ctx->vel, ctx->mass, ctx->uid, ctx->shape, ctx->tag, ctx->send_buf,
                      ctx->radius, ctx->angvel);
        } else {
            PB_LAUNCH(pb_k_pack_border<false>, pb_blocks(ns, 256), 256, ctx->send_offsets[step * 2], ns, ctx->pcap, stride, pb_box(ctx),
                      ctx->send_map, ctx->send_mult, ctx->pos, ctx->vel, ctx->mass, ctx->uid, ctx->shape, ctx->tag, ctx->send_buf,
                      nullptr, nullptr);
        }
        PB_TRY(pb_xprops_pack(ctx, ctx->send_offsets[step * 2], ns, stride, base_elems, ctx->send_map, ctx->send_buf));
    }
    const double *src = nullptr;
    PB_TRY(pb_transport_data(ctx, step, step + 1, stride, &src));
    if(nr > 0) {
        if(ctx->dem) {
            PB_LAUNCH(pb_k_unpack_border<true>, pb_blocks(nr, 256), 256, ctx->recv_offsets[step * 2], nr,
                      ctx->nlocal + ctx->recv_offsets[step * 2], ctx->pcap, stride, src, ctx->pos, ctx->vel, ctx->mass, ctx->type, ctx->flags,
                      ctx->uid, ctx->shape, ctx->tag, ctx->radius, ctx->angvel);
        } else {
            PB_LAUNCH(pb_k_unpack_border<false>, pb_blocks(nr, 256), 256, ctx->recv_offsets[step * 2], nr,
                      ctx->nlocal + ctx->recv_offsets[step * 2], ctx->pcap, stride, src, ctx->pos, ctx->vel, ctx->mass, ctx->type, ctx->flags,
                      ctx->uid, ctx->shape, ctx->tag, nullptr, nullptr);
        }
        PB_TRY(pb_xprops_unpack(ctx, ctx->recv_offsets[step * 2], nr, ctx->nlocal + ctx->recv_offsets[step * 2], stride, base_elems, src));
    }
    return 0;
}

// ---- synchronize -------------------------------------------------------------------------------------------
// Comm.synchronize packs EVERY send entry from the current arrays in one pass, then transfers, then unpacks
// (comm.py:45-54; generated pack_all_ghost_particles / unpack_all_ghost_particles).  A forwarded ghost (edge or
// corner image, whose source is itself a ghost) therefore carries the coordinates its source ghost had BEFORE this
// refresh -- one step stale per forwarding level.  That behaviour is part of the reference's results and is
// reproduced here by construction: one pack kernel over all entries, then one unpack kernel.
__global__ void __launch_bounds__(256) pb_k_pack_sync(int count, int cap, int nlocal, int stride, PbBox box, const int *__restrict__ send_map,
                                                      const int *__restrict__ send_mult, const double4 *__restrict__ pos,
                                                      const double4 *__restrict__ pos_ghost, const double *__restrict__ vel,
                                                      const double *__restrict__ angvel, double *__restrict__ buf) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= count) { return; }
    const int p = send_map[e];
    const double4 x = (p < nlocal) ? pos[p] : pos_ghost[p];     // ghosts: the values of the previous refresh
    double *b = buf + (size_t) e * stride;
    if(angvel != nullptr) {                                     // DEM: angular velocity is refreshed too (comm.py:49)
        b[6] = angvel[p];
        b[7] = angvel[cap + p];
        b[8] = angvel[2 * cap + p];
    }
    b[0] = __dadd_rn(x.x, __dmul_rn((double) send_mult[e * 3 + 0], box.len[0]));
    b[1] = __dadd_rn(x.y, __dmul_rn((double) send_mult[e * 3 + 1], box.len[1]));
    b[2] = __dadd_rn(x.z, __dmul_rn((double) send_mult[e * 3 + 2], box.len[2]));
    b[3] = vel[p];
    b[4] = vel[cap + p];
    b[5] = vel[2 * cap + p];
}

__global__ void __launch_bounds__(256) pb_k_unpack_sync(int count, int dst0, int cap, int stride, const double *__restrict__ buf,
                                                        const double4 *__restrict__ pos_ghost, double4 *__restrict__ pos,
                                                        double *__restrict__ vel, double *__restrict__ angvel) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= count) { return; }
    const double *b = buf + (size_t) k * stride;
    const int p = dst0 + k;
    if(angvel != nullptr) {
        angvel[p] = b[6];
        angvel[cap + p] = b[7];
        angvel[2 * cap + p] = b[8];
    }
    const double w = pos_ghost[p].w;
    pos[p] = make_double4(b[0], b[1], b[2], w);
    vel[p] = b[3];
    vel[cap + p] = b[4];
    vel[2 * cap + p] = b[5];
}

extern "C" int pb_synchronize(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbStage st(ctx, "synchronize");
    const double4 *pos_ghost = ctx->ghosts_in_alt ? ctx->pos_alt : ctx->pos;
    const int stride = ctx->dem ? SYNC_ELEMS_DEM : SYNC_ELEMS;
    double *angvel = ctx->dem ? ctx->angvel : nullptr;
    if(ctx->nsend_all > 0) {
        PB_LAUNCH(pb_k_pack_sync, pb_blocks(ctx->nsend_all, 256), 256, ctx->nsend_all, ctx->pcap, ctx->nlocal, stride, pb_box(ctx), ctx->send_map,
                  ctx->send_mult, ctx->pos, pos_ghost, ctx->vel, angvel, ctx->send_buf);
    }
    const double *src = nullptr;
    PB_TRY(pb_transport_data(ctx, 0, 3, stride, &src));
    if(ctx->nghost > 0) {
        PB_LAUNCH(pb_k_unpack_sync, pb_blocks(ctx->nghost, 256), 256, ctx->nghost, ctx->nlocal, ctx->pcap, stride, src, pos_ghost, ctx->pos, ctx->vel,
                  angvel);
    }
    ctx->ghosts_in_alt = false;
    // inside pb_md_run the mirror of the tile lists follows the refreshed ghosts (tile_lists.cu)
    if(ctx->mirror_scope && ctx->mirror_fresh && ctx->tiles_n == ctx->nlocal) { PB_TRY(pb_tile_mirror_ghosts(ctx)); }
    return 0;
}

// ---- exchange ----------------------------------------------------------------------------------------------
// Single rank in a dimension: "leaving" through a face means re-entering through the opposite one, i.e.
// position += pbc * L applied in place (what pack (comm.py:320-321) + self-copy + unpack produce).  Side 0 and
// side 1 are decided from the position BEFORE the shift, as the reference determines both sides before packing.
__global__ void __launch_bounds__(256) pb_k_wrap(int nlocal, int dim, double lo, double hi, int do_lo, int do_hi, double shift_lo,
                                                 double shift_hi, const int *__restrict__ flags, double4 *__restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= nlocal || (flags[i] & (PB_FLAG_INFINITE | PB_FLAG_GLOBAL)) != 0) { return; }
    double4 p = pos[i];
    const double x = (dim == 0) ? p.x : ((dim == 1) ? p.y : p.z);
    double nx = x;
    if(do_lo && x < lo) { nx = __dadd_rn(x, shift_lo); }
    else if(do_hi && x > hi) { nx = __dadd_rn(x, shift_hi); }
    else { return; }
    // the other two axes get "+ 0 * L" in the reference: x + 0.0 == x bit-for-bit except for -0.0, which cannot be
    // told apart numerically; they are left untouched
    if(dim == 0) { p.x = nx; } else if(dim == 1) { p.y = nx; } else { p.z = nx; }
    pos[i] = p;
}

int pb_exchange_multi(pb_ctx *ctx, int dim);

extern "C" int pb_exchange(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbStage st(ctx, "exchange");
    if(!ctx->domain_set || !ctx->cells_set) { ctx->set_error("pb_exchange: domain / cells not initialised"); return -1; }
    // ghosts are discarded by the reference at this point as well (nghost = 0, comm.py:108)
    ctx->nghost = 0;
    ctx->nsend_all = 0;
    ctx->cells_n = 0;
    ctx->neigh_n = -1;
    ctx->tiles_n = -1;
    ctx->ghosts_in_alt = false;
    for(int dim = 0; dim < 3; dim++) {
        if(ctx->nranks[dim] == 1) {
            if(ctx->nlocal == 0) { continue; }
            const double L = ctx->grid[dim * 2 + 1] - ctx->grid[dim * 2];
            const int do_lo = ctx->pbc_flag[dim] || ctx->pbc[dim * 2] == 0;
            const int do_hi = ctx->pbc_flag[dim] || ctx->pbc[dim * 2 + 1] == 0;
            PB_LAUNCH(pb_k_wrap, pb_blocks(ctx->nlocal, 256), 256, ctx->nlocal, dim, ctx->subdom[dim * 2], ctx->subdom[dim * 2 + 1],
                      do_lo, do_hi, (double) ctx->pbc[dim * 2] * L, (double) ctx->pbc[dim * 2 + 1] * L, ctx->flags, ctx->pos);
        } else {
            PB_TRY(pb_exchange_multi(ctx, dim));
        }
    }
    // cell-order reordering of the locals (north-star item a): the reference permutes locals here too (hole filling,
    // comm.py:445-506); ours is a stable counting sort on the flat cell index
    // (DEM: particles keep their index so that the per-particle contact tables need not move; see dem_kernels.cu)
    if(!ctx->dem) { PB_TRY(pb_sort_locals(ctx)); }
    return 0;
}
