// Cell binning as a deterministic counting sort + cell-order reordering of the local particles.
// Replaces BuildCellListsStencil / BuildCellLists / PartitionCellLists (sim/cell_lists.py:46-171) and the dense
// cell_particles[ncells][cell_capacity] array (25.6 MB, 64-slot rows, atomic slot claims) by
//   particle_cell[i]  -- bit-identical to the reference's value (same fp64 subtract / divide / truncate / clamp)
//   cell_start[c], cell_list[k]  -- CSR cell list; inside a cell particles are ordered by (z sub-bin, index)
// so there is no cell_capacity to overflow and the result is run-to-run reproducible.
// Every cell is split into `zsub` slabs along z (a memory-layout choice, not a reference quantity) and the counting sort runs on
// the combined key cell*zsub + slab: sub_start[] lets the neighbour-list build open each (dx,dy) row of the stencil exactly at
// the z-window that can contain neighbours instead of scanning three whole cells (neighbor.cu).
//
// Kernels (all HBM-bound; algorithmic bytes per binned particle: pos 32 + flags 4 + particle_cell 4 w + slot 4 w,
// then slot 4 r + cell 4 r + list 4 w; the per-cell arrays are ncells*4 B and stay in L2):
//   pb_k_cell_count   cell index + warp-aggregated histogram (one atomic per distinct cell per warp)
//   pb_k_scan_*       exclusive scan of the histogram
//   pb_k_cell_fill    scatter indices to cell_start[c] + slot
//   pb_k_cell_sort    per-cell insertion sort of the (<= ~30) indices -> deterministic order
//   pb_k_reorder_*    gather of every per-particle array through the permutation (cell order)
#include <algorithm>
#include <cmath>

#include "ctx.cuh"
#include "md_math.h"

// ---- BuildCellListsStencil (sim/cell_lists.py:46-87): host, same fp64 expression order ---------------------
extern "C" int pb_setup_cells(pb_ctx *ctx, double spacing) {
    if(!ctx->domain_set) { ctx->set_error("pb_setup_cells: domain not initialised"); return -1; }
    PB_CHECK(cudaSetDevice(ctx->device));
    ctx->spacing = spacing;
    for(int d = 0; d < 3; d++) {
        const double hi = ctx->subdom[d * 2 + 1] + spacing;
        const double lo = ctx->subdom[d * 2 + 0] - spacing;
        const double len = hi - lo;
        const double q = len / spacing;
        ctx->dim_cells[d] = ((int) ceil(q)) + 1;
    }
    const long nc = (long) ctx->dim_cells[0] * ctx->dim_cells[1] * ctx->dim_cells[2] + 1;
    if(nc > 0x7fffff00L) { ctx->set_error("pb_setup_cells: too many cells"); return -1; }
    ctx->ncells = (int) nc;
    int k = 0;
    for(int i = -1; i < 2; i++) {
        for(int j = -1; j < 2; j++) {
            for(int l = -1; l < 2; l++) { ctx->stencil[k++] = (i * ctx->dim_cells[1] + j) * ctx->dim_cells[2] + l; }
        }
    }
    // z slabs per cell: only the Verlet-list build profits from them; DEM rebins every step over many tiny cells -> 1
    ctx->zsub_active = ctx->dem ? 1 : ctx->zsub;
    if((long) ctx->ncells * ctx->zsub_active > 0x7ffffff0L) { ctx->set_error("pb_setup_cells: too many cell slabs"); return -1; }
    const long need = (long) ctx->ncells * ctx->zsub_active + 2;
    if(need > ctx->ccap) {
        for(int **q : {&ctx->cell_count, &ctx->cell_start, &ctx->sub_start}) {
            if(*q != nullptr) { PB_CHECK(cudaFree(*q)); }
            PB_CHECK(cudaMalloc(q, sizeof(int) * (size_t) (need + 2)));
        }
        ctx->ccap = (int) need;
    }
    ctx->cells_set = true;
    ctx->cells_n = 0;
    return 0;
}

extern "C" int pb_get_cells(const pb_ctx *ctx, int dim_cells[3], int *ncells, int stencil[27]) {
    for(int d = 0; d < 3; d++) { dim_cells[d] = ctx->dim_cells[d]; }
    *ncells = ctx->ncells;
    for(int k = 0; k < 27; k++) { stencil[k] = ctx->stencil[k]; }
    return 0;
}

// ---- exclusive scan (3 kernels, 2048 items per block) -----------------------------------------------------
static const int SCAN_T = 512;
static const int SCAN_ITEMS = 4;
static const int SCAN_BLOCK = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ int pb_block_exclusive_scan(int v, int *total, int *warp_sums) {
    // inclusive warp scan
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if(lane >= o) { x += y; }
    }
    if(lane == 31) { warp_sums[wid] = x; }
    __syncthreads();
    if(wid == 0) {
        int w = (lane < (int) (blockDim.x >> 5)) ? warp_sums[lane] : 0;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if(lane >= o) { w += y; }
        }
        warp_sums[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const int warp_off = (wid == 0) ? 0 : warp_sums[wid - 1];
    *total = warp_sums[(blockDim.x >> 5) - 1];
    return warp_off + x - v;
}

__global__ void __launch_bounds__(SCAN_T) pb_k_scan_reduce(const int *__restrict__ in, int n, int *__restrict__ block_sums) {
    __shared__ int warp_sums[32];
    const int base = blockIdx.x * SCAN_BLOCK;
    int s = 0;
#pragma unroll
    for(int k = 0; k < SCAN_ITEMS; k++) {
        const int i = base + k * SCAN_T + threadIdx.x;
        s += (i < n) ? in[i] : 0;
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); }
    if((threadIdx.x & 31) == 0) { warp_sums[threadIdx.x >> 5] = s; }
    __syncthreads();
    if(threadIdx.x < 32) {
        int w = (threadIdx.x < (SCAN_T >> 5)) ? warp_sums[threadIdx.x] : 0;
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) { w += __shfl_down_sync(0xffffffffu, w, o); }
        if(threadIdx.x == 0) { block_sums[blockIdx.x] = w; }
    }
}

// single block: exclusive scan of the block sums in place (serial over chunks of SCAN_T)
__global__ void __launch_bounds__(SCAN_T) pb_k_scan_sums(int *__restrict__ block_sums, int nblocks, int *__restrict__ total_out) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    if(threadIdx.x == 0) { carry_s = 0; }
    __syncthreads();
    for(int base = 0; base < nblocks; base += SCAN_T) {
        const int i = base + threadIdx.x;
        const int v = (i < nblocks) ? block_sums[i] : 0;
        int total;
        const int ex = pb_block_exclusive_scan(v, &total, warp_sums);
        const int carry = carry_s;
        if(i < nblocks) { block_sums[i] = carry + ex; }
        __syncthreads();
        if(threadIdx.x == 0) { carry_s = carry + total; }
        __syncthreads();
    }
    if(threadIdx.x == 0) { *total_out = carry_s; }
}

__global__ void __launch_bounds__(SCAN_T) pb_k_scan_apply(const int *__restrict__ in, int n, const int *__restrict__ block_sums,
                                                           int *__restrict__ out) {
    __shared__ int warp_sums[32];
    const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;   // blocked arrangement: thread owns 4 consecutive items
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for(int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int total;
    int ex = pb_block_exclusive_scan(s, &total, warp_sums) + block_sums[blockIdx.x];
#pragma unroll
    for(int k = 0; k < SCAN_ITEMS; k++) {
        if(base + k < n) { out[base + k] = ex; }
        ex += v[k];
    }
}

// out[0..n-1] = exclusive prefix sums, out[n] = total
int pb_exclusive_scan(pb_ctx *ctx, const int *in, int *out, int n) {
    if(n <= 0) {
        PB_CHECK(cudaMemsetAsync(out, 0, sizeof(int), ctx->stream));
        return 0;
    }
    const int nblocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if(nblocks > ctx->scan_tmp_cap) {
        if(ctx->scan_tmp != nullptr) { PB_CHECK(cudaFree(ctx->scan_tmp)); }
        ctx->scan_tmp_cap = nblocks + 1024;
        PB_CHECK(cudaMalloc(&ctx->scan_tmp, sizeof(int) * (size_t) ctx->scan_tmp_cap));
    }
    PB_LAUNCH(pb_k_scan_reduce, nblocks, SCAN_T, in, n, ctx->scan_tmp);
    PB_LAUNCH(pb_k_scan_sums, 1, SCAN_T, ctx->scan_tmp, nblocks, out + n);
    PB_LAUNCH(pb_k_scan_apply, nblocks, SCAN_T, in, n, ctx->scan_tmp, out);
    return 0;
}

// ---- binning ----------------------------------------------------------------------------------------------
// z slab of a particle inside its cell (consistent with the reference cell index c2 computed above)
__device__ __forceinline__ int pb_zslab(const PbCellGeom &g, double z, int zsub) {
    const double q2 = (z - g.lo[2]) / g.spacing;
    int c2 = (int) q2;
    c2 = (c2 >= 0) ? c2 : 0; c2 = (c2 < g.dim[2]) ? c2 : g.dim[2] - 1;
    int zs = (int) ((q2 - (double) c2) * (double) zsub);
    return min(max(zs, 0), zsub - 1);
}

// Warp-aggregated histogram: lanes of a warp that hit the same cell elect a leader which issues ONE atomicAdd
// for the whole group; every lane derives its slot from the leader's base + its rank inside the group.
__global__ void __launch_bounds__(256) pb_k_cell_count(PbCellGeom g, int first, int n, const double4 *__restrict__ pos,
                                                       const int *__restrict__ flags, int *__restrict__ particle_cell,
                                                       int *__restrict__ cell_count, int *__restrict__ cell_slot,
                                                       int *__restrict__ cell_key, int zsub) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < n;
    int cell = -1;
    if(active) {
        const double4 p = pos[first + k];
        const int c = pb_cell_index(g, p.x, p.y, p.z, flags[first + k]);
        particle_cell[first + k] = c;
        cell = c * zsub + ((c == 0) ? 0 : pb_zslab(g, p.z, zsub));      // counting-sort key: (cell, z slab)
        cell_key[first + k] = cell;
    }
    const unsigned live = __ballot_sync(0xffffffffu, active);
    if(!active) { return; }
    const unsigned peers = __match_any_sync(live, cell);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    const int rank_in_group = __popc(peers & ((1u << lane) - 1u));
    int base = 0;
    if(lane == leader) { base = atomicAdd(&cell_count[cell], __popc(peers)); }
    base = __shfl_sync(peers, base, leader);
    cell_slot[first + k] = base + rank_in_group;
}

__global__ void __launch_bounds__(256) pb_k_cell_fill(int first, int n, const int *__restrict__ cell_key,
                                                      const int *__restrict__ cell_slot, const int *__restrict__ sub_start,
                                                      int *__restrict__ cell_list) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k < n) {
        const int i = first + k;
        cell_list[sub_start[cell_key[i]] + cell_slot[i]] = i;
    }
}

// coarse CSR (one entry per reference cell) out of the slab CSR
__global__ void __launch_bounds__(256) pb_k_cell_coarse(int ncells, int zsub, const int *__restrict__ sub_start, int *__restrict__ cell_start) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if(c <= ncells) { cell_start[c] = sub_start[(size_t) c * zsub]; }
}

// One thread per slab: insertion sort of the slab's (tiny) run by index -- a total order, hence deterministic whatever order
// the atomics delivered.
__global__ void __launch_bounds__(128) pb_k_cell_sort(int nbins, const int *__restrict__ sub_start, int *__restrict__ cell_list) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if(c >= nbins) { return; }
    const int b = sub_start[c], e = sub_start[c + 1];
    for(int i = b + 1; i < e; i++) {
        const int v = cell_list[i];
        int j = i - 1;
        while(j >= b && cell_list[j] > v) {
            cell_list[j + 1] = cell_list[j];
            j--;
        }
        cell_list[j + 1] = v;
    }
}

static PbCellGeom pb_geom(const pb_ctx *ctx) {
    PbCellGeom g;
    for(int d = 0; d < 3; d++) {
        g.lo[d] = ctx->subdom[d * 2] - ctx->spacing;
        g.dim[d] = ctx->dim_cells[d];
    }
    g.spacing = ctx->spacing;
    g.ncells = ctx->ncells;
    return g;
}

// Bins particles [first, first+n): particle_cell, cell_start[0..ncells], cell_list[0..n) (absolute indices).
int pb_bin_particles(pb_ctx *ctx, int first, int n, bool /*write_particle_cell*/) {
    if(!ctx->cells_set) { ctx->set_error("cell lists not set up (pb_setup_cells)"); return -1; }
    const int S = ctx->zsub_active;
    const long nbins = (long) ctx->ncells * S;
    PB_CHECK(cudaMemsetAsync(ctx->cell_count, 0, sizeof(int) * ((size_t) nbins + 1), ctx->stream));
    if(n > 0) {
        PB_LAUNCH(pb_k_cell_count, pb_blocks(n, 256), 256, pb_geom(ctx), first, n, ctx->pos, ctx->flags, ctx->particle_cell,
                  ctx->cell_count, ctx->cell_slot, ctx->cell_key, S);
    }
    PB_TRY(pb_exclusive_scan(ctx, ctx->cell_count, ctx->sub_start, (int) nbins));
    if(n > 0) {
        PB_LAUNCH(pb_k_cell_fill, pb_blocks(n, 256), 256, first, n, ctx->cell_key, ctx->cell_slot, ctx->sub_start, ctx->cell_list);
        PB_LAUNCH(pb_k_cell_sort, pb_blocks(nbins, 128), 128, (int) nbins, ctx->sub_start, ctx->cell_list);
    }
    PB_LAUNCH(pb_k_cell_coarse, pb_blocks(ctx->ncells + 1, 256), 256, ctx->ncells, S, ctx->sub_start, ctx->cell_start);
    return 0;
}

// ---- public: BuildCellLists + PartitionCellLists over locals + ghosts -------------------------------------
extern "C" int pb_build_cell_lists(pb_ctx *ctx) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbStage st(ctx, "build_cell_lists");
    const int n = ctx->nlocal + ctx->nghost;
    PB_TRY(pb_bin_particles(ctx, 0, n, true));
    ctx->cells_n = n;
    return 0;
}

extern "C" int pb_download_cell_lists(pb_ctx *ctx, int *cell_start, int *cell_list) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PB_CHECK(cudaMemcpyAsync(cell_start, ctx->cell_start, sizeof(int) * ((size_t) ctx->ncells + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if(ctx->cells_n > 0) {
        PB_CHECK(cudaMemcpyAsync(cell_list, ctx->cell_list, sizeof(int) * (size_t) ctx->cells_n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- cell-order reordering of the locals ------------------------------------------------------------------
// PART: 3 = every array; 1 = all but velocity and mass, 2 = velocity and mass (the split of an overlapped upload, see below)
template<int PART>
__global__ void __launch_bounds__(256) pb_k_reorder(int n, int cap, const int *__restrict__ perm,
                                                    const double4 *__restrict__ pos, double4 *__restrict__ pos_o,
                                                    const double *__restrict__ vel, double *__restrict__ vel_o,
                                                    const double *__restrict__ mass, double *__restrict__ mass_o,
                                                    const int *__restrict__ type, int *__restrict__ type_o,
                                                    const int *__restrict__ flags, int *__restrict__ flags_o,
                                                    const int *__restrict__ uid, int *__restrict__ uid_o,
                                                    const int *__restrict__ shape, int *__restrict__ shape_o,
                                                    const int *__restrict__ tag, int *__restrict__ tag_o) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= n) { return; }
    const int s = perm[k];
    if(PART & 1) {
        pos_o[k] = pos[s];
        type_o[k] = type[s];
        flags_o[k] = flags[s];
        uid_o[k] = uid[s];
        shape_o[k] = shape[s];
        tag_o[k] = tag[s];
    }
    if(PART & 2) {
        vel_o[k] = vel[s];
        vel_o[cap + k] = vel[cap + s];
        vel_o[2 * cap + k] = vel[2 * cap + s];
        mass_o[k] = mass[s];
    }
}

// Sort the locals into cell order (stable: ties keep ascending previous index).  Called from pb_exchange, before
// ghosts exist; volatile `force` is not carried (it is reset before the next force evaluation, exactly as the
// reference never transfers volatile properties, sim/comm.py:103).
int pb_sort_locals(pb_ctx *ctx) {
    const int n = ctx->nlocal;
    if(n == 0) { return 0; }
    PB_TRY(pb_bin_particles(ctx, 0, n, true));
    if(ctx->upload_pending) {
        // pb_md_run_from_host: velocities and masses are still arriving on comm_stream.  Positions and the integer arrays are
        // permuted here; velocities and masses follow on comm_stream, behind their copies, through a copy of the permutation
        // (cell_list is rewritten by the cell-list build before they are through).
        if((size_t) n > ctx->upload_perm_cap) {
            if(ctx->upload_perm != nullptr) { PB_CHECK(cudaFree(ctx->upload_perm)); ctx->upload_perm = nullptr; ctx->upload_perm_cap = 0; }
            PB_CHECK(cudaMalloc(&ctx->upload_perm, sizeof(int) * (size_t) ctx->pcap));
            ctx->upload_perm_cap = (size_t) ctx->pcap;
        }
        PB_CHECK(cudaMemcpyAsync(ctx->upload_perm, ctx->cell_list, sizeof(int) * (size_t) n, cudaMemcpyDeviceToDevice, ctx->stream));
        PB_CHECK(cudaEventRecord(ctx->ev_io[0], ctx->stream));
        PB_LAUNCH(pb_k_reorder<1>, pb_blocks(n, 256), 256, n, ctx->pcap, ctx->cell_list, ctx->pos, ctx->pos_alt, ctx->vel, ctx->vel_alt,
                  ctx->mass, ctx->mass_alt, ctx->type, ctx->type_alt, ctx->flags, ctx->flags_alt, ctx->uid, ctx->uid_alt,
                  ctx->shape, ctx->shape_alt, ctx->tag, ctx->tag_alt);
        PB_CHECK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_io[0], 0));
        std::swap(ctx->stream, ctx->comm_stream);
        cudaError_t e = cudaSuccess;
        pb_k_reorder<2><<<pb_blocks(n, 256), 256, 0, ctx->stream>>>(n, ctx->pcap, ctx->upload_perm, ctx->pos, ctx->pos_alt, ctx->vel, ctx->vel_alt,
                                                                     ctx->mass, ctx->mass_alt, ctx->type, ctx->type_alt, ctx->flags, ctx->flags_alt,
                                                                     ctx->uid, ctx->uid_alt, ctx->shape, ctx->shape_alt, ctx->tag, ctx->tag_alt);
        e = cudaGetLastError();
        if(e == cudaSuccess) { e = cudaEventRecord(ctx->ev_io[1], ctx->stream); }
        std::swap(ctx->stream, ctx->comm_stream);
        PB_CHECK(e);
    } else {
        PB_LAUNCH(pb_k_reorder<3>, pb_blocks(n, 256), 256, n, ctx->pcap, ctx->cell_list, ctx->pos, ctx->pos_alt, ctx->vel, ctx->vel_alt,
                  ctx->mass, ctx->mass_alt, ctx->type, ctx->type_alt, ctx->flags, ctx->flags_alt, ctx->uid, ctx->uid_alt,
                  ctx->shape, ctx->shape_alt, ctx->tag, ctx->tag_alt);
    }
    std::swap(ctx->pos, ctx->pos_alt);
    std::swap(ctx->vel, ctx->vel_alt);
    std::swap(ctx->mass, ctx->mass_alt);
    std::swap(ctx->type, ctx->type_alt);
    std::swap(ctx->flags, ctx->flags_alt);
    std::swap(ctx->uid, ctx->uid_alt);
    std::swap(ctx->shape, ctx->shape_alt);
    std::swap(ctx->tag, ctx->tag_alt);
    PB_TRY(pb_xprops_permute(ctx, ctx->cell_list, n));     // user-defined properties follow their particle
    return 0;
}
