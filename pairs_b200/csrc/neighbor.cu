// Neighbour-list build: full Verlet lists in padded column-major (ELLPACK) storage.
// Replaces BuildNeighborLists (sim/neighbor_lists.py:21-48, loop nest sim/interaction.py:91-120).
//
// Reference semantics kept exactly: for every local, non-FIXED particle i visit cell 0 (INFINITE particles) and the
// 27 stencil cells `particle_cell[i] + stencil[k]` that satisfy 0 < cell < ncells (flat-index test only, no per-axis
// wrap test), and append every j != i with  (dx*dx + dy*dy) + dz*dz < cutoff^2  (separate fp64 multiplies and adds,
// no FMA contraction) -- so the neighbour SETS are bit-identical to the reference's.  Storage differs:
//   interleaved sliced ELLPACK (PbNeighLayout in ctx.cuh: a warp reads 32 consecutive ints per iteration) instead of
//   AoS [i][k]; inside a list, neighbours are in ascending cell order, ascending index inside a cell (deterministic).
// The three z-adjacent stencil cells of one (dx,dy) row are consecutive flat indices, so each row is ONE
// contiguous run of the CSR cell list: 9 runs + cell 0 per particle, each run narrowed to the z slabs within reach.
//
// Half lists (Simulation.compute_half(), sim/simulation.py:119-120): only partners with a larger index are stored -- each
// local-local pair once, a local-ghost pair at its local end -- as UNORDERED pairs the same set as the reference's.
//
// HBM bytes per local particle: pos 32 + cell 4 + list write 4*K + count 4; candidates (~27 cells * occupancy)
// are served from L1/L2 because the 32 lanes of a warp sit in the same 1-3 cells.
#include <algorithm>

#include "ctx.cuh"

struct PbFaces {
    double lo[3], hi[3];   // subdom_min + margin, subdom_max - margin
};

struct PbBuildGeom {
    double lo[3];          // origin of the cell grid (subdom_min - spacing)
    double spacing, inv_slab;   // cell edge; zsub / spacing
    int dim1, dim2, zsub, ncells;
};

// One thread per local particle.  For each of the 9 (dx,dy) rows of the stencil the three z-adjacent cells are ONE contiguous
// run of the slab CSR; the run is opened only between the z slabs that can hold a neighbour:
//     |z_j - z_i| <= w,   w = sqrt(cutoff^2 - d_xy^2),   d_xy = distance in the xy-plane from i to the row's cell column
// (rows with d_xy >= cutoff are skipped altogether).  The window is widened by a relative 1e-9 and rounded outwards to whole
// slabs, so it is conservative; membership is still decided by the exact reference test below.
// STAGE: each warp collects its 32 lists in shared memory ([k][lane], conflict-free) and writes them out afterwards row by
// row, 128 contiguous bytes per row -- instead of ~75 scattered 4-byte stores per particle whose 32-byte sectors are completed
// by eight different store instructions (ncu: 2.2 GB of DRAM writes for 1.25 GB of lists without staging).
template<bool STAGE>
__global__ void __launch_bounds__(128) pb_k_build_neighbors(int nlocal, int ncap, PbNeighLayout lay, PbBuildGeom g, double cutsq,
                                                            const double4 *__restrict__ pos, const int *__restrict__ flags,
                                                            const int *__restrict__ particle_cell, const int *__restrict__ sub_start,
                                                            const int *__restrict__ cell_list, int *__restrict__ neigh,
                                                            int *__restrict__ numneigh, int *__restrict__ max_count, PbFaces faces,
                                                            int *__restrict__ group_flag, int half) {
    extern __shared__ int s_stage[];                       // [warps per block][ncap][32]
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int *const s_mine = s_stage + (size_t) (threadIdx.x >> 5) * ncap * 32 + lane;
    int count = 0;
    int boundary = 0;     // has a ghost neighbour, or is itself a halo source (within `margin` of a sub-box face)
    if(i < nlocal && (flags[i] & PB_FLAG_FIXED) == 0) {
        const double4 pi = pb_ld_pos(pos + i);
        boundary = (pi.x < faces.lo[0]) | (pi.x > faces.hi[0]) | (pi.y < faces.lo[1]) | (pi.y > faces.hi[1]) |
                   (pi.z < faces.lo[2]) | (pi.z > faces.hi[2]);
        const int pc = particle_cell[i];
        // list slot of neighbour k: base + (k / G) * 32 + k % G  (PbNeighLayout::idx with the per-particle part hoisted)
        int *const out = neigh + ((size_t) (i / lay.A) * lay.T * 32 + (size_t) ((i % lay.A) * lay.G));
        const int G = lay.G;
        // cell coordinates of i (pc = (c0*dim1 + c1)*dim2 + c2 + 1) and its offsets inside the cell
        const int flat = pc - 1;
        const int c2 = flat % g.dim2, c1 = (flat / g.dim2) % g.dim1, c0 = flat / (g.dim2 * g.dim1);
        const double fx = pi.x - (g.lo[0] + c0 * g.spacing), fy = pi.y - (g.lo[1] + c1 * g.spacing);   // in [0, spacing)
        const double zrel = pi.z - g.lo[2];
        const double slack = 1e-9 * g.spacing;
        const int zslabs = g.dim2 * g.zsub;
        // (an INFINITE local particle -- cell 0 -- would get the stencil around flat index 0 in the reference, a handful of
        // cells at the grid origin; such particles are FIXED in practice and skipped above; here they see cell 0 only)
        const int nruns = (pc == 0) ? 1 : 10;
        for(int run = 0; run < nruns; run++) {
            int b, e;
            if(run == 0) {                      // cell 0: INFINITE particles (sim/interaction.py:93-95, disp = -1)
                b = sub_start[0]; e = sub_start[g.zsub];
            } else {
                const int r = run - 1;
                const int dx = r / 3 - 1, dy = r % 3 - 1;
                const double ddx = (dx == 0) ? 0.0 : ((dx < 0) ? fx : g.spacing - fx);
                const double ddy = (dy == 0) ? 0.0 : ((dy < 0) ? fy : g.spacing - fy);
                const double wsq = cutsq - (ddx * ddx + ddy * ddy);
                if(wsq <= 0.0) { continue; }
                const double w = sqrt(wsq) + slack;
                // flat index of the row's cell at z-index 0, +1 for the reserved cell 0; the reference tests only
                // 0 < cell < ncells on the flat index, which for a local particle (never in the outermost cell layer) always holds
                const long col = ((long) (c0 + dx) * g.dim1 + (c1 + dy)) * g.dim2 + 1;
                int gz_lo = (int) floor((zrel - w) * g.inv_slab), gz_hi = (int) floor((zrel + w) * g.inv_slab);
                gz_lo = max(gz_lo, max((c2 - 1) * g.zsub, 0));
                gz_hi = min(gz_hi, min((c2 + 1) * g.zsub + g.zsub - 1, zslabs - 1));
                const long s_lo = col * g.zsub + gz_lo, s_hi = col * g.zsub + gz_hi;
                if(gz_lo > gz_hi || s_lo < g.zsub || s_hi >= (long) g.ncells * g.zsub) { continue; }
                b = sub_start[s_lo]; e = sub_start[s_hi + 1];
            }
            for(int k = b; k < e; k++) {
                const int j = __ldg(cell_list + k);
                const double4 pj = pb_ld_pos(pos + j);
                const double dx = __dsub_rn(pi.x, pj.x);
                const double dy = __dsub_rn(pi.y, pj.y);
                const double dz = __dsub_rn(pi.z, pj.z);
                const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                // compute_half() (sim/interaction.py:107-113) keeps shape[j] > shape[i] || (shape[j] == shape[i] && i < j); every
                // particle on the neighbour-list path has the same shape, which leaves i < j (ghosts sit behind all locals)
                if(rsq < cutsq && (half ? (j > i) : (j != i))) {
                    if(count < ncap) {
                        if(STAGE) { s_mine[count * 32] = j; }
                        else if(G == 1) { out[(size_t) count * 32] = j; }
                        else { out[(size_t) (count / G) * 32 + (count % G)] = j; }
                    }
                    count++;
                    boundary |= (j >= nlocal);
                }
            }
        }
    }
    if(STAGE) {
        // flush (all 32 lanes take part): row k of the warp's slice is 32 consecutive ints in global memory (G == 1 layout)
        __syncwarp();
        const int mine = min(count, ncap);
        int rows = mine;
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) { rows = max(rows, __shfl_xor_sync(0xffffffffu, rows, o)); }
        int *const dst = neigh + ((size_t) (i >> 5) * lay.T * 32 + (size_t) lane);
        for(int k = 0; k < rows; k++) {
            if(k < mine) { dst[(size_t) k * 32] = s_mine[k * 32]; }
        }
    }
    if(i < nlocal) { numneigh[i] = count; }
    // one flag per warp group (= 32 consecutive particles = one warp of the force kernel when G = 1)
    const unsigned any_b = __ballot_sync(0xffffffffu, boundary != 0);
    if((threadIdx.x & 31) == 0 && i < nlocal) { group_flag[i >> 5] = any_b != 0u; }
    // block-wide max -> one atomic per warp
    int m = count;
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) { m = max(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if((threadIdx.x & 31) == 0 && m > 0) { atomicMax(max_count, m); }
}

static PbNeighLayout pb_layout(const pb_ctx *ctx) {
    PbNeighLayout lay;
    lay.G = ctx->lanes;
    lay.A = 32 / ctx->lanes;
    lay.T = (ctx->ncap + ctx->lanes - 1) / ctx->lanes;
    return lay;
}

static int pb_alloc_neigh(pb_ctx *ctx, int n) {
    const PbNeighLayout lay = pb_layout(ctx);
    const size_t groups = ((size_t) n + lay.A - 1) / lay.A;
    const size_t bytes = sizeof(int) * groups * (size_t) lay.T * 32;
    if(bytes > ctx->neigh_bytes) {
        if(ctx->neigh != nullptr) { PB_CHECK(cudaFree(ctx->neigh)); ctx->neigh = nullptr; }
        const size_t want = bytes + bytes / 8;
        PB_CHECK(cudaMalloc(&ctx->neigh, want));
        ctx->neigh_bytes = want;
    }
    return 0;
}

// Ordered compaction of the warp-group ids into an interior list (no ghost neighbour, no halo source: can be computed
// while the ghost refresh is in flight) and a boundary list.
__global__ void __launch_bounds__(256) pb_k_split_groups(int ngroups, const int *__restrict__ flag, const int *__restrict__ scan,
                                                         int *__restrict__ interior, int *__restrict__ boundary) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if(g >= ngroups) { return; }
    if(flag[g]) { boundary[scan[g]] = g; } else { interior[g - scan[g]] = g; }
}

static int pb_split_groups(pb_ctx *ctx, int ngroups) {
    PB_TRY(pb_exclusive_scan(ctx, ctx->group_flag, ctx->group_scan, ngroups));
    PB_LAUNCH(pb_k_split_groups, pb_blocks(ngroups, 256), 256, ngroups, ctx->group_flag, ctx->group_scan, ctx->groups_interior,
              ctx->groups_boundary);
    PB_CHECK(cudaMemcpyAsync(ctx->h_scalars, ctx->group_scan + ngroups, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->n_boundary = ctx->h_scalars[0];
    ctx->n_interior = ngroups - ctx->n_boundary;
    ctx->groups_valid = true;
    return 0;
}

// The per-particle 32-bit lists (sliced ELLPACK): the list format of half lists, several lanes per particle, DEM-free contexts the
// tile planner declines, and of every kernel that walks lists by particle (generated pair kernels, energy / virial, legacy lj).
static int pb_build_neigh32(pb_ctx *ctx, double cutoff) {
    const int n = ctx->nlocal;
    ctx->neigh_n = -1;
    const double cutsq = cutoff * cutoff;
    if(ctx->ncap <= 0) { ctx->ncap = 100; }
    const int ngroups = (n + 31) / 32;
    if(ngroups + 1 > ctx->group_cap) {
        for(int **q : {&ctx->group_flag, &ctx->group_scan, &ctx->groups_interior, &ctx->groups_boundary}) {
            if(*q != nullptr) { PB_CHECK(cudaFree(*q)); }
            PB_CHECK(cudaMalloc(q, sizeof(int) * ((size_t) ngroups + 1024)));
        }
        ctx->group_cap = ngroups + 1023;
    }
    PbFaces faces;
    for(int d = 0; d < 3; d++) {
        faces.lo[d] = ctx->subdom[d * 2] + ctx->spacing;
        faces.hi[d] = ctx->subdom[d * 2 + 1] - ctx->spacing;
    }
    ctx->groups_valid = false;
    PbBuildGeom bg;
    for(int d = 0; d < 3; d++) { bg.lo[d] = ctx->subdom[d * 2] - ctx->spacing; }
    bg.spacing = ctx->spacing;
    bg.inv_slab = (double) ctx->zsub_active / ctx->spacing;
    bg.dim1 = ctx->dim_cells[1];
    bg.dim2 = ctx->dim_cells[2];
    bg.zsub = ctx->zsub_active;
    bg.ncells = ctx->ncells;   // neighbor_capacity default of pairs.simulation() (src/pairs/__init__.py:16)
    for(int attempt = 0; attempt < 8; attempt++) {
        PB_TRY(pb_alloc_neigh(ctx, n));
        ctx->nslots = pb_layout(ctx).T;
        PB_CHECK(cudaMemsetAsync(ctx->d_scalars, 0, sizeof(int), ctx->stream));
        const size_t stage_bytes = (size_t) 4 * ctx->ncap * 32 * sizeof(int);      // 4 warps per block
        if(ctx->stage_lists && ctx->lanes == 1 && stage_bytes <= 96 * 1024) {
            PB_CHECK(cudaFuncSetAttribute(pb_k_build_neighbors<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) stage_bytes));
            pb_k_build_neighbors<true><<<pb_blocks(n, 128), 128, stage_bytes, ctx->stream>>>(
                n, ctx->ncap, pb_layout(ctx), bg, cutsq, ctx->pos, ctx->flags, ctx->particle_cell, ctx->sub_start, ctx->cell_list, ctx->neigh,
                ctx->numneigh, ctx->d_scalars, faces, ctx->group_flag, ctx->half_lists ? 1 : 0);
            ctx->launches++;
            PB_CHECK(cudaGetLastError());
        } else {
            PB_LAUNCH(pb_k_build_neighbors<false>, pb_blocks(n, 128), 128, n, ctx->ncap, pb_layout(ctx), bg, cutsq, ctx->pos, ctx->flags,
                      ctx->particle_cell, ctx->sub_start, ctx->cell_list, ctx->neigh, ctx->numneigh, ctx->d_scalars, faces, ctx->group_flag,
                      ctx->half_lists ? 1 : 0);
        }
        PB_CHECK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
        ctx->max_neigh = ctx->h_scalars[0];
        if(ctx->max_neigh <= ctx->ncap) {
            if(ctx->world > 1 && ctx->overlap_comm && ctx->lanes == 1) { PB_TRY(pb_split_groups(ctx, ngroups)); }
            ctx->neigh_n = n;         // only now: an error return above must not leave half-built lists looking valid
            return 0;
        }
        // capacity-overflow protocol (transformations/modules.py:159-203): grow to twice the need and re-run the module
        ctx->ncap = ctx->max_neigh * 2;
    }
    ctx->set_error("pb_build_neighbor_lists: capacity did not converge");
    return -1;
}

// BuildNeighborLists (sim/neighbor_lists.py:21-48).  Default: tile lists (tile_lists.cu); the per-particle format where those do
// not apply.
extern "C" int pb_build_neighbor_lists(pb_ctx *ctx, double cutoff) {
    PB_CHECK(cudaSetDevice(ctx->device));
    PbStage st(ctx, "build_neighbor_lists");
    if(ctx->cells_n != ctx->nlocal + ctx->nghost) {
        ctx->set_error("pb_build_neighbor_lists: cell lists are stale (call pb_build_cell_lists first)");
        return -1;
    }
    ctx->list_cutoff = cutoff;
    ctx->neigh_n = -1;
    ctx->tiles_n = -1;
    ctx->groups_valid = false;
    ctx->tile_split_valid = false;
    if(ctx->nlocal == 0) { ctx->neigh_n = 0; ctx->max_neigh = 0; return 0; }
    if(cutoff > ctx->spacing * (1.0 + 1e-12)) { ctx->set_error("pb_build_neighbor_lists: cutoff exceeds the cell spacing"); return -1; }
    const int rc = pb_build_tile_lists(ctx, cutoff);
    if(rc <= 0) { return rc; }
    return pb_build_neigh32(ctx, cutoff);
}

int pb_tile_export_ell(pb_ctx *ctx, int *neigh, int T);       // tile_lists.cu

// For the kernels that walk per-particle lists: when the lists of this reneighbouring are tile lists, the 32-bit lists are
// DERIVED from them (slot -> particle index: same sets, same order, whatever the particles have done since the build).
int pb_require_neigh32(pb_ctx *ctx) {
    if(ctx->neigh_n == ctx->nlocal) { return 0; }
    if(ctx->tiles_n != ctx->nlocal) { ctx->set_error("neighbour lists are stale"); return -1; }
    if(ctx->lanes != 1) { ctx->set_error("per-particle lists out of tile lists need lanes_per_particle = 1"); return -1; }
    PbStage st(ctx, "neighbor_lists_32");
    PB_TRY(pb_alloc_neigh(ctx, ctx->nlocal));
    ctx->nslots = pb_layout(ctx).T;
    PB_TRY(pb_tile_export_ell(ctx, ctx->neigh, ctx->nslots));
    ctx->neigh_n = ctx->nlocal;
    return 0;
}

extern "C" int pb_neighbor_capacity(const pb_ctx *ctx) { return ctx->ncap; }
extern "C" int pb_max_neighbors(const pb_ctx *ctx) { return ctx->max_neigh; }

__global__ void pb_k_neigh_to_aos(int n, int cap_out, int ncap, PbNeighLayout lay, const int *__restrict__ neigh,
                                  const int *__restrict__ numneigh, int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) { return; }
    const int c = min(numneigh[i], min(cap_out, ncap));
    for(int k = 0; k < c; k++) { out[(size_t) i * cap_out + k] = neigh[lay.idx(i, k)]; }
    for(int k = c; k < cap_out; k++) { out[(size_t) i * cap_out + k] = -1; }
}

extern "C" int pb_download_neighbors(pb_ctx *ctx, int *out, int capacity) {
    PB_CHECK(cudaSetDevice(ctx->device));
    if(ctx->tiles_n == ctx->nlocal && ctx->neigh_n != ctx->nlocal) { return pb_tile_download_neighbors(ctx, out, capacity); }
    const int n = ctx->neigh_n;
    if(n == 0) { return 0; }
    PbScratch stage_buf;
    PB_CHECK(stage_buf.alloc(sizeof(int) * (size_t) n * (size_t) capacity));
    int *const stage = stage_buf.as<int>();
    PB_LAUNCH(pb_k_neigh_to_aos, pb_blocks(n, 128), 128, n, capacity, ctx->ncap, pb_layout(ctx), ctx->neigh, ctx->numneigh, stage);
    PB_CHECK(cudaMemcpyAsync(out, stage, sizeof(int) * (size_t) n * (size_t) capacity, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
