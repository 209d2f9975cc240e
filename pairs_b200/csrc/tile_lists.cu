// Tile lists: cell tiles staged in shared memory, 16-bit tile-relative neighbour lists -- the list format of the MD hot path
// (BuildNeighborLists sim/neighbor_lists.py:21-48 + the Lennard-Jones pair kernel sim/interaction.py:201-292, examples/md.py:5-8).
//
// Why.  With one 32-byte global gather per neighbour the force kernel is bound by the L1TEX data pipe (round 1: 90 % of peak,
// ~22 wavefronts per warp-wide gather of 32 scattered records, fp64 pipe 51 %).  The neighbours of the particles of a few adjacent
// cells all live in the 3 x 3 cell columns around them, and in cell order each of those columns is ONE contiguous run of the cell
// CSR.  So a CTA takes a TILE -- the cells [za, zb] of a 2 x 2 block of cell columns ("super-column"), cut so that it holds at
// most PB_TILE_M particles -- and
//   1. stages the 4 x 4 columns around it over [za - 1, zb + 1] (16 contiguous CSR runs, <= PB_TILE_CAP particles, 6.3 x the
//      tile's own on average) into shared memory with cp.async: x, y as one 16-byte entry, z as an 8-byte entry per particle
//      (no registers, no L1 allocation by the consumer side),
//   2. walks neighbour lists whose entries are 16-bit SLOTS of that staging order (12 bits slot, 3 bits particle type), four
//      entries per 64-bit word, sliced ELLPACK over list rows (row = tile-major particle order, so a warp reads 256
//      contiguous bytes per four iterations) -- the gather becomes LDS.128 + LDS.64: ~14 data-pipe wavefronts per warp and
//      neighbour instead of ~22, and half the list bytes from HBM.
// The list build uses the same staging: candidates are tested out of shared memory with the reference's exact predicate
// (separate fp64 multiplies and adds, no contraction), so the neighbour SETS stay bit-identical; inside a list the order is the
// per-particle builder's (stencil rows in (dx, dy) order, ascending CSR position), hence with the exact arithmetic the forces
// are bit-identical to the per-particle kernel's, too (tools/micro/tile_force.cu checks both on 4 M atoms).
// Measured on a B200 (4 M atoms, tools/micro/tile_force.cu, profiles/r02_*): force 0.888 -> 0.744 ms with the reference's
// arithmetic, 0.652 ms with fused multiply-adds (option "lj_fma"); list build 2.89 -> 2.62 ms.
//
// What stays per particle: `numneigh[i]`, the force / velocity / position arrays, the fused integrator epilogue
// (md_kernels.cu).  What moves to tiles: the interior / boundary split for comm-compute overlap (a tile is "boundary" if one of
// its particles has a ghost neighbour or is a halo source).  Kernels that walk 32-bit per-particle lists (generated pair
// kernels, energy / virial, half lists, the legacy lj) get them built on demand (pb_require_neigh32, neighbor.cu).
//
// Not applicable -> pb_build_tile_lists returns 1 and the caller takes the per-particle path: half lists, more than one lane
// per particle, DEM contexts, INFINITE particles in cell 0, more than 8 particle types, or a z level of a super-column that
// alone exceeds the tile's particle or staging capacity (a density the 2 x 2 x 1-cell granularity cannot serve).
#include <algorithm>

#include "ctx.cuh"
#include "md_math.h"

static const int PB_TILE_NRUN = 16;
// The last staging slot holds no particle but a point far outside every cutoff: list rows are padded to whole words (four
// entries) with it, so the force kernel needs no "is this entry valid" test -- a padding entry simply fails the cutoff test.
// (Up to two slots per staged run stay empty, see PbTileHdr: the planner keeps 32 slots back for them.)
static const int PB_TILE_DUMMY = PB_TILE_CAP - 1, PB_TILE_STAGED = PB_TILE_CAP - 1 - 2 * PB_TILE_NRUN;
static const double PB_TILE_FAR = 1e150;      // (x - 1e150)^2 * 3 is finite and > any cutoff^2

struct PbTileGeom {
    double lo[3];               // origin of the cell grid (subdom_min - spacing)
    double spacing, inv_slab;   // cell edge; zsub / spacing
    int dim0, dim1, dim2, zsub;
};

// Run table of a tile, computed ONCE per list build (pb_k_tile_headers) and read by every kernel that visits the tile: the 16
// staged runs (columns X0-1 .. X0+2, Y0-1 .. Y0+2 over [za-1, zb+1]; CSR begin, length, first staging slot) and the 4 core runs.
// Staging slots: run r starts at an even slot plus the parity of its CSR begin -- the 8-byte z entries of a run are then 16-byte
// aligned in shared memory exactly where they are in the mirror (TMA bulk copies need 16-byte alignment on both sides).
struct PbTileHdr {
    int total_bytes, ncore, any_local, nslots;
    int run_begin[PB_TILE_NRUN], run_len[PB_TILE_NRUN], run_slot0[PB_TILE_NRUN];
    int core_begin[4], core_off[5];
    int bytes32;      // sum of run_len * 16: the float4 staging of the list build's pre-filter
    int pad[2];
};
static_assert(sizeof(PbTileHdr) == 256, "one tile header = 64 ints");

static PbTileGeom pb_tile_geom(const pb_ctx *ctx) {
    PbTileGeom g;
    for(int d = 0; d < 3; d++) { g.lo[d] = ctx->subdom[d * 2] - ctx->spacing; }
    g.spacing = ctx->spacing;
    g.inv_slab = (double) ctx->zsub_active / ctx->spacing;
    g.dim0 = ctx->dim_cells[0];
    g.dim1 = ctx->dim_cells[1];
    g.dim2 = ctx->dim_cells[2];
    g.zsub = ctx->zsub_active;
    return g;
}

// ---- planner: tiles of <= PB_TILE_M particles and <= PB_TILE_CAP staged particles ----------------------------------------------
// particles per (super-column, z level): in the 2 x 2 core columns and in the 4 x 4 staged columns
__global__ void __launch_bounds__(256) pb_k_tile_levels(int nsx, int nsy, int dim0, int dim1, int dim2, const int *__restrict__ cell_start,
                                                        int *__restrict__ lvl_core, int *__restrict__ lvl_halo) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= nsx * nsy * dim2) { return; }
    const int z = t % dim2, sc = t / dim2;
    const int X0 = (sc / nsy) * 2, Y0 = (sc % nsy) * 2;
    int core = 0, halo = 0;
    for(int a = -1; a <= 2; a++) {
        for(int b = -1; b <= 2; b++) {
            const int X = X0 + a, Y = Y0 + b;
            if(X < 0 || X >= dim0 || Y < 0 || Y >= dim1) { continue; }
            const int c = (X * dim1 + Y) * dim2 + z + 1;
            const int k = cell_start[c + 1] - cell_start[c];
            halo += k;
            if(a >= 0 && a <= 1 && b >= 0 && b <= 1) { core += k; }
        }
    }
    lvl_core[t] = core;
    lvl_halo[t] = halo;
}

// one thread per super-column: greedy walk over z.  WRITE = false counts the tiles, WRITE = true emits them at off[sc].
template<bool WRITE>
__global__ void __launch_bounds__(128) pb_k_tile_plan(int nsc, int nsy, int dim2, const int *__restrict__ lvl_core, const int *__restrict__ lvl_halo,
                                                      int *__restrict__ cnt, const int *__restrict__ off, PbTile *__restrict__ tiles,
                                                      int *__restrict__ pad, int *__restrict__ overflow) {
    const int sc = blockIdx.x * blockDim.x + threadIdx.x;
    if(sc >= nsc) { return; }
    const int *lc = lvl_core + (size_t) sc * dim2, *lh = lvl_halo + (size_t) sc * dim2;
    int ntiles = 0;
    int z = 0;
    while(z < dim2) {
        if(lc[z] == 0) { z++; continue; }
        const int za = z;
        int core = lc[z];
        int staged = ((z > 0) ? lh[z - 1] : 0) + lh[z] + ((z + 1 < dim2) ? lh[z + 1] : 0);
        if(core > PB_TILE_M || staged > PB_TILE_STAGED) { atomicExch(overflow, 1); }
        int zb = z;
        while(zb + 1 < dim2) {
            const int c2 = core + lc[zb + 1];
            const int s2 = staged + ((zb + 2 < dim2) ? lh[zb + 2] : 0);
            if(c2 > PB_TILE_M || s2 > PB_TILE_STAGED) { break; }
            core = c2;
            staged = s2;
            zb++;
        }
        if(WRITE) {
            PbTile t;
            t.X0 = (sc / nsy) * 2; t.Y0 = (sc % nsy) * 2; t.za = za; t.zb = zb; t.row_base = 0; t.ncore = core;
            tiles[off[sc] + ntiles] = t;
            pad[off[sc] + ntiles] = (core + 31) / 32 * 32;
        }
        ntiles++;
        z = zb + 1;
    }
    if(!WRITE) { cnt[sc] = ntiles; }
}

__global__ void __launch_bounds__(256) pb_k_tile_rows(int ntiles, const int *__restrict__ row, PbTile *__restrict__ tiles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t < ntiles) { tiles[t].row_base = row[t]; }
}

// ---- headers, mirror ------------------------------------------------------------------------------------------------------------
// one warp per tile: lanes 0..15 the staged runs, lanes 16..19 the core runs, all lanes look for a LOCAL particle in the core
__global__ void __launch_bounds__(128) pb_k_tile_headers(int ntiles, int nlocal, PbTileGeom g, const PbTile *__restrict__ tiles,
                                                         const int *__restrict__ cell_start, const int *__restrict__ cell_list,
                                                         PbTileHdr *__restrict__ hdrs) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if(t >= ntiles) { return; }
    const PbTile tl = tiles[t];
    int begin = 0, len = 0;
    if(lane < PB_TILE_NRUN) {
        const int X = tl.X0 - 1 + lane / 4, Y = tl.Y0 - 1 + lane % 4;
        if(X >= 0 && X < g.dim0 && Y >= 0 && Y < g.dim1) {
            const int zb = max(tl.za - 1, 0), ze = min(tl.zb + 1, g.dim2 - 1);
            begin = cell_start[(X * g.dim1 + Y) * g.dim2 + zb + 1];
            len = cell_start[(X * g.dim1 + Y) * g.dim2 + ze + 2] - begin;
        }
    } else if(lane < PB_TILE_NRUN + 4) {
        const int q = lane - PB_TILE_NRUN;
        const int X = tl.X0 + q / 2, Y = tl.Y0 + q % 2;
        if(X < g.dim0 && Y < g.dim1) {
            begin = cell_start[(X * g.dim1 + Y) * g.dim2 + tl.za + 1];
            len = cell_start[(X * g.dim1 + Y) * g.dim2 + tl.zb + 2] - begin;
        }
    }
    // staging slots and copy sizes: a sequential walk over the 16 runs (lane 0), core offsets (lane 16)
    PbTileHdr *h = hdrs + t;
    int slot0 = 0, acc = 0, bytes = 0, bytes32 = 0;
    for(int r = 0; r < PB_TILE_NRUN; r++) {
        const int rb = __shfl_sync(0xffffffffu, begin, r), rl = __shfl_sync(0xffffffffu, len, r);
        const int s0 = acc + (rb & 1);
        if(lane == r) { slot0 = s0; }
        if(rl > 0) {
            acc = (s0 + rl + 1) & ~1;
            bytes += rl * 16 + (((rb + rl + 1) & ~1) - (rb & ~1)) * 8;
            bytes32 += rl * 16;
        }
    }
    int off = 0, core_total = 0;
    for(int q = 0; q < 4; q++) {
        const int ql = __shfl_sync(0xffffffffu, len, PB_TILE_NRUN + q);
        if(lane == PB_TILE_NRUN + q) { off = core_total; }
        core_total += ql;
    }
    // any local particle among the core particles?  (tiles of ghosts only are skipped by every kernel)
    int found = 0;
    for(int q = 0; q < 4; q++) {
        const int qb = __shfl_sync(0xffffffffu, begin, PB_TILE_NRUN + q), ql = __shfl_sync(0xffffffffu, len, PB_TILE_NRUN + q);
        for(int k = lane; k < ql; k += 32) { found |= (cell_list[qb + k] < nlocal); }
    }
    found = __any_sync(0xffffffffu, found);
    if(lane < PB_TILE_NRUN) { h->run_begin[lane] = begin; h->run_len[lane] = len; h->run_slot0[lane] = slot0; }
    else if(lane < PB_TILE_NRUN + 4) { h->core_begin[lane - PB_TILE_NRUN] = begin; h->core_off[lane - PB_TILE_NRUN] = off; }
    if(lane == 0) { h->bytes32 = bytes32; h->total_bytes = bytes; h->ncore = core_total; h->any_local = found; h->nslots = acc; h->core_off[4] = core_total; }
}

// The mirror: positions a second time in CSR order (locals and ghosts alike), split into xy (16 bytes) and z (8 bytes), so that
// every staged run is ONE contiguous range of each array.  Written in full here (after every cell-list build, and whenever the
// mirror is not known to be current), kept current inside pb_md_run by the fused force kernel (locals) and pb_tile_mirror_ghosts.
__global__ void __launch_bounds__(256) pb_k_tile_mirror(int nall, int nlocal, const int *__restrict__ cell_list, const double4 *__restrict__ pos,
                                                        double2 *__restrict__ mxy, double *__restrict__ mz, unsigned char *__restrict__ mmeta,
                                                        int *__restrict__ ghost_csr, float4 *__restrict__ m32) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= nall) { return; }
    const int i = __ldg(cell_list + k);
    const double4 p = pb_ld_pos(pos + i);
    mxy[k] = make_double2(p.x, p.y);
    mz[k] = p.z;
    const int meta = (pb_w_type(p.w) & 7) | ((i >= nlocal) ? 8 : 0);
    mmeta[k] = (unsigned char) meta;
    // the list build's pre-filter: the position rounded to fp32, the meta byte in the fourth lane (valid at build time only)
    // (fourth lane: the type where a list entry carries it, bits 12..14, and the ghost flag in bit 15)
    if(m32 != nullptr) {
        m32[k] = make_float4(__double2float_rn(p.x), __double2float_rn(p.y), __double2float_rn(p.z), __int_as_float(((meta & 7) << 12) | ((meta & 8) << 12)));
    }
    if(i >= nlocal) { ghost_csr[i - nlocal] = k; }
}

__global__ void __launch_bounds__(256) pb_k_tile_mirror_ghosts(int nghost, int nlocal, const int *__restrict__ ghost_csr, const double4 *__restrict__ pos,
                                                               double2 *__restrict__ mxy, double *__restrict__ mz) {
    const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
    if(gidx >= nghost) { return; }
    const int k = __ldg(ghost_csr + gidx);
    const double4 p = pb_ld_pos(pos + nlocal + gidx);
    mxy[k] = make_double2(p.x, p.y);
    mz[k] = p.z;
}

// ---- staging: TMA bulk copies, completion on an mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned pb_smem_addr(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void pb_mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pb_smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pb_mbar_expect_tx(unsigned long long *bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pb_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pb_mbar_wait(unsigned long long *bar, int parity) {
    asm volatile("{\n.reg .pred p;\nPB_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra PB_DONE_%=;\nbra PB_WAIT_%=;\nPB_DONE_%=:\n}"
                 ::"r"(pb_smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void pb_bulk_g2s(void *dst, const void *src, int bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pb_smem_addr(dst)), "l"(src), "r"(bytes), "r"(pb_smem_addr(bar)) : "memory");
}

// shared memory: [hdr 256][mbarrier 16][xy: CAP*16][z: CAP*8][meta: CAP (build)]
__device__ __forceinline__ void pb_tile_smem(unsigned char *base, PbTileHdr *&h, unsigned long long *&bar, double2 *&sxy, double *&sz, unsigned char *&smeta) {
    h = reinterpret_cast<PbTileHdr *>(base);
    bar = reinterpret_cast<unsigned long long *>(base + 256);
    sxy = reinterpret_cast<double2 *>(base + 272);
    sz = reinterpret_cast<double *>(sxy + PB_TILE_CAP);
    smeta = reinterpret_cast<unsigned char *>(sz + PB_TILE_CAP);
}
static size_t pb_tile_smem_bytes(bool build) { return 272 + (size_t) PB_TILE_CAP * 24 + (build ? (size_t) PB_TILE_CAP : 0); }

// Header into shared memory, barrier armed, the 2 x 16 copies issued (threads 0..15, one run each).  Returns false -- for the
// whole CTA, before anything is in flight -- when the tile holds ghosts only.  The caller waits with pb_mbar_wait(bar, 0).
__device__ __forceinline__ bool pb_tile_stage(const PbTileHdr *__restrict__ hg, PbTileHdr *h, unsigned long long *bar, const double2 *__restrict__ mxy,
                                              const double *__restrict__ mz, double2 *sxy, double *sz) {
    const int t = threadIdx.x;
    if(t < 64) { reinterpret_cast<int *>(h)[t] = __ldg(reinterpret_cast<const int *>(hg) + t); }
    if(t == 0) {
        pb_mbar_init(bar, 1);
        sxy[PB_TILE_DUMMY] = make_double2(PB_TILE_FAR, PB_TILE_FAR);
        sz[PB_TILE_DUMMY] = PB_TILE_FAR;
    }
    __syncthreads();
    if(!h->any_local) { return false; }
    if(t < PB_TILE_NRUN) {
        const int len = h->run_len[t], begin = h->run_begin[t], slot0 = h->run_slot0[t];
        if(len > 0) {
            pb_bulk_g2s(sxy + slot0, mxy + begin, len * 16, bar);
            const int zb = begin & ~1, ze = (begin + len + 1) & ~1;
            pb_bulk_g2s(sz + (slot0 - (begin & 1)), mz + zb, (ze - zb) * 8, bar);
        }
    }
    if(t == 0) { pb_mbar_expect_tx(bar, h->total_bytes); }
    return true;
}

// core column of the t-th core particle of the tile (-1: none)
__device__ __forceinline__ int pb_tile_core_q(const PbTileHdr *h, int t) {
    if(t >= h->ncore) { return -1; }
    int q = 0;
    if(t >= h->core_off[1]) { q = 1; }
    if(t >= h->core_off[2]) { q = 2; }
    if(t >= h->core_off[3]) { q = 3; }
    return q;
}
// its CSR position, and its own slot in the staging order (core column q is staged run (q / 2 + 1, q % 2 + 1) of the 4 x 4)
__device__ __forceinline__ int pb_tile_core_csr(const PbTileHdr *h, int t, int q) { return h->core_begin[q] + (t - h->core_off[q]); }
__device__ __forceinline__ int pb_tile_self_slot(const PbTileHdr *h, int q, int cs) {
    const int tr = (q / 2 + 1) * 4 + (q % 2 + 1);
    return h->run_slot0[tr] + (cs - h->run_begin[tr]);
}

// word q of list row r (4 entries per word): ((r / 32) * T4 + q) * 32 + r % 32
__device__ __forceinline__ size_t pb_tile_word(int row, int T4, int q) { return ((size_t) (row >> 5) * T4 + (size_t) q) * 32 + (size_t) (row & 31); }

// ---- list build -----------------------------------------------------------------------------------------------------------
struct PbTileFaces { double lo[3], hi[3]; };

// z-window of stencil row r for a particle at (fx, fy) inside its cell column, zrel above the grid origin: CSR range [b, e) of
// the slab CSR (the per-particle builder's window, neighbor.cu: widened by a relative 1e-9 and rounded outwards to whole slabs)
__device__ __forceinline__ bool pb_tile_window(const PbTileGeom &g, int c0, int c1, int c2, double fx, double fy, double zrel, double cutsq, int r,
                                               const int *__restrict__ sub_start, int &b, int &e) {
    const int dx = r / 3 - 1, dy = r % 3 - 1;
    const double ddx = (dx == 0) ? 0.0 : ((dx < 0) ? fx : g.spacing - fx);
    const double ddy = (dy == 0) ? 0.0 : ((dy < 0) ? fy : g.spacing - fy);
    const double wsq = cutsq - (ddx * ddx + ddy * ddy);
    if(wsq <= 0.0) { return false; }
    const int X = c0 + dx, Y = c1 + dy;
    if(X < 0 || X >= g.dim0 || Y < 0 || Y >= g.dim1) { return false; }
    const double w = sqrt(wsq) + 1e-9 * g.spacing;
    const long col = ((long) X * g.dim1 + Y) * g.dim2 + 1;
    int gz_lo = (int) floor((zrel - w) * g.inv_slab), gz_hi = (int) floor((zrel + w) * g.inv_slab);
    gz_lo = max(gz_lo, max((c2 - 1) * g.zsub, 0));
    gz_hi = min(gz_hi, min((c2 + 1) * g.zsub + g.zsub - 1, g.dim2 * g.zsub - 1));
    if(gz_lo > gz_hi) { return false; }
    b = sub_start[col * g.zsub + gz_lo];
    e = sub_start[col * g.zsub + gz_hi + 1];
    return true;
}

// One thread per core particle of the tile:
//   * the z windows of all nine stencil rows are computed first, so their 18 slab-CSR loads are in flight together;
//   * an accepted candidate -- one in five -- enters the pending 64-bit word with one funnel shift; the list order is the
//     per-particle builder's (stencil rows in (dx, dy) order, ascending CSR position).
//   (Measured and dropped: branch-free tests that leave a bit per candidate in a mask, appended afterwards by walking the set
//   bits -- the per-row append loops run to the longest lane's count, 12 of 32 lanes active: 5.5 ms instead of 3.0.)
// REORDER (option "tile_reorder", the default): the finished row is re-read, put into the conflict-aware order below through
// the (then idle) staging area, and written again -- see pb_tile_reorder_row.
struct PbTileBuildArgs {
    int nlocal, ncap, T4, reorder;
    PbTileGeom g;
    double cutsq;
    const PbTile *tiles;
    const PbTileHdr *hdrs;
    const double2 *mxy;
    const double *mz;
    const unsigned char *mmeta;
    const float4 *m32;
    const double4 *pos;
    const int *flags, *particle_cell, *sub_start, *cell_list;
    unsigned long long *words;
    unsigned char *rowsrc;
    int *numneigh, *max_count, *tile_flag;
    PbTileFaces faces;
};

// Conflict-aware order of a list row.  In iteration k of the force kernel the 16 lanes of a half-warp read 16 staged particles:
// two DIFFERENT slots collide in shared memory when they agree modulo 16 (8-byte z entries: 16 bank pairs; the 16-byte xy
// entries of a quarter-warp: modulo 8).  A list is a set, its order is free: lane l puts at position k an entry whose slot is
// (k + l) mod 16 whenever it still has one -- the n-th entry of residue class r goes to k = 16 n + ((r - l) mod 16) -- so the
// lanes of a half-warp ask for 16 different residues in every iteration.  Entries whose class has more members than there are
// positions of its residue fill the holes that smaller classes leave, in ascending order.  Measured (tools/micro/tile_force.cu,
// 4 M atoms): shared-memory wavefronts per launch -38 %, force kernel 0.66 -> 0.58 ms.
//   row: the thread's T4 * 4 entries in shared memory; hist: entries per residue class, 16 x 4 bits, counted in a first pass over
//   the finished row (counting them while the row was built cost a dozen instructions in the accept path, which runs with 5 of 32
//   lanes; here 27 are busy: build 3.4 -> see DESIGN.md section 6).  A class of 16 or more members -- not seen at liquid density --
//   leaves the row in builder order (-> false).  The HOLES -- position
//   16 n + offset of a class with fewer than n + 1 members -- are chained into a list through the row itself; then one pass over
//   the row's words (its own, just written: L2 hits) places every entry; one whose class has run out of positions takes the next hole.
__device__ __forceinline__ bool pb_tile_reorder_row(const unsigned long long *__restrict__ in_words, int nn, int rot, unsigned short *row) {
    unsigned long long hist = 0ull;
    unsigned full = 0u;                  // set once a class with 15 members gets another one
    for(int q = 0; q * 4 < nn; q++) {
        const unsigned long long w = __ldcg(in_words + (size_t) q * 32);
#pragma unroll
        for(int u = 0; u < 4; u++) {
            if(q * 4 + u < nn) {
                const int sh = (int) ((unsigned) (w >> (16 * u)) & 15u) * 4;
                full |= (((unsigned) (hist >> sh) & 15u) == 15u) ? 1u : 0u;
                hist += 1ull << sh;
            }
        }
    }
    if(full != 0u) { return false; }
    int head = 0;
#pragma unroll 1
    for(int r = 0; r < 16; r++) {
        const int c = (int) (hist >> (r * 4)) & 15;
        for(int k = 16 * c + ((r - rot) & 15); k < nn; k += 16) { row[k] = (unsigned short) head; head = k; }
    }
    unsigned long long seen = 0ull;
    for(int q = 0; q * 4 < nn; q++) {
        const unsigned long long w = __ldcg(in_words + (size_t) q * 32);
#pragma unroll
        for(int u = 0; u < 4; u++) {
            if(q * 4 + u < nn) {
                const unsigned e16 = (unsigned) (w >> (16 * u)) & 0xffffu;
                const int r = (int) (e16 & 15u), sh = r * 4;
                const int n = (int) (seen >> sh) & 15;
                seen += 1ull << sh;
                int k = 16 * n + ((r - rot) & 15);
                if(k >= nn) { k = head; head = row[k]; }      // no position of this residue left: the next hole
                row[k] = (unsigned short) e16;
            }
        }
    }
    for(int k = nn; (k & 3) != 0; k++) { row[k] = (unsigned short) PB_TILE_DUMMY; }
    return true;
}


// ---- rows sorted by length ------------------------------------------------------------------------------------------------------
// The force kernel's warp runs as long as its longest row (rows of 60 ... 92 entries, 8 per iteration: a warp of 32 arbitrary rows
// makes 11 iterations for a mean of 9.9 useful ones).  So the build hands the rows of a tile out in ascending order of their
// iteration count: thread t of the force kernel works on the row of core thread rowsrc[row_base + t].  Stable counting sort over
// the tile's core threads (key = iterations, 15 = no list), deterministic: ties keep the thread order.
// s_cnt: 128 ints of shared memory.  Returns the rank (threads outside the core: their own index).
__device__ __forceinline__ int pb_tile_rank(int key, bool in_core, int *s_cnt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if(threadIdx.x < 128) { s_cnt[threadIdx.x] = 0; }
    __syncthreads();
    const unsigned peers = __match_any_sync(0xffffffffu, in_core ? key : 99);
    const int intra = __popc(peers & ((1u << lane) - 1u));
    if(in_core && intra == 0) { s_cnt[key * 8 + warp] = __popc(peers); }      // (key-major: all warps of key 0, then key 1, ...)
    __syncthreads();
    if(warp == 0) {      // exclusive prefix over the 128 counters, four per lane
        int v[4], sum = 0;
#pragma unroll
        for(int k = 0; k < 4; k++) { v[k] = s_cnt[lane * 4 + k]; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) { const int up = __shfl_up_sync(0xffffffffu, incl, o); if(lane >= o) { incl += up; } }
        int run = incl - sum;
#pragma unroll
        for(int k = 0; k < 4; k++) { s_cnt[lane * 4 + k] = run; run += v[k]; }
    }
    __syncthreads();
    return in_core ? s_cnt[key * 8 + warp] + intra : (int) threadIdx.x;
}

// End of a build: rows sorted by length and re-assembled in the conflict-aware order (both only with option "tile_reorder": the
// exact mode keeps the builder's rows where the builder's order puts them).  Every thread of the CTA calls this.
//   out: the thread's row as pass A wrote it (natural position row_base + threadIdx.x); stage: the idle staging area.
__device__ __forceinline__ void pb_tile_finish_rows(const PbTileBuildArgs &a, const PbTile &tl, bool in_core, bool active, int count,
                                                    const unsigned long long *out, unsigned char *stage, int *s_cnt) {
    if(!a.reorder) {
        if(in_core) { a.rowsrc[tl.row_base + threadIdx.x] = (unsigned char) threadIdx.x; }
        return;
    }
    const int key = active ? min((count + 7) >> 3, 14) : 15;
    const int rank = pb_tile_rank(key, in_core, s_cnt);
    if(in_core) { a.rowsrc[tl.row_base + rank] = (unsigned char) threadIdx.x; }
    const bool have = active && count > 0 && count <= a.ncap;
    // thread t assembles its row in its T4 * 8 bytes of the staging area (host: PB_TILE_M * T4 * 8 <= staging bytes)
    unsigned short *const rowbuf = reinterpret_cast<unsigned short *>(stage) + (size_t) threadIdx.x * (size_t) (a.T4 * 4);
    unsigned long long *const rw = reinterpret_cast<unsigned long long *>(rowbuf);
    if(have && !pb_tile_reorder_row(out, count, rank & 15, rowbuf)) {
        for(int q = 0; q * 4 < count; q++) { rw[q] = __ldcg(out + (size_t) q * 32); }      // (a class of 16: builder order)
    }
    __syncthreads();                                                // every row has been read: the rows may change places
    if(have) {
        unsigned long long *const dst = a.words + pb_tile_word(tl.row_base + rank, a.T4, 0);
        for(int q = 0; q * 4 < count; q++) { dst[(size_t) q * 32] = rw[q]; }
    }
}

__global__ void __launch_bounds__(PB_TILE_M) pb_k_tile_build(PbTileBuildArgs a) {
    extern __shared__ __align__(16) unsigned char pb_tile_shared[];
    __shared__ int s_cnt[128];
    PbTileHdr *h; unsigned long long *bar; double2 *sxy; double *sz; unsigned char *smeta;
    pb_tile_smem(pb_tile_shared, h, bar, sxy, sz, smeta);
    const PbTileGeom &g = a.g;
    const int nlocal = a.nlocal, ncap = a.ncap, T4 = a.T4;
    const double cutsq = a.cutsq;
    const PbTile tl = a.tiles[blockIdx.x];
    if(!pb_tile_stage(a.hdrs + blockIdx.x, h, bar, a.mxy, a.mz, sxy, sz)) {      // a tile of ghosts only: nothing to build
        if(threadIdx.x == 0) { a.tile_flag[blockIdx.x] = 0; }
        return;
    }
    // while the positions land: the meta bytes of the staged slots (particle type, 3 bits | 8 for a ghost), warp r runs r, r + 8
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for(int r = warp; r < PB_TILE_NRUN; r += nw) {
            const int len = h->run_len[r], begin = h->run_begin[r], slot0 = h->run_slot0[r];
            for(int k = lane; k < len; k += 32) { smeta[slot0 + k] = __ldg(a.mmeta + begin + k); }
        }
    }
    const int q = pb_tile_core_q(h, threadIdx.x);
    const int cs = (q >= 0) ? pb_tile_core_csr(h, threadIdx.x, q) : -1;
    const int i = (cs >= 0) ? __ldg(a.cell_list + cs) : nlocal;
    const bool live = i < nlocal;                                  // a local particle (ghosts sit in core cells at the faces, too)
    const bool active = live && (a.flags[i] & PB_FLAG_FIXED) == 0;   // FIXED particles get no list (the reference's FIXED filter)
    double4 pi = make_double4(0.0, 0.0, 0.0, 0.0);
    int flat = 0;
    if(active) { pi = pb_ld_pos(a.pos + i); flat = a.particle_cell[i] - 1; }
    pb_mbar_wait(bar, 0);
    __syncthreads();                                               // (the meta bytes)
    int count = 0, boundary = 0;
    const int row = tl.row_base + threadIdx.x;
    unsigned long long *const out = a.words + pb_tile_word(row, T4, 0);
    unsigned long long *outp = out;
    if(active) {
        boundary = (pi.x < a.faces.lo[0]) | (pi.x > a.faces.hi[0]) | (pi.y < a.faces.lo[1]) | (pi.y > a.faces.hi[1]) | (pi.z < a.faces.lo[2]) |
                   (pi.z > a.faces.hi[2]);
        const int c2 = flat % g.dim2, col = flat / g.dim2, c1 = col % g.dim1, c0 = col / g.dim1;
        const double fx = pi.x - (g.lo[0] + c0 * g.spacing), fy = pi.y - (g.lo[1] + c1 * g.spacing), zrel = pi.z - g.lo[2];
        // the particle's own slot in the staging order (its column is staged run `tr0`): what `j != i` becomes
        const int tr0 = (c0 - (tl.X0 - 1)) * 4 + (c1 - (tl.Y0 - 1));
        const int s_self = pb_tile_self_slot(h, q, cs);
        // z windows of the nine stencil rows, as slot ranges of the staging order
        int wb[9], we[9];
#pragma unroll
        for(int r = 0; r < 9; r++) {
            int b, e;
            wb[r] = 0; we[r] = 0;
            if(pb_tile_window(g, c0, c1, c2, fx, fy, zrel, cutsq, r, a.sub_start, b, e)) {
                const int tr = tr0 + (r / 3 - 1) * 4 + (r % 3 - 1);        // the staged run of this stencil row
                const int shift = h->run_slot0[tr] - h->run_begin[tr];
                wb[r] = b + shift; we[r] = e + shift;
            }
        }
        // entries enter a 64-bit word at the top and move down: after four appends the word holds them in list order
        unsigned long long w = 0ull;
        unsigned meta_or = 0u;
#pragma unroll
        for(int r = 0; r < 9; r++) {
#pragma unroll 2
            for(int s = wb[r]; s < we[r]; s++) {
                const double2 xy = sxy[s];
                const double z = sz[s];
                const double dx = __dsub_rn(pi.x, xy.x), dy = __dsub_rn(pi.y, xy.y), dz = __dsub_rn(pi.z, z);
                const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if(rsq < cutsq && (r != 4 || s != s_self)) {
                    const unsigned meta = smeta[s];
                    w = (w >> 16) | ((unsigned long long) ((unsigned) s | ((meta & 7u) << 12)) << 48);
                    count++;
                    if((count & 3) == 0 && count <= ncap) { *outp = w; outp += 32; }
                    meta_or |= meta;
                }
            }
        }
        if((count & 3) != 0 && (count >> 2) < T4) {      // the last word: entries down to the bottom, padded with the dummy slot
            w >>= 16 * (4 - (count & 3));
            w |= 0x0001000100010001ull * (unsigned long long) PB_TILE_DUMMY << (16 * (count & 3));
            out[(size_t) (count >> 2) * 32] = w;
        }
        boundary |= (int) (meta_or >> 3);
    }
    if(live) { a.numneigh[i] = count; }
    const int any_b = __syncthreads_or(boundary);                  // (also: every thread is through with the staged positions)
    if(threadIdx.x == 0) { a.tile_flag[blockIdx.x] = any_b != 0; }
    int m = count;
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) { m = max(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if((threadIdx.x & 31) == 0 && m > 0) { atomicMax(a.max_count, m); }
    pb_tile_finish_rows(a, tl, q >= 0, active, count, out, reinterpret_cast<unsigned char *>(sxy), s_cnt);
}

// The same build with an fp32 PRE-FILTER (option "tile_prefilter", the default).  The staged tile holds float4 (x, y, z rounded to
// fp32, meta byte) -- one 128-bit shared-memory load and seven fp32 instructions per candidate instead of two loads and nine fp64
// instructions.  The filter only ever decides candidates it cannot get wrong: with coordinates |x| <= A the rounded difference is
// off by at most 2 * 2^-24 * A + 2^-24 * |d|, so the fp32 squared distance of a pair closer than the cutoff is within
//     err = 2 sqrt(3) rc (2^-23 A + 2^-24 rc) + 8 * 2^-24 rc^2 + 3 (2^-23 A)^2
// of the exact one.  Candidates below cutsq - 2 err are in, candidates above cutsq + 2 err are out (a pair that is truly inside
// cannot land there), and the few in between -- about one per hundred particles -- are decided by the fp64 expression of the
// reference on the exact positions of the mirror.  A is taken per particle (its own largest coordinate plus the cutoff), so no
// assumption on the box enters.  The lists are the fp64 kernel's, entry for entry.
__global__ void __launch_bounds__(PB_TILE_M) pb_k_tile_build32(PbTileBuildArgs a) {
    extern __shared__ __align__(16) unsigned char pb_tile_shared[];
    __shared__ int s_cnt[128];
    PbTileHdr *h; unsigned long long *bar; double2 *sxy; double *sz; unsigned char *smeta;
    pb_tile_smem(pb_tile_shared, h, bar, sxy, sz, smeta);
    float4 *const s32 = reinterpret_cast<float4 *>(sxy);
    const PbTileGeom &g = a.g;
    const int nlocal = a.nlocal, ncap = a.ncap, T4 = a.T4;
    const double cutsq = a.cutsq;
    const PbTile tl = a.tiles[blockIdx.x];
    {
        const int t = threadIdx.x;
        if(t < 64) { reinterpret_cast<int *>(h)[t] = __ldg(reinterpret_cast<const int *>(a.hdrs + blockIdx.x) + t); }
        if(t == 0) {
            pb_mbar_init(bar, 1);
            s32[PB_TILE_DUMMY] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
        }
        __syncthreads();
        if(!h->any_local) {                                            // a tile of ghosts only: nothing to build
            if(t == 0) { a.tile_flag[blockIdx.x] = 0; }
            return;
        }
        if(t < PB_TILE_NRUN) {
            const int len = h->run_len[t];
            if(len > 0) { pb_bulk_g2s(s32 + h->run_slot0[t], a.m32 + h->run_begin[t], len * 16, bar); }
        }
        if(t == 0) { pb_mbar_expect_tx(bar, h->bytes32); }
    }
    const int q = pb_tile_core_q(h, threadIdx.x);
    const int cs = (q >= 0) ? pb_tile_core_csr(h, threadIdx.x, q) : -1;
    const int i = (cs >= 0) ? __ldg(a.cell_list + cs) : nlocal;
    const bool live = i < nlocal;
    const bool active = live && (a.flags[i] & PB_FLAG_FIXED) == 0;
    double4 pi = make_double4(0.0, 0.0, 0.0, 0.0);
    int flat = 0;
    if(active) { pi = pb_ld_pos(a.pos + i); flat = a.particle_cell[i] - 1; }
    pb_mbar_wait(bar, 0);
    int count = 0, boundary = 0;
    const int row = tl.row_base + threadIdx.x;
    unsigned long long *const out = a.words + pb_tile_word(row, T4, 0);
    int widx = 0;                             // index of the next word of the row (32 apart: sliced layout)
    if(active) {
        boundary = (pi.x < a.faces.lo[0]) | (pi.x > a.faces.hi[0]) | (pi.y < a.faces.lo[1]) | (pi.y > a.faces.hi[1]) | (pi.z < a.faces.lo[2]) |
                   (pi.z > a.faces.hi[2]);
        const int c2 = flat % g.dim2, col = flat / g.dim2, c1 = col % g.dim1, c0 = col / g.dim1;
        const double fx = pi.x - (g.lo[0] + c0 * g.spacing), fy = pi.y - (g.lo[1] + c1 * g.spacing), zrel = pi.z - g.lo[2];
        const int tr0 = (c0 - (tl.X0 - 1)) * 4 + (c1 - (tl.Y0 - 1));
        const int s_self = pb_tile_self_slot(h, q, cs);
        // the nine z windows go to shared memory (begin | end << 16; the float4 staging leaves the upper third of the staging area
        // free until the reorder pass), so that the loop over the stencil rows stays ROLLED: the kernel with nine unrolled copies of
        // the test loop stalled on instruction fetch (ncu: no_instruction 1.5 warps per issue cycle)
        unsigned *const swin = reinterpret_cast<unsigned *>(s32 + PB_TILE_CAP) + threadIdx.x;
#pragma unroll 1
        for(int r = 0; r < 9; r++) {
            int b, e;
            unsigned packed = 0u;
            if(pb_tile_window(g, c0, c1, c2, fx, fy, zrel, cutsq, r, a.sub_start, b, e)) {
                const int tr = tr0 + (r / 3 - 1) * 4 + (r % 3 - 1);
                const int shift = h->run_slot0[tr] - h->run_begin[tr];
                packed = (unsigned) (b + shift) | ((unsigned) (e + shift) << 16);
            }
            swin[r * PB_TILE_M] = packed;
        }
        // the band of the pre-filter for this particle (see above), bounds rounded outwards
        const double rc = sqrt(cutsq), A = fmax(fmax(fabs(pi.x), fabs(pi.y)), fabs(pi.z)) + rc;
        const double u = 5.9604644775390625e-08;                                      // 2^-24
        const double err = 2.0 * 1.7320508075688774 * rc * (2.0 * u * A + u * rc) + 8.0 * u * cutsq + 3.0 * (2.0 * u * A) * (2.0 * u * A);
        const float cut_lo = __double2float_rd(cutsq - 2.0 * err), cut_hi = __double2float_ru(cutsq + 2.0 * err);
        const float xf = __double2float_rn(pi.x), yf = __double2float_rn(pi.y), zf = __double2float_rn(pi.z);
        unsigned long long w = 0ull;
        unsigned meta_or = 0u;
#pragma unroll 1
        for(int r = 0; r < 9; r++) {
            const unsigned packed = swin[r * PB_TILE_M];
            const int s_end = (int) (packed >> 16);
            // (four candidates per trip with their loads and fp32 chains issued together, decisions afterwards, was measured SLOWER:
            // 3.73 vs 3.26 ms per build -- the window tails and the second range check cost more than the chains' latency)
#pragma unroll 4
            for(int s = (int) (packed & 0xffffu); s < s_end; s++) {
                const float4 c = s32[s];
                const float dx = xf - c.x, dy = yf - c.y, dz = zf - c.z;
                const float rsq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if(rsq < cut_hi) {
                    bool in = s != s_self;                             // (a slot belongs to one run: no need to ask for the row)
                    if(rsq > cut_lo) {                                 // too close to call in fp32: the reference's fp64 expression
                        const int tr = tr0 + (r / 3 - 1) * 4 + (r % 3 - 1);
                        const int kk = s - (h->run_slot0[tr] - h->run_begin[tr]);
                        const double2 xy = __ldg(a.mxy + kk);
                        const double z = __ldg(a.mz + kk);
                        const double ex = __dsub_rn(pi.x, xy.x), ey = __dsub_rn(pi.y, xy.y), ez = __dsub_rn(pi.z, z);
                        in = in && __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez)) < cutsq;
                    }
                    if(in) {
                        const unsigned meta = (unsigned) __float_as_int(c.w);
                        w = (w >> 16) | ((unsigned long long) ((unsigned) s | (meta & 0x7000u)) << 48);
                        count++;
                        if((count & 3) == 0 && count <= ncap) { out[widx] = w; widx += 32; }
                        meta_or |= meta;
                    }
                }
            }
        }
        if((count & 3) != 0 && (count >> 2) < T4) {
            w >>= 16 * (4 - (count & 3));
            w |= 0x0001000100010001ull * (unsigned long long) PB_TILE_DUMMY << (16 * (count & 3));
            out[(size_t) (count >> 2) * 32] = w;
        }
        boundary |= (int) (meta_or >> 15);
    }
    if(live) { a.numneigh[i] = count; }
    const int any_b = __syncthreads_or(boundary);                  // (also: every thread is through with the staged positions)
    if(threadIdx.x == 0) { a.tile_flag[blockIdx.x] = any_b != 0; }
    int m = count;
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) { m = max(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if((threadIdx.x & 31) == 0 && m > 0) { atomicMax(a.max_count, m); }
    pb_tile_finish_rows(a, tl, q >= 0, active, count, out, reinterpret_cast<unsigned char *>(sxy), s_cnt);
}

// ---- force ------------------------------------------------------------------------------------------------------------------
struct PbTileLjArgs {
    int nlocal, ncap, T4, cap, ntypes, nsel;
    double cutsq, eps_u, sig6_u, c1_u, c2_u, dt, half_dt;      // c1 = 48 eps sigma6^2, c2 = 24 eps sigma6 (md_math.h pb_lj_fpair_fast)
    const double *eps_t, *sig6_t;
    PbTileGeom g;
    const PbTile *tiles;
    const PbTileHdr *hdrs;
    const int *sel;                 // tile subset of this launch (interior / boundary split), null = all tiles
    const double2 *mxy;             // the mirror: positions in CSR order (this step's), and where the next step's go
    const double *mz;
    double2 *mxy_next;
    double *mz_next;
    const int *flags, *type, *cell_list, *numneigh;
    const unsigned long long *words;
    const unsigned char *rowsrc;      // list row -> core thread of its tile (rows are sorted by length, pb_tile_finish_rows)
    double *force;
    const double *mass;
    double *vel;
    double4 *pos_next;
};

// FUSE / ACCUMULATE: the epilogue of the per-particle kernel (md_kernels.cu pb_k_lennard_jones), operation for operation.
// FMA: the production arithmetic of md_math.h (pb_pair_rsq_fma / pb_lj_fpair_fast, fused accumulation) instead of the reference's
// expression tree.  Either way the pair loop is BRANCH-FREE: the term of a pair outside the cutoff (or of a padding entry) is
// selected to +0 and added like any other -- x + (+-0) = x bit for bit (x = -0 cannot occur: the sums start at +0 and no term
// rounds to zero), so the results are those of the branching loop, but the four pair chains of an iteration are straight-line
// code that the scheduler interleaves (with one divergent region per pair the dependent fp64 chains ran one after the other:
// ncu stall "wait" 3.5 of 11.9 cycles per instruction, 50 instructions per pair; now 35).
template<bool UNIFORM, bool ACCUMULATE, int FUSE, bool FMA>
__global__ void __launch_bounds__(PB_TILE_M, 4) pb_k_tile_lj(PbTileLjArgs a) {
    extern __shared__ __align__(16) unsigned char pb_tile_shared[];
    __shared__ double s_c1[64], s_c2[64];      // per type pair: (48 eps sigma6^2, 24 eps sigma6) with FMA, (sigma6, eps) without
    PbTileHdr *h; unsigned long long *bar; double2 *sxy; double *sz; unsigned char *smeta;
    pb_tile_smem(pb_tile_shared, h, bar, sxy, sz, smeta);
    if(!UNIFORM) {
        for(int k = threadIdx.x; k < a.ntypes * a.ntypes; k += blockDim.x) {
            const double e = a.eps_t[k], s6 = a.sig6_t[k];
            s_c1[k] = FMA ? 48.0 * e * s6 * s6 : s6;
            s_c2[k] = FMA ? 24.0 * e * s6 : e;
        }
    }
    const int tile_id = (a.sel != nullptr) ? __ldg(a.sel + blockIdx.x) : (int) blockIdx.x;
    const int row = __ldg(&a.tiles[tile_id].row_base) + threadIdx.x;
    const int src = (int) __ldg(a.rowsrc + row);      // whose row this thread works on (read before the header is known: the array is padded)
    if(!pb_tile_stage(a.hdrs + tile_id, h, bar, a.mxy, a.mz, sxy, sz)) { return; }      // a tile of ghosts only
    // own data: in flight while the copies land
    const int core_t = ((int) threadIdx.x < h->ncore) ? src : (int) threadIdx.x;
    const int cq = pb_tile_core_q(h, core_t);
    const int cs = (cq >= 0) ? pb_tile_core_csr(h, core_t, cq) : -1;
    const int i = (cs >= 0) ? __ldg(a.cell_list + cs) : a.nlocal;
    const bool live = i < a.nlocal;
    const bool fixed = live && (a.flags[i] & PB_FLAG_FIXED) != 0;
    int nn = 0, type_i = 0;
    const unsigned long long *wp = a.words + pb_tile_word(row, a.T4, 0);
    constexpr int U = FMA ? 8 : 4, W = U / 4;      // pairs per iteration (the exact expression tree needs more registers per pair)
    unsigned long long wnext[W];
#pragma unroll
    for(int w_ = 0; w_ < W; w_++) { wnext[w_] = 0x0001000100010001ull * (unsigned long long) PB_TILE_DUMMY; }
    const int cap = a.cap;
    double m = 1.0, vx = 0.0, vy = 0.0, vz = 0.0;
    if(live) {
        if(!UNIFORM || (FUSE & 2)) { type_i = a.type[i]; }
        if(!fixed) {
            nn = min(a.numneigh[i], a.ncap);
#pragma unroll
            for(int w_ = 0; w_ < W; w_++) { if(w_ * 4 < nn) { wnext[w_] = __ldg(wp + (size_t) w_ * 32); } }
            if(FUSE != 0) {      // the epilogue's operands
                m = a.mass[i];
                vx = a.vel[i]; vy = a.vel[cap + i]; vz = a.vel[2 * (size_t) cap + i];
            }
        }
    }
    pb_mbar_wait(bar, 0);
    if(!live) { return; }
    // the particle's own position comes out of the staged tile as well
    const int s_self = pb_tile_self_slot(h, cq, cs);
    const double2 pxy = sxy[s_self];
    double4 pi = make_double4(pxy.x, pxy.y, sz[s_self], pb_type_w(type_i));
    const int ti = UNIFORM ? 0 : type_i * a.ntypes;
    const double k1 = FMA ? a.c1_u : a.sig6_u, k2 = FMA ? a.c2_u : a.eps_u;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for(int k = 0; k < nn; k += U) {
        // the list words of the iteration after the next: into L1 now.  (The loads of the NEXT iteration's words are issued
        // below, but with 62 of 64 registers live the compiler sinks them to the end of the loop body, ~20 instructions before
        // their use -- ncu r2l: 26 % of the stall samples on that use.  A prefetch has no destination register to economise.
        // It lands in L2 rather than L1 (ncu r2r: 3 % L1 hit rate of the global loads, the use still draws 14 % of the samples);
        // a per-thread cp.async ring in shared memory, one iteration ahead, was measured SLOWER: 0.635 vs 0.572 ms.)
        if(k + 2 * U < nn) {
#pragma unroll
            for(int q = 0; q < W; q++) { asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + (size_t) (((k + 2 * U) >> 2) + q) * 32)); }
        }
        unsigned long long w[W];
#pragma unroll
        for(int q = 0; q < W; q++) {      // a word past the end of the row is replaced by padding entries
            w[q] = wnext[q];
            wnext[q] = 0x0001000100010001ull * (unsigned long long) PB_TILE_DUMMY;
            if(k + U + q * 4 < nn) { wnext[q] = __ldg(wp + (size_t) (((k + U) >> 2) + q) * 32); }
        }
        double xj[U], yj[U], zj[U];
        int tj[U];
#pragma unroll
        for(int u = 0; u < U; u++) {
            const unsigned e16 = (unsigned) (w[u >> 2] >> (16 * (u & 3))) & 0xffffu;
            const int s = (int) (e16 & 0xfffu);
            tj[u] = (int) (e16 >> 12);
            const double2 xy = sxy[s];
            xj[u] = xy.x; yj[u] = xy.y;
            zj[u] = sz[s];
        }
#pragma unroll
        for(int u = 0; u < U; u++) {
            double dx, dy, dz;
            const double p1 = UNIFORM ? k1 : s_c1[ti + tj[u]], p2 = UNIFORM ? k2 : s_c2[ti + tj[u]];
            if(FMA) {
                const double rsq = pb_pair_rsq_fma(pi.x, pi.y, pi.z, xj[u], yj[u], zj[u], &dx, &dy, &dz);
                double f = pb_lj_fpair_fast(rsq, p1, p2);
                f = pb_less_bits(rsq, a.cutsq) ? f : 0.0;
                fx = __fma_rn(dx, f, fx);
                fy = __fma_rn(dy, f, fy);
                fz = __fma_rn(dz, f, fz);
            } else {
                const double rsq = pb_pair_rsq(pi.x, pi.y, pi.z, xj[u], yj[u], zj[u], &dx, &dy, &dz);
                const bool in = rsq < a.cutsq;
                double f = pb_lj_fpair(in ? rsq : 1.0, p1, p2);
                f = in ? f : 0.0;
                fx = __dadd_rn(fx, __dmul_rn(dx, f));
                fy = __dadd_rn(fy, __dmul_rn(dy, f));
                fz = __dadd_rn(fz, __dmul_rn(dz, f));
            }
        }
    }
    // force[i] = force[i] + acc (sim/interaction.py:280-292); a pending reset_volatile_properties is folded in (ACCUMULATE == false)
    if(ACCUMULATE) {
        if(!fixed) {
            fx = __dadd_rn(a.force[i], fx);
            fy = __dadd_rn(a.force[cap + i], fy);
            fz = __dadd_rn(a.force[2 * (size_t) cap + i], fz);
            a.force[i] = fx;
            a.force[cap + i] = fy;
            a.force[2 * (size_t) cap + i] = fz;
        }
    } else {
        fx = __dadd_rn(0.0, fx);
        fy = __dadd_rn(0.0, fy);
        fz = __dadd_rn(0.0, fz);
        a.force[i] = fx;
        a.force[cap + i] = fy;
        a.force[2 * (size_t) cap + i] = fz;
    }
    if(FUSE != 0) {
        if(!fixed) {
            if(FUSE & 1) {
                vx = __dadd_rn(vx, __ddiv_rn(__dmul_rn(a.half_dt, fx), m));
                vy = __dadd_rn(vy, __ddiv_rn(__dmul_rn(a.half_dt, fy), m));
                vz = __dadd_rn(vz, __ddiv_rn(__dmul_rn(a.half_dt, fz), m));
            }
            if(FUSE & 2) {
                vx = __dadd_rn(vx, __ddiv_rn(__dmul_rn(a.half_dt, fx), m));
                vy = __dadd_rn(vy, __ddiv_rn(__dmul_rn(a.half_dt, fy), m));
                vz = __dadd_rn(vz, __ddiv_rn(__dmul_rn(a.half_dt, fz), m));
                pi.x = __dadd_rn(pi.x, __dmul_rn(a.dt, vx));
                pi.y = __dadd_rn(pi.y, __dmul_rn(a.dt, vy));
                pi.z = __dadd_rn(pi.z, __dmul_rn(a.dt, vz));
            }
            a.vel[i] = vx;
            a.vel[cap + i] = vy;
            a.vel[2 * (size_t) cap + i] = vz;
        }
        if(FUSE & 2) {      // the next step's positions: the particle array and the mirror (same CSR position)
            a.pos_next[i] = pi;
            a.mxy_next[cs] = make_double2(pi.x, pi.y);
            a.mz_next[cs] = pi.z;
        }
    }
}

// ---- per-particle view of the lists (pb_download_neighbors: tests, tools) -------------------------------------------------------
template<bool ELL>
__global__ void __launch_bounds__(PB_TILE_M) pb_k_tile_export(int nlocal, int ncap, int T4, int cap_out, const PbTile *__restrict__ tiles,
                                                             const PbTileHdr *__restrict__ hdrs, const int *__restrict__ cell_list,
                                                             const unsigned long long *__restrict__ words, const int *__restrict__ numneigh,
                                                             const unsigned char *__restrict__ rowsrc, int *__restrict__ out) {
    __shared__ PbTileHdr hdr;
    PbTileHdr *h = &hdr;
    const PbTile tl = tiles[blockIdx.x];
    if(threadIdx.x < 64) { reinterpret_cast<int *>(h)[threadIdx.x] = reinterpret_cast<const int *>(hdrs + blockIdx.x)[threadIdx.x]; }
    __syncthreads();
    if(!h->any_local) { return; }                       // (no rows were written for a tile of ghosts only)
    const int core_t = ((int) threadIdx.x < h->ncore) ? (int) rowsrc[tl.row_base + threadIdx.x] : (int) threadIdx.x;
    const int cq = pb_tile_core_q(h, core_t);
    const int cs = (cq >= 0) ? pb_tile_core_csr(h, core_t, cq) : -1;
    const int i = (cs >= 0) ? cell_list[cs] : nlocal;
    if(i >= nlocal) { return; }
    const int row = tl.row_base + threadIdx.x;
    const int c = min(numneigh[i], min(cap_out, ncap));
    for(int k = 0; k < c; k++) {
        const unsigned long long w = words[pb_tile_word(row, T4, k >> 2)];
        const int s = (int) ((w >> (16 * (k & 3))) & 0xfffull);
        int r = 0;
        for(int q = 1; q < PB_TILE_NRUN; q++) { if(s >= h->run_slot0[q]) { r = q; } }      // slot0 is non-decreasing
        const int j = cell_list[h->run_begin[r] + (s - h->run_slot0[r])];
        if(ELL) { out[((size_t) (i >> 5) * cap_out + (size_t) k) * 32 + (size_t) (i & 31)] = j; }      // PbNeighLayout, one lane per particle
        else { out[(size_t) i * cap_out + k] = j; }
    }
    if(!ELL) { for(int k = c; k < cap_out; k++) { out[(size_t) i * cap_out + k] = -1; } }
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
template<typename T>
static int pb_tile_fit(pb_ctx *ctx, T **p, int *cap, size_t need) {
    if(need > (size_t) *cap) {
        if(*p != nullptr) { PB_CHECK(cudaFree(*p)); *p = nullptr; *cap = 0; }
        const size_t want = need + need / 4 + 64;
        PB_CHECK(cudaMalloc(p, sizeof(T) * want));
        *cap = (int) want;
    }
    return 0;
}

static int pb_tile_plan(pb_ctx *ctx, bool *overflow) {
    const PbTileGeom g = pb_tile_geom(ctx);
    const int nsx = (g.dim0 + 1) / 2, nsy = (g.dim1 + 1) / 2, nsc = nsx * nsy;
    const size_t nlvl = (size_t) nsc * g.dim2;
    PB_TRY(pb_tile_fit(ctx, &ctx->tile_lvl, &ctx->tile_lvl_cap, 2 * nlvl));
    if(nsc + 2 > ctx->tile_sc_cap) {
        int c1 = ctx->tile_sc_cap, c2 = ctx->tile_sc_cap;
        PB_TRY(pb_tile_fit(ctx, &ctx->tile_cnt, &c1, (size_t) nsc + 2));
        PB_TRY(pb_tile_fit(ctx, &ctx->tile_off, &c2, (size_t) nsc + 2));
        ctx->tile_sc_cap = std::min(c1, c2);
    }
    int *lvl_core = ctx->tile_lvl, *lvl_halo = ctx->tile_lvl + nlvl;
    int *d_over = ctx->d_scalars + 2;
    PB_CHECK(cudaMemsetAsync(d_over, 0, sizeof(int), ctx->stream));
    PB_LAUNCH(pb_k_tile_levels, pb_blocks((long) nlvl, 256), 256, nsx, nsy, g.dim0, g.dim1, g.dim2, ctx->cell_start, lvl_core, lvl_halo);
    PB_LAUNCH(pb_k_tile_plan<false>, pb_blocks(nsc, 128), 128, nsc, nsy, g.dim2, lvl_core, lvl_halo, ctx->tile_cnt, ctx->tile_off, (PbTile *) nullptr,
              (int *) nullptr, d_over);
    PB_TRY(pb_exclusive_scan(ctx, ctx->tile_cnt, ctx->tile_off, nsc));
    PB_CHECK(cudaMemcpyAsync(ctx->h_scalars + 2, ctx->tile_off + nsc, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaMemcpyAsync(ctx->h_scalars + 3, d_over, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    const int ntiles = ctx->h_scalars[2];
    *overflow = ctx->h_scalars[3] != 0;
    ctx->ntiles = 0;
    if(*overflow || ntiles == 0) { return 0; }
    if(ntiles + 1 > ctx->tiles_cap) {
        int c0 = ctx->tiles_cap, c1 = ctx->tiles_cap, c2 = ctx->tiles_cap;
        PB_TRY(pb_tile_fit(ctx, &ctx->tiles, &c0, (size_t) ntiles + 1));
        PB_TRY(pb_tile_fit(ctx, &ctx->tile_pad, &c1, (size_t) ntiles + 1));
        PB_TRY(pb_tile_fit(ctx, &ctx->tile_row, &c2, (size_t) ntiles + 1));
        ctx->tiles_cap = std::min(c0, std::min(c1, c2));
    }
    PB_LAUNCH(pb_k_tile_plan<true>, pb_blocks(nsc, 128), 128, nsc, nsy, g.dim2, lvl_core, lvl_halo, ctx->tile_cnt, ctx->tile_off, ctx->tiles,
              ctx->tile_pad, d_over);
    PB_TRY(pb_exclusive_scan(ctx, ctx->tile_pad, ctx->tile_row, ntiles));
    PB_LAUNCH(pb_k_tile_rows, pb_blocks(ntiles, 256), 256, ntiles, ctx->tile_row, ctx->tiles);
    PB_CHECK(cudaMemcpyAsync(ctx->h_scalars + 2, ctx->tile_row + ntiles, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->ntiles = ntiles;
    ctx->tile_rows = ctx->h_scalars[2];
    PB_TRY(pb_tile_fit(ctx, &ctx->tile_hdrs, &ctx->tile_hdrs_cap, (size_t) ntiles + 1));
    PB_LAUNCH(pb_k_tile_headers, pb_blocks((long) ntiles * 32, 128), 128, ntiles, ctx->nlocal, g, ctx->tiles, ctx->cell_start, ctx->cell_list, ctx->tile_hdrs);
    return 0;
}

// ---- the mirror (positions in CSR order) ---------------------------------------------------------------------------------------
// everything: after a cell-list build, and before a force evaluation whenever the mirror is not known to be current
int pb_tile_mirror_all(pb_ctx *ctx, bool with_f32) {
    const int nall = ctx->nlocal + ctx->nghost;
    if(nall + 2 > ctx->mirror_cap) {
        for(int b = 0; b < 2; b++) {
            if(ctx->mxy[b] != nullptr) { PB_CHECK(cudaFree(ctx->mxy[b])); ctx->mxy[b] = nullptr; }
            if(ctx->mz[b] != nullptr) { PB_CHECK(cudaFree(ctx->mz[b])); ctx->mz[b] = nullptr; }
        }
        if(ctx->mmeta != nullptr) { PB_CHECK(cudaFree(ctx->mmeta)); ctx->mmeta = nullptr; }
        if(ctx->m32 != nullptr) { PB_CHECK(cudaFree(ctx->m32)); ctx->m32 = nullptr; }
        if(ctx->ghost_csr != nullptr) { PB_CHECK(cudaFree(ctx->ghost_csr)); ctx->ghost_csr = nullptr; }
        ctx->mirror_cap = 0;
        const size_t want = (size_t) nall + nall / 4 + 1024;
        for(int b = 0; b < 2; b++) {
            PB_CHECK(cudaMalloc(&ctx->mxy[b], sizeof(double2) * want));
            PB_CHECK(cudaMalloc(&ctx->mz[b], sizeof(double) * want));
            PB_CHECK(cudaMemsetAsync(ctx->mxy[b], 0, sizeof(double2) * want, ctx->stream));      // (the aligned z copies read up to two entries past a run)
            PB_CHECK(cudaMemsetAsync(ctx->mz[b], 0, sizeof(double) * want, ctx->stream));
        }
        PB_CHECK(cudaMalloc(&ctx->mmeta, want));
        PB_CHECK(cudaMalloc(&ctx->m32, sizeof(float4) * want));
        PB_CHECK(cudaMalloc(&ctx->ghost_csr, sizeof(int) * want));
        ctx->mirror_cap = (int) want;
    }
    if(nall > 0) {
        PB_LAUNCH(pb_k_tile_mirror, pb_blocks(nall, 256), 256, nall, ctx->nlocal, ctx->cell_list, ctx->pos, ctx->mxy[0], ctx->mz[0], ctx->mmeta, ctx->ghost_csr,
                  with_f32 ? ctx->m32 : nullptr);
    }
    ctx->mirror_cur = 0;
    ctx->mirror_n = nall;
    ctx->mirror_fresh = true;
    return 0;
}

// the ghosts only: after the per-step refresh of their positions (pb_synchronize), inside pb_md_run
int pb_tile_mirror_ghosts(pb_ctx *ctx) {
    if(ctx->nghost > 0) {
        PB_LAUNCH(pb_k_tile_mirror_ghosts, pb_blocks(ctx->nghost, 256), 256, ctx->nghost, ctx->nlocal, ctx->ghost_csr, ctx->pos, ctx->mxy[ctx->mirror_cur],
                  ctx->mz[ctx->mirror_cur]);
    }
    return 0;
}

__global__ void __launch_bounds__(256) pb_k_tile_split(int ntiles, const int *__restrict__ flag, const int *__restrict__ scan,
                                                       int *__restrict__ interior, int *__restrict__ boundary) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= ntiles) { return; }
    if(flag[t]) { boundary[scan[t]] = t; } else { interior[t - scan[t]] = t; }
}

// 0: built; 1: not applicable (the caller builds per-particle lists); < 0: error
int pb_build_tile_lists(pb_ctx *ctx, double cutoff) {
    ctx->tiles_n = -1;
    ctx->tile_split_valid = false;
    const int n = ctx->nlocal;
    if(!ctx->tile_lists || ctx->half_lists || ctx->lanes != 1 || ctx->dem || ctx->stage_lists) { return 1; }
    if(ctx->ntypes > 8) { return 1; }
    if(n == 0) { return 1; }
    // INFINITE particles live in cell 0, outside every tile: leave such systems to the per-particle builder
    PB_CHECK(cudaMemcpyAsync(ctx->h_scalars + 4, ctx->cell_start, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    bool overflow = false;
    PB_TRY(pb_tile_plan(ctx, &overflow));           // (synchronises the stream: the two cell_start values have arrived)
    if(ctx->h_scalars[5] - ctx->h_scalars[4] > 0) { return 1; }
    if(overflow || ctx->ntiles == 0) { return 1; }
    const PbTileGeom g = pb_tile_geom(ctx);
    const double cutsq = cutoff * cutoff;
    if(ctx->ncap <= 0) { ctx->ncap = 100; }
    PbTileFaces faces;
    for(int d = 0; d < 3; d++) {
        faces.lo[d] = ctx->subdom[d * 2] + ctx->spacing;
        faces.hi[d] = ctx->subdom[d * 2 + 1] - ctx->spacing;
    }
    if(ctx->ntiles + 1 > ctx->tile_flag_cap) {
        int c[4] = {ctx->tile_flag_cap, ctx->tile_flag_cap, ctx->tile_flag_cap, ctx->tile_flag_cap};
        PB_TRY(pb_tile_fit(ctx, &ctx->tile_flag, &c[0], (size_t) ctx->ntiles + 1));
        PB_TRY(pb_tile_fit(ctx, &ctx->tile_scan, &c[1], (size_t) ctx->ntiles + 1));
        PB_TRY(pb_tile_fit(ctx, &ctx->tiles_interior, &c[2], (size_t) ctx->ntiles + 1));
        PB_TRY(pb_tile_fit(ctx, &ctx->tiles_boundary, &c[3], (size_t) ctx->ntiles + 1));
        ctx->tile_flag_cap = std::min(std::min(c[0], c[1]), std::min(c[2], c[3]));
    }
    const size_t smem = pb_tile_smem_bytes(true);
    PB_CHECK(cudaFuncSetAttribute(pb_k_tile_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    PB_CHECK(cudaFuncSetAttribute(pb_k_tile_build32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    PB_TRY(pb_tile_mirror_all(ctx, ctx->tile_prefilter));                // the tiles are staged out of the mirror
    PB_TRY(pb_tile_fit(ctx, &ctx->tile_rowsrc, &ctx->tile_rowsrc_cap, (size_t) ctx->tile_rows + PB_TILE_M));      // (padded: read before the tile's row count is known)
    for(int attempt = 0; attempt < 8; attempt++) {
        const int T4 = (ctx->ncap + 3) / 4;
        const size_t bytes = sizeof(unsigned long long) * (size_t) (ctx->tile_rows / 32) * (size_t) T4 * 32;
        if(bytes > ctx->twords_bytes) {
            if(ctx->twords != nullptr) { PB_CHECK(cudaFree(ctx->twords)); ctx->twords = nullptr; ctx->twords_bytes = 0; }
            const size_t want = bytes + bytes / 8;
            PB_CHECK(cudaMalloc(&ctx->twords, want));
            ctx->twords_bytes = want;
        }
        ctx->tile_T4 = T4;
        PB_CHECK(cudaMemsetAsync(ctx->d_scalars, 0, sizeof(int), ctx->stream));
        // (the particle type, 3 bits, always rides in the entries: the Lennard-Jones tables may be set after the lists are built)
        PbTileBuildArgs ba;
        ba.nlocal = n; ba.ncap = ctx->ncap; ba.T4 = T4; ba.g = g; ba.cutsq = cutsq; ba.tiles = ctx->tiles; ba.pos = ctx->pos; ba.flags = ctx->flags;
        ba.hdrs = ctx->tile_hdrs; ba.mxy = ctx->mxy[ctx->mirror_cur]; ba.mz = ctx->mz[ctx->mirror_cur]; ba.mmeta = ctx->mmeta;
        ba.particle_cell = ctx->particle_cell; ba.sub_start = ctx->sub_start; ba.cell_list = ctx->cell_list;
        ba.words = ctx->twords; ba.rowsrc = ctx->tile_rowsrc; ba.numneigh = ctx->numneigh; ba.max_count = ctx->d_scalars; ba.tile_flag = ctx->tile_flag; ba.faces = faces;
        // the reorder pass assembles the rows in the staging area (positions + meta bytes): possible while a row fits its share
        ba.reorder = ctx->tile_reorder && (size_t) PB_TILE_M * (size_t) T4 * 8 <= (size_t) PB_TILE_CAP * 25;
        ba.m32 = ctx->m32;
        if(ctx->tile_prefilter) {
            pb_k_tile_build32<<<ctx->ntiles, PB_TILE_M, smem, ctx->stream>>>(ba);
        } else {
            pb_k_tile_build<<<ctx->ntiles, PB_TILE_M, smem, ctx->stream>>>(ba);
        }
        ctx->launches++;
        PB_CHECK(cudaGetLastError());
        const bool split = ctx->world > 1 && ctx->overlap_comm;
        if(split) {       // issued before the read-back so that one synchronisation serves both counters
            PB_TRY(pb_exclusive_scan(ctx, ctx->tile_flag, ctx->tile_scan, ctx->ntiles));
            PB_LAUNCH(pb_k_tile_split, pb_blocks(ctx->ntiles, 256), 256, ctx->ntiles, ctx->tile_flag, ctx->tile_scan, ctx->tiles_interior, ctx->tiles_boundary);
            PB_CHECK(cudaMemcpyAsync(ctx->h_scalars + 1, ctx->tile_scan + ctx->ntiles, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        }
        PB_CHECK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CHECK(cudaStreamSynchronize(ctx->stream));
        ctx->max_neigh = ctx->h_scalars[0];
        if(ctx->max_neigh <= ctx->ncap) {
            if(split) {
                ctx->n_tiles_boundary = ctx->h_scalars[1];
                ctx->n_tiles_interior = ctx->ntiles - ctx->n_tiles_boundary;
                ctx->tile_split_valid = true;
            }
            ctx->tiles_n = n;
            return 0;
        }
        // capacity-overflow protocol (transformations/modules.py:159-203): grow to twice the need and re-run the module
        ctx->ncap = ctx->max_neigh * 2;
    }
    ctx->set_error("pb_build_neighbor_lists: capacity did not converge");
    return -1;
}

template<bool UNIFORM, bool ACCUMULATE, bool FMA>
static int pb_tile_launch(pb_ctx *ctx, const PbTileLjArgs &a, int grid, int fuse) {
    const size_t smem = pb_tile_smem_bytes(false);
#define PB_TILE_GO(F)                                                                                                          \
    {                                                                                                                          \
        PB_CHECK(cudaFuncSetAttribute(pb_k_tile_lj<UNIFORM, ACCUMULATE, F, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
        pb_k_tile_lj<UNIFORM, ACCUMULATE, F, FMA><<<grid, PB_TILE_M, smem, ctx->stream>>>(a);                                     \
    }
    switch(fuse) {
        case 1: PB_TILE_GO(1) break;
        case 2: PB_TILE_GO(2) break;
        case 3: PB_TILE_GO(3) break;
        default: PB_TILE_GO(0) break;
    }
#undef PB_TILE_GO
    ctx->launches++;
    PB_CHECK(cudaGetLastError());
    return 0;
}

// the force evaluation of pb_lennard_jones_fused (md_kernels.cu) over the tile lists; same contract: `fuse` bits, force reset folded
// in when pending, new positions into pos_alt when bit 1 is set (the caller swaps), part 0 = all / 1 = interior / 2 = boundary tiles
int pb_tile_lennard_jones(pb_ctx *ctx, double cutsq, double dt, int fuse, int part) {
    PbTileLjArgs a;
    a.nlocal = ctx->nlocal; a.ncap = ctx->ncap; a.T4 = ctx->tile_T4; a.cap = ctx->pcap; a.ntypes = ctx->ntypes;
    a.cutsq = cutsq; a.eps_u = ctx->h_eps[0]; a.sig6_u = ctx->h_sig6[0]; a.dt = dt; a.half_dt = dt * 0.5;
    a.c1_u = 48.0 * a.eps_u * a.sig6_u * a.sig6_u; a.c2_u = 24.0 * a.eps_u * a.sig6_u;
    a.eps_t = ctx->d_eps; a.sig6_t = ctx->d_sig6;
    a.g = pb_tile_geom(ctx);
    a.tiles = ctx->tiles; a.sel = nullptr; a.nsel = ctx->ntiles;
    int grid = ctx->ntiles;
    if(part != 0) {
        if(!ctx->tile_split_valid) { ctx->set_error("interior/boundary split not available"); return -1; }
        a.sel = (part == 1) ? ctx->tiles_interior : ctx->tiles_boundary;
        grid = (part == 1) ? ctx->n_tiles_interior : ctx->n_tiles_boundary;
        a.nsel = grid;
    }
    if(grid == 0) { return 0; }
    // the mirror is known to be current only inside pb_md_run (which keeps it so); any other caller gets it rebuilt here
    if(!(ctx->mirror_scope && ctx->mirror_fresh && ctx->mirror_n == ctx->nlocal + ctx->nghost)) { PB_TRY(pb_tile_mirror_all(ctx)); }
    a.hdrs = ctx->tile_hdrs;
    a.mxy = ctx->mxy[ctx->mirror_cur]; a.mz = ctx->mz[ctx->mirror_cur];
    a.mxy_next = ctx->mxy[ctx->mirror_cur ^ 1]; a.mz_next = ctx->mz[ctx->mirror_cur ^ 1];
    a.flags = ctx->flags; a.type = ctx->type; a.cell_list = ctx->cell_list; a.numneigh = ctx->numneigh;
    a.words = ctx->twords; a.rowsrc = ctx->tile_rowsrc; a.force = ctx->force; a.mass = ctx->mass; a.vel = ctx->vel; a.pos_next = ctx->pos_alt;
    const bool acc = !ctx->force_is_zero;
    const bool uni = ctx->lj_uniform;
    if(ctx->lj_fma) {
        if(uni) { return acc ? pb_tile_launch<true, true, true>(ctx, a, grid, fuse) : pb_tile_launch<true, false, true>(ctx, a, grid, fuse); }
        return acc ? pb_tile_launch<false, true, true>(ctx, a, grid, fuse) : pb_tile_launch<false, false, true>(ctx, a, grid, fuse);
    }
    if(uni) { return acc ? pb_tile_launch<true, true, false>(ctx, a, grid, fuse) : pb_tile_launch<true, false, false>(ctx, a, grid, fuse); }
    return acc ? pb_tile_launch<false, true, false>(ctx, a, grid, fuse) : pb_tile_launch<false, false, false>(ctx, a, grid, fuse);
}

int pb_tile_download_neighbors(pb_ctx *ctx, int *out, int capacity) {
    const int n = ctx->tiles_n;
    if(n <= 0) { return 0; }
    PbScratch stage_buf;
    PB_CHECK(stage_buf.alloc(sizeof(int) * (size_t) n * (size_t) capacity));
    int *const stage = stage_buf.as<int>();
    PB_CHECK(cudaMemsetAsync(stage, 0xff, sizeof(int) * (size_t) n * (size_t) capacity, ctx->stream));      // FIXED particles: no list, all -1
    PB_LAUNCH(pb_k_tile_export<false>, ctx->ntiles, PB_TILE_M, n, ctx->ncap, ctx->tile_T4, capacity, ctx->tiles, ctx->tile_hdrs, ctx->cell_list,
              ctx->twords, ctx->numneigh, ctx->tile_rowsrc, stage);
    PB_CHECK(cudaMemcpyAsync(out, stage, sizeof(int) * (size_t) n * (size_t) capacity, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// the tile lists as 32-bit per-particle lists in the sliced-ELLPACK layout (T slots per particle), for pb_require_neigh32
int pb_tile_export_ell(pb_ctx *ctx, int *neigh, int T) {
    if(ctx->tiles_n <= 0) { return 0; }
    PB_LAUNCH(pb_k_tile_export<true>, ctx->ntiles, PB_TILE_M, ctx->tiles_n, ctx->ncap, ctx->tile_T4, T, ctx->tiles, ctx->tile_hdrs, ctx->cell_list,
              ctx->twords, ctx->numneigh, ctx->tile_rowsrc, neigh);
    return 0;
}
