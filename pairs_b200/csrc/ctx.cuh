// Internal context of libpairs_b200.so (see include/pairs_b200.h for the public C-ABI).
//
// Device data layout (all owned by the context; the reference's AoS host layout exists only at the
// upload/download boundary):
//   pos    double4[pcap]      x, y, z, w = particle type in the low 32 bits (so a neighbour gather is ONE
//                             32-byte sector: position and the index into the epsilon/sigma6 tables)
//   vel    double[3][pcap]    SoA, coalesced streaming in the integrators
//   force  double[3][pcap]    SoA
//   mass   double[pcap];  type/flags/uid/shape/tag  int[pcap]
//   [0, nlocal) are this rank's particles in CELL ORDER (sorted by the reference's flat cell index,
//   ties by previous index: a stable, deterministic counting sort); [nlocal, nlocal+nghost) are ghosts.
//   neigh  int[ncap][pitch]   padded column-major (ELLPACK) neighbour lists, pitch = nlocal rounded to 32
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdint>
#include <algorithm>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/pairs_b200.h"

#define PB_CHECK(call)                                                                                         \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if(e_ != cudaSuccess) {                                                                                \
            ctx->set_error(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +        \
                           std::to_string(__LINE__) + ")");                                                    \
            return -1;                                                                                         \
        }                                                                                                      \
    } while(0)

#define PB_TRY(expr)                   \
    do {                               \
        int rc_ = (expr);              \
        if(rc_ < 0) { return rc_; }    \
    } while(0)

static const int PB_XPROP_MAX_COMPS = 9;   // a matrix property
static const int PB_XPROP_MAX_ROWS = 48;   // rows of user-defined properties per context (row lists are kernel arguments)

struct PbTimer {
    double ms = 0.0;
    long calls = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;   // recorded, not yet read back (no sync on the hot path)
};

struct pb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    long launches = 0;

    // ---- domain (runtime/domain/regular_6d_stencil.cpp) ----
    bool domain_set = false;
    double grid[6] = {0, 0, 0, 0, 0, 0};
    int pbc_flag[3] = {1, 1, 1};
    int partitioner = 0, world = 1, rank = 0;
    int nranks[3] = {1, 1, 1}, coords[3] = {0, 0, 0};
    int neighbor_ranks[6] = {0, 0, 0, 0, 0, 0}, pbc[6] = {1, -1, 1, -1, 1, -1};
    double subdom[6] = {0, 0, 0, 0, 0, 0};

    // ---- particles ----
    int pcap = 0, nlocal = 0, nghost = 0;
    int tag_base = 0;
    double4 *pos = nullptr, *pos_alt = nullptr;
    double *vel = nullptr, *vel_alt = nullptr, *force = nullptr;
    double *mass = nullptr, *mass_alt = nullptr;
    int *type = nullptr, *type_alt = nullptr, *flags = nullptr, *flags_alt = nullptr;
    int *uid = nullptr, *uid_alt = nullptr, *shape = nullptr, *shape_alt = nullptr, *tag = nullptr, *tag_alt = nullptr;
    bool ghosts_in_alt = false;   // after a fused initial_integrate: ghosts of the last refresh still live in pos_alt
    bool dem_fuse = true;         // pb_dem_run folds usage reset / volatile reset / gravity / history clean-up into the contact kernel
    int dem_sort_every = 200;     // DEM: re-sort the locals into cell order every so many iterations (0 = never; option "dem_sort_every")
    bool half_lists = false;      // compute_half(): lists hold j with i < j only, the force kernel updates both partners
    bool fuse_integrate = true;   // pb_md_run folds the integrator halves into the force kernel's epilogue
    bool force_is_zero = false;   // reset_volatile requested and not yet materialised (fused into the force kernel)

    // ---- cells (sim/cell_lists.py) ----
    bool cells_set = false;
    double spacing = 0.0;
    int dim_cells[3] = {0, 0, 0}, ncells = 0, stencil[27];
    int ccap = 0;                 // capacity of per-cell arrays
    int *particle_cell = nullptr; // [pcap]
    int *cell_count = nullptr;    // [ccap+1]
    int *cell_start = nullptr;    // [ncells+1] coarse CSR (one entry per reference cell)
    int *sub_start = nullptr;     // [ncells*zsub+1] CSR over (cell, z slab)
    int zsub = 8, zsub_active = 1;   // z slabs per cell (option "cell_zsub"); 1 in DEM mode
    int *cell_slot = nullptr;     // [pcap] slot of the particle inside its cell (arrival order, sorted later)
    int *cell_list = nullptr;     // [pcap]
    int *sel_blocks = nullptr;    // [2][blocks] per-block selection counts of pb_borders + 2 totals
    int sel_blocks_cap = 0;
    int *cell_key = nullptr;      // [pcap] counting-sort key cell*zsub + slab
    int *scan_tmp = nullptr;      // block sums for the scan
    int scan_tmp_cap = 0;
    int cells_n = 0;              // number of particles binned by the last pb_build_cell_lists

    // ---- neighbour lists (sim/neighbor_lists.py) ----
    int ncap = 0, pitch = 0, max_neigh = 0;
    int lanes = 1;                // lanes of a warp that share one particle's list in the force kernel (1,2,4,8,16)
    bool stage_lists = false;     // neighbour build: collect lists in shared memory first (measured slower: 4.9 vs 3.1 ms)
    int lj_unroll = 4;            // independent gathers in flight per lane
    int nslots = 0;               // list slots per group row: ceil(ncap / lanes)
    size_t neigh_bytes = 0;
    int *neigh = nullptr, *numneigh = nullptr;
    int neigh_n = 0;              // nlocal at build time
    // interior / boundary split of the warp groups (32 particles each) for comm / compute overlap
    int *group_flag = nullptr, *group_scan = nullptr, *groups_interior = nullptr, *groups_boundary = nullptr;
    int group_cap = 0, n_interior = 0, n_boundary = 0;
    bool groups_valid = false;
    const int *lj_groups = nullptr;   // group list of the force launch being issued (null = all groups)
    int lj_ngroups = 0;
    bool overlap_comm = true;     // multi-rank: refresh ghosts on comm_stream while the interior groups compute
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_prev = nullptr, ev_sync = nullptr;
    // pb_md_run_from_host: velocities and masses are still on their way (comm_stream) while the first list build runs on the
    // positions; ev_io[0] main -> comm_stream (positions copied / permutation ready), ev_io[1] comm_stream -> main (arrays complete)
    bool upload_pending = false;
    cudaEvent_t ev_io[2] = {nullptr, nullptr};
    int *upload_perm = nullptr;
    size_t upload_perm_cap = 0;

    // persistent staging area of the bulk transfers (pb_upload_particles, pb_download_real): no allocation, no free and hence no
    // device-wide synchronisation inside a transfer
    double *io_stage = nullptr;
    size_t io_stage_bytes = 0;

    // ---- tile lists (tile_lists.cu): the list format of the MD hot path.  A tile = the cells [za, zb] of a 2 x 2 block of cell
    //      columns holding <= PB_TILE_M particles; its CTA stages the 4 x 4 columns around it in shared memory and walks 16-bit
    //      tile-relative neighbour lists (4 entries per 64-bit word, sliced ELLPACK over list ROWS = tile-major particle order).
    //      The 32-bit per-particle lists above are then built only on demand (pb_require_neigh32). ----
    bool tile_lists = true;       // option "tile_lists"
    bool tile_prefilter = true;   // option "tile_prefilter": the tile build tests candidates in fp32 first, exact fp64 only near the cutoff (same lists)
    float4 *m32 = nullptr;        // fp32 copy of the mirror + meta byte, valid during a list build only
    bool tile_reorder = true;     // option "tile_reorder": list rows in the conflict-aware order (tile_lists.cu pb_tile_reorder_row)
    struct PbTileHdr *tile_hdrs = nullptr;   // [ntiles] run tables (tile_lists.cu)
    int tile_hdrs_cap = 0;
    unsigned char *tile_rowsrc = nullptr;    // [tile_rows + PB_TILE_M] list row -> the core thread (= core particle of its tile) the row belongs to
    int tile_rowsrc_cap = 0;
    // the mirror: positions in CSR order, split xy / z, double-buffered like pos / pos_alt; meta byte (type | 8 for a ghost) and the
    // CSR position of every ghost.  `mirror_fresh` is honoured only while `mirror_scope` is set (inside pb_md_run).
    double2 *mxy[2] = {nullptr, nullptr};
    double *mz[2] = {nullptr, nullptr};
    unsigned char *mmeta = nullptr;
    int *ghost_csr = nullptr;
    int mirror_cap = 0, mirror_cur = 0, mirror_n = -1;
    bool mirror_scope = false, mirror_fresh = false;
    bool lj_fma = true;           // option "lj_fma": fused multiply-adds + Newton reciprocal in the pair term (md_math.h)
    struct PbTile *tiles = nullptr;
    int tiles_cap = 0, ntiles = 0, tile_rows = 0, tile_T4 = 0;
    int *tile_lvl = nullptr;      // [2][nsc * dim2] particles per (super-column, z level): core cells / staged cells
    int *tile_cnt = nullptr, *tile_off = nullptr;   // [nsc + 1] tiles per super-column and their prefix sums
    int *tile_pad = nullptr, *tile_row = nullptr;   // [ntiles + 1] rows per tile (core rounded up to 32) and their prefix sums
    int tile_sc_cap = 0, tile_lvl_cap = 0;
    unsigned long long *twords = nullptr;
    size_t twords_bytes = 0;
    int tiles_n = -1;             // nlocal when the tile lists were built (-1: none)
    double list_cutoff = 0.0;     // cutoff of the last pb_build_neighbor_lists (a lazy 32-bit build uses it)
    int *tile_flag = nullptr, *tile_scan = nullptr, *tiles_interior = nullptr, *tiles_boundary = nullptr;   // overlap split, per tile
    int tile_flag_cap = 0, n_tiles_interior = 0, n_tiles_boundary = 0;
    bool tile_split_valid = false;

    // ---- DEM (examples/dem.py): extra particle properties + per-particle contact history ----
    bool dem = false;
    int ccontacts = 0;            // contact capacity per particle (neighbor_capacity of pairs.simulation(), 20 in dem.py)
    double *radius = nullptr;     // [pcap]
    double *angvel = nullptr, *torque = nullptr, *normal = nullptr;   // SoA [3][pcap]
    double *inv_inertia = nullptr, *rotmat = nullptr;                 // SoA [9][pcap]
    double *quat = nullptr;                                           // SoA [4][pcap]
    int *num_contacts = nullptr;  // [pcap]
    int *contact_uid = nullptr, *contact_used = nullptr, *contact_stick = nullptr;   // [ccontacts][pcap]
    double *contact_tsd = nullptr;   // [3][ccontacts][pcap]
    double *contact_ivm = nullptr;   // [ccontacts][pcap]
    int dem_ntypes = 1;
    double *d_fric_static = nullptr, *d_fric_dynamic = nullptr;
    double dem_params[16];        // PbDemParams, see dem_math.h
    int *d_dem_flag = nullptr;    // [0]: contact capacity overflow
    // contact properties beyond examples/dem.py's three (one integer, one vector, one real): `cx` further double lanes per contact,
    // layout [lane][slot][particle] like the tangential displacement; only contact models compiled at run time use them
    double *contact_x = nullptr;
    int cx = 0;
    double cx_default[16] = {0};

    // ---- user-defined properties (props.cu): what add_property() declares beyond the built-in set, as SoA rows
    //      xdata[xrows][pcap] (a real = 1 row, a vector = 3 rows).  Non-volatile rows travel with their particle (cell-order
    //      sort, migration, ghost creation); volatile rows are zeroed by reset_volatile. ----
    struct XProp {
        std::string name;
        int comps = 1, row0 = 0;
        bool is_volatile = false;
        double dflt[PB_XPROP_MAX_COMPS] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    };
    std::vector<XProp> xprops;
    int xrows = 0, xrows_nv = 0;  // all rows / rows of non-volatile properties
    double *xdata = nullptr, *xdata_alt = nullptr;

    // ---- LJ feature properties ----
    int ntypes = 0;
    bool lj_uniform = false;
    double *d_eps = nullptr, *d_sig6 = nullptr;
    double h_eps[64], h_sig6[64];

    // ---- comm (sim/comm.py) ----
    int send_cap = 0;             // entries
    int nsend_all = 0, nsend[6] = {0, 0, 0, 0, 0, 0}, nrecv[6] = {0, 0, 0, 0, 0, 0};
    int send_offsets[6] = {0, 0, 0, 0, 0, 0}, recv_offsets[6] = {0, 0, 0, 0, 0, 0};
    int *send_map = nullptr;      // [send_cap] source index of every send entry (locals or ghosts)
    int *send_mult = nullptr;     // [send_cap][3]
    double *send_buf = nullptr, *recv_buf = nullptr; // [send_cap][PB_MAX_ELEMS]
    int recv_cap = 0;
    int *sel_flag = nullptr, *sel_scan = nullptr; // [pcap+1] compaction scratch
    void *jit = nullptr;          // table of NVRTC-compiled user kernels (jit.cu)
    void *dem_user_force = nullptr;   // DEM contact kernel built around a user-defined contact model (jit.cu), null = examples/dem.py's
    int dem_force_maxreg = 0;         // tuning option "dem_force_maxreg": > 0 re-builds the contact kernel at run time with this register cap
    void *nccl = nullptr;         // NcclState* (comm_nccl.cu), null on a single rank

    // ---- reductions / host mirrors ----
    double *d_partial = nullptr;  // thermo partial sums
    int d_partial_cap = 0;
    int *d_scalars = nullptr;     // small device scratch: [0] max neighbours, [1..] counters
    int *h_scalars = nullptr;     // pinned mirror

    // ---- timers: CUDA events on the launching stream, collected lazily (pb_timers_get) ----
    bool timers_on = false;
    bool nvtx = false;            // option "profiler": every stage is also an NVTX range (what enable_profiler()'s LIKWID markers are in the reference)
    std::map<std::string, PbTimer> timers;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // pb_stream_timer_start / stop

    cudaEvent_t get_event() {
        if(!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }

    void set_error(const std::string &e) { err = e; }
};

// tile lists (tile_lists.cu)
struct PbTile { int X0, Y0, za, zb, row_base, ncore; };
static const int PB_TILE_M = 256;          // threads per CTA = most particles of a tile
static const int PB_TILE_CAP = 2048;       // staged particles per tile (slots are 12-bit: <= 4096)
int pb_build_tile_lists(pb_ctx *ctx, double cutoff);        // 0 built, 1 not applicable here (caller takes the per-particle path)
int pb_tile_lennard_jones(pb_ctx *ctx, double cutsq, double dt, int fuse, int part);
int pb_tile_finish_split(pb_ctx *ctx);
int pb_tile_download_neighbors(pb_ctx *ctx, int *out, int capacity);
int pb_io_stage(pb_ctx *ctx, size_t bytes, double **out);
int pb_tile_mirror_all(pb_ctx *ctx, bool with_f32 = false);
int pb_tile_mirror_ghosts(pb_ctx *ctx);
int pb_require_neigh32(pb_ctx *ctx);                        // per-particle 32-bit lists for the kernels that walk them (built lazily)
static inline bool pb_lists_valid(const pb_ctx *ctx) { return ctx->tiles_n == ctx->nlocal || ctx->neigh_n == ctx->nlocal; }

static const int PB_MAX_ELEMS = 16;   // doubles per packed particle record in MD (exchange: 12, borders: 11/15, sync: 6)
// DEM exchange record: 12 base + radius 1 + angvel 3 + normal 3 + inv_inertia 9 + rotmat 9 + quat 4 + num_contacts 1 + 6 per slot
// user-defined properties: their non-volatile rows follow the built-in elements of a record (MD: 12 exchange / 11 borders,
// DEM: 42 + 6 C exchange / 15 borders)
static inline int pb_exchange_base_elems(const pb_ctx *ctx) { return ctx->dem ? 42 + (6 + ctx->cx) * ctx->ccontacts : 12; }
static inline int pb_record_elems(const pb_ctx *ctx) { return std::max(PB_MAX_ELEMS, pb_exchange_base_elems(ctx) + ctx->xrows_nv); }
static const int PB_NSCALARS = 16;

// ---- user-defined properties (props.cu) ----
struct PbXRows {                   // the rows of the non-volatile user properties, in declaration order
    int n;
    int row[PB_XPROP_MAX_ROWS];
};
PbXRows pb_xprops_nv_rows(const pb_ctx *ctx);
int pb_xprops_grow(pb_ctx *ctx, size_t oldcap, size_t newcap, size_t used);
int pb_xprops_defaults(pb_ctx *ctx);                                   // every slot back to the declared default
int pb_xprops_permute(pb_ctx *ctx, const int *perm, int n);            // new index k <- old index perm[k], locals [0, n)
int pb_xprops_reset_volatile(pb_ctx *ctx);
// wire records: the non-volatile rows follow the `offset` built-in elements of a record of `stride` doubles
int pb_xprops_pack(pb_ctx *ctx, int first, int count, int stride, int offset, const int *send_map, double *buf);
int pb_xprops_pack_leavers(pb_ctx *ctx, int n, int stride, int offset, const int *rec, double *buf);
int pb_xprops_unpack(pb_ctx *ctx, int first_rec, int count, int dst0, int stride, int offset, const double *buf);
int pb_xprops_move(pb_ctx *ctx, int max_count, const int *count, const int *src_idx, const int *dst_idx);

// ---- helpers shared by the .cu files ----
int pb_ensure_particle_capacity(pb_ctx *ctx, int needed);
int pb_ensure_send_capacity(pb_ctx *ctx, int needed);
int pb_exclusive_scan(pb_ctx *ctx, const int *in, int *out, int n);   // out[0..n], out[n] = total
int pb_bin_particles(pb_ctx *ctx, int first, int n, bool write_particle_cell);
// Scratch device memory of the set-up / test paths (upload, download): released on every exit path, error returns included.
struct PbScratch {
    void *p = nullptr;
    PbScratch() = default;
    PbScratch(const PbScratch &) = delete;
    PbScratch &operator=(const PbScratch &) = delete;
    ~PbScratch() { if(p != nullptr) { cudaFree(p); } }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
    template<typename T> T *as() const { return (T *) p; }
    template<typename T> T *release() { T *q = (T *) p; p = nullptr; return q; }     // the caller keeps the allocation
};

int pb_dem_grow(pb_ctx *ctx, size_t oldcap, size_t newcap, size_t used);
int pb_dem_sort_locals(pb_ctx *ctx);
int pb_sort_locals(pb_ctx *ctx);
template<typename T> int pb_regrow(pb_ctx *ctx, T **p, size_t old_count, size_t new_count, bool keep);
int pb_regrow_soa(pb_ctx *ctx, double **p, int comps, size_t old_cap, size_t new_cap, size_t used, bool keep);

// Brackets one stage with two events from the pool; nothing is synchronised here.
struct PbStage {
    pb_ctx *ctx;
    const char *name;
    cudaEvent_t a = nullptr;
    bool ranged = false;
    PbStage(pb_ctx *c, const char *n) : ctx(c), name(n) {
        if(ctx->nvtx) { nvtxRangePushA(n); ranged = true; }
        if(ctx->timers_on) { a = ctx->get_event(); cudaEventRecord(a, ctx->stream); }
    }
    ~PbStage() {
        if(ranged) { nvtxRangePop(); }
        if(a != nullptr) {
            cudaEvent_t b = ctx->get_event();
            cudaEventRecord(b, ctx->stream);
            ctx->timers[name].pending.emplace_back(a, b);
        }
    }
};

static inline int pb_blocks(long n, int threads) { return (int) ((n + threads - 1) / threads); }

#define PB_LAUNCH(kernel, grid, block, ...)                                    \
    do {                                                                       \
        kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__);              \
        ctx->launches++;                                                       \
        PB_CHECK(cudaGetLastError());                                          \
    } while(0)

// One 256-bit read-only load (sm_100: LDG.E.256): a particle's double4 is exactly one 32-byte sector, fetched with a
// single instruction / single L1 tag lookup instead of two 128-bit halves.
__device__ __forceinline__ double4 pb_ld_pos(const double4 *p) {
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

// Neighbour-list layout: "interleaved sliced ELLPACK".  A warp serves A = 32/G particles, G lanes each.  The k-th
// neighbour of particle i lives at
//     ((i / A) * T + k / G) * 32 + (i % A) * G + k % G          T = ceil(capacity / G)
// so that in iteration t the 32 lanes of the warp read 32 CONSECUTIVE ints (one 128-byte line), and the G lanes of
// one particle fetch G consecutive neighbours of its (cell-sorted, hence memory-adjacent) list -> their 32-byte
// position gathers fall into the same one or two 128-byte lines.  G = 1 is the classic slice-32 ELLPACK.
struct PbNeighLayout {
    int G, A, T;
    __host__ __device__ __forceinline__ size_t idx(int i, int k) const {
        return ((size_t) (i / A) * T + (size_t) (k / G)) * 32 + (size_t) ((i % A) * G + (k % G));
    }
};

// particle type rides in the low 32 bits of pos.w
__device__ __forceinline__ int pb_w_type(double w) { return (int) (__double_as_longlong(w) & 0xffffffffLL); }
__device__ __forceinline__ double pb_type_w(int t) { return __longlong_as_double((long long) (unsigned int) t); }
