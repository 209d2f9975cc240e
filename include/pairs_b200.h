/* pairs_b200.h -- C-ABI of libpairs_b200.so: a B200 (sm_100a) execution backend for the per-timestep
 * particle pipeline of P4IRS/"pairs" (cell binning -> neighbour lists -> pair force -> integrate ->
 * halo/migration).
 *
 * This is the LOWER boundary of the drop-in: in the reference every building
 * block is a generated module `void <name>(PairsSimulation *pairs, scalars..., arrays...)`
 * (code_gen/cgen.py:144-200) calling into the C++ runtime (runtime/pairs.hpp:27-297).  Here each
 * module is one entry point on an opaque context that owns all device memory.  Python
 * (pairs_b200/, the unchanged `pairs` DSL surface) binds these with ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - every entry returns int: 0 = ok, < 0 = error (text via pb_last_error), > 0 only where documented.
 *   - host arrays use the reference's layouts: vectors are AoS [n][3] doubles, ints are int32.
 *   - `grid`  = {xmin, xmax, ymin, ymax, zmin, zmax} (argument order of PairsSimulation::initDomain,
 *     runtime/pairs.cpp:14-16); `subdom` likewise per rank.
 *   - the library never exits the process and has NO CPU fallback: without a CUDA device pb_create fails.
 *   - one context per GPU / per process rank; entry points are not re-entrant per context.
 */
#ifndef PAIRS_B200_H
#define PAIRS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_ctx pb_ctx;

/* particle flags (runtime/pairs.hpp:20-23, src/pairs/sim/flags.py) and shapes (src/pairs/sim/shapes.py) */
#define PB_FLAG_INFINITE 1
#define PB_FLAG_GHOST 2
#define PB_FLAG_FIXED 4
#define PB_FLAG_GLOBAL 8
#define PB_SHAPE_SPHERE 0
#define PB_SHAPE_HALFSPACE 1
#define PB_SHAPE_POINTMASS 2

/* domain partitioners (runtime/pairs.cpp:18-26: Regular = {1,1,1}, RegularXY = {1,1,0}) */
#define PB_PARTITION_REGULAR 0
#define PB_PARTITION_REGULAR_XY 1

/* ---- lifecycle: replaces `new PairsSimulation(...)` / `delete pairs` (code_gen/cgen.py:128) ---- */
int pb_create(pb_ctx **out, int device);
void pb_destroy(pb_ctx *ctx);
const char *pb_last_error(const pb_ctx *ctx); /* ctx may be NULL: error of the last failed pb_create */
const char *pb_version(void);

/* ---- domain: PairsSimulation::initDomain + Regular6DStencil::{setConfig,setBoundingBox,fillArrays}
 *      (runtime/pairs.cpp:14-29, runtime/domain/regular_6d_stencil.cpp:10-111), without MPI:
 *      `rank`/`world_size` come from the launcher (torchrun env). ---- */
int pb_init_domain(pb_ctx *ctx, const double grid[6], const int pbc[3], int partitioner, int world_size, int rank);
int pb_get_decomposition(const pb_ctx *ctx, int nranks[3], int neighbor_ranks[6], int pbc[6], double subdom[6]);
/* pure helper = Regular6DStencil::setConfig */
int pb_rank_grid(int world_size, const double grid[6], int partitioner, int nranks[3]);

/* ---- capacities: particle_capacity / neighbor_capacity of pairs.simulation() (src/pairs/__init__.py:9-17).
 *      Both grow automatically (the reference's resize protocol, transformations/modules.py:159-203). ---- */
int pb_reserve(pb_ctx *ctx, int particle_capacity, int neighbor_capacity);

/* ---- set-up (host side, bit-identical to the reference's runtime) ----
 * pb_copper_fcc_lattice = pairs::copper_fcc_lattice (runtime/copper_fcc_lattice.hpp:64-145); returns nlocal
 * of this rank in *nlocal.  pb_adjust_thermo = pairs::adjust_thermo (runtime/thermo.hpp:53-97). */
int pb_copper_fcc_lattice(pb_ctx *ctx, int nx, int ny, int nz, double rho, int ntypes, int *nlocal);
int pb_adjust_thermo(pb_ctx *ctx, double temp);
/* Bulk upload of local particles (replaces property initialisation + copy*ToDevice, runtime/pairs.cpp:161-327).
 * Any pointer except position may be NULL (defaults: velocity 0, mass 1, type 0, flags 0, uid 0, shape point mass). */
int pb_upload_particles(pb_ctx *ctx, int n, const double *position, const double *velocity, const double *mass,
                        const int *type, const int *flags, const int *uid, const int *shape);

/* ---- state download (tests, output): device order; `tag` = index the particle had at upload / lattice time
 *      (+ first tag of the rank), ghosts carry the tag of their source.  n = nlocal (+ nghost if with_ghosts). ---- */
int pb_counts(const pb_ctx *ctx, int *nlocal, int *nghost);
int pb_download_real(pb_ctx *ctx, const char *name, double *out, int with_ghosts); /* position|linear_velocity|force [n][3], mass [n] */
int pb_download_int(pb_ctx *ctx, const char *name, int *out, int with_ghosts);     /* type|flags|uid|shape|tag|particle_cell|numneighs */
/* neighbour lists as the reference stores them on the CPU: out[i*capacity + k] (sim/neighbor_lists.py:14), device indices */
int pb_download_neighbors(pb_ctx *ctx, int *out, int capacity);
int pb_neighbor_capacity(const pb_ctx *ctx);
int pb_max_neighbors(const pb_ctx *ctx);
/* ghost bookkeeping (send_map / send_mult of sim/comm.py:223-259, one entry per ghost this rank OWNS as receiver):
 * src[g] = device index of the source particle on the sending rank, mult[g][3] = PBC multipliers */
int pb_download_ghost_map(pb_ctx *ctx, int *src, int *mult);

/* ---- cell lists ----
 * pb_setup_cells  = BuildCellListsStencil (sim/cell_lists.py:46-87): dim_cells, ncells, 27-cell stencil.
 * pb_build_cell_lists = BuildCellLists + PartitionCellLists (sim/cell_lists.py:90-171) over nlocal+nghost particles:
 *   particle_cell[] is bit-identical to the reference's; storage is a CSR cell list (counting sort). */
int pb_setup_cells(pb_ctx *ctx, double spacing);
int pb_get_cells(const pb_ctx *ctx, int dim_cells[3], int *ncells, int stencil[27]);
int pb_build_cell_lists(pb_ctx *ctx);
/* CSR cell list download: cell_start[ncells+1], cell_list[nlocal+nghost] */
int pb_download_cell_lists(pb_ctx *ctx, int *cell_start, int *cell_list);

/* ---- neighbour lists: BuildNeighborLists (sim/neighbor_lists.py:21-48), full lists, ghosts included as j,
 *      padded column-major (ELLPACK) storage on the device. ---- */
int pb_build_neighbor_lists(pb_ctx *ctx, double cutoff);

/* ---- kernels of examples/md.py ---- */
/* feature properties epsilon/sigma6 [ntypes*ntypes] (add_feature_property, sim/simulation.py:179) */
int pb_set_lj_params(pb_ctx *ctx, int ntypes, const double *epsilon, const double *sigma6);
int pb_reset_volatile(pb_ctx *ctx);                       /* ResetVolatileProperties, sim/properties.py:61-70 */
int pb_lennard_jones(pb_ctx *ctx, double cutoff);         /* examples/md.py:5-8 via ParticleInteraction */
int pb_initial_integrate(pb_ctx *ctx, double dt);         /* examples/md.py:11-13 */
int pb_final_integrate(pb_ctx *ctx, double dt);           /* examples/md.py:16-17 */
/* kernels of examples/lj_onetype.py (older P4IRS API): scalar epsilon/sigma6, explicit Euler */
int pb_lj_legacy(pb_ctx *ctx, double cutoff, double epsilon, double sigma6);
int pb_euler_legacy(pb_ctx *ctx, double dt);
/* potential energy and virial of the LJ system (rank-local sums over the current lists; an addition, the reference computes neither) */
int pb_lj_energy_virial(pb_ctx *ctx, double cutoff, double *epot, double *virial);
/* pairs::compute_thermo (runtime/thermo.hpp:11-51): T and P over ALL ranks' locals (rank-local sums are returned
 * in *sum_mv2 / *natoms when world_size > 1 and no communicator is attached) */
int pb_compute_thermo(pb_ctx *ctx, double *temperature, double *pressure);
int pb_thermo_partial(pb_ctx *ctx, double *sum_mv2, int *natoms);

/* ---- communication (sim/comm.py): exchange = migration + PBC wrap (+ cell-order reordering of the locals),
 *      borders = ghost creation, synchronize = per-step ghost refresh ---- */
int pb_exchange(pb_ctx *ctx);
int pb_borders(pb_ctx *ctx);
int pb_synchronize(pb_ctx *ctx);

/* ---- user-defined particle properties: Simulation.add_property() beyond the built-in set (sim/simulation.py:166-168; one array per
 *      property in the reference, sim/properties.py:13-36).  ncomps = 1 (real), 3 (vector), up to 9; host layout [n][ncomps].
 *      Non-volatile properties travel with their particle through the cell-order sort, migration (Comm.exchange carries every
 *      non-volatile property, sim/comm.py:100-103) and ghost creation; volatile ones are zeroed by pb_reset_volatile
 *      (sim/properties.py:61-70).  New particles start from `defaults` (NULL = zeros).  Generated kernels (pb_jit_*) address
 *      component d as PbJitArgs.xdata[(row0 + d) * cap + i]; pb_property_info returns row0.  Available on the md.py and the DEM path. ---- */
int pb_add_property(pb_ctx *ctx, const char *name, int ncomps, int is_volatile, const double *defaults, int *prop_id);
int pb_property_count(const pb_ctx *ctx);
int pb_property_info(const pb_ctx *ctx, int prop_id, int *ncomps, int *row0, int *is_volatile);
int pb_upload_property(pb_ctx *ctx, int prop_id, int n, const double *values);        /* the first n locals */
int pb_download_property(pb_ctx *ctx, int prop_id, double *out, int with_ghosts);     /* device order, like pb_download_real */

/* ---- DEM path of examples/dem.py (spheres + half-spaces, contact history, cell-list traversal), one or several GPUs. ----
 * pb_dem_enable            use_contact_history=True + the DEM property set (contact_capacity = neighbor_capacity of pairs.simulation())
 * pb_dem_set_params        symbols of the kernels (examples/dem.py:189-201) + feature properties friction_static/dynamic
 * pb_dem_sc_grid           pairs::dem_sc_grid (runtime/dem_sc_grid.hpp:62-172), host arrays out (NULL pointers: count only)
 * pb_dem_upload_real/download_real   radius | angular_velocity | torque | normal | inv_inertia | rotation_matrix | rotation_quat |
 *                                    force | mass | linear_velocity, n particles from index `first` (AoS host layout)
 * pb_dem_upload/download_contacts    contact history in the reference's layout: num[n], uid/used/sticking [n][C], tsd [n][C][3], ivm [n][C]
 * pb_dem_update_mass_and_inertia     setup() function of examples/dem.py:6-15
 * pb_dem_reset_contact_usage / pb_dem_clear_unused_contacts   sim/contact_history.py:75-127
 * pb_dem_gravity / pb_dem_linear_spring_dashpot / pb_dem_euler  compute() kernels of examples/dem.py:18-91
 * pb_dem_run               the generated DEM loop (exchange, borders, cell lists, ..., euler, clear) for ts in [ts_begin, ts_end) */
int pb_dem_enable(pb_ctx *ctx, int contact_capacity);
int pb_dem_set_params(pb_ctx *ctx, double dt, double pi, double kappa, double ln_dry_res_coeff, double collision_time,
                      double density_particle, double density_fluid, double gravity, int ntypes, const double *friction_static,
                      const double *friction_dynamic);
int pb_dem_sc_grid(pb_ctx *ctx, double xmax, double ymax, double zmax, double spacing, double diameter, double min_diameter,
                   double max_diameter, double initial_velocity, double particle_density, int ntypes, int capacity, int *uid, int *type,
                   double *mass, double *radius, double *position, double *velocity, int *count);
int pb_dem_upload_real(pb_ctx *ctx, const char *name, int first, int n, const double *data);
int pb_dem_download_real(pb_ctx *ctx, const char *name, int first, int n, double *out);
int pb_set_counts(pb_ctx *ctx, int nlocal, int nghost);
int pb_dem_upload_contacts(pb_ctx *ctx, int n, const int *num, const int *uid, const int *sticking, const double *tsd, const double *ivm);
int pb_dem_download_contacts(pb_ctx *ctx, int n, int *num, int *uid, int *used, int *sticking, double *tsd, double *ivm);
int pb_dem_update_mass_and_inertia(pb_ctx *ctx);
int pb_dem_reset_contact_usage(pb_ctx *ctx);
int pb_dem_clear_unused_contacts(pb_ctx *ctx);
int pb_dem_gravity(pb_ctx *ctx);
int pb_dem_linear_spring_dashpot(pb_ctx *ctx);
int pb_dem_euler(pb_ctx *ctx);
int pb_dem_contact_overflow(pb_ctx *ctx);
/* contact-capacity growth (the reference's resize protocol for the contact-history arrays, transformations/modules.py:159-203):
   reads the contact kernel's high-water mark and doubles the capacity (all ranks together, <= 64) once a row is nearly full;
   < 0 if a contact was lost before that.  pb_dem_run calls it every 8 iterations; module-by-module loops call it themselves. */
int pb_dem_check_contacts(pb_ctx *ctx);
/* contact properties beyond examples/dem.py's three (add_contact_property, sim/simulation.py:185; mapping/funcs.py:230-263 gives
   every declared contact property its own array): `extra_lanes` further doubles per contact (a real = 1, a vector = 3, an integer
   = 1 exact double), `extra_defaults` = what a fresh contact starts from.  They migrate, sort and are cleaned up with the row; only
   contact models compiled at run time (pb_jit_set_dem_model) read or write them.  pb_dem_contact_extras moves them between the
   host layout [n][contact_capacity][extra_lanes] and the device. */
int pb_dem_enable_ex(pb_ctx *ctx, int contact_capacity, int extra_lanes, const double *extra_defaults);
int pb_dem_contact_extras(pb_ctx *ctx, int n, double *values, int upload);
int pb_dem_contact_extra_lanes(const pb_ctx *ctx);
int pb_dem_contact_capacity(const pb_ctx *ctx);
int pb_dem_run(pb_ctx *ctx, double cell_spacing, int ts_begin, int ts_end);

/* multi-GPU: NCCL communicator over the ranks of pb_init_domain.  id = 128-byte ncclUniqueId made by rank 0
 * (pb_nccl_unique_id) and distributed by the launcher. */
int pb_nccl_unique_id(void *id128);
int pb_nccl_init(pb_ctx *ctx, const void *id128);

/* ---- whole timestep loop (sim/timestep.py:9-72 + sim/simulation.py:387-417), iterations ts in [ts_begin, ts_end):
 *      guards ((ts+1)%every==0)||ts==0 and ts>0 exactly as generated.  thermo_out (may be NULL) receives
 *      {ts, T, P} triples for every thermo step, up to thermo_cap triples; *n_thermo = number written. ---- */
typedef struct pb_md_params {
    double dt, cutoff_force, cutoff_lists, cell_spacing;
    int reneighbor_every, thermo_every;
} pb_md_params;
int pb_md_run(pb_ctx *ctx, const pb_md_params *p, int ts_begin, int ts_end, double *thermo_out, int thermo_cap, int *n_thermo);
/* pb_upload_particles + pb_md_run in one call for state that lives in host memory (the reference keeps host and device copies of
 * every property and moves them around each module, runtime/pairs.hpp copyArrayToDevice / copyPropertyToDevice): velocities and
 * masses are copied while the first neighbour-list build already runs on the positions.  Same results as the two calls, bit for
 * bit; falls back to them when the loop does not start with a reneighbouring iteration (ts_begin != 0), on several ranks, for DEM
 * contexts and with user-defined properties.  Arrays as in pb_upload_particles (pinned host memory makes the overlap real). */
int pb_md_run_from_host(pb_ctx *ctx, const pb_md_params *p, int n, const double *position, const double *velocity, const double *mass,
                        const int *type, const int *flags, const int *uid, const int *shape, int ts_begin, int ts_end,
                        double *thermo_out, int thermo_cap, int *n_thermo);

/* ---- host-only self-test of the shared-memory count board that replaces communicateSizes between the ranks of a node
 *      (runtime/domain/regular_6d_stencil.cpp:113-127): `rounds` messages per dimension on a periodic ring of `world`
 *      processes; returns 0 on success.  Needs no GPU; pb_board_unlink removes the named segment afterwards. ---- */
int pb_board_selftest(const char *name, int world, int rank, int rounds);
int pb_board_unlink(const char *name);

/* ---- user-defined kernels (the reference generates code for arbitrary kernel bodies: src/pairs/mapping/funcs.py:39-334,
 *      code_gen/cgen.py).  `source` is CUDA C++ printed by pairs_b200/kernelgen.py from the Python kernel; it starts with
 *      pb_jit_prelude() and defines  extern "C" __global__ void <kernel_name>(PbJitArgs).  Compiled at run time with NVRTC for
 *      sm_100a (--fmad=false).  pb_jit_check compiles only (no GPU needed): returns the cubin size, or -1 with the compiler
 *      log in `log`.  pb_jit_launch: kind 0 = pair kernel over the current neighbour lists with interaction cutoff `cutoff`,
 *      kind 1 = per-particle kernel, kind 2 = pair kernel over the current CELL lists (scripts without build_neighbor_lists,
 *      sim/interaction.py:92-118), kind 3 = pair kernel generated for compute_half() over half lists (ir/apply.py:111-125). ---- */
const char *pb_jit_prelude(void);
int pb_jit_check(const char *source, char *log, int log_cap);
int pb_jit_compile(pb_ctx *ctx, const char *source, const char *kernel_name, int *handle);
int pb_jit_launch(pb_ctx *ctx, int handle, int kind, double cutoff);

/* User-defined DEM contact models: the contact KERNEL (detection, contact history keyed by the partner's uid, usage marks, clean-up,
 * force / torque accumulation) stays the library's, the per-pair model of examples/dem.py:18-74 is exchanged for a device function
 * `bool <model_name>(xi, vi, wi, mi, ri, xj, vj, wj, mj, rj, n, cp, delta, tij, tsd, ivm, sticking, cx, F, T)` printed by
 * pairs_b200/kernelgen.py from the user's kernel body (the reference generates code for any body: mapping/funcs.py:39-334, contact
 * properties :230-263).  pb_jit_check_dem_model compiles only (cubin size or -1 + log, no GPU); pb_jit_set_dem_model installs the
 * model for pb_dem_linear_spring_dashpot / pb_dem_run (source NULL: back to the built-in). */
int pb_jit_check_dem_model(const char *model_source, const char *model_name, char *log, int log_cap);
int pb_jit_set_dem_model(pb_ctx *ctx, const char *model_source, const char *model_name);

/* Options.  Behaviour: "compute_half" (0/1) = Simulation.compute_half() (sim/simulation.py:119-120): half neighbour lists,
 * pair terms applied to both partners (ir/apply.py:111-125); applies from the next neighbour-list build.
 * Tuning knobs: "lanes_per_particle" (1,2,4,8; applies from the next build), "lj_unroll" (2,4,8), "fuse_integrate" (0/1),
 * "overlap_comm" (0/1), "profiler" (0/1: every stage also opens an NVTX range named like the reference's timers -- Simulation.enable_profiler(),
 * sim/simulation.py:116-117, LIKWID markers there), "cell_zsub" (1..32, applies from the next pb_setup_cells), "stage_lists" (0/1), "dem_sort_every"
 * (DEM: iterations between two spatial re-sorts of the locals, 0 = never, default 200), "dem_fuse" (0/1: pb_dem_run folds the
 * per-particle modules around the contact evaluation into the contact kernel; results are identical), "tile_lists" (0/1, default 1: the force kernel works on cell TILES staged in shared memory by TMA with 16-bit tile-relative
 * neighbour lists; 0 = per-particle 32-bit lists; applies from the next list build), "tile_reorder" (0/1, default 1: list rows in a
 * shared-memory-conflict-aware order -- same sets, another summation order; 0 = the reference's list order), "tile_prefilter" (0/1,
 * default 1: the tile build tests candidates in fp32 first and decides in fp64 only inside the rounding-error band of the cutoff;
 * the lists are identical either way), "lj_fma" (0/1, default 1: fused-multiply-add pair arithmetic, within 1e-12 of the
 * reference's expression tree; 0 = that expression tree operation for operation, bit-identical forces), "dem_force_maxreg" (0 = off: the DEM
 * contact kernel is re-built at run time with this register cap -- occupancy experiments; after pb_dem_enable). */
int pb_set_option(pb_ctx *ctx, const char *name, int value);

/* ---- streams / timing ---- */
int pb_synchronize_device(pb_ctx *ctx);
/* per-stage device time in ms accumulated with CUDA events when enabled (names follow the reference's timers) */
int pb_timers_enable(pb_ctx *ctx, int on);
int pb_timers_get(pb_ctx *ctx, const char *name, double *ms, long *calls);
int pb_timers_reset(pb_ctx *ctx);
/* CUDA-event bracket on the context's stream around an arbitrary region (bench.py's timed region) */
int pb_stream_timer_start(pb_ctx *ctx);
int pb_stream_timer_stop(pb_ctx *ctx, double *ms);
/* page-lock a caller-owned host buffer (cudaHostRegister) so uploads/downloads are DMA transfers */
int pb_host_register(pb_ctx *ctx, void *ptr, size_t bytes);
int pb_host_unregister(pb_ctx *ctx, void *ptr);
/* number of kernels launched by this context so far */
long pb_kernel_launches(const pb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
