/* pairs_oracle.c -- CPU restatement of the reference's MD hot path (P4IRS "pairs", examples/md.py).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this; the product (pairs_b200/) never does.
 *
 * Pinning: the reference ships no tests or golden vectors.  This restatement is
 * pinned against the reference ITSELF run in this container: oracle/_ref/libref_*.so is the
 * reference generator's own serial C++ output compiled from /root/reference (oracle/build_ref.py);
 * tests/test_oracle_pin.py checks this file bit-for-bit against it (per-step positions, velocities,
 * forces, neighbour lists), and against fixtures under tests/golden/ produced from it.
 *
 * The restatement follows the *serial* semantics of the generated program (the only valid
 * reference target): every loop runs in ascending index order, atomics are
 * plain increments.  Unlike the reference (one MPI process per rank) it can hold R ranks in one
 * process and move messages with memcpy, which is what MPI_Sendrecv does between ranks
 * (runtime/domain/regular_6d_stencil.cpp:113-238); with R = 1 it is the reference's single-rank
 * run.  Arithmetic is written operation by operation in the order the generator emits it
 * (cited per function); build with -ffp-contract=off.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PO_FLAG_INFINITE 1   /* src/pairs/sim/flags.py, runtime/pairs.hpp:20-23 */
#define PO_FLAG_GHOST 2
#define PO_FLAG_FIXED 4
#define PO_FLAG_GLOBAL 8
#define PO_SHAPE_POINTMASS 2 /* src/pairs/sim/shapes.py */

typedef struct po_rank {
    int rank, coords[3];
    int neighbor_ranks[6], pbc[6];
    double subdom[6];
    int nlocal, nghost, pcap;
    int *uid, *shape, *flags, *type;
    double *position, *mass, *velocity, *force;
    /* cell lists (sim/cell_lists.py:19-43) */
    int ncells, dim_cells[3], stencil[27], nstencil, cell_capacity;
    int *cell_particles, *cell_sizes, *nshapes, *particle_cell;
    /* neighbour lists (sim/neighbor_lists.py:12-19) */
    int neighbor_capacity;
    int *neighborlists, *numneighs;
    /* comm (sim/comm.py:19-43) */
    int nsend_all, nsend[6], nrecv[6], send_offsets[6], recv_offsets[6], send_capacity;
    int *send_map, *send_mult, *exchg_flag, *exchg_copy_to;
    double *send_buffer, *recv_buffer;
} po_rank;

typedef struct po_sim {
    int world, nranks[3], part_flags[3], pbc_flag[3];
    double grid[6]; /* xmin,xmax,ymin,ymax,zmin,zmax (initDomain argument order, runtime/pairs.cpp:14-16) */
    double cell_spacing, cutoff_lists, cutoff_force, dt;
    int ntypes;
    double epsilon[64], sigma6[64];
    int reneigh_every, thermo_every;
    int compute_half; /* Simulation.compute_half(), sim/simulation.py:119-120 */
    po_rank *r;
} po_sim;

/* ---- domain partitioning: runtime/domain/regular_6d_stencil.cpp:10-54 (setConfig) ---- */
void po_set_config(int world_size, const double grid[6], const int part_flags[3], int nranks[3]) {
    double area[3];
    double best_surf = 0.0;
    int d = 0;
    for(int d1 = 0; d1 < 3; d1++) {
        nranks[d1] = 1;
        for(int d2 = d1 + 1; d2 < 3; d2++) {
            area[d] = (grid[d1 * 2 + 1] - grid[d1 * 2]) * (grid[d2 * 2 + 1] - grid[d2 * 2]);
            best_surf += 2.0 * area[d];
            d++;
        }
    }
    for(int i = 1; i <= world_size; i++) {
        if(world_size % i == 0) {
            const int rem_yz = world_size / i;
            for(int j = 1; j <= rem_yz; j++) {
                if(rem_yz % j == 0) {
                    const int k = rem_yz / j;
                    if((part_flags[0] || i == 1) && (part_flags[1] || j == 1) && (part_flags[2] || k == 1)) {
                        const double surf = (area[0] / i / j) + (area[1] / i / k) + (area[2] / j / k);
                        if(surf < best_surf) {
                            nranks[0] = i; nranks[1] = j; nranks[2] = k;
                            best_surf = surf;
                        }
                    }
                }
            }
        }
    }
}

/* MPI_Cart_create(reorder=0) numbers ranks row-major over coords; MPI_Cart_shift(d, +1) gives
 * (coord-1, coord+1) with periodic wrap.  runtime/domain/regular_6d_stencil.cpp:56-82 (setBoundingBox). */
static int po_cart_rank(const int n[3], int c0, int c1, int c2) { return (c0 * n[1] + c1) * n[2] + c2; }

static void po_set_bounding_box(po_sim *s, po_rank *r) {
    int c[3];
    int rem = r->rank;
    c[2] = rem % s->nranks[2]; rem /= s->nranks[2];
    c[1] = rem % s->nranks[1]; rem /= s->nranks[1];
    c[0] = rem;
    for(int d = 0; d < 3; d++) {
        const double rank_length = (s->grid[d * 2 + 1] - s->grid[d * 2]) / (double) s->nranks[d];
        int cp[3] = {c[0], c[1], c[2]}, cn[3] = {c[0], c[1], c[2]};
        cp[d] = (c[d] - 1 + s->nranks[d]) % s->nranks[d];
        cn[d] = (c[d] + 1) % s->nranks[d];
        r->coords[d] = c[d];
        r->neighbor_ranks[d * 2 + 0] = po_cart_rank(s->nranks, cp[0], cp[1], cp[2]);
        r->neighbor_ranks[d * 2 + 1] = po_cart_rank(s->nranks, cn[0], cn[1], cn[2]);
        r->pbc[d * 2 + 0] = (c[d] == 0) ? 1 : 0;
        r->pbc[d * 2 + 1] = (c[d] == s->nranks[d] - 1) ? -1 : 0;
        r->subdom[d * 2 + 0] = s->grid[d * 2] + rank_length * (double) c[d];
        r->subdom[d * 2 + 1] = r->subdom[d * 2 + 0] + rank_length;
    }
}

/* runtime/domain/regular_6d_stencil.cpp:96-100 (isWithinSubdomain), SMALL = 1e-5 (.hpp:6) */
static int po_within_subdomain(const po_rank *r, double x, double y, double z) {
    return x >= r->subdom[0] && x < r->subdom[1] - 0.00001 &&
           y >= r->subdom[2] && y < r->subdom[3] - 0.00001 &&
           z >= r->subdom[4] && z < r->subdom[5] - 0.00001;
}

po_sim *po_create(int world_size, const double grid[6], const int part_flags[3], const int pbc_flag[3],
                  int particle_capacity, int neighbor_capacity, int cell_capacity, int send_capacity) {
    po_sim *s = (po_sim *) calloc(1, sizeof(po_sim));
    s->world = world_size;
    memcpy(s->grid, grid, sizeof(double) * 6);
    memcpy(s->part_flags, part_flags, sizeof(int) * 3);
    memcpy(s->pbc_flag, pbc_flag, sizeof(int) * 3);
    po_set_config(world_size, s->grid, s->part_flags, s->nranks);
    s->r = (po_rank *) calloc((size_t) world_size, sizeof(po_rank));
    for(int k = 0; k < world_size; k++) {
        po_rank *r = &s->r[k];
        r->rank = k;
        po_set_bounding_box(s, r);
        r->pcap = particle_capacity;
        r->uid = (int *) calloc((size_t) particle_capacity, sizeof(int));
        r->shape = (int *) calloc((size_t) particle_capacity, sizeof(int));
        r->flags = (int *) calloc((size_t) particle_capacity, sizeof(int));
        r->type = (int *) calloc((size_t) particle_capacity, sizeof(int));
        r->position = (double *) calloc((size_t) particle_capacity * 3, sizeof(double));
        r->mass = (double *) calloc((size_t) particle_capacity, sizeof(double));
        r->velocity = (double *) calloc((size_t) particle_capacity * 3, sizeof(double));
        r->force = (double *) calloc((size_t) particle_capacity * 3, sizeof(double));
        r->particle_cell = (int *) calloc((size_t) particle_capacity, sizeof(int));
        r->neighbor_capacity = neighbor_capacity;
        r->neighborlists = (int *) calloc((size_t) particle_capacity * (size_t) neighbor_capacity, sizeof(int));
        r->numneighs = (int *) calloc((size_t) particle_capacity, sizeof(int));
        r->cell_capacity = cell_capacity;
        r->send_capacity = send_capacity;
        r->send_map = (int *) calloc((size_t) send_capacity, sizeof(int));
        r->send_mult = (int *) calloc((size_t) send_capacity * 3, sizeof(int));
        r->exchg_flag = (int *) calloc((size_t) particle_capacity, sizeof(int));
        r->exchg_copy_to = (int *) calloc((size_t) send_capacity, sizeof(int));
        r->send_buffer = (double *) calloc((size_t) send_capacity * 11, sizeof(double));
        r->recv_buffer = (double *) calloc((size_t) send_capacity * 11, sizeof(double));
    }
    return s;
}

void po_destroy(po_sim *s) {
    for(int k = 0; k < s->world; k++) {
        po_rank *r = &s->r[k];
        free(r->uid); free(r->shape); free(r->flags); free(r->type); free(r->position); free(r->mass);
        free(r->velocity); free(r->force); free(r->particle_cell); free(r->neighborlists); free(r->numneighs);
        free(r->send_map); free(r->send_mult); free(r->exchg_flag); free(r->exchg_copy_to);
        free(r->send_buffer); free(r->recv_buffer); free(r->cell_particles); free(r->cell_sizes); free(r->nshapes);
    }
    free(s->r);
    free(s);
}

void po_set_params(po_sim *s, double cell_spacing, double cutoff_lists, double cutoff_force, double dt, int ntypes,
                   const double *epsilon, const double *sigma6, int reneigh_every, int thermo_every) {
    s->cell_spacing = cell_spacing;
    s->cutoff_lists = cutoff_lists;
    s->cutoff_force = cutoff_force;
    s->dt = dt;
    s->ntypes = ntypes;
    memcpy(s->epsilon, epsilon, sizeof(double) * (size_t) (ntypes * ntypes));
    memcpy(s->sigma6, sigma6, sizeof(double) * (size_t) (ntypes * ntypes));
    s->reneigh_every = reneigh_every;
    s->thermo_every = thermo_every;
}

void po_set_compute_half(po_sim *s, int on) { s->compute_half = on; }

/* ---- set-up: runtime/copper_fcc_lattice.hpp:18-26 (Park-Miller RNG) and :64-145 (lattice walk) ---- */
static double po_myrandom(int *seed) {
    int k = (*seed) / 127773;
    double ans;
    *seed = 16807 * (*seed - k * 127773) - 2836 * k;
    if(*seed < 0) { *seed += 2147483647; }
    ans = (1.0 / 2147483647) * (*seed);
    return ans;
}

static int po_imax(int a, int b) { return a > b ? a : b; }
static int po_imin(int a, int b) { return a < b ? a : b; }

/* Every reference rank is its own process with glibc rand() in its initial state (seed 1), hence srand(1) per rank. */
void po_copper_fcc_lattice(po_sim *s, int nx, int ny, int nz, double rho, int ntypes) {
    for(int kr = 0; kr < s->world; kr++) {
        po_rank *r = &s->r[kr];
        const double xlo = 0.0, xhi = s->grid[1], ylo = 0.0, yhi = s->grid[3], zlo = 0.0, zhi = s->grid[5];
        int natoms = 0;
        double alat = pow((4.0 / rho), (1.0 / 3.0));
        int ilo = (int) (xlo / (0.5 * alat) - 1);
        int ihi = (int) (xhi / (0.5 * alat) + 1);
        int jlo = (int) (ylo / (0.5 * alat) - 1);
        int jhi = (int) (yhi / (0.5 * alat) + 1);
        int klo = (int) (zlo / (0.5 * alat) - 1);
        int khi = (int) (zhi / (0.5 * alat) + 1);
        ilo = po_imax(ilo, 0); ihi = po_imin(ihi, 2 * nx - 1);
        jlo = po_imax(jlo, 0); jhi = po_imin(jhi, 2 * ny - 1);
        klo = po_imax(klo, 0); khi = po_imin(khi, 2 * nz - 1);
        int sx = 0, sy = 0, sz = 0, ox = 0, oy = 0, oz = 0;
        const int subboxdim = 8;
        srand(1);
        while(oz * subboxdim <= khi) {
            const int k = oz * subboxdim + sz;
            const int j = oy * subboxdim + sy;
            const int i = ox * subboxdim + sx;
            if(((i + j + k) % 2 == 0) && (i >= ilo) && (i <= ihi) && (j >= jlo) && (j <= jhi) && (k >= klo) && (k <= khi)) {
                const double xtmp = 0.5 * alat * i;
                const double ytmp = 0.5 * alat * j;
                const double ztmp = 0.5 * alat * k;
                if(po_within_subdomain(r, xtmp, ytmp, ztmp)) {
                    int n = k * (2 * ny) * (2 * nx) + j * (2 * nx) + i + 1;
                    for(int m = 0; m < 5; m++) { po_myrandom(&n); }
                    const double vxtmp = po_myrandom(&n);
                    for(int m = 0; m < 5; m++) { po_myrandom(&n); }
                    const double vytmp = po_myrandom(&n);
                    for(int m = 0; m < 5; m++) { po_myrandom(&n); }
                    const double vztmp = po_myrandom(&n);
                    r->mass[natoms] = 1.0;
                    r->position[natoms * 3 + 0] = xtmp;
                    r->position[natoms * 3 + 1] = ytmp;
                    r->position[natoms * 3 + 2] = ztmp;
                    r->velocity[natoms * 3 + 0] = vxtmp;
                    r->velocity[natoms * 3 + 1] = vytmp;
                    r->velocity[natoms * 3 + 2] = vztmp;
                    r->type[natoms] = rand() % ntypes;
                    r->flags[natoms] = 0;
                    r->shape[natoms] = PO_SHAPE_POINTMASS;
                    natoms++;
                }
            }
            sx++;
            if(sx == subboxdim) { sx = 0; sy++; }
            if(sy == subboxdim) { sy = 0; sz++; }
            if(sz == subboxdim) { sz = 0; ox++; }
            if(ox * subboxdim > ihi) { ox = 0; oy++; }
            if(oy * subboxdim > jhi) { oy = 0; oz++; }
        }
        r->nlocal = natoms;
        r->nghost = 0;
    }
}

/* runtime/thermo.hpp:11-51.  Per-rank partial sums in ascending index order, combined in rank order
 * (MPI_Allreduce's combination order is implementation-defined; R = 1 has no combination). */
double po_compute_thermo(po_sim *s, double *pressure) {
    int natoms = 0;
    double t = 0.0;
    for(int kr = 0; kr < s->world; kr++) {
        po_rank *r = &s->r[kr];
        double tr = 0.0;
        natoms += r->nlocal;
        for(int i = 0; i < r->nlocal; i++) {
            const double *v = &r->velocity[i * 3];
            tr += r->mass[i] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        }
        t = (s->world > 1) ? (t + tr) : tr;
    }
    const double xprd = s->grid[1] - s->grid[0], yprd = s->grid[3] - s->grid[2], zprd = s->grid[5] - s->grid[4];
    const double mvv2e = 1.0;
    const double dof_boltz = (natoms * 3 - 3);
    const double t_scale = mvv2e / dof_boltz;
    const double p_scale = 1.0 / 3 / xprd / yprd / zprd;
    t = t * t_scale;
    if(pressure != NULL) { *pressure = (t * dof_boltz) * p_scale; }
    return t;
}

/* runtime/thermo.hpp:53-97 */
void po_adjust_thermo(po_sim *s, double temp) {
    double vxtot = 0.0, vytot = 0.0, vztot = 0.0;
    int natoms = 0;
    for(int kr = 0; kr < s->world; kr++) {
        po_rank *r = &s->r[kr];
        double vx = 0.0, vy = 0.0, vz = 0.0;
        for(int i = 0; i < r->nlocal; i++) {
            vx += r->velocity[i * 3 + 0];
            vy += r->velocity[i * 3 + 1];
            vz += r->velocity[i * 3 + 2];
        }
        natoms += r->nlocal;
        if(s->world > 1) { vxtot += vx; vytot += vy; vztot += vz; } else { vxtot = vx; vytot = vy; vztot = vz; }
    }
    vxtot /= natoms; vytot /= natoms; vztot /= natoms;
    for(int kr = 0; kr < s->world; kr++) {
        po_rank *r = &s->r[kr];
        for(int i = 0; i < r->nlocal; i++) {
            r->velocity[i * 3 + 0] -= vxtot;
            r->velocity[i * 3 + 1] -= vytot;
            r->velocity[i * 3 + 2] -= vztot;
        }
    }
    const double t = po_compute_thermo(s, NULL);
    const double factor = sqrt(temp / t);
    for(int kr = 0; kr < s->world; kr++) {
        po_rank *r = &s->r[kr];
        for(int i = 0; i < r->nlocal; i++) {
            r->velocity[i * 3 + 0] *= factor;
            r->velocity[i * 3 + 1] *= factor;
            r->velocity[i * 3 + 2] *= factor;
        }
    }
}

/* ---- cell lists ---- */
/* sim/cell_lists.py:46-87 (BuildCellListsStencil): dim = ceil(((max+s)-(min-s))/s)+1, ncells = prod+1,
 * stencil offsets in x-slowest / z-fastest order. */
void po_build_cell_lists_stencil(po_sim *s, po_rank *r) {
    const double sp = s->cell_spacing;
    for(int d = 0; d < 3; d++) {
        const double hi = r->subdom[d * 2 + 1] + sp;
        const double lo = r->subdom[d * 2 + 0] - sp;
        const double len = hi - lo;
        const double q = len / sp;
        r->dim_cells[d] = ((int) ceil(q)) + 1;
    }
    r->ncells = r->dim_cells[0] * r->dim_cells[1] * r->dim_cells[2] + 1;
    r->nstencil = 0;
    for(int i = -1; i < 2; i++) {
        for(int j = -1; j < 2; j++) {
            const int a = (i * r->dim_cells[1] + j) * r->dim_cells[2];
            for(int k = -1; k < 2; k++) { r->stencil[r->nstencil++] = a + k; }
        }
    }
    free(r->cell_particles); free(r->cell_sizes); free(r->nshapes);
    r->cell_particles = (int *) calloc((size_t) r->ncells * (size_t) r->cell_capacity, sizeof(int));
    r->cell_sizes = (int *) calloc((size_t) r->ncells, sizeof(int));
    r->nshapes = (int *) calloc((size_t) r->ncells, sizeof(int));
}

/* sim/cell_lists.py:90-134 (BuildCellLists).  Returns needed capacity (>0) on overflow, like resizes[0]
 * (transformations/modules.py:70-126, runtime/devices/device.hpp:63-71). */
int po_cell_index(const po_sim *s, const po_rank *r, const double *x, int flags) {
    if(flags & PO_FLAG_INFINITE) { return 0; }
    int flat = 0;
    for(int d = 0; d < 3; d++) {
        const double lo = r->subdom[d * 2] - s->cell_spacing;
        const double rel = x[d] - lo;
        const double q = rel / s->cell_spacing;
        int c = ((int) q >= 0) ? (int) q : 0;
        c = (c < r->dim_cells[d]) ? c : r->dim_cells[d] - 1;
        flat = (d == 0) ? c : flat * r->dim_cells[d] + c;
    }
    return flat + 1;
}

int po_build_cell_lists(po_sim *s, po_rank *r) {
    int resize = 0;
    for(int c = 0; c < r->ncells; c++) { r->cell_sizes[c] = 0; }
    for(int i = 0; i < r->nlocal + r->nghost; i++) {
        const int cell = po_cell_index(s, r, &r->position[i * 3], r->flags[i]);
        if(cell >= 0 && cell < r->ncells) {
            r->particle_cell[i] = cell;
            /* atomic_add_resize_check, runtime/devices/device.hpp:63-71 */
            const int slot = r->cell_sizes[cell];
            if(slot + 1 >= r->cell_capacity) {
                resize = slot + 1;               /* counter is NOT advanced; the module is re-run after the grow */
            } else {
                r->cell_sizes[cell] = slot + 1;
            }
            r->cell_particles[cell * r->cell_capacity + slot] = i;
        }
    }
    return resize;
}

/* sim/cell_lists.py:137-171 (PartitionCellLists) with one shape (point mass): two-pointer partition. */
void po_partition_cell_lists(po_rank *r, int shape_id) {
    for(int c = 0; c < r->ncells; c++) {
        int start = 0, end = r->cell_sizes[c] - 1;
        int *cp = &r->cell_particles[c * r->cell_capacity];
        const int shape_start = start;
        r->nshapes[c] = 0;
        while(start <= end) {
            const int p = cp[start];
            if(r->shape[p] != shape_id) {
                if(start != end) { cp[start] = cp[end]; cp[end] = p; }
                end--;
            } else {
                start++;
                r->nshapes[c] = start - shape_start;
            }
        }
    }
}

/* sim/neighbor_lists.py:21-48 + sim/interaction.py:91-120: cell 0 first, then the 27 stencil cells with
 * 0 < cell < ncells; j != i; r2 < cutoff^2 with (dx*dx + dy*dy) + dz*dz.  Returns needed capacity on overflow. */
int po_build_neighbor_lists(po_sim *s, po_rank *r) {
    int resize = 0;
    const double cutsq = s->cutoff_lists * s->cutoff_lists;
    for(int i = 0; i < r->nlocal; i++) { r->numneighs[i] = 0; }
    for(int i = 0; i < r->nlocal; i++) {
        if((r->flags[i] & PO_FLAG_FIXED) != 0) { continue; }
        const int pc = r->particle_cell[i];
        for(int k = -1; k < r->nstencil; k++) {
            const int cell = (k < 0) ? 0 : pc + r->stencil[k];
            if(!(k < 0 || (cell > 0 && cell < r->ncells))) { continue; }
            const int ns = r->nshapes[cell];
            const double xi = r->position[i * 3], yi = r->position[i * 3 + 1], zi = r->position[i * 3 + 2];
            for(int m = 0; m < ns; m++) {
                const int j = r->cell_particles[cell * r->cell_capacity + m];
                if(s->compute_half) { /* sim/interaction.py:107-113: shape[j] > shape_i || (shape[j] == shape_i && i < j) */
                    if(!(r->shape[j] > r->shape[i] || (r->shape[j] == r->shape[i] && i < j))) { continue; }
                } else if(j == i) {
                    continue;
                }
                const double dx = xi - r->position[j * 3];
                const double dy = yi - r->position[j * 3 + 1];
                const double dz = zi - r->position[j * 3 + 2];
                const double a = dx * dx;
                const double b = dy * dy;
                const double ab = a + b;
                const double c = dz * dz;
                const double rsq = ab + c;
                if(rsq < cutsq) {
                    const int n = r->numneighs[i];
                    r->neighborlists[(size_t) i * r->neighbor_capacity + n] = j;
                    if(n + 2 >= r->neighbor_capacity) {
                        resize = n + 1;
                    } else {
                        r->numneighs[i] = n + 1;
                    }
                }
            }
        }
    }
    return resize;
}

/* ---- kernels ---- */
/* examples/md.py:5-8 + sim/interaction.py:201-292; operation order as the generator emits it. */
void po_lennard_jones(po_sim *s, po_rank *r) {
    const double cutsq = s->cutoff_force * s->cutoff_force;
    for(int i = 0; i < r->nlocal; i++) {
        if((r->flags[i] & PO_FLAG_FIXED) != 0) { continue; }
        double fx = 0.0, fy = 0.0, fz = 0.0;
        const double xi = r->position[i * 3], yi = r->position[i * 3 + 1], zi = r->position[i * 3 + 2];
        const int ti = r->type[i] * s->ntypes;
        for(int k = 0; k < r->numneighs[i]; k++) {
            const int j = r->neighborlists[(size_t) i * r->neighbor_capacity + k];
            const double dx = xi - r->position[j * 3];
            const double dy = yi - r->position[j * 3 + 1];
            const double dz = zi - r->position[j * 3 + 2];
            const double a = dx * dx;
            const double b = dy * dy;
            const double ab = a + b;
            const double c = dz * dz;
            const double rsq = ab + c;
            if(rsq < cutsq) {
                const double sr2 = 1.0 / rsq;
                const double sr4 = sr2 * sr2;
                const double sr6a = sr4 * sr2;
                const double sr6 = sr6a * s->sigma6[ti + r->type[j]];
                const double f48 = 48.0 * sr6;
                const double m05 = sr6 - 0.5;
                const double p = f48 * m05;
                const double q = p * sr2;
                const double f = q * s->epsilon[ti + r->type[j]];
                fx = fx + dx * f;
                fy = fy + dy * f;
                fz = fz + dz * f;
                /* ir/apply.py:111-125: the partner of a half-list pair gets the opposite term, unless it is a ghost or FIXED
                 * (atomic_add in the generated code; the serial build performs them in loop order) */
                if(s->compute_half && j < r->nlocal && (r->flags[j] & PO_FLAG_FIXED) == 0) {
                    r->force[j * 3 + 0] += -(dx * f);
                    r->force[j * 3 + 1] += -(dy * f);
                    r->force[j * 3 + 2] += -(dz * f);
                }
            }
        }
        r->force[i * 3 + 0] = r->force[i * 3 + 0] + fx;
        r->force[i * 3 + 1] = r->force[i * 3 + 1] + fy;
        r->force[i * 3 + 2] = r->force[i * 3 + 2] + fz;
    }
}

/* examples/md.py:11-13: v += ((dt*0.5)*f)/m ; x += dt*v */
void po_initial_integrate(po_sim *s, po_rank *r) {
    const double hdt = s->dt * 0.5;
    for(int i = 0; i < r->nlocal; i++) {
        if((r->flags[i] & PO_FLAG_FIXED) != 0) { continue; }
        for(int d = 0; d < 3; d++) {
            const double t = hdt * r->force[i * 3 + d];
            const double u = t / r->mass[i];
            r->velocity[i * 3 + d] = r->velocity[i * 3 + d] + u;
        }
        for(int d = 0; d < 3; d++) {
            const double t = s->dt * r->velocity[i * 3 + d];
            r->position[i * 3 + d] = r->position[i * 3 + d] + t;
        }
    }
}

/* examples/md.py:16-17 */
void po_final_integrate(po_sim *s, po_rank *r) {
    const double hdt = s->dt * 0.5;
    for(int i = 0; i < r->nlocal; i++) {
        if((r->flags[i] & PO_FLAG_FIXED) != 0) { continue; }
        for(int d = 0; d < 3; d++) {
            const double t = hdt * r->force[i * 3 + d];
            const double u = t / r->mass[i];
            r->velocity[i * 3 + d] = r->velocity[i * 3 + d] + u;
        }
    }
}

/* sim/properties.py:61-70 */
void po_reset_volatile(po_rank *r) {
    for(int i = 0; i < r->nlocal; i++) { r->force[i * 3] = 0.0; r->force[i * 3 + 1] = 0.0; r->force[i * 3 + 2] = 0.0; }
}

/* ---- communication ---- */
/* sim/comm.py:223-259 + sim/domain_partitioning.py:31-65.  `offset` = 0 for exchange, cell spacing for
 * borders.  Side 0 ("pos < min+offset", goes to prev) is scanned completely before side 1. */
static void po_determine(po_sim *s, po_rank *r, int dim, double offset, int is_exchange) {
    if(is_exchange) { for(int i = 0; i < r->nlocal; i++) { r->exchg_flag[i] = 0; } }
    for(int side = 0; side < 2; side++) {
        const int j = dim * 2 + side;
        if(!s->pbc_flag[dim] && r->pbc[j] != 0) { continue; }
        for(int i = 0; i < r->nlocal + r->nghost; i++) {
            if((r->flags[i] & (PO_FLAG_INFINITE | PO_FLAG_GLOBAL)) != 0) { continue; }
            const double x = r->position[i * 3 + dim];
            const int hit = (side == 0) ? (x < r->subdom[j] + offset) : (x > r->subdom[j] - offset);
            if(hit) {
                const int idx = r->nsend_all++;
                if(idx >= r->send_capacity) { fprintf(stderr, "pairs_oracle: send_capacity exceeded\n"); abort(); }
                r->send_map[idx] = i;
                if(is_exchange) { r->exchg_flag[i] = 1; }
                for(int d = 0; d < 3; d++) { r->send_mult[idx * 3 + d] = (d == dim) ? r->pbc[j] : 0; }
                r->nsend[j]++;
            }
        }
    }
}

/* runtime/domain/regular_6d_stencil.cpp:113-127 (communicateSizes) over all ranks */
static void po_communicate_sizes(po_sim *s, int dim) {
    for(int kr = 0; kr < s->world; kr++) {
        po_rank *r = &s->r[kr];
        po_rank *prev = &s->r[r->neighbor_ranks[dim * 2 + 0]];
        po_rank *next = &s->r[r->neighbor_ranks[dim * 2 + 1]];
        prev->nrecv[dim * 2 + 0] = r->nsend[dim * 2 + 0]; /* what I send to prev arrives in its "from next" slot */
        next->nrecv[dim * 2 + 1] = r->nsend[dim * 2 + 1];
    }
}

/* sim/comm.py:262-288 (SetCommunicationOffsets) */
static void po_set_offsets(po_rank *r, int step) {
    int isend = 0, irecv = 0;
    for(int i = 0; i < step; i++) {
        for(int j = i * 2; j < i * 2 + 2; j++) { isend += r->nsend[j]; irecv += r->nrecv[j]; }
    }
    for(int j = step * 2; j < step * 2 + 2; j++) {
        r->send_offsets[j] = isend; r->recv_offsets[j] = irecv;
        isend += r->nsend[j]; irecv += r->nrecv[j];
    }
}

/* runtime/domain/regular_6d_stencil.cpp:129-181 (communicateData) over all ranks */
static void po_communicate_data(po_sim *s, int dim, int elem) {
    for(int kr = 0; kr < s->world; kr++) {
        po_rank *r = &s->r[kr];
        for(int side = 0; side < 2; side++) {
            const int j = dim * 2 + side;
            po_rank *dst = &s->r[r->neighbor_ranks[j]];
            memcpy(&dst->recv_buffer[(size_t) dst->recv_offsets[j] * elem], &r->send_buffer[(size_t) r->send_offsets[j] * elem],
                   sizeof(double) * (size_t) r->nsend[j] * elem);
        }
    }
}

static double po_len(const po_sim *s, int d) { return s->grid[d * 2 + 1] - s->grid[d * 2]; }

/* sim/comm.py:291-330 (PackGhostParticles): exchange list = uid,shape,flags,position,mass,linear_velocity,type
 * (11 doubles); borders list = uid,type,mass,position,linear_velocity,shape (10 doubles).  Appendix A.1. */
static void po_pack(po_sim *s, po_rank *r, int dim, int is_exchange) {
    const int elem = is_exchange ? 11 : 10;
    const int beg = r->send_offsets[dim * 2];
    const int end = beg + r->nsend[dim * 2] + r->nsend[dim * 2 + 1];
    for(int i = beg; i < end; i++) {
        double *b = &r->send_buffer[(size_t) i * elem];
        const int p = r->send_map[i];
        double x[3];
        for(int d = 0; d < 3; d++) {
            const double sh = r->send_mult[i * 3 + d] * po_len(s, d);
            x[d] = r->position[p * 3 + d] + sh;
        }
        if(is_exchange) {
            b[0] = (double) r->uid[p]; b[1] = (double) r->shape[p]; b[2] = (double) r->flags[p];
            b[3] = x[0]; b[4] = x[1]; b[5] = x[2]; b[6] = r->mass[p];
            b[7] = r->velocity[p * 3]; b[8] = r->velocity[p * 3 + 1]; b[9] = r->velocity[p * 3 + 2];
            b[10] = (double) r->type[p];
        } else {
            b[0] = (double) r->uid[p]; b[1] = (double) r->type[p]; b[2] = r->mass[p];
            b[3] = x[0]; b[4] = x[1]; b[5] = x[2];
            b[6] = r->velocity[p * 3]; b[7] = r->velocity[p * 3 + 1]; b[8] = r->velocity[p * 3 + 2];
            b[9] = (double) r->shape[p];
        }
    }
}

/* sim/comm.py:333-366 (UnpackGhostParticles): append at nlocal + i.  Ghost `flags` are never transmitted by
 * borders (Appendix A.1): the slot keeps whatever value it had. */
static void po_unpack(po_rank *r, int dim, int is_exchange) {
    const int elem = is_exchange ? 11 : 10;
    const int beg = r->recv_offsets[dim * 2];
    const int end = beg + r->nrecv[dim * 2] + r->nrecv[dim * 2 + 1];
    for(int i = beg; i < end; i++) {
        const double *b = &r->recv_buffer[(size_t) i * elem];
        const int p = r->nlocal + i;
        if(p >= r->pcap) { fprintf(stderr, "pairs_oracle: particle_capacity exceeded\n"); abort(); }
        if(is_exchange) {
            r->uid[p] = (int) b[0]; r->shape[p] = (int) b[1]; r->flags[p] = (int) b[2];
            r->position[p * 3] = b[3]; r->position[p * 3 + 1] = b[4]; r->position[p * 3 + 2] = b[5];
            r->mass[p] = b[6];
            r->velocity[p * 3] = b[7]; r->velocity[p * 3 + 1] = b[8]; r->velocity[p * 3 + 2] = b[9];
            r->type[p] = (int) b[10];
        } else {
            r->uid[p] = (int) b[0]; r->type[p] = (int) b[1]; r->mass[p] = b[2];
            r->position[p * 3] = b[3]; r->position[p * 3 + 1] = b[4]; r->position[p * 3 + 2] = b[5];
            r->velocity[p * 3] = b[6]; r->velocity[p * 3 + 1] = b[7]; r->velocity[p * 3 + 2] = b[8];
            r->shape[p] = (int) b[9];
        }
    }
}

/* sim/comm.py:445-466 (pt1, host) and :469-506 (pt2): fill holes left by leavers from the tail.
 * `copy_to > 0` (not >= 0) is the reference's filter (comm.py:481). */
static void po_remove_exchanged(po_rank *r) {
    int tail = r->nlocal - 1;
    for(int i = 0; i < r->nsend_all; i++) {
        if(r->send_map[i] < r->nlocal - r->nsend_all) {
            while(r->exchg_flag[tail] == 1) { tail--; }
            r->exchg_copy_to[i] = tail;
            tail--;
        } else {
            r->exchg_copy_to[i] = -1;
        }
    }
    for(int i = 0; i < r->nsend_all; i++) {
        const int src = r->exchg_copy_to[i];
        if(src > 0) {
            const int dst = r->send_map[i];
            r->uid[dst] = r->uid[src]; r->shape[dst] = r->shape[src]; r->flags[dst] = r->flags[src];
            for(int d = 0; d < 3; d++) { r->position[dst * 3 + d] = r->position[src * 3 + d]; }
            r->mass[dst] = r->mass[src];
            for(int d = 0; d < 3; d++) { r->velocity[dst * 3 + d] = r->velocity[src * 3 + d]; }
            r->type[dst] = r->type[src];
        }
    }
    r->nlocal = r->nlocal - r->nsend_all;
}

/* sim/comm.py:100-151 (Comm.exchange) */
void po_exchange(po_sim *s) {
    for(int step = 0; step < 3; step++) {
        for(int kr = 0; kr < s->world; kr++) {
            po_rank *r = &s->r[kr];
            r->nsend_all = 0; r->nghost = 0;
            for(int j = 0; j < (step + 1) * 2; j++) { r->nsend[j] = 0; r->nrecv[j] = 0; r->send_offsets[j] = 0; r->recv_offsets[j] = 0; }
            po_determine(s, r, step, 0.0, 1);
        }
        po_communicate_sizes(s, step);
        for(int kr = 0; kr < s->world; kr++) {
            po_rank *r = &s->r[kr];
            po_set_offsets(r, step);
            po_pack(s, r, step, 1);
            po_remove_exchanged(r);
        }
        po_communicate_data(s, step, 11);
        for(int kr = 0; kr < s->world; kr++) {
            po_rank *r = &s->r[kr];
            po_unpack(r, step, 1);
            r->nlocal = r->nlocal + (r->nrecv[step * 2] + r->nrecv[step * 2 + 1]); /* ChangeSizeAfterExchange */
        }
    }
}

/* sim/comm.py:56-98 (Comm.borders) */
void po_borders(po_sim *s) {
    for(int kr = 0; kr < s->world; kr++) { s->r[kr].nsend_all = 0; s->r[kr].nghost = 0; }
    for(int step = 0; step < 3; step++) {
        for(int kr = 0; kr < s->world; kr++) {
            po_rank *r = &s->r[kr];
            r->nsend[step * 2] = 0; r->nsend[step * 2 + 1] = 0; r->nrecv[step * 2] = 0; r->nrecv[step * 2 + 1] = 0;
            po_determine(s, r, step, s->cell_spacing, 0);
        }
        po_communicate_sizes(s, step);
        for(int kr = 0; kr < s->world; kr++) {
            po_rank *r = &s->r[kr];
            po_set_offsets(r, step);
            po_pack(s, r, step, 0);
        }
        po_communicate_data(s, step, 10);
        for(int kr = 0; kr < s->world; kr++) {
            po_rank *r = &s->r[kr];
            po_unpack(r, step, 0);
            r->nghost = r->nghost + (r->nrecv[step * 2] + r->nrecv[step * 2 + 1]);
        }
    }
}

/* sim/comm.py:45-54, 369-442 (Comm.synchronize): ALL send entries are packed from the current arrays in one
 * pass, then moved, then unpacked -- so a ghost that is itself forwarded (edge/corner images) carries the
 * position its source ghost had BEFORE this refresh (one step stale per forwarding level). */
void po_synchronize(po_sim *s) {
    for(int kr = 0; kr < s->world; kr++) {
        po_rank *r = &s->r[kr];
        for(int i = 0; i < r->nsend_all; i++) {
            double *b = &r->send_buffer[(size_t) i * 6];
            const int p = r->send_map[i];
            for(int d = 0; d < 3; d++) {
                const double sh = r->send_mult[i * 3 + d] * po_len(s, d);
                b[d] = r->position[p * 3 + d] + sh;
            }
            b[3] = r->velocity[p * 3]; b[4] = r->velocity[p * 3 + 1]; b[5] = r->velocity[p * 3 + 2];
        }
    }
    for(int d = 0; d < 3; d++) { po_communicate_data(s, d, 6); }
    for(int kr = 0; kr < s->world; kr++) {
        po_rank *r = &s->r[kr];
        int n = 0;
        for(int j = 0; j < 6; j++) { n += r->nrecv[j]; }
        for(int i = 0; i < n; i++) {
            const double *b = &r->recv_buffer[(size_t) i * 6];
            const int p = r->nlocal + i;
            r->position[p * 3] = b[0]; r->position[p * 3 + 1] = b[1]; r->position[p * 3 + 2] = b[2];
            r->velocity[p * 3] = b[3]; r->velocity[p * 3 + 1] = b[4]; r->velocity[p * 3 + 2] = b[5];
        }
    }
}

static void po_grow_cells(po_rank *r, int needed) {
    r->cell_capacity = needed * 2; /* transformations/modules.py:159-203: capacity = 2 * resizes[k] */
    free(r->cell_particles);
    r->cell_particles = (int *) calloc((size_t) r->ncells * (size_t) r->cell_capacity, sizeof(int));
}

static void po_grow_neigh(po_rank *r, int needed) {
    r->neighbor_capacity = needed * 2;
    free(r->neighborlists);
    r->neighborlists = (int *) calloc((size_t) r->pcap * (size_t) r->neighbor_capacity, sizeof(int));
}

void po_setup_cells(po_sim *s) {
    for(int kr = 0; kr < s->world; kr++) { po_build_cell_lists_stencil(s, &s->r[kr]); }
}

/* One iteration `ts` of the generated timestep loop: sim/simulation.py:387-417 (order of the per-step
 * procedure list) and sim/timestep.py:36-61 (guards: every -> ((ts+1)%n==0)||ts==0 ; skip_first -> ts>0). */
void po_md_step(po_sim *s, int ts) {
    const int reneigh = (((ts + 1) % s->reneigh_every) == 0) || (ts == 0);
    if(ts > 0) { for(int kr = 0; kr < s->world; kr++) { po_initial_integrate(s, &s->r[kr]); } }
    if(reneigh) {
        po_exchange(s);
        po_borders(s);
    } else {
        po_synchronize(s);
    }
    for(int kr = 0; kr < s->world; kr++) {
        po_rank *r = &s->r[kr];
        if(reneigh) {
            int need;
            while((need = po_build_cell_lists(s, r)) > 0) { po_grow_cells(r, need); }
            po_partition_cell_lists(r, PO_SHAPE_POINTMASS);
            while((need = po_build_neighbor_lists(s, r)) > 0) { po_grow_neigh(r, need); }
        }
        po_reset_volatile(r);
        po_lennard_jones(s, r);
        if(ts > 0) { po_final_integrate(s, r); }
    }
}

/* ---- accessors for ctypes ---- */
po_rank *po_rank_ptr(po_sim *s, int k) { return &s->r[k]; }
int po_world(po_sim *s) { return s->world; }
void po_get_nranks(po_sim *s, int *out) { out[0] = s->nranks[0]; out[1] = s->nranks[1]; out[2] = s->nranks[2]; }
int po_nlocal(po_rank *r) { return r->nlocal; }
int po_nghost(po_rank *r) { return r->nghost; }
int po_ncells(po_rank *r) { return r->ncells; }
int po_neighbor_capacity(po_rank *r) { return r->neighbor_capacity; }
int po_cell_capacity(po_rank *r) { return r->cell_capacity; }
void po_get_decomposition(po_rank *r, int *neighbor_ranks, int *pbc, double *subdom, int *dim_cells) {
    memcpy(neighbor_ranks, r->neighbor_ranks, sizeof(int) * 6);
    memcpy(pbc, r->pbc, sizeof(int) * 6);
    memcpy(subdom, r->subdom, sizeof(double) * 6);
    memcpy(dim_cells, r->dim_cells, sizeof(int) * 3);
}
int *po_int_array(po_rank *r, const char *name) {
    if(!strcmp(name, "uid")) return r->uid;
    if(!strcmp(name, "shape")) return r->shape;
    if(!strcmp(name, "flags")) return r->flags;
    if(!strcmp(name, "type")) return r->type;
    if(!strcmp(name, "particle_cell")) return r->particle_cell;
    if(!strcmp(name, "numneighs")) return r->numneighs;
    if(!strcmp(name, "neighborlists")) return r->neighborlists;
    if(!strcmp(name, "cell_sizes")) return r->cell_sizes;
    if(!strcmp(name, "cell_particles")) return r->cell_particles;
    if(!strcmp(name, "send_map")) return r->send_map;
    if(!strcmp(name, "send_mult")) return r->send_mult;
    if(!strcmp(name, "stencil")) return r->stencil;
    return NULL;
}
double *po_real_array(po_rank *r, const char *name) {
    if(!strcmp(name, "position")) return r->position;
    if(!strcmp(name, "mass")) return r->mass;
    if(!strcmp(name, "linear_velocity")) return r->velocity;
    if(!strcmp(name, "force")) return r->force;
    return NULL;
}
void po_set_counts(po_rank *r, int nlocal, int nghost) { r->nlocal = nlocal; r->nghost = nghost; }
int po_nsend_all(po_rank *r) { return r->nsend_all; }
