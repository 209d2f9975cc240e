// CPU build of the PRODUCT's MD arithmetic (pairs_b200/csrc/md_math.h) for unit tests against the reference's generated modules:
// the cell index of every particle, and the Lennard-Jones force over caller-supplied neighbour lists in the caller's order
// (AoS host layout of the reference).  Test code only; compile with -ffp-contract=off.
#include "md_math.h"

extern "C" {

void host_md_cell_index(const double *lo, double spacing, const int *dim, int n, const double *pos, const int *flags, int *out) {
    PbCellGeom g;
    for(int d = 0; d < 3; d++) { g.lo[d] = lo[d]; g.dim[d] = dim[d]; }
    g.spacing = spacing;
    g.ncells = dim[0] * dim[1] * dim[2] + 1;
    for(int i = 0; i < n; i++) { out[i] = pb_cell_index(g, pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2], flags[i]); }
}

// same loop shape as pb_k_lennard_jones (thread = particle, accumulate in registers, force[i] = force[i] + acc once)
void host_md_lennard_jones(int nlocal, int neighbor_capacity, const int *numneighs, const int *neighborlists, const int *flags,
                           const double *pos, const int *type, int ntypes, const double *sigma6, const double *epsilon, double cutsq,
                           double *force) {
    for(int i = 0; i < nlocal; i++) {
        if(flags[i] & 4) { continue; }
        double fx = 0.0, fy = 0.0, fz = 0.0;
        for(int k = 0; k < numneighs[i]; k++) {
            const int j = neighborlists[(long) i * neighbor_capacity + k];
            double dx, dy, dz;
            const double rsq = pb_pair_rsq(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2], pos[j * 3], pos[j * 3 + 1], pos[j * 3 + 2], &dx, &dy, &dz);
            if(rsq < cutsq) {
                const int t = type[i] * ntypes + type[j];
                const double f = pb_lj_fpair(rsq, sigma6[t], epsilon[t]);
                fx = fx + dx * f;
                fy = fy + dy * f;
                fz = fz + dz * f;
            }
        }
        force[i * 3 + 0] = force[i * 3 + 0] + fx;
        force[i * 3 + 1] = force[i * 3 + 1] + fy;
        force[i * 3 + 2] = force[i * 3 + 2] + fz;
    }
}

}
