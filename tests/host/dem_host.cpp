// CPU build of the PRODUCT's DEM arithmetic (pairs_b200/csrc/dem_math.h) for unit tests against the reference's generated
// code: brute-force pair sweep in the reference's order (sphere sweep, then half-space sweep; partners in ascending index),
// gravity and euler on arrays in the reference's AoS host layout.  Test code only.
#include <cstring>
#include "dem_math.h"

extern "C" {

void host_dem_params(PbDemParams *P, double dt, double pi, double kappa, double ln, double ct, double rho_p, double rho_f, double g) {
    P->dt = dt; P->c_sum = pi * pi + ln * ln; P->ct2 = ct * ct; P->ct = ct; P->ln_coeff = ln; P->kappa = kappa;
    P->sqrt_kappa = sqrt(kappa); P->grav_coeff = rho_p - rho_f; P->gravity = g; P->pi = pi;
}

// contacts: num[n], c_uid/c_used/c_stick [n][C], c_tsd [n][C][3], c_ivm [n][C]
void host_dem_contacts(const PbDemParams *P, int nlocal, int ntotal, int C, int ntypes, const double *pos, const double *vel,
                       const double *angvel, const double *mass, const double *radius, const double *normal, const int *flags,
                       const int *shape, const int *uid, const int *type, const double *fs, const double *fd, int *num, int *c_uid,
                       int *c_used, int *c_stick, double *c_tsd, double *c_ivm, double *force, double *torque) {
    for(int i = 0; i < nlocal; i++) {
        if(flags[i] & 4) { continue; }
        double acc[2][2][3];
        memset(acc, 0, sizeof(acc));
        const double inv_mi = 1.0 / mass[i];
        for(int sh = 0; sh < 2; sh++) {
            for(int j = 0; j < ntotal; j++) {
                if(j == i || shape[j] != sh) { continue; }
                double n[3], cp[3], delta;
                const int hit = (sh == 0) ? pb_dem_geom_sphere(&pos[i * 3], radius[i], &pos[j * 3], radius[j], n, cp, &delta)
                                          : pb_dem_geom_halfspace(&pos[i * 3], radius[i], &pos[j * 3], &normal[j * 3], n, cp, &delta);
                if(!hit) { continue; }
                int slot = -1;
                for(int c = 0; c < num[i]; c++) { if(c_uid[i * C + c] == uid[j]) { slot = c; } }
                if(slot == -1) {
                    slot = num[i]++;
                    c_uid[i * C + slot] = uid[j];
                    c_stick[i * C + slot] = 0;
                    c_tsd[(i * C + slot) * 3] = c_tsd[(i * C + slot) * 3 + 1] = c_tsd[(i * C + slot) * 3 + 2] = 0.0;
                    c_ivm[i * C + slot] = 0.0;
                }
                c_used[i * C + slot] = 1;
                double F[3], T[3];
                pb_dem_pair_force(*P, &pos[i * 3], &vel[i * 3], &angvel[i * 3], inv_mi, &pos[j * 3], &vel[j * 3], &angvel[j * 3], mass[j], n, cp,
                                  delta, fs[type[i] * ntypes + type[j]], fd[type[i] * ntypes + type[j]], &c_tsd[(i * C + slot) * 3],
                                  &c_ivm[i * C + slot], &c_stick[i * C + slot], F, T);
                for(int d = 0; d < 3; d++) { acc[sh][0][d] = acc[sh][0][d] + F[d]; acc[sh][1][d] = acc[sh][1][d] + T[d]; }
            }
        }
        for(int d = 0; d < 3; d++) {
            force[i * 3 + d] = force[i * 3 + d] + (acc[0][0][d] + acc[1][0][d]);
            torque[i * 3 + d] = torque[i * 3 + d] + (acc[0][1][d] + acc[1][1][d]);
        }
    }
}

void host_dem_gravity(const PbDemParams *P, int nlocal, const int *flags, const double *radius, double *force) {
    for(int i = 0; i < nlocal; i++) {
        if(flags[i] & 4) { continue; }
        force[i * 3 + 2] = pb_dem_gravity(*P, radius[i], force[i * 3 + 2]);
    }
}

void host_dem_euler(const PbDemParams *P, int nlocal, const int *flags, const double *mass, const double *force, const double *torque,
                    const double *Iinv, double *pos, double *vel, double *angvel, double *quat, double *rotmat) {
    for(int i = 0; i < nlocal; i++) {
        if(flags[i] & 4) { continue; }
        pb_dem_euler(P->dt, mass[i], &force[i * 3], &torque[i * 3], &Iinv[i * 9], &pos[i * 3], &vel[i * 3], &angvel[i * 3], &quat[i * 4],
                     &rotmat[i * 9]);
    }
}

void host_dem_sphere_inv_inertia(double mass, double radius, double *Iinv) { pb_dem_sphere_inv_inertia(mass, radius, Iinv); }

}
