// DEM contact model of examples/dem.py as host+device inline functions.
//
// These are the per-pair / per-particle bodies of the reference's DEM kernels:
//   linear_spring_dashpot   examples/dem.py:18-74   (Hooke normal spring + dashpot, tangential spring with Coulomb cap and
//                                                     sticking state, contact-history read/modify/write)
//   contact geometry        sim/interaction.py:234-264 (sphere-sphere, sphere-half-space)
//   euler                   examples/dem.py:77-85   (semi-implicit Euler + quaternion rotation update)
//   gravity                 examples/dem.py:88-90
//   update_mass_and_inertia examples/dem.py:6-15
// written operation by operation in the order the reference's expression trees evaluate (Python precedence, vector
// operations component-wise, dot = (a0*b0 + a1*b1) + a2*b2, keyword lowering of mapping/keywords.py), so that with
// contraction disabled (nvcc --fmad=false / g++ -ffp-contract=off) every pair term is bit-identical to the reference's
// generated code.  The same header is compiled into the CUDA kernels (dem_kernels.cu) and, for CPU-side unit tests
// against the reference, into a host library -- it is product code either way.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD static inline
#endif

struct PbDemParams {
    double dt;
    double c_sum;        // pi*pi + lnDryResCoeff*lnDryResCoeff
    double ct2;          // collisionTime * collisionTime
    double ct;           // collisionTime
    double ln_coeff;     // lnDryResCoeff
    double kappa;
    double sqrt_kappa;   // sqrt(kappa)
    double grav_coeff;   // (densityParticle - densityFluid)
    double gravity;      // gravity_SI
    double pi;
};

PB_HD double pb_dot3(const double *a, const double *b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

PB_HD void pb_cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// ---- contact geometry (sim/interaction.py:234-264) ----------------------------------------------------------------
// sphere i - sphere j.  Returns 1 if the pair passes the cutoff filter r2 < (ri+rj)*((ri+rj)+0) AND skip_when(-pen < 0)
// does not skip it.  Outputs: contact normal n, contact point cp, delta = -penetration_depth.
PB_HD int pb_dem_geom_sphere(const double *xi, double ri, const double *xj, double rj, double *n, double *cp, double *delta) {
    const double dx = xi[0] - xj[0], dy = xi[1] - xj[1], dz = xi[2] - xj[2];
    const double rsq = (dx * dx + dy * dy) + dz * dz;
    const double sep = ri + rj;
    const double sep2 = sep + 0.0;          // contact_threshold = 0.0
    if(!(rsq < sep * sep2)) { return 0; }
    const double dist = sqrt(rsq);
    const double pen = (dist - ri) - rj;
    const double d = -(pen);
    if(d < 0.0) { return 0; }               // skip_when(delta_ij < 0.0)
    const double inv = 1.0 / dist;
    n[0] = dx * inv; n[1] = dy * inv; n[2] = dz * inv;
    const double k = rj + 0.5 * pen;
    cp[0] = xj[0] + n[0] * k; cp[1] = xj[1] + n[1] * k; cp[2] = xj[2] + n[2] * k;
    *delta = d;
    return 1;
}

// sphere i - half-space j (position xj, unit normal nj)
PB_HD int pb_dem_geom_halfspace(const double *xi, double ri, const double *xj, const double *nj, double *n, double *cp, double *delta) {
    const double k = (nj[0] * xi[0] + nj[1] * xi[1]) + nj[2] * xi[2];
    const double km = k - ri;
    const double dd = (nj[0] * xj[0] + nj[1] * xj[1]) + nj[2] * xj[2];
    const double pen = km - dd;
    if(!(pen < 0.0)) { return 0; }
    const double d = -(pen);
    if(d < 0.0) { return 0; }
    const double tmp = ri + pen;
    n[0] = nj[0]; n[1] = nj[1]; n[2] = nj[2];
    cp[0] = xi[0] - tmp * nj[0]; cp[1] = xi[1] - tmp * nj[1]; cp[2] = xi[2] - tmp * nj[2];
    *delta = d;
    return 1;
}

// ---- linear_spring_dashpot (examples/dem.py:18-74) ----------------------------------------------------------------
// Contact state (tangential_spring_displacement, impact_velocity_magnitude, is_sticking) is read and overwritten.
// F = partial force on i, T = cross(contact_point - x_i, F).
PB_HD void pb_dem_pair_force(const PbDemParams &P, const double *xi, const double *vi, const double *wi, double inv_mi,
                             const double *xj, const double *vj, const double *wj, double mj, const double *n, const double *cp,
                             double delta, double fric_static, double fric_dynamic, double *tsd, double *ivm, int *sticking,
                             double *F, double *T) {
    const double inv_mj = 1.0 / mj;
    const double meff = 1.0 / (inv_mi + inv_mj);
    const double stiffness_norm = (meff * P.c_sum) / P.ct2;
    const double stiffness_tan = P.kappa * stiffness_norm;
    const double damping_norm = (((-(2.0)) * meff) * P.ln_coeff) / P.ct;
    const double damping_tan = P.sqrt_kappa * damping_norm;

    double ri_[3] = {cp[0] - xi[0], cp[1] - xi[1], cp[2] - xi[2]}, ci[3];
    pb_cross3(wi, ri_, ci);
    const double vwi[3] = {vi[0] + ci[0], vi[1] + ci[1], vi[2] + ci[2]};
    double rj_[3] = {cp[0] - xj[0], cp[1] - xj[1], cp[2] - xj[2]}, cj[3];
    pb_cross3(wj, rj_, cj);
    const double vwj[3] = {vj[0] + cj[0], vj[1] + cj[1], vj[2] + cj[2]};

    const double rel[3] = {-(vwi[0] - vwj[0]), -(vwi[1] - vwj[1]), -(vwi[2] - vwj[2])};
    const double rn = pb_dot3(rel, n);
    const double rel_n[3] = {rn * n[0], rn * n[1], rn * n[2]};
    const double rel_t[3] = {rel[0] - rel_n[0], rel[1] - rel_n[1], rel[2] - rel_n[2]};
    const double sd = stiffness_norm * delta;
    const double fN[3] = {sd * n[0] + damping_norm * rel_n[0], sd * n[1] + damping_norm * rel_n[1], sd * n[2] + damping_norm * rel_n[2]};

    const double t0[3] = {tsd[0], tsd[1], tsd[2]};
    const double ivm_old = *ivm;
    const double impact_magnitude = (ivm_old > 0.0) ? ivm_old : sqrt(pb_dot3(rel, rel));
    const int stick = *sticking;

    const double tn = pb_dot3(t0, n);
    const double rot[3] = {t0[0] - n[0] * tn, t0[1] - n[1] * tn, t0[2] - n[2] * tn};
    const double rot_len2 = pb_dot3(rot, rot);
    const double scale = sqrt(pb_dot3(t0, t0) / rot_len2);      // evaluated unconditionally, as select() does
    const bool rot_zero = rot_len2 <= 0.0;
    double nt[3];
    for(int d = 0; d < 3; d++) { nt[d] = P.dt * rel_t[d] + (rot_zero ? 0.0 : rot[d] * scale); }

    const double dtv[3] = {damping_tan * rel_t[0], damping_tan * rel_t[1], damping_tan * rel_t[2]};
    const double fTLS[3] = {stiffness_tan * nt[0] + dtv[0], stiffness_tan * nt[1] + dtv[1], stiffness_tan * nt[2] + dtv[2]};
    const double fTLS_len = sqrt(pb_dot3(fTLS, fTLS));
    const double inv_len = 1.0 / fTLS_len;
    const bool len_pos = fTLS_len > 0.0;
    const double t[3] = {len_pos ? fTLS[0] * inv_len : 0.0, len_pos ? fTLS[1] * inv_len : 0.0, len_pos ? fTLS[2] * inv_len : 0.0};

    const double fN_len = sqrt(pb_dot3(fN, fN));
    const double f_static = fric_static * fN_len;
    const double f_dynamic = fric_dynamic * fN_len;
    const double rel_t_len = sqrt(pb_dot3(rel_t, rel_t));
    const bool cond1 = (stick == 1) && (rel_t_len < 1e-08) && (fTLS_len < f_static);
    const bool cond2 = (stick == 1) && (fTLS_len < f_dynamic);
    const double f_abs = cond1 ? f_static : f_dynamic;
    const int n_sticking = (cond1 || cond2 || (fTLS_len < f_dynamic)) ? 1 : 0;
    const bool relax = (!cond1) && (!cond2) && (stiffness_tan > 0.0);
    for(int d = 0; d < 3; d++) {
        const double alt = ((f_abs * t[d]) - dtv[d]) / stiffness_tan;
        tsd[d] = relax ? alt : nt[d];
    }
    *ivm = impact_magnitude;
    *sticking = n_sticking;

    const double fTabs = (f_abs < fTLS_len) ? f_abs : fTLS_len;      // min(fTLS_len, f_friction_abs)
    for(int d = 0; d < 3; d++) { F[d] = fN[d] + fTabs * t[d]; }
    const double r[3] = {cp[0] - xi[0], cp[1] - xi[1], cp[2] - xi[2]};
    pb_cross3(r, F, T);
}

// ---- euler (examples/dem.py:77-85) -------------------------------------------------------------------------------
// x, v, w, q (w, x, y, z), R (row-major 3x3) are updated in place.  sin/cos come from the platform's libm (CUDA's and
// glibc's agree to ~1 ulp, not bit for bit): the angular part is parity-tested to 1e-12, the linear part bit-exactly.
PB_HD void pb_dem_euler(double dt, double mass, const double *f, const double *tau, const double *Iinv, double *x, double *v,
                        double *w, double *q, double *R) {
    const double inv_mass = 1.0 / mass;
    const double h = 0.5 * inv_mass;
    for(int d = 0; d < 3; d++) { x[d] = x[d] + ((((h * f[d]) * dt) * dt) + v[d] * dt); }
    for(int d = 0; d < 3; d++) { v[d] = v[d] + (inv_mass * f[d]) * dt; }
    double a[3], b[3], wdot[3];
    for(int r = 0; r < 3; r++) { a[r] = (Iinv[r * 3] * tau[0] + Iinv[r * 3 + 1] * tau[1]) + Iinv[r * 3 + 2] * tau[2]; }   // inv_inertia * torque
    for(int r = 0; r < 3; r++) { b[r] = (R[r * 3] * a[0] + R[r * 3 + 1] * a[1]) + R[r * 3 + 2] * a[2]; }                  // R * (...)
    // (vector) * transposed(R): out[c] = sum_k b[k] * Rt[c*3+k] with Rt = transposed(R) -> Rt[c*3+k] = R[k*3+c]
    for(int c = 0; c < 3; c++) { wdot[c] = (b[0] * R[0 * 3 + c] + b[1] * R[1 * 3 + c]) + b[2] * R[2 * 3 + c]; }
    double phi[3];
    for(int d = 0; d < 3; d++) { phi[d] = w[d] * dt + ((0.5 * wdot[d]) * dt) * dt; }
    const double len = sqrt(pb_dot3(phi, phi));
    double dq[4] = {1.0, 0.0, 0.0, 0.0};
    if(!(fabs(len) < 1e-06)) {
        const double half = len * 0.5;
        const double ca = cos(half), sa = sin(half);
        const double il = 1.0 / len;
        dq[0] = ca; dq[1] = sa * (phi[0] * il); dq[2] = sa * (phi[1] * il); dq[3] = sa * (phi[2] * il);
    }
    const double r0 = ((dq[0] * q[0] - dq[1] * q[1]) - dq[2] * q[2]) - dq[3] * q[3];
    const double r1 = ((dq[0] * q[1] + dq[1] * q[0]) + dq[2] * q[3]) - dq[3] * q[2];
    const double r2 = ((dq[0] * q[2] + dq[2] * q[0]) + dq[3] * q[1]) - dq[1] * q[3];
    const double r3 = ((dq[0] * q[3] + dq[3] * q[0]) + dq[1] * q[2]) - dq[2] * q[1];
    const double len2 = ((r0 * r0 + r1 * r1) + r2 * r2) + r3 * r3;
    const double ilen = (fabs(len2 - 1.0) < 1e-08) ? 1.0 : 1.0 / sqrt(len2);
    q[0] = r0 * ilen; q[1] = r1 * ilen; q[2] = r2 * ilen; q[3] = r3 * ilen;
    R[0] = (1.0 - (2.0 * q[2]) * q[2]) - (2.0 * q[3]) * q[3];
    R[1] = 2.0 * (q[1] * q[2] - q[0] * q[3]);
    R[2] = 2.0 * (q[1] * q[3] + q[0] * q[2]);
    R[3] = 2.0 * (q[1] * q[2] + q[0] * q[3]);
    R[4] = (1.0 - (2.0 * q[1]) * q[1]) - (2.0 * q[3]) * q[3];
    R[5] = 2.0 * (q[2] * q[3] - q[0] * q[1]);
    R[6] = 2.0 * (q[1] * q[3] - q[0] * q[2]);
    R[7] = 2.0 * (q[2] * q[3] + q[0] * q[1]);
    R[8] = (1.0 - (2.0 * q[1]) * q[1]) - (2.0 * q[2]) * q[2];
    for(int d = 0; d < 3; d++) { w[d] = w[d] + wdot[d] * dt; }
}

// ---- gravity (examples/dem.py:88-90): force_z -= (rho_p - rho_f) * volume * g ------------------------------------
PB_HD double pb_dem_gravity(const PbDemParams &P, double radius, double fz) {
    const double volume = ((((4.0 / 3.0) * P.pi) * radius) * radius) * radius;
    return fz - (P.grav_coeff * volume) * P.gravity;
}

// ---- update_mass_and_inertia (examples/dem.py:6-15), sphere branch: inversed(diagonal_matrix(0.4*m*r*r)) ---------
PB_HD void pb_dem_sphere_inv_inertia(double mass, double radius, double *Iinv) {
    const double I = ((0.4 * mass) * radius) * radius;
    const double m[9] = {I, 0.0, 0.0, 0.0, I, 0.0, 0.0, 0.0, I};
    const double det = (m[0] * ((m[4] * m[8]) - (m[7] * m[5])) + m[1] * ((m[5] * m[6]) - (m[8] * m[3]))) + m[2] * ((m[3] * m[7]) - (m[4] * m[6]));
    const double id = 1.0 / det;
    Iinv[0] = id * ((m[4] * m[8]) - (m[5] * m[7]));
    Iinv[1] = id * ((m[7] * m[2]) - (m[8] * m[1]));
    Iinv[2] = id * ((m[1] * m[5]) - (m[2] * m[4]));
    Iinv[3] = id * ((m[5] * m[6]) - (m[3] * m[8]));
    Iinv[4] = id * ((m[8] * m[0]) - (m[6] * m[2]));
    Iinv[5] = id * ((m[2] * m[3]) - (m[0] * m[5]));
    Iinv[6] = id * ((m[3] * m[7]) - (m[4] * m[6]));
    Iinv[7] = id * ((m[6] * m[1]) - (m[7] * m[0]));
    Iinv[8] = id * ((m[0] * m[4]) - (m[1] * m[3]));
}
