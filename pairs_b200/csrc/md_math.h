// Per-particle / per-pair arithmetic of the MD path as host+device inline functions: the cell index of BuildCellLists and
// the Lennard-Jones pair term, each written operation by operation in the order the reference's generator emits it
// (generated md.cpp: build_cell_lists, lennard_jones).  On the device the operations are the explicit round-to-nearest
// intrinsics (never contracted); on the host -- where the header is compiled into tests/host/md_host.cpp so that the CPU
// test-suite can run the PRODUCT's arithmetic against the reference's generated modules -- they are the plain operators,
// built with -ffp-contract=off.  Same role as dem_math.h for the DEM path.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PB_MD_HD __host__ __device__ __forceinline__
#else
#define PB_MD_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define PB_ADD(a, b) __dadd_rn((a), (b))
#define PB_SUB(a, b) __dsub_rn((a), (b))
#define PB_MUL(a, b) __dmul_rn((a), (b))
#define PB_DIV(a, b) __ddiv_rn((a), (b))
#else
#define PB_ADD(a, b) ((a) + (b))
#define PB_SUB(a, b) ((a) - (b))
#define PB_MUL(a, b) ((a) * (b))
#define PB_DIV(a, b) ((a) / (b))
#endif

#ifndef PB_FLAG_INFINITE
#define PB_FLAG_INFINITE 1
#endif

struct PbCellGeom {
    double lo[3];       // subdom_min - spacing
    double spacing;
    int dim[3];
    int ncells;
};

// BuildCellLists index arithmetic (sim/cell_lists.py:111-127; generated md.cpp build_cell_lists):
//   c_d = clamp((int)((x_d - (min_d - s)) / s), 0, dim_d - 1);  flat = (c0*dim1 + c1)*dim2 + c2 + 1;  INFINITE -> 0
PB_MD_HD int pb_cell_index(const PbCellGeom &g, double x, double y, double z, int flags) {
    if(flags & PB_FLAG_INFINITE) { return 0; }
    const double q0 = PB_DIV(PB_SUB(x, g.lo[0]), g.spacing);
    const double q1 = PB_DIV(PB_SUB(y, g.lo[1]), g.spacing);
    const double q2 = PB_DIV(PB_SUB(z, g.lo[2]), g.spacing);
    int c0 = (int) q0, c1 = (int) q1, c2 = (int) q2;      // truncation toward zero, as the C cast
    c0 = (c0 >= 0) ? c0 : 0; c0 = (c0 < g.dim[0]) ? c0 : g.dim[0] - 1;
    c1 = (c1 >= 0) ? c1 : 0; c1 = (c1 < g.dim[1]) ? c1 : g.dim[1] - 1;
    c2 = (c2 >= 0) ? c2 : 0; c2 = (c2 < g.dim[2]) ? c2 : g.dim[2] - 1;
    return (c0 * g.dim[1] + c1) * g.dim[2] + c2 + 1;
}

// squared distance of a pair, (dx*dx + dy*dy) + dz*dz (sim/interaction.py squared_distance, generated md.cpp)
PB_MD_HD double pb_pair_rsq(double xi, double yi, double zi, double xj, double yj, double zj, double *dx, double *dy, double *dz) {
    *dx = PB_SUB(xi, xj);
    *dy = PB_SUB(yi, yj);
    *dz = PB_SUB(zi, zj);
    return PB_ADD(PB_ADD(PB_MUL(*dx, *dx), PB_MUL(*dy, *dy)), PB_MUL(*dz, *dz));
}

// examples/md.py:5-8: sr2 = 1/rsq; sr6 = sr2*sr2*sr2*sigma6; f = 48*sr6*(sr6 - 0.5)*sr2*epsilon  (scalar factor of delta)
PB_MD_HD double pb_lj_fpair(double rsq, double sig6, double eps) {
    const double sr2 = PB_DIV(1.0, rsq);
    const double sr6 = PB_MUL(PB_MUL(PB_MUL(sr2, sr2), sr2), sig6);
    return PB_MUL(PB_MUL(PB_MUL(PB_MUL(48.0, sr6), PB_SUB(sr6, 0.5)), sr2), eps);
}

// ---- production arithmetic (option "lj_fma", the default of the force kernels) ---------------------------------------------------
// The force kernel is co-limited by the fp64 pipe, the shared-memory gathers and the issue slots, so the
// production variant spends as few fp64 instructions per pair as the 1e-12 parity budget allows -- 17 instead of the 33 of the
// reference's expression tree:
//   rsq = fma(dz, dz, fma(dy, dy, dx * dx))                                   3 + 3 (the differences)
//   cutoff test on the BIT PATTERNS (rsq >= +0: doubles order like their bits)  0   (two integer compares on the ALU, no DSETP)
//   sr2 = 1 / rsq: MUFU.RCP64H (~2^-20) + one CUBIC step y (1 + e + e^2)        3   (error e^3 ~ 2^-60; not correctly rounded)
//   a = sr2^3;  f = sr2 * a * (c1 * a - c2),  c1 = 48 eps sigma^12, c2 = 24 eps sigma^6    5
//   force += delta * f                                                        3 fmas
// Every result is within a few ulp of the reference's expression: measured <= 2e-15 of the largest force component on 4 M atoms
// (tools/micro/tile_force.cu), against the 1e-12 of the parity contract.  The bit-exact expressions above stay the ones the
// parity tests pin (option "lj_fma" = 0).
PB_MD_HD double pb_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

PB_MD_HD double pb_rcp_cubic(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));      // MUFU.RCP64H: ~20 bits
    const double e = __fma_rn(-x, y, 1.0);
    const double t = __fma_rn(e, e, e);
    return __fma_rn(y, t, y);
#else
    return 1.0 / x;
#endif
}

PB_MD_HD double pb_pair_rsq_fma(double xi, double yi, double zi, double xj, double yj, double zj, double *dx, double *dy, double *dz) {
    *dx = PB_SUB(xi, xj);
    *dy = PB_SUB(yi, yj);
    *dz = PB_SUB(zi, zj);
    return pb_fma(*dz, *dz, pb_fma(*dy, *dy, PB_MUL(*dx, *dx)));
}

// rsq < cutsq for rsq >= +0 (NaN never: rsq is a sum of squares of finite numbers)
PB_MD_HD bool pb_less_bits(double rsq, double cutsq) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(rsq) < __double_as_longlong(cutsq);
#else
    return rsq < cutsq;
#endif
}

// c1 = 48 eps sigma6^2, c2 = 24 eps sigma6 (host side: pb_lj_fast_coeff)
PB_MD_HD double pb_lj_fpair_fast(double rsq, double c1, double c2) {
    const double sr2 = pb_rcp_cubic(rsq);
    const double a = PB_MUL(PB_MUL(sr2, sr2), sr2);
    return PB_MUL(PB_MUL(sr2, a), pb_fma(c1, a, -c2));
}
