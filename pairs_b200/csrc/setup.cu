// Host-side set-up with the reference runtime's exact integer / fp64 sequences, so that the initial state is
// bit-identical to the reference's:
//   pb_copper_fcc_lattice  = pairs::copper_fcc_lattice (runtime/copper_fcc_lattice.hpp:18-26 RNG, :64-145 walk)
//   pb_adjust_thermo       = pairs::adjust_thermo       (runtime/thermo.hpp:53-97)
// Particle types come from glibc rand() in its initial state (seed 1) in the reference (one process per rank); here
// a private random_r state seeded with 1 gives the same stream without touching the host program's rand().
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.cuh"

int pb_allreduce_thermo(pb_ctx *ctx, double *sum_mv2, long *natoms);
int pb_allreduce_sum(pb_ctx *ctx, double *vals, int n);

struct PbHostStage {
    std::vector<double> pos, vel, mass;
    std::vector<int> type;
};

static std::map<pb_ctx *, PbHostStage> g_stage;   // lives until the next upload

static double pb_myrandom(int *seed) {
    const int k = (*seed) / 127773;
    *seed = 16807 * (*seed - k * 127773) - 2836 * k;
    if(*seed < 0) { *seed += 2147483647; }
    return (1.0 / 2147483647) * (*seed);
}

static int pb_within_subdomain(const pb_ctx *ctx, double x, double y, double z) {
    // Regular6DStencil::isWithinSubdomain, SMALL = 1e-5 (runtime/domain/regular_6d_stencil.{hpp:6,cpp:96-100})
    return x >= ctx->subdom[0] && x < ctx->subdom[1] - 0.00001 && y >= ctx->subdom[2] && y < ctx->subdom[3] - 0.00001 &&
           z >= ctx->subdom[4] && z < ctx->subdom[5] - 0.00001;
}

extern "C" int pb_copper_fcc_lattice(pb_ctx *ctx, int nx, int ny, int nz, double rho, int ntypes, int *nlocal) {
    if(!ctx->domain_set) { ctx->set_error("pb_copper_fcc_lattice: domain not initialised"); return -1; }
    PbHostStage &hs = g_stage[ctx];
    hs.pos.clear(); hs.vel.clear(); hs.mass.clear(); hs.type.clear();
    const double xlo = 0.0, xhi = ctx->grid[1], ylo = 0.0, yhi = ctx->grid[3], zlo = 0.0, zhi = ctx->grid[5];
    const double alat = pow((4.0 / rho), (1.0 / 3.0));
    int ilo = (int) (xlo / (0.5 * alat) - 1), ihi = (int) (xhi / (0.5 * alat) + 1);
    int jlo = (int) (ylo / (0.5 * alat) - 1), jhi = (int) (yhi / (0.5 * alat) + 1);
    int klo = (int) (zlo / (0.5 * alat) - 1), khi = (int) (zhi / (0.5 * alat) + 1);
    ilo = std::max(ilo, 0); ihi = std::min(ihi, 2 * nx - 1);
    jlo = std::max(jlo, 0); jhi = std::min(jhi, 2 * ny - 1);
    klo = std::max(klo, 0); khi = std::min(khi, 2 * nz - 1);
    struct random_data rd;
    char statebuf[128];
    memset(&rd, 0, sizeof(rd));
    memset(statebuf, 0, sizeof(statebuf));
    initstate_r(1, statebuf, sizeof(statebuf), &rd);   // == the process-initial state of rand()
    int sx = 0, sy = 0, sz = 0, ox = 0, oy = 0, oz = 0;
    const int subboxdim = 8;
    while(oz * subboxdim <= khi) {
        const int k = oz * subboxdim + sz, j = oy * subboxdim + sy, i = ox * subboxdim + sx;
        if(((i + j + k) % 2 == 0) && (i >= ilo) && (i <= ihi) && (j >= jlo) && (j <= jhi) && (k >= klo) && (k <= khi)) {
            const double xtmp = 0.5 * alat * i, ytmp = 0.5 * alat * j, ztmp = 0.5 * alat * k;
            if(pb_within_subdomain(ctx, xtmp, ytmp, ztmp)) {
                int n = k * (2 * ny) * (2 * nx) + j * (2 * nx) + i + 1;
                for(int m = 0; m < 5; m++) { pb_myrandom(&n); }
                const double vx = pb_myrandom(&n);
                for(int m = 0; m < 5; m++) { pb_myrandom(&n); }
                const double vy = pb_myrandom(&n);
                for(int m = 0; m < 5; m++) { pb_myrandom(&n); }
                const double vz = pb_myrandom(&n);
                hs.mass.push_back(1.0);
                hs.pos.push_back(xtmp); hs.pos.push_back(ytmp); hs.pos.push_back(ztmp);
                hs.vel.push_back(vx); hs.vel.push_back(vy); hs.vel.push_back(vz);
                int32_t r = 0;
                random_r(&rd, &r);
                hs.type.push_back((int) (r % ntypes));
            }
        }
        sx++;
        if(sx == subboxdim) { sx = 0; sy++; }
        if(sy == subboxdim) { sy = 0; sz++; }
        if(sz == subboxdim) { sz = 0; ox++; }
        if(ox * subboxdim > ihi) { ox = 0; oy++; }
        if(oy * subboxdim > jhi) { oy = 0; oz++; }
    }
    const int n = (int) hs.mass.size();
    // global tags: rank-major numbering needs the counts of lower ranks; a per-rank stride keeps tags unique
    ctx->tag_base = ctx->rank * (int) (((long) 4 * nx * ny * nz + ctx->world - 1) / ctx->world * 2);
    PB_TRY(pb_upload_particles(ctx, n, hs.pos.data(), hs.vel.data(), hs.mass.data(), hs.type.data(), nullptr, nullptr, nullptr));
    *nlocal = n;
    return 0;
}

extern "C" int pb_adjust_thermo(pb_ctx *ctx, double temp) {
    auto it = g_stage.find(ctx);
    if(it == g_stage.end() || (int) it->second.mass.size() != ctx->nlocal) {
        ctx->set_error("pb_adjust_thermo: must directly follow pb_copper_fcc_lattice");
        return -1;
    }
    PbHostStage &hs = it->second;
    const int nlocal = ctx->nlocal;
    double v[4] = {0.0, 0.0, 0.0, (double) nlocal};
    for(int i = 0; i < nlocal; i++) {
        v[0] += hs.vel[i * 3 + 0];
        v[1] += hs.vel[i * 3 + 1];
        v[2] += hs.vel[i * 3 + 2];
    }
    if(ctx->world > 1) { PB_TRY(pb_allreduce_sum(ctx, v, 4)); }
    const long natoms = (long) (v[3] + 0.5);
    const double vxtot = v[0] / natoms, vytot = v[1] / natoms, vztot = v[2] / natoms;
    for(int i = 0; i < nlocal; i++) {
        hs.vel[i * 3 + 0] -= vxtot;
        hs.vel[i * 3 + 1] -= vytot;
        hs.vel[i * 3 + 2] -= vztot;
    }
    // compute_thermo(print = 0), serial left-to-right as the reference
    double t = 0.0;
    for(int i = 0; i < nlocal; i++) {
        t += hs.mass[i] * (hs.vel[i * 3] * hs.vel[i * 3] + hs.vel[i * 3 + 1] * hs.vel[i * 3 + 1] + hs.vel[i * 3 + 2] * hs.vel[i * 3 + 2]);
    }
    if(ctx->world > 1) {
        double tt[1] = {t};
        PB_TRY(pb_allreduce_sum(ctx, tt, 1));
        t = tt[0];
    }
    const double dof_boltz = (double) (natoms * 3 - 3);
    t = t * (1.0 / dof_boltz);
    const double factor = sqrt(temp / t);
    for(int i = 0; i < nlocal; i++) {
        hs.vel[i * 3 + 0] *= factor;
        hs.vel[i * 3 + 1] *= factor;
        hs.vel[i * 3 + 2] *= factor;
    }
    const int keep_base = ctx->tag_base;
    PB_TRY(pb_upload_particles(ctx, nlocal, hs.pos.data(), hs.vel.data(), hs.mass.data(), hs.type.data(), nullptr, nullptr, nullptr));
    ctx->tag_base = keep_base;
    g_stage.erase(it);
    return 0;
}
