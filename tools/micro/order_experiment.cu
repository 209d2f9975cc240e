// Experiment: how much does the PHYSICAL ORDER of the particles change the cost of the neighbour-list force kernel?
// The force kernel is L1TEX-wavefront bound (one wavefront per distinct 128-byte line a warp-wide gather touches), so the
// question is how many distinct lines the 32 lanes (= 32 consecutive particles) touch when each gathers its k-th neighbour
// (lists sorted by ascending index).  Orders compared on a jittered FCC liquid (rho 0.8442, list cutoff 2.8, force cutoff 2.5):
//   0  coarse cell (2.8, z fastest) + 4 z-slabs per cell        (the production order of round 1)
//   1  xy column of width 2.8, exact z inside the column
//   2  xy column of width 1.4, exact z
//   3  xy column of width 0.933, exact z
//   4  fine cells 1.4^3, z fastest
//   5  Morton order of 0.7^3 cells
// Build: nvcc -O3 -std=c++17 --fmad=false -gencode arch=compute_100a,code=sm_100a -o order_experiment order_experiment.cu
#include <cuda_runtime.h>
#include <thrust/device_vector.h>
#include <thrust/sort.h>
#include <thrust/sequence.h>
#include <thrust/gather.h>
#include <thrust/binary_search.h>
#include <thrust/iterator/counting_iterator.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while(0)

static const int T = 112;      // list capacity

__device__ __forceinline__ double4 ld256(const double4 *p) {
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ unsigned long long spread3(unsigned v) {
    unsigned long long x = v & 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

__global__ void k_keys(int n, int mode, double L, const double4 *pos, unsigned long long *key) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double4 p = pos[i];
    unsigned long long k = 0;
    if(mode == 0) {
        int nc = (int) ceil(L / 2.8);
        int cx = min((int) (p.x / 2.8), nc - 1), cy = min((int) (p.y / 2.8), nc - 1), cz = min((int) (p.z / 2.8), nc - 1);
        int slab = min((int) ((p.z - cz * 2.8) / 0.7), 3);
        k = ((unsigned long long) ((cx * nc + cy) * nc + cz)) * 4 + slab;
    } else if(mode >= 1 && mode <= 3) {
        double w = (mode == 1) ? 2.8 : ((mode == 2) ? 1.4 : 2.8 / 3.0);
        int nc = (int) ceil(L / w);
        int cx = min((int) (p.x / w), nc - 1), cy = min((int) (p.y / w), nc - 1);
        unsigned zq = (unsigned) (p.z / L * 4.0e9);
        k = ((unsigned long long) (cx * nc + cy) << 32) | zq;
    } else if(mode == 4) {
        double w = 1.4;
        int nc = (int) ceil(L / w);
        int cx = min((int) (p.x / w), nc - 1), cy = min((int) (p.y / w), nc - 1), cz = min((int) (p.z / w), nc - 1);
        k = ((unsigned long long) (cx * nc + cy)) * nc + cz;
    } else {
        double w = 0.7;
        k = spread3((unsigned) (p.x / w)) << 2 | spread3((unsigned) (p.y / w)) << 1 | spread3((unsigned) (p.z / w));
    }
    key[i] = k;
}

__global__ void k_cell(int n, int nc, const double4 *pos, int *cell) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double4 p = pos[i];
    int cx = min((int) (p.x / 2.8), nc - 1), cy = min((int) (p.y / 2.8), nc - 1), cz = min((int) (p.z / 2.8), nc - 1);
    cell[i] = (cx * nc + cy) * nc + cz;
}

// brute-force build over the coarse grid; list sorted ascending by index; ELLPACK slice-32
__global__ void __launch_bounds__(128) k_build(int n, int nc, const double4 *pos, const int *cell_start, const int *cell_list, int *numneigh, int *neigh) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double4 p = pos[i];
    int cx = min((int) (p.x / 2.8), nc - 1), cy = min((int) (p.y / 2.8), nc - 1), cz = min((int) (p.z / 2.8), nc - 1);
    int nb[T];
    int c = 0;
    for(int dx = -1; dx <= 1; dx++) for(int dy = -1; dy <= 1; dy++) for(int dz = -1; dz <= 1; dz++) {
        int x = cx + dx, y = cy + dy, z = cz + dz;
        if(x < 0 || y < 0 || z < 0 || x >= nc || y >= nc || z >= nc) continue;
        int cc = (x * nc + y) * nc + z;
        for(int s = cell_start[cc]; s < cell_start[cc + 1]; s++) {
            int j = cell_list[s];
            if(j == i) continue;
            double4 q = pos[j];
            double ddx = p.x - q.x, ddy = p.y - q.y, ddz = p.z - q.z;
            double r2 = ddx * ddx + ddy * ddy + ddz * ddz;
            if(r2 < 7.84 && c < T) nb[c++] = j;
        }
    }
    for(int a = 1; a < c; a++) {      // insertion sort, ascending
        int v = nb[a], b = a - 1;
        while(b >= 0 && nb[b] > v) { nb[b + 1] = nb[b]; b--; }
        nb[b + 1] = v;
    }
    numneigh[i] = c;
    int *out = neigh + (size_t) (i / 32) * T * 32 + (i % 32);
    for(int k = 0; k < c; k++) out[(size_t) k * 32] = nb[k];
}

__global__ void __launch_bounds__(128) k_force(int n, const int *__restrict__ numneigh, const int *__restrict__ neigh, const double4 *__restrict__ pos,
                                               double *__restrict__ force) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const double4 p = ld256(pos + i);
    const int nn = numneigh[i];
    const int *nb = neigh + (size_t) (i / 32) * T * 32 + (i % 32);
    double fx = 0, fy = 0, fz = 0;
    int k = 0;
    for(; k + 4 <= nn; k += 4) {
        int j[4]; double4 q[4];
#pragma unroll
        for(int u = 0; u < 4; u++) j[u] = __ldg(nb + (size_t) (k + u) * 32);
#pragma unroll
        for(int u = 0; u < 4; u++) q[u] = ld256(pos + j[u]);
#pragma unroll
        for(int u = 0; u < 4; u++) {
            double dx = p.x - q[u].x, dy = p.y - q[u].y, dz = p.z - q[u].z;
            double r2 = dx * dx + dy * dy + dz * dz;
            if(r2 < 6.25) {
                double sr2 = 1.0 / r2, sr6 = sr2 * sr2 * sr2;
                double f = 48.0 * sr6 * (sr6 - 0.5) * sr2;
                fx += dx * f; fy += dy * f; fz += dz * f;
            }
        }
    }
    for(; k < nn; k++) {
        int j = __ldg(nb + (size_t) k * 32);
        double4 q = ld256(pos + j);
        double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        double r2 = dx * dx + dy * dy + dz * dz;
        if(r2 < 6.25) {
            double sr2 = 1.0 / r2, sr6 = sr2 * sr2 * sr2;
            double f = 48.0 * sr6 * (sr6 - 0.5) * sr2;
            fx += dx * f; fy += dy * f; fz += dz * f;
        }
    }
    force[i] = fx; force[n + i] = fy; force[2 * (size_t) n + i] = fz;
}

// the same kernel with one of every TEXMOD gathers routed through the texture path (2 x tex1Dfetch<int4>)
__device__ __forceinline__ double4 ldtex(cudaTextureObject_t t, int j) {
    int4 a = tex1Dfetch<int4>(t, 2 * j), b = tex1Dfetch<int4>(t, 2 * j + 1);
    double4 r;
    r.x = __hiloint2double(a.y, a.x); r.y = __hiloint2double(a.w, a.z);
    r.z = __hiloint2double(b.y, b.x); r.w = __hiloint2double(b.w, b.z);
    return r;
}

template<int TEXMOD>
__global__ void __launch_bounds__(128) k_force_mix(int n, const int *__restrict__ numneigh, const int *__restrict__ neigh, const double4 *__restrict__ pos,
                                                   cudaTextureObject_t tex, double *__restrict__ force) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const double4 p = ld256(pos + i);
    const int nn = numneigh[i];
    const int *nb = neigh + (size_t) (i / 32) * T * 32 + (i % 32);
    double fx = 0, fy = 0, fz = 0;
    int k = 0;
    for(; k + 4 <= nn; k += 4) {
        int j[4]; double4 q[4];
#pragma unroll
        for(int u = 0; u < 4; u++) j[u] = __ldg(nb + (size_t) (k + u) * 32);
#pragma unroll
        for(int u = 0; u < 4; u++) q[u] = (u % TEXMOD == 0) ? ldtex(tex, j[u]) : ld256(pos + j[u]);
#pragma unroll
        for(int u = 0; u < 4; u++) {
            double dx = p.x - q[u].x, dy = p.y - q[u].y, dz = p.z - q[u].z;
            double r2 = dx * dx + dy * dy + dz * dz;
            if(r2 < 6.25) {
                double sr2 = 1.0 / r2, sr6 = sr2 * sr2 * sr2;
                double f = 48.0 * sr6 * (sr6 - 0.5) * sr2;
                fx += dx * f; fy += dy * f; fz += dz * f;
            }
        }
    }
    for(; k < nn; k++) {
        int j = __ldg(nb + (size_t) k * 32);
        double4 q = ld256(pos + j);
        double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        double r2 = dx * dx + dy * dy + dz * dz;
        if(r2 < 6.25) {
            double sr2 = 1.0 / r2, sr6 = sr2 * sr2 * sr2;
            double f = 48.0 * sr6 * (sr6 - 0.5) * sr2;
            fx += dx * f; fy += dy * f; fz += dz * f;
        }
    }
    force[i] = fx; force[n + i] = fy; force[2 * (size_t) n + i] = fz;
}

// the same kernel with SoA positions (x[], y[], z[]): 16 particles per 128-byte line instead of 4, three 64-bit gathers per neighbour
__global__ void __launch_bounds__(128) k_force_soa(int n, const int *__restrict__ numneigh, const int *__restrict__ neigh, const double *__restrict__ X,
                                                   const double *__restrict__ Y, const double *__restrict__ Z, double *__restrict__ force) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const double px = X[i], py = Y[i], pz = Z[i];
    const int nn = numneigh[i];
    const int *nb = neigh + (size_t) (i / 32) * T * 32 + (i % 32);
    double fx = 0, fy = 0, fz = 0;
    int k = 0;
    for(; k + 4 <= nn; k += 4) {
        int j[4]; double qx[4], qy[4], qz[4];
#pragma unroll
        for(int u = 0; u < 4; u++) j[u] = __ldg(nb + (size_t) (k + u) * 32);
#pragma unroll
        for(int u = 0; u < 4; u++) { qx[u] = __ldg(X + j[u]); qy[u] = __ldg(Y + j[u]); qz[u] = __ldg(Z + j[u]); }
#pragma unroll
        for(int u = 0; u < 4; u++) {
            double dx = px - qx[u], dy = py - qy[u], dz = pz - qz[u];
            double r2 = dx * dx + dy * dy + dz * dz;
            if(r2 < 6.25) {
                double sr2 = 1.0 / r2, sr6 = sr2 * sr2 * sr2;
                double f = 48.0 * sr6 * (sr6 - 0.5) * sr2;
                fx += dx * f; fy += dy * f; fz += dz * f;
            }
        }
    }
    for(; k < nn; k++) {
        int j = __ldg(nb + (size_t) k * 32);
        double dx = px - __ldg(X + j), dy = py - __ldg(Y + j), dz = pz - __ldg(Z + j);
        double r2 = dx * dx + dy * dy + dz * dz;
        if(r2 < 6.25) {
            double sr2 = 1.0 / r2, sr6 = sr2 * sr2 * sr2;
            double f = 48.0 * sr6 * (sr6 - 0.5) * sr2;
            fx += dx * f; fy += dy * f; fz += dz * f;
        }
    }
    force[i] = fx; force[n + i] = fy; force[2 * (size_t) n + i] = fz;
}

__global__ void k_split(int n, const double4 *pos, double *X, double *Y, double *Z) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) { double4 p = pos[i]; X[i] = p.x; Y[i] = p.y; Z[i] = p.z; }
}

// distinct 128-byte lines per warp-wide gather
__global__ void __launch_bounds__(128) k_lines(int n, const int *__restrict__ numneigh, const int *__restrict__ neigh, unsigned long long *stats) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int nn = (i < n) ? numneigh[i] : 0;
    const int *nb = neigh + (size_t) (i / 32) * T * 32 + (i % 32);
    int lane = threadIdx.x & 31;
    unsigned long long lines = 0, gathers = 0, lines16 = 0;
    for(int k = 0; k < T; k++) {
        unsigned act = __ballot_sync(0xffffffffu, k < nn);
        if(act == 0) break;
        if(k < nn) {
            int j = nb[(size_t) k * 32];
            unsigned m = __match_any_sync(act, j >> 2);
            if(__ffs(m) - 1 == lane) lines++;
            unsigned m2 = __match_any_sync(act, j >> 4);
            if(__ffs(m2) - 1 == lane) lines16++;
            if(__ffs(act) - 1 == lane) gathers++;
        }
    }
    atomicAdd(stats, lines);
    atomicAdd(stats + 1, gathers);
    atomicAdd(stats + 2, lines16);
}

int main(int argc, char **argv) {
    const int nx = (argc > 1) ? atoi(argv[1]) : 64;
    const double a = pow(4.0 / 0.8442, 1.0 / 3.0), L = nx * a;
    const int n = 4 * nx * nx * nx;
    std::vector<double4> h(n);
    srand(3);
    const double basis[4][3] = {{0, 0, 0}, {0.5, 0.5, 0}, {0.5, 0, 0.5}, {0, 0.5, 0.5}};
    int c = 0;
    for(int x = 0; x < nx; x++) for(int y = 0; y < nx; y++) for(int z = 0; z < nx; z++) for(int b = 0; b < 4; b++) {
        double j[3];
        for(int d = 0; d < 3; d++) j[d] = 0.5 * (rand() / (double) RAND_MAX - 0.5);
        h[c].x = fmin(fmax((x + basis[b][0]) * a + 0.2 + j[0], 0.0), L - 1e-9);
        h[c].y = fmin(fmax((y + basis[b][1]) * a + 0.2 + j[1], 0.0), L - 1e-9);
        h[c].z = fmin(fmax((z + basis[b][2]) * a + 0.2 + j[2], 0.0), L - 1e-9);
        h[c].w = 0;
        c++;
    }
    thrust::device_vector<double4> pos0(h.begin(), h.end()), pos(n);
    thrust::device_vector<unsigned long long> key(n);
    thrust::device_vector<int> perm(n), cell(n), cell_list(n), numneigh(n);
    const int nc = (int) ceil(L / 2.8);
    thrust::device_vector<int> cell_start(nc * nc * nc + 1);
    thrust::device_vector<int> neigh((size_t) ((n + 31) / 32) * T * 32);
    thrust::device_vector<double> force(3 * (size_t) n);
    thrust::device_vector<unsigned long long> stats(3);
    thrust::device_vector<double> X(n), Y(n), Z(n);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("n = %d, L = %.3f, coarse cells %d^3\n", n, L, nc);
    for(int mode = 0; mode < 6; mode++) {
        k_keys<<<(n + 255) / 256, 256>>>(n, mode, L, thrust::raw_pointer_cast(pos0.data()), thrust::raw_pointer_cast(key.data()));
        thrust::sequence(perm.begin(), perm.end());
        thrust::stable_sort_by_key(key.begin(), key.end(), perm.begin());
        thrust::gather(perm.begin(), perm.end(), pos0.begin(), pos.begin());
        k_cell<<<(n + 255) / 256, 256>>>(n, nc, thrust::raw_pointer_cast(pos.data()), thrust::raw_pointer_cast(cell.data()));
        thrust::sequence(cell_list.begin(), cell_list.end());
        thrust::stable_sort_by_key(cell.begin(), cell.end(), cell_list.begin());
        thrust::lower_bound(cell.begin(), cell.end(), thrust::counting_iterator<int>(0), thrust::counting_iterator<int>(nc * nc * nc + 1), cell_start.begin());
        k_build<<<(n + 127) / 128, 128>>>(n, nc, thrust::raw_pointer_cast(pos.data()), thrust::raw_pointer_cast(cell_start.data()),
                                          thrust::raw_pointer_cast(cell_list.data()), thrust::raw_pointer_cast(numneigh.data()),
                                          thrust::raw_pointer_cast(neigh.data()));
        CK(cudaDeviceSynchronize());
        thrust::fill(stats.begin(), stats.end(), 0ULL);
        k_lines<<<(n + 127) / 128, 128>>>(n, thrust::raw_pointer_cast(numneigh.data()), thrust::raw_pointer_cast(neigh.data()), thrust::raw_pointer_cast(stats.data()));
        CK(cudaDeviceSynchronize());
        unsigned long long hs[3];
        cudaMemcpy(hs, thrust::raw_pointer_cast(stats.data()), 24, cudaMemcpyDeviceToHost);
        long long tot = thrust::reduce(numneigh.begin(), numneigh.end(), 0LL);
        for(int w = 0; w < 3; w++) k_force<<<(n + 127) / 128, 128>>>(n, thrust::raw_pointer_cast(numneigh.data()), thrust::raw_pointer_cast(neigh.data()), thrust::raw_pointer_cast(pos.data()), thrust::raw_pointer_cast(force.data()));
        cudaEventRecord(e0);
        const int R = 20;
        for(int r = 0; r < R; r++) k_force<<<(n + 127) / 128, 128>>>(n, thrust::raw_pointer_cast(numneigh.data()), thrust::raw_pointer_cast(neigh.data()), thrust::raw_pointer_cast(pos.data()), thrust::raw_pointer_cast(force.data()));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double fsum = thrust::reduce(force.begin(), force.end(), 0.0);
        k_split<<<(n + 255) / 256, 256>>>(n, thrust::raw_pointer_cast(pos.data()), thrust::raw_pointer_cast(X.data()), thrust::raw_pointer_cast(Y.data()), thrust::raw_pointer_cast(Z.data()));
        for(int w = 0; w < 3; w++) k_force_soa<<<(n + 127) / 128, 128>>>(n, thrust::raw_pointer_cast(numneigh.data()), thrust::raw_pointer_cast(neigh.data()), thrust::raw_pointer_cast(X.data()), thrust::raw_pointer_cast(Y.data()), thrust::raw_pointer_cast(Z.data()), thrust::raw_pointer_cast(force.data()));
        cudaEventRecord(e0);
        for(int r = 0; r < R; r++) k_force_soa<<<(n + 127) / 128, 128>>>(n, thrust::raw_pointer_cast(numneigh.data()), thrust::raw_pointer_cast(neigh.data()), thrust::raw_pointer_cast(X.data()), thrust::raw_pointer_cast(Y.data()), thrust::raw_pointer_cast(Z.data()), thrust::raw_pointer_cast(force.data()));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms2;
        cudaEventElapsedTime(&ms2, e0, e1);
        double fsum2 = thrust::reduce(force.begin(), force.end(), 0.0);
        if(mode == 0) {
            cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = thrust::raw_pointer_cast(pos.data());
            rd.res.linear.desc = cudaCreateChannelDesc<int4>(); rd.res.linear.sizeInBytes = (size_t) n * 32;
            cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
            cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
            auto timeit = [&](const char *name, auto kern) {
                for(int w = 0; w < 3; w++) kern<<<(n + 127) / 128, 128>>>(n, thrust::raw_pointer_cast(numneigh.data()), thrust::raw_pointer_cast(neigh.data()), thrust::raw_pointer_cast(pos.data()), tex, thrust::raw_pointer_cast(force.data()));
                cudaEventRecord(e0);
                for(int r = 0; r < R; r++) kern<<<(n + 127) / 128, 128>>>(n, thrust::raw_pointer_cast(numneigh.data()), thrust::raw_pointer_cast(neigh.data()), thrust::raw_pointer_cast(pos.data()), tex, thrust::raw_pointer_cast(force.data()));
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float t; cudaEventElapsedTime(&t, e0, e1);
                printf("        %s: %.4f ms, checksum %.3e\n", name, t / R, thrust::reduce(force.begin(), force.end(), 0.0));
            };
            timeit("TEX 1 of 4", k_force_mix<4>);
            timeit("TEX 1 of 2", k_force_mix<2>);
            timeit("TEX all   ", k_force_mix<1>);
            cudaDestroyTextureObject(tex);
        }
        printf("mode %d: mean neighbours %.2f, lines per warp gather %.2f (AoS 32 B) / %.2f (SoA 8 B), force kernel AoS %.4f ms (%.3e atoms/s), SoA %.4f ms (%.3e atoms/s), checksums %.3e %.3e\n", mode,
               tot / (double) n, hs[0] / (double) hs[1], hs[2] / (double) hs[1], ms / R, n / (ms / R * 1e-3), ms2 / R, n / (ms2 / R * 1e-3), fsum, fsum2);
    }
    return 0;
}
