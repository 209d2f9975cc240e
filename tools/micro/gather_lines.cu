// Microbenchmark: cost of a warp-wide 32-byte-per-lane gather (LDG.256) as a function of how many distinct 128-byte
// lines / 32-byte sectors the 32 lanes touch.  Records are 32 B (4 per line).  Pattern per warp and iteration:
// lanes are split into groups of S lanes; each group reads S consecutive records starting at a random 128B-aligned
// (S = 4: one full line per group; S = 2: half a line; S = 1: every lane its own random line) or the same record (B).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ double4 ld256(const double4* p) {
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(128) k_gather(int n, int K, const int* __restrict__ idx, const double4* __restrict__ pos, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const int* nb = idx + (size_t)(i / 32) * K * 32 + (i % 32);
    double s = 0;
    for(int k = 0; k + 4 <= K; k += 4) {
        int j[4]; double4 p[4];
#pragma unroll
        for(int u = 0; u < 4; u++) j[u] = __ldg(nb + (size_t)(k + u) * 32);
#pragma unroll
        for(int u = 0; u < 4; u++) p[u] = ld256(pos + j[u]);
#pragma unroll
        for(int u = 0; u < 4; u++) s += p[u].x + p[u].y + p[u].z + p[u].w;
    }
    out[i] = s;
}

int main() {
    const int n = 2000000, K = 76, W = 1500;
    int* d_idx; double4* d_pos; double* d_out;
    size_t ni = (size_t)((n + 31) / 32) * K * 32;
    cudaMalloc(&d_idx, ni * sizeof(int)); cudaMalloc(&d_pos, (size_t)n * 32); cudaMalloc(&d_out, (size_t)n * 8);
    cudaMemset(d_pos, 0, (size_t)n * 32);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    std::vector<int> h(ni);
    // mode: S lanes per group share consecutive records; S=0 -> all 32 lanes the same record
    int modes[] = {1, 2, 4, 8, 32, 0};
    for(int S : modes) {
        srand(7);
        for(int w = 0; w < (n + 31) / 32; w++) {
            for(int k = 0; k < K; k++) {
                int* row = &h[((size_t)w * K + k) * 32];
                if(S == 0) { long j = (long)w * 32 - W + rand() % (2 * W); j = (j % n + n) % n; for(int l = 0; l < 32; l++) row[l] = (int)j; continue; }
                for(int g = 0; g < 32 / S; g++) {
                    long j = (long)w * 32 - W + rand() % (2 * W);
                    j = ((j % n + n) % n) / S * S;      // aligned so that S consecutive records are 32*S bytes contiguous
                    if(j + S > n) j = n - S;
                    for(int l = 0; l < S; l++) row[g * S + l] = (int)(j + l);
                }
            }
        }
        cudaMemcpy(d_idx, h.data(), ni * sizeof(int), cudaMemcpyHostToDevice);
        for(int w = 0; w < 3; w++) k_gather<<<(n + 127) / 128, 128>>>(n, K, d_idx, d_pos, d_out);
        cudaEventRecord(a);
        for(int r = 0; r < 10; r++) k_gather<<<(n + 127) / 128, 128>>>(n, K, d_idx, d_pos, d_out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double per = ms / 10 * 1e-3 * 1.9e9 * 148 / ((double)(n / 32) * K);   // SM-cycles per warp gather at ~1.9 GHz
        printf("S=%2d lanes per contiguous run: %.3f ms per launch, ~%.1f SM-cycles per warp-gather (incl. id load)\n", S, ms / 10, per);
    }
    return 0;
}
