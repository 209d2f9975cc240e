// Microbenchmark: does the TEX path have gather throughput that is independent of the LSU path on B200?
// Each thread gathers K 32-byte records at pseudo-random indices inside a window around its own index
// (the access pattern of a cell-sorted neighbour list) via (a) LDG.256, (b) two tex1Dfetch<int4>, (c) a mix.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ double4 ld256(const double4* p) {
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double4 ldtex(cudaTextureObject_t t, int j) {
    int4 a = tex1Dfetch<int4>(t, 2 * j), b = tex1Dfetch<int4>(t, 2 * j + 1);
    return make_double4(__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x), __hiloint2double(b.w, b.z));
}

template<int MODE, int TEXMOD>
__global__ void __launch_bounds__(128) k_gather(int n, int K, const int* __restrict__ idx, const double4* __restrict__ pos,
                                                cudaTextureObject_t tex, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const int* nb = idx + (size_t)(i / 32) * K * 32 + (i % 32);
    double sx = 0, sy = 0, sz = 0;
    for(int k = 0; k + 4 <= K; k += 4) {
        int j[4]; double4 p[4];
#pragma unroll
        for(int u = 0; u < 4; u++) j[u] = __ldg(nb + (size_t)(k + u) * 32);
#pragma unroll
        for(int u = 0; u < 4; u++) {
            bool usetex = (MODE == 1) || (MODE == 2 && (u % TEXMOD) == 0);
            p[u] = usetex ? ldtex(tex, j[u]) : ld256(pos + j[u]);
        }
#pragma unroll
        for(int u = 0; u < 4; u++) { sx += p[u].x; sy += p[u].y; sz += p[u].z + p[u].w; }
    }
    out[i] = sx + sy + sz;
}

int main() {
    const int n = 4000000, K = 76, W = 1500;
    std::vector<int> h((size_t)((n + 31) / 32) * K * 32);
    srand(1);
    for(int i = 0; i < n; i++) {
        // sorted pseudo-neighbours inside a +-W window, as in a cell-sorted list
        std::vector<int> js(K);
        for(int k = 0; k < K; k++) { long j = (long)i - W + (long)(2.0 * W * k / K) + rand() % (2 * W / K); if(j < 0) j += n; if(j >= n) j -= n; js[k] = (int)j; }
        for(int k = 0; k < K; k++) h[(size_t)(i / 32) * K * 32 + (size_t)k * 32 + (i % 32)] = js[k];
    }
    int* d_idx; double4* d_pos; double* d_out;
    cudaMalloc(&d_idx, h.size() * sizeof(int)); cudaMalloc(&d_pos, (size_t)n * 32); cudaMalloc(&d_out, (size_t)n * 8);
    cudaMemcpy(d_idx, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemset(d_pos, 0, (size_t)n * 32);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = d_pos;
    rd.res.linear.desc = cudaCreateChannelDesc<int4>(); rd.res.linear.sizeInBytes = (size_t)n * 32;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    printf("texture object: %s\n", cudaGetErrorString(e));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char* name, auto kern) {
        for(int w = 0; w < 3; w++) kern<<<(n + 127) / 128, 128>>>(n, K, d_idx, d_pos, tex, d_out);
        cudaEventRecord(a);
        for(int r = 0; r < 10; r++) kern<<<(n + 127) / 128, 128>>>(n, K, d_idx, d_pos, tex, d_out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-28s %.3f ms per launch (%s)\n", name, ms / 10, cudaGetErrorString(cudaGetLastError()));
    };
    run("LDG.256 only", k_gather<0, 1>);
    run("TEX only (2x int4)", k_gather<1, 1>);
    run("mix: 1 of 2 via TEX", k_gather<2, 2>);
    run("mix: 1 of 4 via TEX", k_gather<2, 4>);
    return 0;
}
