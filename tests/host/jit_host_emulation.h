// Host-side stand-ins for the few CUDA constructs the generated user kernels use (pairs_b200/kernelgen.py + the prelude of
// csrc/jit.cu), so that the CPU test-suite can EXECUTE generated kernels -- one "thread" at a time -- and compare them bit for
// bit with the oracle.  Test infrastructure only.  Compile with -ffp-contract=off (the device build uses --fmad=false).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>

struct double2 { double x, y; };
struct double4 { double x, y, z, w; };
struct PbHostIdx { int x; };
static PbHostIdx blockIdx, blockDim, threadIdx;
#define PB_INFINITY INFINITY
static inline double atomicAdd(double *p, double v) { const double old = *p; *p = old + v; return old; }     // one "thread" at a time
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(n)
#define __restrict__
template<typename T> static inline T __ldg(const T *p) { return *p; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
