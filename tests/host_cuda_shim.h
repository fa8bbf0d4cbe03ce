// Just enough of the CUDA language for g++ to compile selected kernel SOURCE TEXT of csrc/ as host code
// (tests/test_kernel_logic_host.py).  A kernel runs as one "thread" of a 1 x 1 grid, so its grid-stride loop
// visits every item; rounding intrinsics map to the IEEE operations they name (compile with -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstdint>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
struct HostDim3 { unsigned x = 0, y = 0, z = 0; };
static const HostDim3 blockIdx{0, 0, 0}, threadIdx{0, 0, 0};
static const HostDim3 blockDim{1, 1, 1}, gridDim{1, 1, 1};
using std::isfinite;
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int atomicAdd(int* p, int v) { const int old = *p; *p += v; return old; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long old = *p; *p += v; return old; }
// the one emulated thread already holds the whole grid's partial: every other lane of a warp reduction contributes zero
template <class T> static inline T __shfl_down_sync(unsigned, T, int) { return T(0); }
