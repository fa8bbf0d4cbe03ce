// CSR SpMV building block ("CSR-stream"): a block takes SP_ROWS consecutive rows, streams their
// contiguous nnz range (values + column indices, fully coalesced, evict-first) through shared
// memory as products val*x[col], then one thread per row sums its products left to right.
// Rows of this problem are short and uniform (3*(neighbours+1), ~11 on average), so this keeps
// every lane busy on the bandwidth-critical stream and makes the per-row sum order fixed
// (bit-reproducible).  The epilogue functor fuses what follows the SpMV in the caller
// (CG: +reg*p and the p.Ap partial dot; Dirichlet RHS; true residual).
#pragma once
#include "common.cuh"

constexpr int SP_THREADS = 256;
constexpr int SP_ROWS = 256;    // rows per tile == threads, one row per thread in the reduce phase
constexpr int SP_CAP = 4608;    // products staged per tile (36 KB) -> 18 nnz/row on average

__device__ __forceinline__ double myc_block_reduce(double v, double* s_warp) {
  // fixed-order tree: shuffle within warps, then warp 0 adds the 8 warp sums in order.
  // Result valid in thread 0 only.
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                      // s_warp may still be read from a previous call
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < SP_THREADS / 32; ++w) t += s_warp[w];
  }
  return t;
}

// Last-arriving block sums the per-block partials (fixed order) into out[0..NACC).
template <int NACC>
__device__ __forceinline__ void myc_finalize_partials(double* partials, const double (&acc)[NACC],
                                                      unsigned* counter, double* out, double* s_warp) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) partials[(size_t)blockIdx.x * NACC + j] = acc[j];
    __threadfence();
    const unsigned t = atomicInc(counter, gridDim.x - 1);   // wraps to 0 after the last block
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      double s = 0.0;
      for (unsigned b = threadIdx.x; b < gridDim.x; b += SP_THREADS) s += __ldcg(&partials[(size_t)b * NACC + j]);
      s = myc_block_reduce(s, s_warp);
      if (threadIdx.x == 0) out[j] = s;
    }
  }
}

// Epi interface:
//   static constexpr int NACC;                     number of fused partial sums
//   struct Pre;  __device__ Pre load(int64_t r) const;     per-row operands, requested EARLY so
//                                                          that their latency overlaps the tile
//   __device__ void row(int64_t r, double sum, const Pre&, double (&acc)[NACC]) const;
template <class Epi>
__device__ __forceinline__ void myc_spmv_tiles(int64_t n_rows, const int32_t* __restrict__ rp,
                                               const int32_t* __restrict__ ci,
                                               const double* __restrict__ v,
                                               const double* __restrict__ x, const Epi& epi,
                                               double (&acc)[Epi::NACC == 0 ? 1 : Epi::NACC],
                                               double* s_prod, int32_t* s_rp) {
  const int64_t n_tiles = (n_rows + SP_ROWS - 1) / SP_ROWS;
  const int t = threadIdx.x;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * SP_ROWS;
    const int rows = (int)((n_rows - r0) < SP_ROWS ? (n_rows - r0) : SP_ROWS);
    if (t <= rows) s_rp[t] = rp[r0 + t];
    if (t == 0 && rows == SP_ROWS) s_rp[SP_ROWS] = rp[r0 + SP_ROWS];
    __syncthreads();
    const int32_t start = s_rp[0];
    const int nnz = s_rp[rows] - start;
    if (nnz <= SP_CAP) {
      const int32_t* cit = ci + start;
      const double* vt = v + start;
      int k = t;
      for (; k + 3 * SP_THREADS < nnz; k += 4 * SP_THREADS) {
        const int32_t c0 = __ldcs(cit + k), c1 = __ldcs(cit + k + SP_THREADS),
                      c2 = __ldcs(cit + k + 2 * SP_THREADS), c3 = __ldcs(cit + k + 3 * SP_THREADS);
        const double a0 = __ldcs(vt + k), a1 = __ldcs(vt + k + SP_THREADS),
                     a2 = __ldcs(vt + k + 2 * SP_THREADS), a3 = __ldcs(vt + k + 3 * SP_THREADS);
        const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
        s_prod[k] = a0 * x0;
        s_prod[k + SP_THREADS] = a1 * x1;
        s_prod[k + 2 * SP_THREADS] = a2 * x2;
        s_prod[k + 3 * SP_THREADS] = a3 * x3;
      }
      for (; k < nnz; k += SP_THREADS) s_prod[k] = __ldcs(vt + k) * __ldg(x + __ldcs(cit + k));
      __syncthreads();
      if (t < rows) {
        double s = 0.0;
        const int e = s_rp[t + 1] - start;
        for (int j = s_rp[t] - start; j < e; ++j) s += s_prod[j];
        epi.row(r0 + t, s, epi.load(r0 + t), acc);
      }
    } else {
      // tile too dense for the staging buffer: one warp per row, lanes stride the row,
      // fixed shuffle tree (still deterministic)
      const int lane = t & 31, warp = t >> 5;
      for (int r = warp; r < rows; r += SP_THREADS / 32) {
        double s = 0.0;
        for (int32_t j = s_rp[r] + lane; j < s_rp[r + 1]; j += 32) s = __dadd_rn(s, __dmul_rn(v[j], __ldg(x + ci[j])));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) epi.row(r0 + r, s, epi.load(r0 + r), acc);
      }
    }
    __syncthreads();
  }
}

// Device-resident scalars of one PCG solve (also used by the one-shot fused reductions).
struct PcgScalars {
  double pAp;       // p.Ap of the current iteration
  double red[2];    // {r.z (new), r.r} written by the update kernel
  double rz_old;    // r.z of the previous iteration
  double bb;        // b.b
  double tol2;      // convergence threshold on r.r
  double out[4];    // generic outputs of one-shot reductions
  double rr_final;  // r.r at the moment `done` was set (later all-reduces keep summing red[])
  int done;         // set once r.r <= tol2
  int breakdown;    // p.Ap <= 0 or not finite
  long long iters;  // iterations performed when `done` was set
  unsigned counter; // last-block election (self-resetting)
  unsigned pad;
};

struct EpiPlain {
  static constexpr int NACC = 0;
  double* y;
  struct Pre {};
  __device__ __forceinline__ Pre load(int64_t) const { return Pre{}; }
  __device__ __forceinline__ void row(int64_t r, double s, const Pre&, double (&)[1]) const { y[r] = s; }
};

// Generic SpMV kernel: grid-stride over row tiles, optional fused reductions finalised by the
// last block into `out` (device).  `done` (may be null) makes the launch a no-op once set.
template <class Epi>
__global__ void __launch_bounds__(SP_THREADS)
myc_spmv_kernel(int64_t n_rows, const int32_t* __restrict__ rp, const int32_t* __restrict__ ci,
                const double* __restrict__ v, const double* __restrict__ x, Epi epi,
                double* partials, unsigned* counter, double* out, const int* done) {
  __shared__ double s_prod[SP_CAP];
  __shared__ int32_t s_rp[SP_ROWS + 1];
  __shared__ double s_warp[SP_THREADS / 32];
  if (done && *done) return;
  double acc[Epi::NACC == 0 ? 1 : Epi::NACC];
#pragma unroll
  for (int j = 0; j < (Epi::NACC == 0 ? 1 : Epi::NACC); ++j) acc[j] = 0.0;
  myc_spmv_tiles<Epi>(n_rows, rp, ci, v, x, epi, acc, s_prod, s_rp);
  if constexpr (Epi::NACC > 0) {
#pragma unroll
    for (int j = 0; j < Epi::NACC; ++j) acc[j] = myc_block_reduce(acc[j], s_warp);
    myc_finalize_partials<Epi::NACC>(partials, acc, counter, out, s_warp);
  }
}

constexpr int SP_BLOCKS_PER_SM = 4;
