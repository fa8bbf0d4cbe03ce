// K1: batched element stiffness, the stand-alone form of bar_stiffness_bulk
// (src/fea_solver.py:30-68).  One thread evaluates one element's 6 unique entries of S; the
// block then expands them to the (6,6) row-major layout through shared memory so that the
// 288 B/element output is written with fully coalesced stores.  No tensor cores: each
// element is ~70 flops against 345 B of traffic (HBM-write-bound).
#include "common.cuh"
#include "ke.cuh"

namespace {

constexpr int KE_THREADS = 128;

__global__ void __launch_bounds__(KE_THREADS)
ke_batch_kernel(const double* __restrict__ p1s, const double* __restrict__ p2s, int64_t n,
                double E, double A, double I, double* __restrict__ out_ke, double* __restrict__ out_L) {
  __shared__ double s6[KE_THREADS][7];   // padded: 7 doubles per element -> conflict-free reads
  const BarConsts c = myc_bar_consts(E, A, I);
  const int64_t n_tiles = (n + KE_THREADS - 1) / KE_THREADS;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t e = tile * KE_THREADS + threadIdx.x;
    if (e < n) {
      double L;
      const Sym3 s = myc_bar_block(p1s[3 * e], p1s[3 * e + 1], p1s[3 * e + 2], p2s[3 * e],
                                   p2s[3 * e + 1], p2s[3 * e + 2], c, &L);
      out_L[e] = L;
      s6[threadIdx.x][0] = s.xx; s6[threadIdx.x][1] = s.xy; s6[threadIdx.x][2] = s.xz;
      s6[threadIdx.x][3] = s.yy; s6[threadIdx.x][4] = s.yz; s6[threadIdx.x][5] = s.zz;
    }
    __syncthreads();
    const int64_t tile_elems = (n - tile * KE_THREADS) < KE_THREADS ? (n - tile * KE_THREADS) : KE_THREADS;
    double* dst = out_ke + tile * KE_THREADS * 36;
    for (int idx = threadIdx.x; idx < (int)tile_elems * 36; idx += KE_THREADS) {
      const int el = idx / 36, ij = idx - el * 36;
      const int i = ij / 6, j = ij - i * 6;
      const int a = i % 3, b = j % 3;
      const int lo = a < b ? a : b, hi = a < b ? b : a;
      const int u = lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);   // (0,*)->0..2 (1,1)->3 (1,2)->4 (2,2)->5
      const double v = s6[el][u];
      dst[idx] = ((i < 3) == (j < 3)) ? v : -v;
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int myc_bar_stiffness_bulk(myc_ctx* ctx, const double* d_p1s, const double* d_p2s,
                                      int64_t n, double E, double A, double I, double* d_out_ke,
                                      double* d_out_L, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n < 0 || (n > 0 && (!d_p1s || !d_p2s || !d_out_ke || !d_out_L)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "bar_stiffness_bulk: null pointer or negative n");
  if (n == 0) return MYC_OK;
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int grid = grid_for(ctx, ceil_div64(n, KE_THREADS), 8);
  ke_batch_kernel<<<grid, KE_THREADS, 0, (cudaStream_t)stream>>>(d_p1s, d_p2s, n, E, A, I, d_out_ke, d_out_L);
  MYC_LAUNCHED(ctx);
  return MYC_OK;
}
