// K5, single-kernel form: the whole Jacobi-PCG solve is ONE persistent cooperative launch
// (one 1024-thread block per SM).  Motivation (profiles/r1_pcg512_launches.md): on the
// 512^2 benchmark operator (L2-resident) an iteration of the three-kernel PCG costs ~47 us of
// which the SpMV body is only about half; the rest is launch gaps, per-kernel prologues and
// last-block reductions.  Here an iteration is
//
//   A. w = K u + reg u   (TMA-pipelined sweep, spmv_tma.cuh) with gamma = r.u, delta = w.u and
//      r.r folded into the epilogue                                   -> grid barrier + reduce
//   B. beta = gamma/gamma_old, alpha = gamma/(delta - beta gamma/alpha_old)
//      p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;  u = M^-1 r   -> grid barrier
//
// i.e. the single-reduction recurrence of Chronopoulos & Gear (one global reduction and two
// grid barriers per iteration instead of two reductions and four launches).  The matrix stream
// never stops: each warp requests the first tile of the next sweep before it enters phase B.
// Scalars are reduced from per-block partials in a fixed order by every block, so all blocks
// take identical decisions and the result is bit-reproducible for a given grid.
// The gathered vector u is rewritten every iteration by other SMs: it is read with ordinary
// (L1-coherent-after-fence) loads, never through ld.global.nc, and the acquiring
// __threadfence() of the grid barrier invalidates the SM's L1.
#include "common.cuh"
#include "spmv.cuh"
#include "spmv_tma.cuh"

namespace {

constexpr int FU_WARPS = 32;
constexpr int FU_THREADS = 32 * FU_WARPS;
constexpr size_t FU_SMEM_BYTES = FU_WARPS * TM_SMEM_PER_WARP + FU_WARPS * TM_STAGES * sizeof(uint64_t) + 128;

struct FusedArgs {
  int64_t n_rows;
  const int32_t* rp;
  const int32_t* ci;
  const double* v;
  const double* dinv;
  double* x;
  double* r;
  double* u;      // gathered vector (n_rows)
  double* w;
  double* p;
  double* s;
  double reg;
  long long maxit;
  double* partials;     // [gridDim][3]
  unsigned* bar_counter;
  PcgScalars* sc;       // in: bb, tol2 ; out: iters, rr_final, done, breakdown
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();                        // release: this block's stores are visible gpu-wide
    atomicAdd(counter, 1u);
    unsigned spins = 0;
    while (ld_acquire_u32(counter) < target) {
      if (++spins > (1u << 27)) __trap();   // co-residency is guaranteed by the cooperative launch;
    }                                       // a lost block must fail loudly, not hang the GPU
    __threadfence();                        // acquire + L1 invalidate (CCTL.IVALL)
  }
  __syncthreads();
}

struct EpiFused {   // w = K u + reg u ; acc = {r.u, w.u, r.r}
  static constexpr int NACC = 3;
  double* w;
  const double* u;
  const double* r;
  double reg;
  struct Pre { double ui, ri; };
  __device__ __forceinline__ Pre load(int64_t i) const { return Pre{u[i], r[i]}; }
  __device__ __forceinline__ void row(int64_t i, double sum, const Pre& pre, double (&acc)[3]) const {
    const double ui = pre.ui, ri = pre.ri;
    const double wi = sum + reg * ui;
    w[i] = wi;
    acc[0] += ri * ui;
    acc[1] += wi * ui;
    acc[2] += ri * ri;
  }
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#ifdef MYC_FUSED_TIMING
#define FT_MARK(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long n_ = gtimer(); tacc[k] += n_ - tlast; tlast = n_; } } while (0)
#else
#define FT_MARK(k) do { } while (0)
#endif

__global__ void __launch_bounds__(FU_THREADS, 1) pcg_fused_kernel(FusedArgs a) {
  extern __shared__ __align__(128) unsigned char fu_smem[];
  __shared__ double s_red[FU_WARPS][3];
  __shared__ double s_tot[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n = a.n_rows;
  const int64_t gtid = (int64_t)blockIdx.x * FU_THREADS + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * FU_THREADS;
  const double tol2 = a.sc->tol2;
  unsigned bar_target = 0;

  TmPipe pp;
  tm_pipe_init(pp, fu_smem, FU_WARPS, warp, lane);
  const int32_t nnz_total = a.rp[n];
  const int64_t gw = (int64_t)blockIdx.x * FU_WARPS + warp;
  const int64_t n_warps = (int64_t)gridDim.x * FU_WARPS;

  // init: u = M^-1 r, p = s = 0
  for (int64_t i = gtid; i < n; i += gstride) {
    a.u[i] = a.dinv[i] * a.r[i];
    a.p[i] = 0.0;
    a.s[i] = 0.0;
  }
  grid_barrier(a.bar_counter, bar_target);

#ifdef MYC_FUSED_TIMING
  unsigned long long tacc[4] = {0, 0, 0, 0}, tlast = gtimer();
#endif
  double gamma_old = 1.0, alpha_old = 1.0, rr = 0.0;
  long long it = 0;
  int status = 0;   // 1 converged, 2 breakdown, 0 maxit
  EpiFused epi{a.w, a.u, a.r, a.reg};
  for (;;) {
    // ---- phase A: w = A u, partial dots
    double acc[3] = {0.0, 0.0, 0.0};
    tm_warp_sweep<EpiFused, true, true>(pp, n, a.rp, a.ci, a.v, a.u, epi, acc, gw, n_warps, lane, nnz_total);
    FT_MARK(0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double t = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (lane == 0) s_red[warp][j] = t;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = 0.0;
      for (int wq = 0; wq < FU_WARPS; ++wq) t += s_red[wq][threadIdx.x];
      a.partials[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
    grid_barrier(a.bar_counter, bar_target);
    // every block sums the per-block partials in the same fixed order
    if (warp == 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double t = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) t += __ldcg(&a.partials[(size_t)b * 3 + j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) s_tot[j] = t;
      }
    }
    __syncthreads();
    FT_MARK(1);
    const double gamma = s_tot[0], delta = s_tot[1];
    rr = s_tot[2];
    if (!(rr > tol2)) { status = 1; break; }             // converged (x, r are consistent)
    if (it >= a.maxit) { status = 0; break; }
    const double beta = (it == 0) ? 0.0 : gamma / gamma_old;
    const double denom = (it == 0) ? delta : delta - beta * gamma / alpha_old;
    if (!(denom > 0.0) || !isfinite(gamma)) { status = 2; break; }
    const double alpha = gamma / denom;
    // ---- phase B: all vector recurrences in one pass
    for (int64_t i = gtid; i < n; i += gstride) {
      const double d = a.dinv[i];
      const double pi = a.u[i] + beta * a.p[i];
      const double si = a.w[i] + beta * a.s[i];
      a.p[i] = pi;
      a.s[i] = si;
      a.x[i] += alpha * pi;
      const double ri = d != 0.0 ? a.r[i] - alpha * si : 0.0;
      a.r[i] = ri;
      a.u[i] = d * ri;
    }
    gamma_old = gamma;
    alpha_old = alpha;
    ++it;
    FT_MARK(2);
    grid_barrier(a.bar_counter, bar_target);
    FT_MARK(3);
  }
  // drain the prefetched head tile so that no bulk copy is in flight when the block exits
  if (pp.head_in_flight) {
    const int64_t n_tiles = (n + TM_ROWS - 1) / TM_ROWS;
    if (gw < n_tiles) {
      const int64_t r0 = gw * TM_ROWS;
      const int32_t lo = a.rp[r0];
      const int64_t re = r0 + TM_ROWS < n ? r0 + TM_ROWS : n;
      const int32_t hi = a.rp[re];
      const int32_t a0 = lo & ~3;
      int32_t a1 = (hi + 3) & ~3;
      if (a1 > (nnz_total & ~3)) a1 = nnz_total & ~3;
      if (hi > lo && hi - a0 <= TM_CAP && a1 > a0) tm_mbar_wait(&pp.bars[0], pp.phase_bits & 1u);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.sc->iters = it;
    a.sc->rr_final = rr;
    a.sc->red[1] = rr;
    a.sc->done = (status == 1);
    a.sc->breakdown = (status == 2);
#ifdef MYC_FUSED_TIMING
    for (int k = 0; k < 4; ++k) a.sc->out[k] = (double)tacc[k] / (double)(it > 0 ? it : 1);   // ns per iteration
#endif
  }
}

}  // namespace

// Returns MYC_OK and fills *handled = 1 if the fused path ran; *handled = 0 means "not applicable
// here" (caller falls back to the multi-kernel PCG).  On entry r = b - A x0 is in ctx->vec[1] and
// sc->bb / sc->tol2 / sc->done are set (pcg.cu does that for both paths).
int myc_pcg_fused_try(myc_ctx* ctx, int64_t n_rows, const int32_t* d_row_ptr, const int32_t* d_col_idx,
                      const double* d_val, const double* d_dinv, double reg, int64_t maxit, double* d_x,
                      cudaStream_t st, int* handled) {
  *handled = 0;
  if (ctx->world > 1 || ctx->no_fused_pcg || n_rows == 0) return MYC_OK;
  if ((((uintptr_t)d_col_idx | (uintptr_t)d_val) & 15u) != 0) return MYC_OK;
  static int max_blocks_per_sm = -1;
  if (max_blocks_per_sm < 0) {
    MYC_CUDA(ctx, cudaFuncSetAttribute(pcg_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FU_SMEM_BYTES));
    MYC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, pcg_fused_kernel, FU_THREADS,
                                                                FU_SMEM_BYTES));
  }
  int coop = 0;
  MYC_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  if (max_blocks_per_sm < 1 || !coop) return MYC_OK;
  // work vectors: u, w, p, s (r is vec[1])
  MYC_TRY(myc_ensure(ctx, ctx->vec[0], (size_t)(n_rows + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->vec[2], (size_t)(n_rows + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->vec[3], (size_t)(n_rows + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->vec[5], (size_t)(n_rows + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->misc, 256));
  const int64_t n_tiles = ceil_div64(n_rows, TM_ROWS);
  int grid = ctx->sm_count;
  if (ceil_div64(n_tiles, FU_WARPS) < grid) grid = (int)ceil_div64(n_tiles, FU_WARPS);
  if (grid < 1) grid = 1;
  MYC_TRY(myc_ensure(ctx, ctx->partials, (size_t)ctx->sm_count * 16 * 4 * sizeof(double)));
  unsigned* bar = (unsigned*)((char*)ctx->misc.p + 224);
  MYC_CUDA(ctx, cudaMemsetAsync(bar, 0, sizeof(unsigned), st));
  FusedArgs a;
  a.n_rows = n_rows;
  a.rp = d_row_ptr;
  a.ci = d_col_idx;
  a.v = d_val;
  a.dinv = d_dinv;
  a.x = d_x;
  a.r = (double*)ctx->vec[1].p;
  a.u = (double*)ctx->vec[0].p;
  a.w = (double*)ctx->vec[2].p;
  a.p = (double*)ctx->vec[3].p;
  a.s = (double*)ctx->vec[5].p;
  a.reg = reg;
  a.maxit = (long long)maxit;
  a.partials = (double*)ctx->partials.p;
  a.bar_counter = bar;
  a.sc = (PcgScalars*)ctx->scalars.p;
  void* params[] = {&a};
  MYC_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)pcg_fused_kernel, dim3(grid), dim3(FU_THREADS), params,
                                            FU_SMEM_BYTES, st));
  ctx->launches++;
  *handled = 1;
  return MYC_OK;
}
