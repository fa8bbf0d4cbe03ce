// K5: preconditioned conjugate gradients on the Dirichlet-eliminated operator
//   A = P K P + reg*P   (P = projector on free DOFs; K_ff + reg*I of src/fea_solver.py:118-125)
// replacing spsolve (src/fea_solver.py:128) / KSPCG (src/fea_petsc.cpp:323-341).
//
// Three kernels per iteration, each one pass over its operands:
//   1. Ap = K p + reg p  fused with the p.Ap partial dot          (CSR-stream SpMV, spmv.cuh)
//   2. alpha = rz/pAp;  x += alpha p;  r -= alpha Ap (0 on known rows);  z = M^-1 r;
//      partial dots r.z and r.r                                  (one pass over x, r, p, Ap)
//   3. beta = rz'/rz;  p = z + beta p                            (z recomputed for Jacobi)
// Known rows carry dinv == 0, which keeps r, z, p identically zero there, so the SpMV needs no
// mask.  All scalars (alpha, beta, convergence flag) stay on the device: reductions are finished
// by the last block of the producing kernel in a fixed order (bit-reproducible for a given
// grid), the host only polls a pinned copy every few dozen iterations, and once `done` is set
// every queued kernel returns immediately.  On a distributed context the three scalars are
// NCCL all-reduced in place and p's halo is refreshed before each SpMV.
#include <stdlib.h>

#include "common.cuh"
#include "spmv.cuh"
#include "spmv_tma.cuh"
#include "amg.cuh"

int myc_dist_allreduce_dev(myc_ctx* ctx, double* d_buf, int n, cudaStream_t st);   // dist.cu
int myc_dist_halo(myc_ctx* ctx, double* d_x_global, cudaStream_t st);              // dist.cu
int myc_pcg_fused_try(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                      const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                      const double* d_dinv, const double* d_binv, int pc, double reg, int64_t maxit, double* d_x,
                      cudaStream_t st, int* handled, int* op_used);                 // pcg_fused.cu

namespace {

constexpr int VEC_THREADS = 256;

// ---- SpMV epilogues ---------------------------------------------------------------------
struct EpiCgAp {   // Ap = K p + reg p ; acc0 = p.Ap
  static constexpr int NACC = 1;
  double* Ap;
  const double* pg;
  int64_t row_offset;
  double reg;
  struct Pre { double pi; };
  __device__ __forceinline__ Pre load(int64_t r) const { return Pre{pg[row_offset + r]}; }
  __device__ __forceinline__ void row(int64_t r, double s, const Pre& pre, double (&acc)[1]) const {
    const double pi = pre.pi;
    const double y = s + reg * pi;
    Ap[r] = y;
    acc[0] += pi * y;
  }
};

struct EpiResid {  // r = b - (K x + reg x) on free rows, 0 on known rows ; acc = {r.r, b.b}
  static constexpr int NACC = 2;
  double* r_out;
  const double* b;
  const double* dinv;
  const double* xg;
  int64_t row_offset;
  double reg;
  struct Pre { double bi, di, xi; };
  __device__ __forceinline__ Pre load(int64_t r) const { return Pre{b[r], dinv[r], xg[row_offset + r]}; }
  __device__ __forceinline__ void row(int64_t r, double s, const Pre& pre, double (&acc)[2]) const {
    const double bi = pre.bi;
    const double res = pre.di != 0.0 ? bi - (s + reg * pre.xi) : 0.0;
    r_out[r] = res;
    acc[0] += res * res;
    acc[1] += bi * bi;
  }
};

// ---- vector kernels ------------------------------------------------------------------------
// z = M^-1 r for the rows of node `nd` (block3) or row i (jacobi)
__device__ __forceinline__ void apply_block3(const double* __restrict__ binv, int64_t nd,
                                             const double r[3], double z[3]) {
  const double* m = binv + 9 * nd;
  z[0] = m[0] * r[0] + m[1] * r[1] + m[2] * r[2];
  z[1] = m[3] * r[0] + m[4] * r[1] + m[5] * r[2];
  z[2] = m[6] * r[0] + m[7] * r[1] + m[8] * r[2];
}

// z = M^-1 r, p = z, partial r.z; also publishes tol2 / done for a zero right-hand side
template <bool BLOCK3>
__global__ void __launch_bounds__(VEC_THREADS)
pcg_init_kernel(int64_t n_rows, int64_t row_offset, const double* __restrict__ r,
                const double* __restrict__ dinv, const double* __restrict__ binv,
                double* __restrict__ pg, double* __restrict__ zv, double* partials, PcgScalars* sc) {
  __shared__ double s_warp[VEC_THREADS / 32];
  double acc[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if constexpr (BLOCK3) {
    for (int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nd < n_rows / 3; nd += stride) {
      const double rr[3] = {r[3 * nd], r[3 * nd + 1], r[3 * nd + 2]};
      double z[3];
      apply_block3(binv, nd, rr, z);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        zv[3 * nd + c] = z[c];
        pg[row_offset + 3 * nd + c] = z[c];
        acc[0] += rr[c] * z[c];
      }
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += stride) {
      const double ri = r[i], z = dinv[i] * ri;
      pg[row_offset + i] = z;
      acc[0] += ri * z;
    }
  }
  acc[0] = myc_block_reduce(acc[0], s_warp);
  myc_finalize_partials<1>(partials, acc, &sc->counter, &sc->red[0], s_warp);
}

// after the (all-reduced) r.r and b.b are known: tolerance and the trivial-convergence case
__global__ void pcg_set_tol_kernel(PcgScalars* sc, double rtol, double atol) {
  const double t = fmax(rtol * rtol * sc->bb, atol * atol);
  sc->tol2 = t;
  sc->red[1] = sc->out[0];
  sc->rr_final = sc->out[0];
  if (sc->out[0] <= t) { sc->done = 1; sc->iters = 0; }
}

// roll r.z -> rz_old (runs between the update of iteration k and the SpMV of k+1 would race
// with readers, so it is its own 1-thread launch right before the SpMV)
__global__ void pcg_roll_kernel(PcgScalars* sc) {
  if (sc->done) return;
  sc->rz_old = sc->red[0];
}

template <bool BLOCK3>
__global__ void __launch_bounds__(VEC_THREADS)
pcg_update_kernel(int64_t n_rows, int64_t row_offset, const double* __restrict__ pg,
                  const double* __restrict__ Ap, const double* __restrict__ dinv,
                  const double* __restrict__ binv, double* __restrict__ x, double* __restrict__ r,
                  double* __restrict__ zv, double* partials, PcgScalars* sc) {
  __shared__ double s_warp[VEC_THREADS / 32];
  if (sc->done) return;
  const double pAp = sc->pAp;
  double alpha = sc->rz_old / pAp;
  if (!(pAp > 0.0) || !isfinite(alpha)) {          // breakdown: freeze, let the host report it
    alpha = 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) sc->breakdown = 1;
  }
  double acc[2] = {0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if constexpr (BLOCK3) {
    for (int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nd < n_rows / 3; nd += stride) {
      double rr[3], z[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int64_t i = 3 * nd + c;
        x[i] += alpha * pg[row_offset + i];
        rr[c] = dinv[i] != 0.0 ? r[i] - alpha * Ap[i] : 0.0;
        r[i] = rr[c];
      }
      apply_block3(binv, nd, rr, z);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        zv[3 * nd + c] = z[c];
        acc[0] += rr[c] * z[c];
        acc[1] += rr[c] * rr[c];
      }
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += stride) {
      const double d = dinv[i];
      x[i] += alpha * pg[row_offset + i];
      const double ri = d != 0.0 ? r[i] - alpha * Ap[i] : 0.0;
      r[i] = ri;
      acc[0] += ri * (d * ri);
      acc[1] += ri * ri;
    }
  }
  acc[0] = myc_block_reduce(acc[0], s_warp);
  acc[1] = myc_block_reduce(acc[1], s_warp);
  myc_finalize_partials<2>(partials, acc, &sc->counter, &sc->red[0], s_warp);
}

template <bool BLOCK3>
__global__ void __launch_bounds__(VEC_THREADS)
pcg_direction_kernel(int64_t n_rows, int64_t row_offset, const double* __restrict__ r,
                     const double* __restrict__ dinv, const double* __restrict__ zv,
                     double* __restrict__ pg, PcgScalars* sc, long long iter_done) {
  if (sc->done) return;
  if (sc->red[1] <= sc->tol2 || sc->breakdown) {   // uniform across the grid
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      sc->iters = iter_done;
      sc->rr_final = sc->red[1];
      __threadfence();
      sc->done = 1;
    }
    // Other blocks may still be reading sc->done == 0 above; they take this same branch
    // because red[1]/tol2/breakdown are stable during this launch.
    return;
  }
  const double beta = sc->red[0] / sc->rz_old;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += stride) {
    const double z = BLOCK3 ? zv[i] : dinv[i] * r[i];
    pg[row_offset + i] = z + beta * pg[row_offset + i];
  }
}

__global__ void __launch_bounds__(VEC_THREADS)
copy_into_global_kernel(int64_t n_rows, int64_t row_offset, const double* __restrict__ src,
                        double* __restrict__ dst_global) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += stride)
    dst_global[row_offset + i] = src[i];
}

struct HostScalars {  // pinned mirror
  PcgScalars s;
};

}  // namespace

int myc_launch_spmv(myc_ctx* ctx, int64_t n_rows, const int32_t* rp, const int32_t* ci,
                    const double* v, const double* x, double* y, cudaStream_t st) {
  if (n_rows == 0) return MYC_OK;
  EpiPlain epi{y};
  return myc_launch_spmv_epi<EpiPlain>(ctx, n_rows, rp, ci, v, x, epi, nullptr, nullptr, nullptr, nullptr, st);
}

extern "C" int myc_spmv(myc_ctx* ctx, int64_t n_rows, const int32_t* d_row_ptr, const int32_t* d_col_idx,
                        const double* d_val, const double* d_x, double* d_y, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n_rows < 0 || !d_row_ptr || (n_rows > 0 && (!d_x || !d_y)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "spmv: bad argument");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  return myc_launch_spmv(ctx, n_rows, d_row_ptr, d_col_idx, d_val, d_x, d_y, (cudaStream_t)stream);
}

static int prepare_solver_buffers(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, bool block3) {
  MYC_TRY(myc_ensure(ctx, ctx->vec[0], (size_t)(n_cols_global + 1) * sizeof(double)));   // p (global)
  MYC_TRY(myc_ensure(ctx, ctx->vec[1], (size_t)(n_rows + 1) * sizeof(double)));          // r
  MYC_TRY(myc_ensure(ctx, ctx->vec[2], (size_t)(n_rows + 1) * sizeof(double)));          // Ap
  if (block3) MYC_TRY(myc_ensure(ctx, ctx->vec[3], (size_t)(n_rows + 1) * sizeof(double)));  // z
  MYC_TRY(myc_ensure(ctx, ctx->partials, (size_t)ctx->sm_count * 16 * 4 * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->scalars, sizeof(PcgScalars)));
  return MYC_OK;
}

// r = b - A x (x local; uses vec[0] as the global-length staging of x), sums to sc->out[0..1]
static int launch_residual(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                           const int32_t* rp, const int32_t* ci, const double* v, const double* b,
                           const double* dinv, double reg, const double* x, double* r_out,
                           cudaStream_t st) {
  double* pg = (double*)ctx->vec[0].p;
  PcgScalars* sc = (PcgScalars*)ctx->scalars.p;
  const int vgrid = grid_for(ctx, ceil_div64(n_rows, VEC_THREADS), 8);
  if (ctx->world > 1) MYC_CUDA(ctx, cudaMemsetAsync(pg, 0, (size_t)n_cols_global * sizeof(double), st));
  if (n_rows > 0) {
    copy_into_global_kernel<<<vgrid, VEC_THREADS, 0, st>>>(n_rows, row_offset, x, pg);
    MYC_LAUNCHED(ctx);
  }
  if (ctx->world > 1) MYC_TRY(myc_dist_halo(ctx, pg, st));
  EpiResid epi{r_out, b, dinv, pg, row_offset, reg};
  MYC_TRY(myc_launch_spmv_epi<EpiResid>(ctx, n_rows, rp, ci, v, pg, epi, (double*)ctx->partials.p, &sc->counter,
                                        &sc->out[0], nullptr, st));
  if (ctx->world > 1) MYC_TRY(myc_dist_allreduce_dev(ctx, &sc->out[0], 2, st));
  return MYC_OK;
}

extern "C" int myc_pcg_solve(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                             const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                             const double* d_rhs, const double* d_dinv, const double* d_binv,
                             int precond, double reg, double rtol, double atol, int64_t maxit,
                             double* d_x, int64_t* h_out_iters, double* h_out_relres, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n_rows == 0 && (precond == MYC_PC_BLOCK6 || precond == MYC_PC_BLOCK12)) precond = MYC_PC_JACOBI;   // nothing to solve
  const bool block3 = precond == MYC_PC_BLOCK3;
  const bool group = precond == MYC_PC_BLOCK6 || precond == MYC_PC_BLOCK12;     // fused single-GPU kernel only
  const bool amg = precond == MYC_PC_AMG;                                       // hierarchy from myc_amg_setup
  if (n_rows < 0 || n_cols_global < n_rows || row_offset < 0 || row_offset + n_rows > n_cols_global ||
      !d_row_ptr || (n_rows > 0 && (!d_rhs || !d_dinv || !d_x)) || maxit < 0 ||
      (precond != MYC_PC_JACOBI && !block3 && !group && !amg) || (block3 && (!d_binv || n_rows % 3)) ||
      (group && n_rows > 0 && !d_binv))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "pcg_solve: bad argument");
  // the persistent kernel's barrier epochs are 32-bit counters of (barriers x blocks): 2 barriers per
  // iteration x 148 blocks wrap after ~14.5 M iterations
  if (maxit > MYC_MAXIT_LIMIT)
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "pcg_solve: maxit %lld exceeds the supported %lld", (long long)maxit, (long long)MYC_MAXIT_LIMIT);
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  MYC_TRY(prepare_solver_buffers(ctx, n_rows, n_cols_global, block3));
  double* pg = (double*)ctx->vec[0].p;
  double* r = (double*)ctx->vec[1].p;
  double* Ap = (double*)ctx->vec[2].p;
  double* zv = block3 ? (double*)ctx->vec[3].p : nullptr;
  double* partials = (double*)ctx->partials.p;
  PcgScalars* sc = (PcgScalars*)ctx->scalars.p;
  PcgScalars* h_sc = (PcgScalars*)ctx->h_pinned;       // two pinned slots, 512 B apart
  const bool dist = ctx->world > 1;

  const int64_t vec_items = block3 ? n_rows / 3 : n_rows;
  const int vgrid = grid_for(ctx, ceil_div64(vec_items, VEC_THREADS), 8);
  const int dgrid = grid_for(ctx, ceil_div64(n_rows, VEC_THREADS), 8);

  MYC_CUDA(ctx, cudaMemsetAsync(sc, 0, sizeof(PcgScalars), st));
  // r0 = b - A x0, r.r -> out[0], b.b -> out[1]
  MYC_TRY(launch_residual(ctx, n_rows, n_cols_global, row_offset, d_row_ptr, d_col_idx, d_val, d_rhs,
                          d_dinv, reg, d_x, r, st));
  MYC_CUDA(ctx, cudaMemcpyAsync(&sc->bb, &sc->out[1], sizeof(double), cudaMemcpyDeviceToDevice, st));
  pcg_set_tol_kernel<<<1, 1, 0, st>>>(sc, rtol, atol);
  MYC_LAUNCHED(ctx);
  // ---- the whole iteration loop is one persistent cooperative kernel per GPU (NVLink peer
  // memory between GPUs); falls through to the multi-kernel / NCCL loop when not applicable
  {
    int handled = 0, op_used = 0;
    if (ctx->prof_on) MYC_CUDA(ctx, cudaEventRecord(ctx->prof_ev[0], st));
    if (amg) {
      op_used = 2;
      MYC_TRY(myc_pcg_amg_try(ctx, n_rows, n_cols_global, row_offset, d_row_ptr, d_dinv, reg, maxit, d_x, st, &handled));
    } else {
      MYC_TRY(myc_pcg_fused_try(ctx, n_rows, n_cols_global, row_offset, d_row_ptr, d_col_idx, d_val, d_dinv,
                                (block3 || group) ? d_binv : nullptr, precond, reg, maxit, d_x, st, &handled, &op_used));
    }
    if (handled) {
      if (ctx->prof_on) MYC_CUDA(ctx, cudaEventRecord(ctx->prof_ev[1], st));
      MYC_CUDA(ctx, cudaMemcpyAsync(h_sc, sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
      MYC_CUDA(ctx, cudaStreamSynchronize(st));
      const PcgScalars fin = *h_sc;
      if (dist && !amg) {
        ctx->peer_epoch_red = (unsigned)fin.pAp;
        ctx->peer_epoch_halo = (unsigned)fin.rz_old;
      }
      if (dist && amg) {
        ctx->amg_epoch_red = (unsigned)fin.pAp;
        ctx->amg_epoch_halo = (unsigned)fin.rz_old;
        ctx->amg_epoch_seam = (unsigned)fin.out[3];
      }
      if (getenv("MYC_FUSED_TIMING_PRINT"))
        fprintf(stderr, "[fused] block0 ns/iter: sweep %.0f  barrier+reduce %.0f  vector %.0f  barrier %.0f  (iters %lld)\n",
                fin.out[0], fin.out[1], fin.out[2], fin.out[3], (long long)fin.iters);
      const double rel = fin.bb > 0.0 ? sqrt(fin.rr_final / fin.bb) : 0.0;
      if (ctx->prof_on) {
        float ms = 0.f;
        int32_t h_nnz = 0;
        MYC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[0], ctx->prof_ev[1]));
        MYC_CUDA(ctx, cudaMemcpy(&h_nnz, d_row_ptr + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost));
        ctx->prof_ms += ms;
        ctx->prof_samples += 1;
        ctx->prof_launches += 1;
        // bytes the sweep streams per iteration: CSR 12 B/nnz, symmetric block view 52 B per 9 nnz
        const double mat = op_used == 2 ? (52.0 / 9.0) * h_nnz : 12.0 * h_nnz;
        if (amg) ctx->prof_bytes += ((double)fin.iters + 1.0) * myc_amg_bytes_per_iteration(ctx);
        else ctx->prof_bytes += ((double)fin.iters + 1.0) * (mat + 20.0 * (double)n_rows) +
                           (double)fin.iters * (double)n_rows *           // + the inverse blocks of the preconditioner
                               // algorithmic minimum: the symmetric inverse, R(R+1)/2 doubles per R rows
                               // (the 6x6 blocks are STORED row by row, 48 B per row; not counted)
                               (precond == MYC_PC_BLOCK12 ? 148.0 : precond == MYC_PC_BLOCK6 ? 124.0 : block3 ? 120.0 : 96.0);
        ctx->prof_op = op_used;
      }
      if (h_out_iters) *h_out_iters = (int64_t)fin.iters;
      if (h_out_relres) *h_out_relres = rel;
      if (fin.breakdown || !(rel == rel) || !isfinite(fin.rr_final))
        MYC_FAIL(ctx, MYC_ERR_BREAKDOWN, "pcg_solve (fused): breakdown after %lld iterations, r.r = %g",
                 (long long)fin.iters, fin.rr_final);
      if (!fin.done)
        MYC_FAIL(ctx, MYC_ERR_NOT_CONVERGED, "pcg_solve: %lld iterations, ||r||/||b|| = %.3e > rtol",
                 (long long)fin.iters, rel);
      return MYC_OK;
    }
  }
  if (amg) MYC_FAIL(ctx, MYC_ERR_STATE, "pcg_solve(MYC_PC_AMG): the persistent multigrid solver kernel is unavailable here");
  if (group)
    MYC_FAIL(ctx, MYC_ERR_STATE, "pcg_solve: MYC_PC_BLOCK6 / MYC_PC_BLOCK12 run only in the single-GPU persistent solver "
                                 "kernel, which is unavailable here (multi-GPU context, MYC_NO_FUSED_PCG, unaligned arrays "
                                 "or no cooperative launch): use MYC_PC_BLOCK3");
  if (block3)
    pcg_init_kernel<true><<<vgrid, VEC_THREADS, 0, st>>>(n_rows, row_offset, r, d_dinv, d_binv, pg, zv, partials, sc);
  else
    pcg_init_kernel<false><<<vgrid, VEC_THREADS, 0, st>>>(n_rows, row_offset, r, d_dinv, d_binv, pg, zv, partials, sc);
  MYC_LAUNCHED(ctx);
  if (dist) MYC_TRY(myc_dist_allreduce_dev(ctx, &sc->red[0], 1, st));

  // ---- iteration loop: enqueue `chunk` iterations, then look at the scalars of the chunk
  // before (one chunk of latency, so the GPU never waits for the host)
  const int64_t chunk = 32;
  int64_t it = 0;
  int64_t n_snap = 0;
  int prof_used = 0;
  PcgScalars last{};
  for (;;) {
    const int64_t upto = (it + chunk < maxit) ? it + chunk : maxit;
    for (; it < upto; ++it) {
      pcg_roll_kernel<<<1, 1, 0, st>>>(sc);
      MYC_LAUNCHED(ctx);
      if (dist) MYC_TRY(myc_dist_halo(ctx, pg, st));
      EpiCgAp epi{Ap, pg, row_offset, reg};
      const bool sample = ctx->prof_on && (it % 32) == 8 && prof_used < myc_ctx::PROF_PAIRS;
      if (sample) MYC_CUDA(ctx, cudaEventRecord(ctx->prof_ev[2 * prof_used], st));
      MYC_TRY(myc_launch_spmv_epi<EpiCgAp>(ctx, n_rows, d_row_ptr, d_col_idx, d_val, pg, epi, partials, &sc->counter,
                                           &sc->pAp, &sc->done, st));
      if (sample) MYC_CUDA(ctx, cudaEventRecord(ctx->prof_ev[2 * prof_used++ + 1], st));
      if (dist) MYC_TRY(myc_dist_allreduce_dev(ctx, &sc->pAp, 1, st));
      if (block3)
        pcg_update_kernel<true><<<vgrid, VEC_THREADS, 0, st>>>(n_rows, row_offset, pg, Ap, d_dinv, d_binv, d_x, r, zv, partials, sc);
      else
        pcg_update_kernel<false><<<vgrid, VEC_THREADS, 0, st>>>(n_rows, row_offset, pg, Ap, d_dinv, d_binv, d_x, r, zv, partials, sc);
      MYC_LAUNCHED(ctx);
      if (dist) MYC_TRY(myc_dist_allreduce_dev(ctx, &sc->red[0], 2, st));
      if (block3)
        pcg_direction_kernel<true><<<dgrid, VEC_THREADS, 0, st>>>(n_rows, row_offset, r, d_dinv, zv, pg, sc, (long long)(it + 1));
      else
        pcg_direction_kernel<false><<<dgrid, VEC_THREADS, 0, st>>>(n_rows, row_offset, r, d_dinv, zv, pg, sc, (long long)(it + 1));
      MYC_LAUNCHED(ctx);
    }
    // snapshot this chunk's scalars into pinned slot (n_snap & 1) ...
    MYC_CUDA(ctx, cudaMemcpyAsync((char*)h_sc + 512 * (n_snap & 1), sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaEventRecord(ctx->ev[n_snap & 1], st));
    ++n_snap;
    // ... and look at the snapshot of the chunk before it
    if (n_snap >= 2) {
      MYC_CUDA(ctx, cudaEventSynchronize(ctx->ev[(n_snap - 2) & 1]));
      last = *(const PcgScalars*)((const char*)h_sc + 512 * ((n_snap - 2) & 1));
      if (last.done) break;
    }
    if (it >= maxit) break;
  }
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  if (!last.done && n_snap >= 1) last = *(const PcgScalars*)((const char*)h_sc + 512 * ((n_snap - 1) & 1));
  if (ctx->prof_on) {
    // only launches that ran before convergence did the work (later ones return immediately)
    const int64_t live = last.done ? (int64_t)last.iters : it;
    int64_t nnz_local = 0;
    {
      int32_t h_nnz = 0;
      MYC_CUDA(ctx, cudaMemcpy(&h_nnz, d_row_ptr + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost));
      nnz_local = h_nnz;
    }
    for (int k = 0; k < prof_used; ++k) {
      if ((int64_t)k * 32 + 8 >= live) break;
      float ms = 0.f;
      MYC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[2 * k], ctx->prof_ev[2 * k + 1]));
      ctx->prof_ms += ms;
      ctx->prof_bytes += 12.0 * (double)nnz_local + 20.0 * (double)n_rows;
      ctx->prof_samples++;
    }
    ctx->prof_launches += live;
  }
  const double rr_last = last.done ? last.rr_final : last.red[1];
  const double relres = last.bb > 0.0 ? sqrt(rr_last / last.bb) : 0.0;
  if (h_out_iters) *h_out_iters = last.done ? (int64_t)last.iters : it;
  if (h_out_relres) *h_out_relres = relres;
  if (last.breakdown || !(relres == relres))
    MYC_FAIL(ctx, MYC_ERR_BREAKDOWN, "pcg_solve: breakdown (p.Ap = %g, r.r = %g) after %lld iterations",
             last.pAp, rr_last, (long long)(last.done ? last.iters : it));
  if (!last.done) {
    if (rr_last <= last.tol2) return MYC_OK;         // converged exactly at maxit
    MYC_FAIL(ctx, MYC_ERR_NOT_CONVERGED, "pcg_solve: %lld iterations, ||r||/||b|| = %.3e > rtol", (long long)it, relres);
  }
  return MYC_OK;
}

extern "C" int myc_true_residual(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                                 const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                                 const double* d_rhs, const double* d_dinv, double reg, const double* d_x,
                                 double* h_out_relres, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n_rows < 0 || n_cols_global < n_rows || !d_row_ptr || !h_out_relres ||
      (n_rows > 0 && (!d_rhs || !d_dinv || !d_x)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "true_residual: bad argument");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  MYC_TRY(prepare_solver_buffers(ctx, n_rows, n_cols_global, false));
  PcgScalars* sc = (PcgScalars*)ctx->scalars.p;
  MYC_CUDA(ctx, cudaMemsetAsync(sc, 0, sizeof(PcgScalars), st));
  MYC_TRY(launch_residual(ctx, n_rows, n_cols_global, row_offset, d_row_ptr, d_col_idx, d_val, d_rhs,
                          d_dinv, reg, d_x, (double*)ctx->vec[2].p, st));
  double* h = (double*)ctx->h_pinned;
  MYC_CUDA(ctx, cudaMemcpyAsync(h, &sc->out[0], 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  *h_out_relres = h[1] > 0.0 ? sqrt(h[0] / h[1]) : 0.0;
  return MYC_OK;
}
