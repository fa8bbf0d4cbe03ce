// K4: Dirichlet elimination in the full index space (replaces src/fea_solver.py:113-125 and
// plays the role of MatZeroRowsColumnsIS, src/fea_petsc.cpp:309), the reaction gather-sum
// (src/fea_solver.py:263-264), the solution merge (:131-133) and the explicit K_ff extraction
// used for structure parity (:118).
#include "common.cuh"
#include "spmv.cuh"
#include "spmv_tma.cuh"

namespace {

// [host-test-begin dirichlet_kernels]  (tests/test_kernel_logic_host.py compiles this text with g++)
constexpr int DI_THREADS = 256;

__global__ void __launch_bounds__(DI_THREADS)
scatter_known_kernel(const int64_t* __restrict__ dofs, const double* __restrict__ vals, int64_t n_known,
                     int64_t n_cols_global, int64_t row_offset, int64_t n_rows, double* __restrict__ ubc,
                     double* __restrict__ dinv, int* __restrict__ bad_flag) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_known;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t dof = dofs[k];
    if (dof < 0 || dof >= n_cols_global) { *bad_flag = 1; continue; }
    ubc[dof] = vals[k];
    const int64_t loc = dof - row_offset;
    if (loc >= 0 && loc < n_rows) dinv[loc] = 0.0;        // 0 marks a known row
  }
}

__global__ void __launch_bounds__(DI_THREADS)
fill_kernel(double* __restrict__ p, int64_t n, double v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// dinv[i] = 1/(K_ii + reg) on free rows (dinv still 1 from the fill), untouched (0) on known rows
__global__ void __launch_bounds__(DI_THREADS)
jacobi_kernel(int64_t n_rows, int64_t row_offset, const int32_t* __restrict__ rp,
              const int32_t* __restrict__ ci, const double* __restrict__ v, double reg,
              double* __restrict__ dinv) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (dinv[i] == 0.0) continue;
    const int32_t want = (int32_t)(row_offset + i);
    double d = 0.0;
    for (int32_t j = rp[i]; j < rp[i + 1]; ++j)
      if (ci[j] == want) d += v[j];
    dinv[i] = 1.0 / (d + reg);
  }
}

// [host-test-end dirichlet_kernels]

struct EpiRhs {   // b = -(K u_bc) on free rows, 0 on known rows
  static constexpr int NACC = 0;
  double* b;
  const double* dinv;
  struct Pre { double di; };
  __device__ __forceinline__ Pre load(int64_t r) const { return Pre{dinv[r]}; }
  __device__ __forceinline__ void row(int64_t r, double s, const Pre& pre, double (&)[1]) const {
    b[r] = pre.di != 0.0 ? -s : 0.0;
  }
};

__global__ void __launch_bounds__(DI_THREADS)
merge_kernel(int64_t n_rows, int64_t row_offset, const double* __restrict__ x,
             const double* __restrict__ dinv, const double* __restrict__ ubc, double* __restrict__ U) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows;
       i += (int64_t)gridDim.x * blockDim.x)
    U[row_offset + i] = dinv[i] != 0.0 ? x[i] : ubc[row_offset + i];
}

// single block, fixed order: thread t sums idx t, t+T, ... then the block tree
__global__ void __launch_bounds__(SP_THREADS)
gather_sum_kernel(const double* __restrict__ v, const int64_t* __restrict__ idx, int64_t n, double* out) {
  __shared__ double s_warp[SP_THREADS / 32];
  double s = 0.0;
  for (int64_t k = threadIdx.x; k < n; k += SP_THREADS) s += v[idx[k]];
  s = myc_block_reduce(s, s_warp);
  if (threadIdx.x == 0) *out = s;
}

// ---- explicit reduced matrix -----------------------------------------------------------
__global__ void __launch_bounds__(DI_THREADS)
free_flag_kernel(int64_t n, const double* __restrict__ dinv, int32_t* __restrict__ flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = dinv[i] != 0.0 ? 1 : 0;
}

__global__ void __launch_bounds__(DI_THREADS)
reduced_count_kernel(int64_t n_rows, const int32_t* __restrict__ rp, const int32_t* __restrict__ ci,
                     const double* __restrict__ dinv, const int32_t* __restrict__ free_index,
                     int32_t* __restrict__ cnt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (dinv[i] == 0.0) continue;
    int c = 0;
    for (int32_t j = rp[i]; j < rp[i + 1]; ++j) c += dinv[ci[j]] != 0.0 ? 1 : 0;
    cnt[free_index[i]] = c;
  }
}

__global__ void __launch_bounds__(DI_THREADS)
reduced_fill_kernel(int64_t n_rows, const int32_t* __restrict__ rp, const int32_t* __restrict__ ci,
                    const double* __restrict__ v, const double* __restrict__ dinv,
                    const int32_t* __restrict__ free_index, const int32_t* __restrict__ rrp,
                    int32_t* __restrict__ rci, double* __restrict__ rv) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (dinv[i] == 0.0) continue;
    int32_t o = rrp[free_index[i]];
    for (int32_t j = rp[i]; j < rp[i + 1]; ++j) {
      const int32_t c = ci[j];
      if (dinv[c] != 0.0) { rci[o] = free_index[c]; rv[o] = v[j]; ++o; }
    }
  }
}

// ---- structure check: does this CSR have the 3x3 node-block structure? ------------------------
__global__ void __launch_bounds__(DI_THREADS)
block3_check_kernel(int64_t n_nodes, const int32_t* __restrict__ rp, const int32_t* __restrict__ ci, int* __restrict__ bad) {
  for (int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nd < n_nodes; nd += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s0 = rp[3 * nd], s1 = rp[3 * nd + 1], s2 = rp[3 * nd + 2], s3 = rp[3 * nd + 3];
    const int32_t w = s1 - s0;
    bool ok = (s2 - s1 == w) && (s3 - s2 == w) && (w % 3 == 0);
    for (int32_t k = 0; ok && k < w; k += 3) {
      const int32_t c = ci[s0 + k];
      ok = (c % 3 == 0) && ci[s0 + k + 1] == c + 1 && ci[s0 + k + 2] == c + 2;
      for (int a = 1; ok && a < 3; ++a) {
        const int32_t o = s0 + a * w + k;
        ok = ci[o] == c && ci[o + 1] == c + 1 && ci[o + 2] == c + 2;
      }
    }
    if (!ok) *bad = 1;
  }
}

// ---- 3x3 node-block inverse --------------------------------------------------------------
__global__ void __launch_bounds__(DI_THREADS)
block3_kernel(int64_t n_nodes_local, int64_t row_offset, const int32_t* __restrict__ rp,
              const int32_t* __restrict__ ci, const double* __restrict__ v,
              const double* __restrict__ dinv, double reg, double* __restrict__ binv) {
  for (int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nd < n_nodes_local;
       nd += (int64_t)gridDim.x * blockDim.x) {
    double m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const int32_t c0 = (int32_t)(row_offset + 3 * nd);
    for (int a = 0; a < 3; ++a) {
      const int64_t i = 3 * nd + a;
      for (int32_t j = rp[i]; j < rp[i + 1]; ++j) {
        const int32_t c = ci[j] - c0;
        if (c >= 0 && c < 3) m[a][c] += v[j];
      }
      m[a][a] += reg;
    }
    bool fr[3];
    for (int a = 0; a < 3; ++a) fr[a] = dinv[3 * nd + a] != 0.0;
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c)
        if (!fr[a] || !fr[c]) m[a][c] = (a == c) ? 1.0 : 0.0;   // known DOF: identity row/col
    // symmetric 3x3 inverse by cofactors
    const double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
    const double c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
    const double c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
    const double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
    const double id = 1.0 / det;
    double inv[3][3];
    inv[0][0] = c00 * id;
    inv[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
    inv[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
    inv[1][0] = c01 * id;
    inv[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
    inv[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
    inv[2][0] = c02 * id;
    inv[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
    inv[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) binv[9 * nd + 3 * a + c] = (fr[a] && fr[c]) ? inv[a][c] : 0.0;
  }
}

// ---- aligned R x R diagonal-block inverse (R = 6, 12), symmetric-packed --------------------------
// One thread per block: gather the block from the CSR rows, regularise, replace known / padded
// rows and columns by identity, invert in place by Gauss-Jordan (the block is a principal submatrix
// of K_ff + reg I, hence SPD: no pivoting needed), store the symmetrised upper triangle.  A block
// that is singular in floating point keeps only the inverses of its 3x3 node blocks.
// [host-test-begin block_inverse_kernel]  (tests/test_kernel_logic_host.py compiles this text with g++)
template <int R>
__global__ void __launch_bounds__(128)
block_inverse_kernel(int64_t n_blocks, int64_t n_rows, int64_t row_offset, const int32_t* __restrict__ rp,
                     const int32_t* __restrict__ ci, const double* __restrict__ v,
                     const double* __restrict__ dinv, double reg, double* __restrict__ pinv) {
  for (int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; blk < n_blocks;
       blk += (int64_t)gridDim.x * blockDim.x) {
    double m[R][R];
    bool fr[R];
    const int64_t r0 = blk * R;
    const int64_t c0 = row_offset + r0;
    for (int a = 0; a < R; ++a) {
      for (int c = 0; c < R; ++c) m[a][c] = 0.0;
      const int64_t i = r0 + a;
      fr[a] = i < n_rows && dinv[i] != 0.0;
      if (fr[a]) {
        for (int32_t j = rp[i]; j < rp[i + 1]; ++j) {
          const int64_t c = (int64_t)ci[j] - c0;
          if (c >= 0 && c < R) m[a][c] += v[j];
        }
        m[a][a] += reg;
      }
    }
    double m0[R][R];                                   // the block before elimination (fallback below)
    for (int a = 0; a < R; ++a)
      for (int c = 0; c < R; ++c) {
        if (!fr[a] || !fr[c]) m[a][c] = (a == c) ? 1.0 : 0.0;
        m0[a][c] = m[a][c];
      }
    bool bad = false;
    for (int k = 0; k < R; ++k) {
      // a pivot that cancelled below 1e-13 of its diagonal entry carries no digits: meshes with
      // coincident nodes couple two nodes with ~1e28 and the coupled block is singular in fp64
      if (!(m[k][k] > 1e-13 * m0[k][k])) { bad = true; break; }
      const double piv = 1.0 / m[k][k];
      m[k][k] = 1.0;
      for (int c = 0; c < R; ++c) m[k][c] *= piv;
      for (int a = 0; a < R; ++a) {
        if (a == k) continue;
        const double f = m[a][k];
        m[a][k] = 0.0;
        for (int c = 0; c < R; ++c) m[a][c] -= f * m[k][c];
      }
    }
    if (!bad)
      for (int a = 0; a < R; ++a)
        for (int c = 0; c < R; ++c)
          if (!isfinite(m[a][c])) bad = true;
    if (bad) {   // this block falls back to its 3x3 node blocks (what MYC_PC_BLOCK3 applies)
      for (int a = 0; a < R; ++a)
        for (int c = 0; c < R; ++c) m[a][c] = 0.0;
      for (int q = 0; q < R; q += 3) {
        const double a00 = m0[q][q], a01 = m0[q][q + 1], a02 = m0[q][q + 2];
        const double a10 = m0[q + 1][q], a11 = m0[q + 1][q + 1], a12 = m0[q + 1][q + 2];
        const double a20 = m0[q + 2][q], a21 = m0[q + 2][q + 1], a22 = m0[q + 2][q + 2];
        const double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
        const double id = 1.0 / (a00 * c00 + a01 * c01 + a02 * c02);
        m[q][q] = c00 * id;
        m[q][q + 1] = (a02 * a21 - a01 * a22) * id;
        m[q][q + 2] = (a01 * a12 - a02 * a11) * id;
        m[q + 1][q] = c01 * id;
        m[q + 1][q + 1] = (a00 * a22 - a02 * a20) * id;
        m[q + 1][q + 2] = (a02 * a10 - a00 * a12) * id;
        m[q + 2][q] = c02 * id;
        m[q + 2][q + 1] = (a01 * a20 - a00 * a21) * id;
        m[q + 2][q + 2] = (a00 * a11 - a01 * a10) * id;
      }
    }
    double* out = pinv + blk * myc_block_inverse_stride(R);
    for (int a = 0; a < R; ++a)
      for (int c = a; c < R; ++c) {
        const double e = (fr[a] && fr[c]) ? 0.5 * (m[a][c] + m[c][a]) : 0.0;
        if (R == 6 && MYC_B6_FULL) {
          out[a * R + c] = e;
          out[c * R + a] = e;
        } else {
          out[myc_sympack(R, a, c)] = e;
        }
      }
  }
}
// [host-test-end block_inverse_kernel]

}  // namespace

extern "C" int myc_apply_dirichlet(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                                   const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                                   const int64_t* d_known_dofs, const double* d_known_vals, int64_t n_known,
                                   double reg, double* d_out_ubc, double* d_out_rhs, double* d_out_dinv,
                                   void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n_rows < 0 || n_cols_global < n_rows || row_offset < 0 || row_offset + n_rows > n_cols_global ||
      n_known < 0 || !d_row_ptr || (n_known > 0 && (!d_known_dofs || !d_known_vals)) ||
      (n_cols_global > 0 && !d_out_ubc) || (n_rows > 0 && (!d_out_rhs || !d_out_dinv)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "apply_dirichlet: bad argument");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  MYC_TRY(myc_ensure(ctx, ctx->misc, 256));
  int* bad_flag = (int*)ctx->misc.p;
  MYC_CUDA(ctx, cudaMemsetAsync(bad_flag, 0, sizeof(int), st));
  MYC_CUDA(ctx, cudaMemsetAsync(d_out_ubc, 0, (size_t)n_cols_global * sizeof(double), st));
  if (n_rows == 0) return MYC_OK;
  const int g_rows = grid_for(ctx, ceil_div64(n_rows, DI_THREADS), 8);
  fill_kernel<<<g_rows, DI_THREADS, 0, st>>>(d_out_dinv, n_rows, 1.0);
  MYC_LAUNCHED(ctx);
  if (n_known > 0) {
    scatter_known_kernel<<<grid_for(ctx, ceil_div64(n_known, DI_THREADS), 8), DI_THREADS, 0, st>>>(
        d_known_dofs, d_known_vals, n_known, n_cols_global, row_offset, n_rows, d_out_ubc, d_out_dinv, bad_flag);
    MYC_LAUNCHED(ctx);
  }
  jacobi_kernel<<<g_rows, DI_THREADS, 0, st>>>(n_rows, row_offset, d_row_ptr, d_col_idx, d_val, reg, d_out_dinv);
  MYC_LAUNCHED(ctx);
  EpiRhs epi{d_out_rhs, d_out_dinv};
  MYC_TRY(myc_launch_spmv_epi<EpiRhs>(ctx, n_rows, d_row_ptr, d_col_idx, d_val, d_out_ubc, epi, nullptr, nullptr,
                                      nullptr, nullptr, st));
  int* h = (int*)ctx->h_pinned;
  MYC_CUDA(ctx, cudaMemcpyAsync(h, bad_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  if (*h) MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "apply_dirichlet: known DOF outside [0, n_dof)");
  return MYC_OK;
}

extern "C" int myc_csr_is_block3(myc_ctx* ctx, int64_t n_rows, const int32_t* d_row_ptr, const int32_t* d_col_idx,
                                 int* h_out_is_block3, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n_rows < 0 || !d_row_ptr || !h_out_is_block3) MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "csr_is_block3: bad argument");
  *h_out_is_block3 = 0;
  if (n_rows % 3 != 0) return MYC_OK;
  if (n_rows == 0) { *h_out_is_block3 = 1; return MYC_OK; }
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  MYC_TRY(myc_ensure(ctx, ctx->misc, 512));
  int* bad = (int*)ctx->misc.p;
  MYC_CUDA(ctx, cudaMemsetAsync(bad, 0, sizeof(int), st));
  block3_check_kernel<<<grid_for(ctx, ceil_div64(n_rows / 3, DI_THREADS), 8), DI_THREADS, 0, st>>>(n_rows / 3, d_row_ptr,
                                                                                                  d_col_idx, bad);
  MYC_LAUNCHED(ctx);
  int* h = (int*)ctx->h_pinned;
  MYC_CUDA(ctx, cudaMemcpyAsync(h, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  *h_out_is_block3 = *h ? 0 : 1;
  return MYC_OK;
}

extern "C" int myc_block3_inverse(myc_ctx* ctx, int64_t n_rows, int64_t row_offset, const int32_t* d_row_ptr,
                                  const int32_t* d_col_idx, const double* d_val, const double* d_dinv,
                                  double reg, double* d_out_binv, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n_rows < 0 || n_rows % 3 || !d_row_ptr || (n_rows > 0 && (!d_dinv || !d_out_binv)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "block3_inverse: bad argument (n_rows must be a multiple of 3)");
  if (n_rows == 0) return MYC_OK;
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  block3_kernel<<<grid_for(ctx, ceil_div64(n_rows / 3, DI_THREADS), 8), DI_THREADS, 0, (cudaStream_t)stream>>>(
      n_rows / 3, row_offset, d_row_ptr, d_col_idx, d_val, d_dinv, reg, d_out_binv);
  MYC_LAUNCHED(ctx);
  return MYC_OK;
}

extern "C" int64_t myc_block_inverse_size(int nodes_per_block, int64_t n_rows) {
  if ((nodes_per_block != 2 && nodes_per_block != 4) || n_rows < 0) return -1;
  const int R = 3 * nodes_per_block;
  return ceil_div64(n_rows, R) * myc_block_inverse_stride(R);
}

extern "C" int myc_block_inverse_packed(myc_ctx* ctx, int nodes_per_block, int64_t n_rows, int64_t row_offset,
                                        const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                                        const double* d_dinv, double reg, double* d_out_pinv, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  const int R = 3 * nodes_per_block;
  if ((nodes_per_block != 2 && nodes_per_block != 4) || n_rows < 0 || row_offset < 0 || row_offset % R ||
      !d_row_ptr || (n_rows > 0 && (!d_col_idx || !d_val || !d_dinv || !d_out_pinv)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "block_inverse_packed: bad argument (nodes_per_block 2 or 4, row_offset a multiple of the block size)");
  if (n_rows == 0) return MYC_OK;
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n_blocks = ceil_div64(n_rows, R);
  const int grid = grid_for(ctx, ceil_div64(n_blocks, 128), 8);
  if (nodes_per_block == 2)
    block_inverse_kernel<6><<<grid, 128, 0, (cudaStream_t)stream>>>(n_blocks, n_rows, row_offset, d_row_ptr, d_col_idx,
                                                                    d_val, d_dinv, reg, d_out_pinv);
  else
    block_inverse_kernel<12><<<grid, 128, 0, (cudaStream_t)stream>>>(n_blocks, n_rows, row_offset, d_row_ptr, d_col_idx,
                                                                     d_val, d_dinv, reg, d_out_pinv);
  MYC_LAUNCHED(ctx);
  return MYC_OK;
}

extern "C" int myc_merge_solution(myc_ctx* ctx, int64_t n_rows, int64_t row_offset, const double* d_x,
                                  const double* d_dinv, const double* d_ubc, double* d_out_U, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n_rows < 0 || (n_rows > 0 && (!d_x || !d_dinv || !d_ubc || !d_out_U)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "merge_solution: bad argument");
  if (n_rows == 0) return MYC_OK;
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  merge_kernel<<<grid_for(ctx, ceil_div64(n_rows, DI_THREADS), 8), DI_THREADS, 0, (cudaStream_t)stream>>>(
      n_rows, row_offset, d_x, d_dinv, d_ubc, d_out_U);
  MYC_LAUNCHED(ctx);
  return MYC_OK;
}

extern "C" int myc_gather_sum(myc_ctx* ctx, const double* d_v, const int64_t* d_idx, int64_t n,
                              double* h_out_sum, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n < 0 || !h_out_sum || (n > 0 && (!d_v || !d_idx))) MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "gather_sum: bad argument");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  MYC_TRY(myc_ensure(ctx, ctx->misc, 256));
  double* d_out = (double*)((char*)ctx->misc.p + 128);
  gather_sum_kernel<<<1, SP_THREADS, 0, st>>>(d_v, d_idx, n, d_out);
  MYC_LAUNCHED(ctx);
  double* h = (double*)ctx->h_pinned;
  MYC_CUDA(ctx, cudaMemcpyAsync(h, d_out, sizeof(double), cudaMemcpyDeviceToHost, st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  *h_out_sum = *h;
  return MYC_OK;
}

extern "C" int myc_reduce_csr(myc_ctx* ctx, int64_t n_rows, const int32_t* d_row_ptr, const int32_t* d_col_idx,
                              const double* d_val, const double* d_dinv, int32_t* d_out_free_index,
                              int32_t* d_out_row_ptr, int32_t* d_out_col_idx, double* d_out_val,
                              int64_t* h_out_n_free, int64_t* h_out_nnz, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n_rows < 0 || !d_row_ptr || !d_out_free_index || !d_out_row_ptr || !h_out_n_free || !h_out_nnz ||
      (n_rows > 0 && !d_dinv))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "reduce_csr: bad argument");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  MYC_TRY(myc_ensure(ctx, ctx->misc, 256));
  int64_t* d_tot = (int64_t*)((char*)ctx->misc.p + 64);
  int64_t* h = (int64_t*)ctx->h_pinned;
  const int g = grid_for(ctx, ceil_div64(n_rows, DI_THREADS), 8);
  // free_index = exclusive scan of the free flags (the compacted numbering of np.setdiff1d)
  if (n_rows > 0) {
    free_flag_kernel<<<g, DI_THREADS, 0, st>>>(n_rows, d_dinv, d_out_free_index);
    MYC_LAUNCHED(ctx);
  }
  MYC_TRY(myc_exclusive_scan_i32(ctx, d_out_free_index, d_out_free_index, n_rows, false, d_tot, st));
  MYC_CUDA(ctx, cudaMemcpyAsync(h, d_tot, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  const int64_t n_free = h[0];
  *h_out_n_free = n_free;
  if (d_out_col_idx == nullptr) {   // phase 1: row_ptr + nnz
    MYC_CUDA(ctx, cudaMemsetAsync(d_out_row_ptr, 0, (size_t)(n_free + 1) * sizeof(int32_t), st));
    if (n_rows > 0) {
      reduced_count_kernel<<<g, DI_THREADS, 0, st>>>(n_rows, d_row_ptr, d_col_idx, d_dinv, d_out_free_index, d_out_row_ptr);
      MYC_LAUNCHED(ctx);
    }
    MYC_TRY(myc_exclusive_scan_i32(ctx, d_out_row_ptr, d_out_row_ptr, n_free, true, d_tot, st));
    MYC_CUDA(ctx, cudaMemcpyAsync(h, d_tot, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    *h_out_nnz = h[0];
    return MYC_OK;
  }
  if (!d_out_val) MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "reduce_csr: d_out_val is null");
  if (n_rows > 0) {
    reduced_fill_kernel<<<g, DI_THREADS, 0, st>>>(n_rows, d_row_ptr, d_col_idx, d_val, d_dinv, d_out_free_index,
                                                  d_out_row_ptr, d_out_col_idx, d_out_val);
    MYC_LAUNCHED(ctx);
  }
  return MYC_OK;
}
