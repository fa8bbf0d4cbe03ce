// Aggregation-multigrid hierarchy for the PCG (the GPU path's answer to the "gamg / icc / sor" rows of the
// reference's PETSc menu, src/fea_petsc_solverAndPC.cpp:330-331; algorithm restated in oracle/amg_oracle.py).
//
// Per level, all on the caller's stream, integer work + fixed-order sums only (deterministic):
//   propose  every active node points at its most strongly coupled active neighbour, w = -trace(block) > 0
//   accept   mutual pointers become a pair
//   join     an unpaired node joins the pair of its strongest paired neighbour; root = smallest node of the pair
//   keep     aggregates with no block to an active node outside themselves are not represented further down
//   number   kept aggregates in root order (exclusive scan)  -> agg[]
//   members  stable radix sort of the nodes by aggregate      -> mptr / mlist (restriction gathers)
//   coarse   radix sort of the fine blocks by (agg row, agg column), runs summed in fine block order
//            -> brp / bcol / bval of the next level, again symmetric 3x3 blocks
//   dinv     symmetric inverse of (diagonal block + reg I) for the 3x3-block Jacobi smoother
#include "amg.cuh"
#include "spmv_sym3.cuh"

namespace {

constexpr int AM_THREADS = 256;

__device__ __forceinline__ bool am_active(const uint8_t* __restrict__ act, int64_t i) { return !act || act[i]; }

// level 0: block row pointer and activity (a node is active iff all three DOFs are free; mixed nodes -> *bad)
__global__ void __launch_bounds__(AM_THREADS)
am_level0_kernel(int64_t n, const int32_t* __restrict__ rp, const double* __restrict__ dinv, int32_t* __restrict__ brp,
                 uint8_t* __restrict__ act, int* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
    brp[i] = rp[3 * i] / 9;
    if (i < n) {
      const int f = (dinv[3 * i] != 0.0) + (dinv[3 * i + 1] != 0.0) + (dinv[3 * i + 2] != 0.0);
      act[i] = f == 3;
      if (f != 0 && f != 3) *bad = 1;
    }
  }
}

// strongest eligible neighbour of node i (ties: smaller column = first met); `need_paired`: only paired ones
__device__ __forceinline__ int32_t am_strongest(int64_t i, int64_t n, int64_t node_off, const int32_t* __restrict__ brp,
                                                const int32_t* __restrict__ bcol, const double* __restrict__ bval,
                                                const uint8_t* __restrict__ act, const int32_t* __restrict__ paired) {
  int32_t bj = -1;
  double bw = 0.0;
  for (int32_t b = brp[i]; b < brp[i + 1]; ++b) {
    const int64_t j = (int64_t)(bcol[b] / 3) - node_off;
    if (j == i || j < 0 || j >= n || !am_active(act, j)) continue;     // remote nodes never join a local aggregate
    if (paired && paired[j] < 0) continue;
    const double* v = bval + 6 * (size_t)b;
    const double w = -((v[0] + v[3]) + v[5]);
    if (w > bw) { bw = w; bj = (int32_t)j; }
  }
  return bj;
}

__global__ void __launch_bounds__(AM_THREADS)
am_propose_kernel(int64_t n, int64_t node_off, const int32_t* __restrict__ brp, const int32_t* __restrict__ bcol,
                  const double* __restrict__ bval, const uint8_t* __restrict__ act, int32_t* __restrict__ best) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    best[i] = am_active(act, i) ? am_strongest(i, n, node_off, brp, bcol, bval, act, nullptr) : -1;
}

__global__ void __launch_bounds__(AM_THREADS)
am_accept_kernel(int64_t n, const int32_t* __restrict__ best, int32_t* __restrict__ paired) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = best[i];
    paired[i] = (b >= 0 && best[b] == (int32_t)i) ? b : -1;
  }
}

__global__ void __launch_bounds__(AM_THREADS)
am_root_kernel(int64_t n, int64_t node_off, const int32_t* __restrict__ brp, const int32_t* __restrict__ bcol,
               const double* __restrict__ bval, const uint8_t* __restrict__ act, const int32_t* __restrict__ paired,
               int32_t* __restrict__ root) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t r = (int32_t)i;
    if (am_active(act, i)) {
      const int32_t p = paired[i];
      if (p >= 0) {
        r = p < r ? p : r;
      } else {
        const int32_t j = am_strongest(i, n, node_off, brp, bcol, bval, act, paired);
        if (j >= 0) r = j < paired[j] ? j : paired[j];
      }
    }
    root[i] = r;
  }
}

// keep[root] = 1 if the aggregate has a block to an active node outside itself (benign same-value races).
// act_global (level 0 on several GPUs): activity of every node of the level, indexed by global id.
__global__ void __launch_bounds__(AM_THREADS)
am_keep_kernel(int64_t n, int64_t node_off, const int32_t* __restrict__ brp, const int32_t* __restrict__ bcol,
               const uint8_t* __restrict__ act, const uint8_t* __restrict__ act_global,
               const int32_t* __restrict__ root, int32_t* __restrict__ keep) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!am_active(act, i)) continue;
    const int32_t ri = root[i];
    bool ext = false;
    for (int32_t b = brp[i]; b < brp[i + 1] && !ext; ++b) {
      const int64_t c = bcol[b] / 3, j = c - node_off;
      if (j == i) continue;
      if (j < 0 || j >= n) ext = !act_global || act_global[c];
      else ext = am_active(act, j) && root[j] != ri;
    }
    if (ext) keep[ri] = 1;
  }
}

__global__ void __launch_bounds__(AM_THREADS)
am_lead_kernel(int64_t n, const uint8_t* __restrict__ act, const int32_t* __restrict__ root,
               const int32_t* __restrict__ keep, int32_t* __restrict__ flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (am_active(act, i) && root[i] == (int32_t)i && keep[i]) ? 1 : 0;
}

__global__ void __launch_bounds__(AM_THREADS)
am_assign_kernel(int64_t n, const uint8_t* __restrict__ act, const int32_t* __restrict__ root,
                 const int32_t* __restrict__ keep, const int32_t* __restrict__ cid, int32_t* __restrict__ agg,
                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int32_t n_c) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = root[i];
    const int32_t a = (am_active(act, i) && keep[r]) ? cid[r] : -1;
    agg[i] = a;
    keys[i] = (uint64_t)(a >= 0 ? a : n_c);        // members: nodes grouped by aggregate, unrepresented ones last
    vals[i] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(AM_THREADS)
am_members_kernel(int64_t n, int32_t n_c, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                  int32_t* __restrict__ mptr, int32_t* __restrict__ mlist) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t k = (int32_t)keys[p];
    mlist[p] = (int32_t)vals[p];
    if (p == 0 || (int32_t)keys[p - 1] != k) mptr[k] = (int32_t)p;      // k == n_c: end of the last member list
    if (p == n - 1 && k < n_c) mptr[n_c] = (int32_t)n;
  }
}

// one key per fine block: (aggregate of the row node, GLOBAL aggregate id of the column node); blocks that do
// not survive (unrepresented or inactive end) sort behind everything else.
// agg_global (several GPUs): global aggregate id (or -1) of every node of the level, indexed by global node id.
__global__ void __launch_bounds__(AM_THREADS)
am_coarse_emit_kernel(int64_t n, int64_t node_off, const int32_t* __restrict__ brp, const int32_t* __restrict__ bcol,
                      const uint8_t* __restrict__ act, const int32_t* __restrict__ agg,
                      const int32_t* __restrict__ agg_global, int64_t cnode_off, int32_t n_c, int cbits,
                      uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t I = agg[i];
    for (int32_t b = brp[i]; b < brp[i + 1]; ++b) {
      const int64_t c = bcol[b] / 3, j = c - node_off;
      int64_t J = -1;
      if (j >= 0 && j < n) { if (am_active(act, j) && agg[j] >= 0) J = cnode_off + agg[j]; }
      else if (agg_global) J = agg_global[c];
      keys[b] = (I >= 0 && J >= 0) ? (((uint64_t)I << cbits) | (uint64_t)J) : ((uint64_t)n_c << cbits);
      vals[b] = (uint32_t)b;
    }
  }
}

__global__ void __launch_bounds__(AM_THREADS)
am_head_kernel(int64_t nb, int32_t n_c, int cbits, const uint64_t* __restrict__ keys, int32_t* __restrict__ flag) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nb; p += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[p];
    flag[p] = ((int64_t)(k >> cbits) < n_c && (p == 0 || keys[p - 1] != k)) ? 1 : 0;
  }
}

// one thread per coarse block (run head): sum the run in sorted (= fine block) order
__global__ void __launch_bounds__(AM_THREADS)
am_coarse_fill_kernel(int64_t nb, int32_t n_c, int cbits, const uint64_t* __restrict__ keys,
                      const uint32_t* __restrict__ vals, const int32_t* __restrict__ flag,
                      const int32_t* __restrict__ uidx, const double* __restrict__ bval_f, int32_t n_cb,
                      int32_t* __restrict__ brp_c, int32_t* __restrict__ bcol_c, double* __restrict__ bval_c) {
  const uint64_t cmask = ((uint64_t)1 << cbits) - 1;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nb; p += (int64_t)gridDim.x * blockDim.x) {
    if (p == 0) brp_c[n_c] = n_cb;
    if (!flag[p]) continue;
    const uint64_t k = keys[p];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0;
    for (int64_t q = p; q < nb && keys[q] == k; ++q) {
      const double* v = bval_f + 6 * (size_t)vals[q];
      a0 += v[0]; a1 += v[1]; a2 += v[2]; a3 += v[3]; a4 += v[4]; a5 += v[5];
    }
    const int32_t u = uidx[p];
    const int32_t I = (int32_t)(k >> cbits);
    double* o = bval_c + 6 * (size_t)u;
    o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3; o[4] = a4; o[5] = a5;
    bcol_c[u] = 3 * (int32_t)(k & cmask);
    if (p == 0 || (int32_t)(keys[p - 1] >> cbits) != I) brp_c[I] = u;
  }
}

// symmetric inverse of (diagonal block + reg I) by cofactors; zeros for inactive nodes and singular blocks
__global__ void __launch_bounds__(AM_THREADS)
am_dinv_kernel(int64_t n, int64_t node_off, const int32_t* __restrict__ brp, const int32_t* __restrict__ bcol,
               const double* __restrict__ bval, const uint8_t* __restrict__ act, double reg, double* __restrict__ dinv) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
    const int32_t self = 3 * (int32_t)(node_off + i);
    for (int32_t b = brp[i]; b < brp[i + 1]; ++b)
      if (bcol[b] == self) {
        const double* v = bval + 6 * (size_t)b;
        xx = v[0]; xy = v[1]; xz = v[2]; yy = v[3]; yz = v[4]; zz = v[5];
      }
    xx += reg; yy += reg; zz += reg;
    const double c00 = yy * zz - yz * yz, c01 = yz * xz - xy * zz, c02 = xy * yz - yy * xz;
    const double det = xx * c00 + xy * c01 + xz * c02;
    double* o = dinv + 6 * (size_t)i;
    if (am_active(act, i) && det > 0.0 && isfinite(det)) {
      const double id = 1.0 / det;
      o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
      o[3] = (xx * zz - xz * xz) * id; o[4] = (xz * xy - xx * yz) * id; o[5] = (xx * yy - xy * xy) * id;
    } else {
      o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.0;
    }
  }
}

int am_bits_for(int64_t n) {
  int b = 1;
  while (((int64_t)1 << b) < n) ++b;
  return b;
}

}  // namespace

int myc_amg_destroy(myc_ctx* ctx) {
  AmgState* s = ctx->amg;
  if (!s) return MYC_OK;
  auto drop = [](DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; };
  for (AmgLevelHost& L : s->lv) {
    drop(L.brp); drop(L.bcol); drop(L.bval); drop(L.dinv); drop(L.agg); drop(L.mptr); drop(L.mlist); drop(L.r); drop(L.t);
  }
  drop(s->lv_dev); drop(s->brp0); drop(s->arena); drop(s->act0);
  for (DevBuf& b : s->work) drop(b);
  delete s;
  ctx->amg = nullptr;
  return MYC_OK;
}

extern "C" int myc_amg_setup(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                             const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                             const double* d_dinv, double reg, int* h_out_levels, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (h_out_levels) *h_out_levels = 0;
  if (n_rows < 0 || n_cols_global < n_rows || row_offset < 0 || !d_row_ptr || (n_rows > 0 && (!d_col_idx || !d_val || !d_dinv)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "amg_setup: bad argument");
  if (!ctx->amg) ctx->amg = new AmgState();
  AmgState* S = ctx->amg;
  S->valid = false;
  S->n_levels = 0;
  // the hierarchy is built on the symmetric 3x3 node-block view of K: needs the node-block structure
  if (!ctx->csr_block3 || n_rows % 3 != 0 || row_offset % 3 != 0 || n_rows == 0 ||
      (((uintptr_t)d_col_idx | (uintptr_t)d_val) & 15u) != 0)
    return MYC_OK;
  if (ctx->world > 1) return MYC_OK;        // the row-partitioned hierarchy is set up by myc_amg_setup_dist
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  ctx->plan_valid = false;                  // the sort buffers of the assembly plan are reused below
  MYC_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
  const int64_t n0 = n_rows / 3;
  int64_t* h_pin = (int64_t*)ctx->h_pinned;
  MYC_TRY(myc_ensure(ctx, ctx->misc, 512));
  int* bad = (int*)((char*)ctx->misc.p + 320);
  int64_t* d_total = (int64_t*)((char*)ctx->misc.p + 64);

  // ---- level 0: symmetric block view of K (also what the solver sweeps), block row pointer, activity
  int32_t h_nnz = 0;
  MYC_CUDA(ctx, cudaMemcpyAsync(&h_nnz, d_row_ptr + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  const int64_t nb0 = h_nnz / 9;
  MYC_TRY(myc_ensure(ctx, ctx->sym_val, ((size_t)nb0 + 4) * 6 * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->sym_col, ((size_t)nb0 + 4) * sizeof(int32_t)));
  MYC_TRY(myc_ensure(ctx, S->brp0, (size_t)(n0 + 1) * sizeof(int32_t)));
  MYC_TRY(myc_ensure(ctx, S->act0, (size_t)(n0 + 1)));
  MYC_CUDA(ctx, cudaMemsetAsync(bad, 0, 2 * sizeof(int), st));
  myc_sym3_convert_kernel<<<grid_for(ctx, ceil_div64(n0, 256), 8), 256, 0, st>>>(
      n0, d_row_ptr, d_col_idx, d_val, (double*)ctx->sym_val.p, (int32_t*)ctx->sym_col.p, bad);
  MYC_LAUNCHED(ctx);
  am_level0_kernel<<<grid_for(ctx, ceil_div64(n0 + 1, AM_THREADS), 8), AM_THREADS, 0, st>>>(
      n0, d_row_ptr, d_dinv, (int32_t*)S->brp0.p, (uint8_t*)S->act0.p, bad + 1);
  MYC_LAUNCHED(ctx);
  int* h_bad = (int*)(h_pin + 8);
  MYC_CUDA(ctx, cudaMemcpyAsync(h_bad, bad, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  if (h_bad[0] || h_bad[1]) return MYC_OK;  // K not blockwise symmetric, or a node with a partial Dirichlet set
  ctx->sym_owner = 1;                       // sym_val / sym_col now belong to this hierarchy

  for (DevBuf& w : S->work) MYC_TRY(myc_ensure(ctx, w, (size_t)(n0 + 2) * sizeof(int32_t)));
  int32_t* best = (int32_t*)S->work[0].p;
  int32_t* paired = (int32_t*)S->work[1].p;
  int32_t* root = (int32_t*)S->work[2].p;
  int32_t* keep = (int32_t*)S->work[3].p;
  int32_t* flag = (int32_t*)S->work[4].p;   // regrown below when a level has more blocks than level 0 has nodes

  // level 0 views (not owned by the level: K's block view lives in ctx->sym_*)
  AmgLevelHost* L = &S->lv[0];
  L->n = n0; L->nb = nb0; L->n_global = n_cols_global / 3; L->node_off = row_offset / 3;
  const int32_t* brp = (const int32_t*)S->brp0.p;
  const int32_t* bcol = (const int32_t*)ctx->sym_col.p;
  const double* bval = (const double*)ctx->sym_val.p;
  const uint8_t* act = (const uint8_t*)S->act0.p;
  int lv = 0;
  for (;;) {
    L = &S->lv[lv];
    const int64_t n = L->n, nb = L->nb;
    const int g_n = grid_for(ctx, ceil_div64(n, AM_THREADS), 8);
    MYC_TRY(myc_ensure(ctx, L->dinv, (size_t)(n + 1) * 6 * sizeof(double)));
    am_dinv_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, brp, bcol, bval, act, reg, (double*)L->dinv.p);
    MYC_LAUNCHED(ctx);
    if (lv + 1 >= AMG_MAX_LEVELS || L->n_global <= AMG_MIN_NODES) break;
    // ---- aggregates
    am_propose_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, brp, bcol, bval, act, best);
    MYC_LAUNCHED(ctx);
    am_accept_kernel<<<g_n, AM_THREADS, 0, st>>>(n, best, paired);
    MYC_LAUNCHED(ctx);
    am_root_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, brp, bcol, bval, act, paired, root);
    MYC_LAUNCHED(ctx);
    MYC_CUDA(ctx, cudaMemsetAsync(keep, 0, (size_t)n * sizeof(int32_t), st));
    am_keep_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, brp, bcol, act, nullptr, root, keep);
    MYC_LAUNCHED(ctx);
    am_lead_kernel<<<g_n, AM_THREADS, 0, st>>>(n, act, root, keep, flag);
    MYC_LAUNCHED(ctx);
    MYC_TRY(myc_exclusive_scan_i32(ctx, flag, flag, n, false, d_total, st));
    MYC_CUDA(ctx, cudaMemcpyAsync(h_pin, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t n_c = h_pin[0];
    if (n_c == 0 || (double)n_c > AMG_MAX_RATIO * (double)n) break;
    // ---- agg[], member lists
    const size_t items = (size_t)(nb > n ? nb : n) + 1;
    for (int k = 0; k < 2; ++k) {
      MYC_TRY(myc_ensure(ctx, ctx->sort_keys[k], items * sizeof(uint64_t)));
      MYC_TRY(myc_ensure(ctx, ctx->sort_vals[k], items * sizeof(uint32_t)));
    }
    MYC_TRY(myc_ensure(ctx, L->agg, (size_t)(n + 1) * sizeof(int32_t)));
    AmgLevelHost* C = &S->lv[lv + 1];
    MYC_TRY(myc_ensure(ctx, C->mptr, (size_t)(n_c + 2) * sizeof(int32_t)));
    MYC_TRY(myc_ensure(ctx, C->mlist, (size_t)(n + 1) * sizeof(int32_t)));
    am_assign_kernel<<<g_n, AM_THREADS, 0, st>>>(n, act, root, keep, flag, (int32_t*)L->agg.p,
                                                 (uint64_t*)ctx->sort_keys[0].p, (uint32_t*)ctx->sort_vals[0].p, (int32_t)n_c);
    MYC_LAUNCHED(ctx);
    int sorted = 0;
    MYC_TRY(myc_radix_sort_pairs(ctx, n, 0, am_bits_for(n_c + 1), &sorted, st));
    am_members_kernel<<<g_n, AM_THREADS, 0, st>>>(n, (int32_t)n_c, (const uint64_t*)ctx->sort_keys[sorted].p,
                                                  (const uint32_t*)ctx->sort_vals[sorted].p, (int32_t*)C->mptr.p,
                                                  (int32_t*)C->mlist.p);
    MYC_LAUNCHED(ctx);
    // ---- Galerkin operator of the aggregates
    const int cbits = am_bits_for(n_c + 1);
    am_coarse_emit_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, brp, bcol, act, (const int32_t*)L->agg.p, nullptr,
                                                      0, (int32_t)n_c, cbits, (uint64_t*)ctx->sort_keys[0].p,
                                                      (uint32_t*)ctx->sort_vals[0].p);
    MYC_LAUNCHED(ctx);
    MYC_TRY(myc_radix_sort_pairs(ctx, nb, 0, 2 * cbits, &sorted, st));
    const uint64_t* skeys = (const uint64_t*)ctx->sort_keys[sorted].p;
    const uint32_t* svals = (const uint32_t*)ctx->sort_vals[sorted].p;
    MYC_TRY(myc_ensure(ctx, S->work[4], (size_t)(nb + 2) * sizeof(int32_t)));
    MYC_TRY(myc_ensure(ctx, S->work[5], (size_t)(nb + 2) * sizeof(int32_t)));
    flag = (int32_t*)S->work[4].p;
    int32_t* uidx = (int32_t*)S->work[5].p;
    const int g_b = grid_for(ctx, ceil_div64(nb, AM_THREADS), 8);
    am_head_kernel<<<g_b, AM_THREADS, 0, st>>>(nb, (int32_t)n_c, cbits, skeys, flag);
    MYC_LAUNCHED(ctx);
    MYC_TRY(myc_exclusive_scan_i32(ctx, flag, uidx, nb, false, d_total, st));
    MYC_CUDA(ctx, cudaMemcpyAsync(h_pin, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t n_cb = h_pin[0];
    MYC_TRY(myc_ensure(ctx, C->brp, (size_t)(n_c + 2) * sizeof(int32_t)));
    MYC_TRY(myc_ensure(ctx, C->bcol, (size_t)(n_cb + 4) * sizeof(int32_t)));
    MYC_TRY(myc_ensure(ctx, C->bval, (size_t)(n_cb + 4) * 6 * sizeof(double)));
    am_coarse_fill_kernel<<<g_b, AM_THREADS, 0, st>>>(nb, (int32_t)n_c, cbits, skeys, svals, flag, uidx, bval, (int32_t)n_cb,
                                                      (int32_t*)C->brp.p, (int32_t*)C->bcol.p, (double*)C->bval.p);
    MYC_LAUNCHED(ctx);
    C->n = n_c; C->nb = n_cb; C->n_global = n_c; C->node_off = 0;
    brp = (const int32_t*)C->brp.p;
    bcol = (const int32_t*)C->bcol.p;
    bval = (const double*)C->bval.p;
    act = nullptr;
    ++lv;
  }
  S->n_levels = lv + 1;
  // ---- vectors: r, t per level (level 0's r is the CG residual); arena of the gathered correction vectors
  int64_t off = 0;
  for (int l = 0; l < S->n_levels; ++l) {
    AmgLevelHost& H = S->lv[l];
    if (l > 0) MYC_TRY(myc_ensure(ctx, H.r, (size_t)(3 * H.n + 4) * sizeof(double)));
    MYC_TRY(myc_ensure(ctx, H.t, (size_t)(3 * H.n + 4) * sizeof(double)));
    for (int k = 0; k < 2; ++k) {
      H.e_off[k] = off;
      off += (3 * H.n_global + 15) / 16 * 16;
    }
  }
  S->arena_doubles = off;
  MYC_TRY(myc_ensure(ctx, S->arena, (size_t)(off + 16) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, S->lv_dev, sizeof(AmgLevelDev) * AMG_MAX_LEVELS));
  MYC_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  MYC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
  S->setup_ms = ms;
  S->n_rows0 = n_rows;
  S->row_offset0 = row_offset;
  S->key_rp = d_row_ptr;
  S->key_val = d_val;
  S->key_dinv = d_dinv;
  S->reg = reg;
  S->valid = true;
  if (h_out_levels) *h_out_levels = S->n_levels;
  return MYC_OK;
}

double myc_amg_bytes_per_iteration(const myc_ctx* ctx) {
  const AmgState* S = ctx->amg;
  if (!S || !S->valid) return 0.0;
  double bytes = 0.0;
  for (int l = 0; l < S->n_levels; ++l) {
    const double nb = (double)S->lv[l].nb, rows = 3.0 * (double)S->lv[l].n;
    const double mat = 52.0 * nb + 4.0 * rows / 3.0;          // block view + block row pointer
    const bool coarsest = l == S->n_levels - 1;
    // sweeps: matrix + gathered e (8) + r (8) + own e (8) + write (8) per row; smoothing sweeps also read dinv (16)
    const double sweeps = coarsest ? (AMG_COARSE_SWEEPS - 1) : 2;
    bytes += sweeps * (mat + 32.0 * rows) + (coarsest ? sweeps : 1.0) * 16.0 * rows;
    if (l == 0) {
      bytes += mat + 32.0 * rows;                              // w = A u: u gathered, u own, r read, w written
      bytes += (13.0 * 8.0) * rows;                            // D0: u w p s x r mask dinv(2) read, p s x r e written
    } else {
      bytes += (8.0 + 8.0 + 16.0 + 8.0) * rows + 4.0 * rows;   // restriction: t of the members, r and e written, dinv, lists
    }
    if (!coarsest) bytes += (8.0 + 8.0 + 8.0) * rows + 4.0 * rows / 3.0;   // prolongation: e read + written, coarse e, agg
  }
  return bytes;
}

// Introspection for tests and reports: out[0] = owned nodes, out[1] = owned blocks of `level`; out[2] = levels;
// out[3] = setup time in microseconds.  d_out_agg (may be NULL): the level's aggregate map, n int32.
extern "C" int myc_amg_level_info(myc_ctx* ctx, int level, int64_t* h_out4, int32_t* d_out_agg, void* stream) {
  if (!ctx || !h_out4) return MYC_ERR_BAD_ARG;
  AmgState* S = ctx->amg;
  if (!S || !S->valid || level < 0 || level >= S->n_levels) MYC_FAIL(ctx, MYC_ERR_STATE, "amg_level_info: no such level");
  h_out4[0] = S->lv[level].n;
  h_out4[1] = S->lv[level].nb;
  h_out4[2] = S->n_levels;
  h_out4[3] = (int64_t)(S->setup_ms * 1e3);
  if (d_out_agg) {
    if (level + 1 >= S->n_levels) MYC_FAIL(ctx, MYC_ERR_STATE, "amg_level_info: the coarsest level has no aggregates");
    MYC_CUDA(ctx, cudaMemcpyAsync(d_out_agg, S->lv[level].agg.p, (size_t)S->lv[level].n * sizeof(int32_t),
                                  cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  }
  return MYC_OK;
}
