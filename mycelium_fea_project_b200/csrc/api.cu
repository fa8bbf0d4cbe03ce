// Context lifetime and the host-buffer end-to-end entry point of the C-ABI.
#include <stdlib.h>

#include "common.cuh"

int myc_dist_destroy(myc_ctx* ctx);   // dist.cu
int myc_amg_destroy(myc_ctx* ctx);    // amg_setup.cu

static char g_create_err[256] = "no error";

extern "C" int myc_abi_version(void) { return MYC_ABI_VERSION; }

extern "C" int myc_create(int device_ordinal, myc_ctx** out_ctx) {
  if (!out_ctx) return MYC_ERR_BAD_ARG;
  *out_ctx = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    snprintf(g_create_err, sizeof(g_create_err), "no CUDA device: %s (this library has no CPU path)",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return MYC_ERR_CUDA;
  }
  if (device_ordinal < 0 || device_ordinal >= n_dev) {
    snprintf(g_create_err, sizeof(g_create_err), "device ordinal %d out of range (0..%d)", device_ordinal, n_dev - 1);
    return MYC_ERR_BAD_ARG;
  }
  myc_ctx* ctx = new myc_ctx();
  ctx->device = device_ordinal;
  {
    const char* e = getenv("MYC_FORCE_PLAIN_SPMV");
    ctx->force_plain_spmv = e && e[0] == '1';
    const char* f = getenv("MYC_NO_FUSED_PCG");
    ctx->no_fused_pcg = f && f[0] == '1';
    const char* g = getenv("MYC_NO_BLOCK3_SPMV");
    ctx->no_block3_spmv = g && g[0] == '1';
    const char* ss = getenv("MYC_ASM_FULL_SORT");
    ctx->asm_full_sort = ss && ss[0] == '1';
    const char* sh = getenv("MYC_ASM_SHORT_SORT");
    ctx->asm_short_sort = sh && sh[0] == '1';
    const char* d = getenv("MYC_ASM_DIRECT_FILL");
    ctx->asm_direct_fill = d && d[0] == '1';
    const char* y = getenv("MYC_NO_SYM3");
    ctx->no_sym3 = y && y[0] == '1';
    // measured on 2 GPUs (512^2 per GPU): 37.3 us/iteration with the gated sweep against 35.4 us with
    // the plain halo barrier -- the all-rank reduction barrier that follows absorbs the wait either
    // way -- so the gate is opt-in
    const char* z = getenv("MYC_HALO_OVERLAP");
    ctx->no_halo_overlap = !(z && z[0] == '1');
    const char* f64 = getenv("MYC_AMG_FP64");
    ctx->amg_fp64 = f64 && f64[0] == '1';
    const char* rn = getenv("MYC_AMG_REPLICATE_NODES");
    if (rn && rn[0]) ctx->amg_replicate_nodes = atoll(rn);
  }
  e = cudaSetDevice(device_ordinal);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device_ordinal);
  if (e == cudaSuccess && prop.major < 10) {
    snprintf(g_create_err, sizeof(g_create_err), "device %d is sm_%d%d; this library is built for sm_100a only",
             device_ordinal, prop.major, prop.minor);
    delete ctx;
    return MYC_ERR_CUDA;
  }
  if (e == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_pinned, 4096, cudaHostAllocDefault);
  for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
  if (e != cudaSuccess) {
    snprintf(g_create_err, sizeof(g_create_err), "context setup failed: %s", cudaGetErrorString(e));
    delete ctx;
    return MYC_ERR_CUDA;
  }
  *out_ctx = ctx;
  return MYC_OK;
}

extern "C" int myc_destroy(myc_ctx* ctx) {
  if (!ctx) return MYC_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  myc_dist_destroy(ctx);
  myc_amg_destroy(ctx);
  DevBuf* all[] = {&ctx->scan_tmp, &ctx->sort_keys[0], &ctx->sort_keys[1], &ctx->sort_vals[0], &ctx->sort_vals[1],
                   &ctx->sort_table, &ctx->edge_cnt, &ctx->node_deg, &ctx->node_bc, &ctx->partials, &ctx->scalars,
                   &ctx->vec[0], &ctx->vec[1], &ctx->vec[2], &ctx->vec[3], &ctx->vec[4], &ctx->vec[5], &ctx->misc,
                   &ctx->sym_val, &ctx->sym_col, &ctx->xchg};
  for (DevBuf* b : all) if (b->p) cudaFree(b->p);
  for (DevBuf& b : ctx->lc) if (b.p) cudaFree(b.p);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  for (int i = 0; i < 6; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (cudaEvent_t e : ctx->prof_ev) if (e) cudaEventDestroy(e);
  delete ctx;
  return MYC_OK;
}

extern "C" int myc_set_csr_hint(myc_ctx* ctx, int node_block3) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  ctx->csr_block3 = node_block3 != 0;
  return MYC_OK;
}

extern "C" int myc_profile_reset(myc_ctx* ctx, int enable) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  ctx->prof_on = enable != 0;
  ctx->prof_ms = ctx->prof_bytes = 0.0;
  ctx->prof_samples = ctx->prof_launches = 0;
  if (ctx->prof_on && !ctx->prof_ev[0]) {
    MYC_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < 2 * myc_ctx::PROF_PAIRS; ++i) MYC_CUDA(ctx, cudaEventCreate(&ctx->prof_ev[i]));
  }
  return MYC_OK;
}

extern "C" int myc_profile_get(myc_ctx* ctx, double* h_out4) {
  if (!ctx || !h_out4) return MYC_ERR_BAD_ARG;
  h_out4[0] = ctx->prof_ms;
  h_out4[1] = (double)ctx->prof_samples;
  h_out4[2] = ctx->prof_bytes;
  h_out4[3] = (double)ctx->prof_launches;
  return MYC_OK;
}

extern "C" const char* myc_last_error(const myc_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int64_t myc_launch_count(const myc_ctx* ctx) { return ctx ? ctx->launches : 0; }

// -------------------------------------------------------------------------------------------
extern "C" int myc_load_case_host(myc_ctx* ctx, const double* h_coords, const int32_t* h_n1, const int32_t* h_n2,
                                  const uint8_t* h_active, int64_t n_elem, int64_t n_nodes, double E, double A,
                                  double I, const int64_t* h_known_dofs, const double* h_known_vals,
                                  int64_t n_known, double reg, int precond, double rtol, int64_t maxit,
                                  const int64_t* h_react_idx, int64_t n_react, double* h_out_U,
                                  double* h_out_force, int64_t* h_out_iters, double* h_out_relres,
                                  int64_t* h_out_nnz, double* h_out_ms_assemble, double* h_out_ms_solve) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (ctx->world > 1) MYC_FAIL(ctx, MYC_ERR_STATE, "load_case_host is single-GPU; use the device API on a distributed context");
  if (n_nodes < 0 || n_elem < 0 || n_known < 0 || n_react < 0 || !h_out_U ||
      (n_nodes > 0 && !h_coords) || (n_elem > 0 && (!h_n1 || !h_n2)) ||
      (n_known > 0 && (!h_known_dofs || !h_known_vals)) || (n_react > 0 && (!h_react_idx || !h_out_force)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "load_case_host: bad argument");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = 0;
  const int64_t n_dof = 3 * n_nodes;
  enum { B_COORDS, B_N1, B_N2, B_ACT, B_KD, B_KV, B_RI, B_RP, B_CI, B_VAL, B_UBC, B_RHS, B_DINV, B_X };
  DevBuf* b = ctx->lc;
  MYC_TRY(myc_ensure(ctx, b[B_COORDS], (size_t)(n_dof + 1) * 8));
  MYC_TRY(myc_ensure(ctx, b[B_N1], (size_t)(n_elem + 1) * 4));
  MYC_TRY(myc_ensure(ctx, b[B_N2], (size_t)(n_elem + 1) * 4));
  MYC_TRY(myc_ensure(ctx, b[B_ACT], (size_t)(n_elem + 1)));
  MYC_TRY(myc_ensure(ctx, b[B_KD], (size_t)(n_known + 1) * 8));
  MYC_TRY(myc_ensure(ctx, b[B_KV], (size_t)(n_known + 1) * 8));
  MYC_TRY(myc_ensure(ctx, b[B_RI], (size_t)(n_react + 1) * 8));
  MYC_TRY(myc_ensure(ctx, b[B_RP], (size_t)(n_dof + 1) * 4));
  MYC_TRY(myc_ensure(ctx, b[B_UBC], (size_t)(n_dof + 1) * 8));
  MYC_TRY(myc_ensure(ctx, b[B_RHS], (size_t)(n_dof + 1) * 8));   // reused for F = K U and for U
  MYC_TRY(myc_ensure(ctx, b[B_DINV], (size_t)(n_dof + 1) * 8));
  MYC_TRY(myc_ensure(ctx, b[B_X], (size_t)(n_dof + 1) * 8));
  double* d_coords = (double*)b[B_COORDS].p;
  int32_t *d_n1 = (int32_t*)b[B_N1].p, *d_n2 = (int32_t*)b[B_N2].p, *d_rp = (int32_t*)b[B_RP].p;
  uint8_t* d_act = h_active ? (uint8_t*)b[B_ACT].p : nullptr;
  int64_t *d_kd = (int64_t*)b[B_KD].p, *d_ri = (int64_t*)b[B_RI].p;
  double *d_kv = (double*)b[B_KV].p, *d_ubc = (double*)b[B_UBC].p, *d_rhs = (double*)b[B_RHS].p,
         *d_dinv = (double*)b[B_DINV].p, *d_x = (double*)b[B_X].p;

  MYC_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
  MYC_CUDA(ctx, cudaMemcpyAsync(d_coords, h_coords, (size_t)n_dof * 8, cudaMemcpyHostToDevice, st));
  MYC_CUDA(ctx, cudaMemcpyAsync(d_n1, h_n1, (size_t)n_elem * 4, cudaMemcpyHostToDevice, st));
  MYC_CUDA(ctx, cudaMemcpyAsync(d_n2, h_n2, (size_t)n_elem * 4, cudaMemcpyHostToDevice, st));
  if (h_active) MYC_CUDA(ctx, cudaMemcpyAsync(d_act, h_active, (size_t)n_elem, cudaMemcpyHostToDevice, st));
  MYC_CUDA(ctx, cudaMemcpyAsync(d_kd, h_known_dofs, (size_t)n_known * 8, cudaMemcpyHostToDevice, st));
  MYC_CUDA(ctx, cudaMemcpyAsync(d_kv, h_known_vals, (size_t)n_known * 8, cudaMemcpyHostToDevice, st));
  if (n_react) MYC_CUDA(ctx, cudaMemcpyAsync(d_ri, h_react_idx, (size_t)n_react * 8, cudaMemcpyHostToDevice, st));

  int64_t nnz = 0;
  MYC_TRY(myc_assemble_symbolic(ctx, d_n1, d_n2, d_act, n_elem, n_nodes, 0, n_nodes, d_rp, &nnz, st));
  MYC_TRY(myc_ensure(ctx, b[B_CI], (size_t)(nnz + 1) * 4));
  MYC_TRY(myc_ensure(ctx, b[B_VAL], (size_t)(nnz + 1) * 8));
  int32_t* d_ci = (int32_t*)b[B_CI].p;
  double* d_val = (double*)b[B_VAL].p;
  MYC_TRY(myc_assemble_numeric(ctx, d_coords, d_n1, d_n2, E, A, I, nnz, d_rp, d_ci, d_val, st));
  // what the assembler emits always has the node-block structure; the caller's sticky hint is restored on
  // every exit path (including the early error returns below)
  struct HintGuard {
    myc_ctx* c;
    bool before;
    ~HintGuard() { c->csr_block3 = before; }
  } hint_guard{ctx, ctx->csr_block3};
  ctx->csr_block3 = true;
  MYC_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
  if (n_dof == 0) {                                   // nothing to solve
    if (h_out_iters) *h_out_iters = 0;
    if (h_out_relres) *h_out_relres = 0.0;
    if (h_out_nnz) *h_out_nnz = 0;
    if (h_out_force) *h_out_force = 0.0;
    return MYC_OK;
  }
  MYC_TRY(myc_apply_dirichlet(ctx, n_dof, n_dof, 0, d_rp, d_ci, d_val, d_kd, d_kv, n_known, reg, d_ubc, d_rhs,
                              d_dinv, st));
  double* d_binv = nullptr;
  if (precond == MYC_PC_AMG) {
    int levels = 0;
    MYC_TRY(myc_amg_setup(ctx, n_dof, n_dof, 0, d_rp, d_ci, d_val, d_dinv, reg, &levels, st));
    if (levels == 0) precond = MYC_PC_BLOCK6;        // hierarchy not applicable to this system (see myc_amg_setup)
  }
  if (precond == MYC_PC_BLOCK3) {
    MYC_TRY(myc_ensure(ctx, ctx->vec[4], (size_t)(3 * n_dof + 9) * 8));
    d_binv = (double*)ctx->vec[4].p;
    MYC_TRY(myc_block3_inverse(ctx, n_dof, 0, d_rp, d_ci, d_val, d_dinv, reg, d_binv, st));
  } else if (precond == MYC_PC_BLOCK6 || precond == MYC_PC_BLOCK12) {
    const int npb = precond == MYC_PC_BLOCK6 ? 2 : 4;
    MYC_TRY(myc_ensure(ctx, ctx->vec[4], (size_t)(myc_block_inverse_size(npb, n_dof) + 2) * 8));
    d_binv = (double*)ctx->vec[4].p;
    MYC_TRY(myc_block_inverse_packed(ctx, npb, n_dof, 0, d_rp, d_ci, d_val, d_dinv, reg, d_binv, st));
  }
  MYC_CUDA(ctx, cudaMemsetAsync(d_x, 0, (size_t)n_dof * 8, st));
  int64_t iters = 0;
  double relres = 0.0;
  int rc = myc_pcg_solve(ctx, n_dof, n_dof, 0, d_rp, d_ci, d_val, d_rhs, d_dinv, d_binv, precond, reg, rtol, 0.0,
                         maxit, d_x, &iters, &relres, st);
  if (h_out_iters) *h_out_iters = iters;
  if (h_out_relres) *h_out_relres = relres;
  if (h_out_nnz) *h_out_nnz = nnz;
  if (rc != MYC_OK && rc != MYC_ERR_NOT_CONVERGED) return rc;
  // U (into the rhs buffer), reactions F = K U (into the dinv buffer after the merge)
  double* d_U = d_rhs;
  MYC_TRY(myc_merge_solution(ctx, n_dof, 0, d_x, d_dinv, d_ubc, d_U, st));
  MYC_CUDA(ctx, cudaMemcpyAsync(h_out_U, d_U, (size_t)n_dof * 8, cudaMemcpyDeviceToHost, st));
  if (n_react) {
    double* d_F = d_dinv;
    MYC_TRY(myc_launch_spmv(ctx, n_dof, d_rp, d_ci, d_val, d_U, d_F, st));
    MYC_TRY(myc_gather_sum(ctx, d_F, d_ri, n_react, h_out_force, st));
  }
  MYC_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  float ms_a = 0.f, ms_s = 0.f;
  MYC_CUDA(ctx, cudaEventElapsedTime(&ms_a, ctx->ev[2], ctx->ev[3]));
  MYC_CUDA(ctx, cudaEventElapsedTime(&ms_s, ctx->ev[3], ctx->ev[0]));
  if (h_out_ms_assemble) *h_out_ms_assemble = ms_a;
  if (h_out_ms_solve) *h_out_ms_solve = ms_s;
  return rc;
}
