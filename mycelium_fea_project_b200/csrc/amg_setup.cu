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

// ---- several GPUs -------------------------------------------------------------------------------------------
// own section of the level's global aggregate map (global aggregate id, or -1)
__global__ void __launch_bounds__(AM_THREADS)
am_agg_global_kernel(int64_t n, int64_t node_off, const int32_t* __restrict__ agg, int32_t cnode_off,
                     int32_t* __restrict__ agg_global) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    agg_global[node_off + i] = agg[i] >= 0 ? cnode_off + agg[i] : -1;
}

// a[i] += shift where a[i] >= 0
__global__ void __launch_bounds__(AM_THREADS)
am_shift_kernel(int64_t n, int32_t* __restrict__ a, int32_t shift) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (a[i] >= 0) a[i] += shift;
}

// own section of a replicated level's block row pointer; every rank also writes the end marker
__global__ void __launch_bounds__(AM_THREADS)
am_brp_section_kernel(int64_t n_c, int64_t cnode_off, const int32_t* __restrict__ brp_local, int32_t blk_off,
                      int64_t n_c_global, int32_t nb_global, int32_t* __restrict__ brp_full) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_c; i += (int64_t)gridDim.x * blockDim.x)
    brp_full[cnode_off + i] = blk_off + brp_local[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) brp_full[n_c_global] = nb_global;
}

struct AmOffsets { int64_t off[MYC_MAX_WORLD + 1]; int world, rank; };
// per peer q: the half-open range of q's nodes that appear as block columns of this rank's rows
// (integer min / max: order independent).  lo[] preset to INT_MAX, hi[] to 0.
__global__ void __launch_bounds__(AM_THREADS)
am_halo_range_kernel(int64_t nb, const int32_t* __restrict__ bcol, AmOffsets o, int32_t* __restrict__ lo, int32_t* __restrict__ hi) {
  int32_t mylo[MYC_MAX_WORLD], myhi[MYC_MAX_WORLD];
#pragma unroll
  for (int q = 0; q < MYC_MAX_WORLD; ++q) { mylo[q] = 0x7fffffff; myhi[q] = 0; }
  const int64_t own_lo = o.off[o.rank], own_hi = o.off[o.rank + 1];
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = bcol[b] / 3;
    if (c >= own_lo && c < own_hi) continue;
#pragma unroll
    for (int q = 0; q < MYC_MAX_WORLD; ++q)
      if (q < o.world && c >= o.off[q] && c < o.off[q + 1]) {
        mylo[q] = c < mylo[q] ? c : mylo[q];
        myhi[q] = c + 1 > myhi[q] ? c + 1 : myhi[q];
      }
  }
#pragma unroll
  for (int q = 0; q < MYC_MAX_WORLD; ++q) {
    int32_t l = mylo[q], h = myhi[q];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const int32_t l2 = __shfl_xor_sync(0xffffffffu, l, d), h2 = __shfl_xor_sync(0xffffffffu, h, d);
      l = l2 < l ? l2 : l;
      h = h2 > h ? h2 : h;
    }
    if ((threadIdx.x & 31) == 0 && h > 0) { atomicMin(&lo[q], l); atomicMax(&hi[q], h); }
  }
}

// FP32 copy of a level's block values for the V-cycle's sweeps (amg_sweep.cuh)
__global__ void __launch_bounds__(AM_THREADS)
am_to_f32_kernel(int64_t n, const double* __restrict__ in, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (float)in[i];
}

int am_bits_for(int64_t n) {
  int b = 1;
  while (((int64_t)1 << b) < n) ++b;
  return b;
}

}  // namespace

size_t myc_amg_peer_tail_bytes();      // pcg_amg.cu: the AgPeerSync block behind the arena

namespace {
void am_drop(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

// CUDA IPC rule: the importing processes close their mappings BEFORE the exporting process frees the allocation.
// So teardown is two steps with a rank barrier in between (am_peer_ensure, myc_dist_release_peers + myc_destroy).
void am_close_imports(myc_ctx* ctx) {
  for (int q = 0; q < MYC_MAX_WORLD; ++q) {
    if (ctx->amg_peer_base[q] && ctx->amg_peer_base[q] != ctx->amg_peer_own) cudaIpcCloseMemHandle(ctx->amg_peer_base[q]);
    ctx->amg_peer_base[q] = nullptr;
  }
  if (ctx->amg) ctx->amg->valid = false;          // a multi-GPU hierarchy without its arena cannot be solved with
}
void am_free_own(myc_ctx* ctx) {
  if (ctx->amg_peer_own) cudaFree(ctx->amg_peer_own);
  ctx->amg_peer_own = nullptr;
  ctx->amg_peer_cap = 0;
}
void am_close_peers(myc_ctx* ctx) {
  am_close_imports(ctx);
  am_free_own(ctx);
}

// Several GPUs: the vector arena lives in one IPC-shared allocation per rank (vectors, then the AgPeerSync block).
// Collective; `doubles` is the same number on every rank, so every rank takes the same branch.  *ok = 0 if a peer
// mapping could not be opened on SOME rank (then no rank uses the multigrid path).
int am_peer_ensure(myc_ctx* ctx, int64_t doubles, cudaStream_t st, int* ok) {
  *ok = 1;
  if (ctx->amg_peer_own && ctx->amg_peer_cap >= doubles) return MYC_OK;
  MYC_CUDA(ctx, cudaDeviceSynchronize());
  // nobody may still be storing into the old buffers: all ranks pass this exchange before any of them unmaps,
  // and every rank has unmapped its imports before any owner frees
  int64_t token = 1, all[MYC_MAX_WORLD];
  MYC_TRY(myc_dist_allgather_host_i64(ctx, &token, 1, all, st));
  am_close_imports(ctx);
  MYC_TRY(myc_dist_allgather_host_i64(ctx, &token, 1, all, st));
  am_free_own(ctx);
  const int64_t cap = doubles + doubles / 4 + 1024;
  const size_t vec_bytes = ((size_t)cap * sizeof(double) + 255) / 256 * 256;
  const size_t bytes = vec_bytes + myc_amg_peer_tail_bytes();
  MYC_CUDA(ctx, cudaMalloc(&ctx->amg_peer_own, bytes));
  MYC_CUDA(ctx, cudaMemset(ctx->amg_peer_own, 0, bytes));
  MYC_CUDA(ctx, cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  MYC_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->amg_peer_own));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  unsigned char handles[64 * MYC_MAX_WORLD];
  MYC_TRY(myc_dist_allgather_host_64b(ctx, &h, handles, st));
  int64_t mine_ok = 1;
  for (int q = 0; q < ctx->world; ++q) {
    if (q == ctx->rank) { ctx->amg_peer_base[q] = ctx->amg_peer_own; continue; }
    cudaIpcMemHandle_t hq;
    memcpy(&hq, handles + 64 * (size_t)q, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      mine_ok = 0;
      break;
    }
    ctx->amg_peer_base[q] = p;
  }
  MYC_TRY(myc_dist_allgather_host_i64(ctx, &mine_ok, 1, all, st));
  for (int q = 0; q < ctx->world; ++q) if (!all[q]) *ok = 0;
  if (!*ok) {
    am_close_peers(ctx);
    return MYC_OK;
  }
  ctx->amg_peer_cap = cap;
  ctx->amg_epoch_red = ctx->amg_epoch_halo = ctx->amg_epoch_seam = 0;      // the flag block is new (zeroed) on every rank
  return MYC_OK;
}
}  // namespace

void* myc_amg_peer_sync_of(const myc_ctx* ctx, int q) {
  const size_t vec_bytes = ((size_t)ctx->amg_peer_cap * sizeof(double) + 255) / 256 * 256;
  return (char*)ctx->amg_peer_base[q] + vec_bytes;
}

void myc_amg_close_imports(myc_ctx* ctx) { am_close_imports(ctx); }

int myc_amg_destroy(myc_ctx* ctx) {
  am_close_peers(ctx);
  AmgState* s = ctx->amg;
  if (!s) return MYC_OK;
  for (AmgLevelHost& L : s->lv) {
    am_drop(L.brp); am_drop(L.bcol); am_drop(L.bval); am_drop(L.dinv); am_drop(L.agg); am_drop(L.mptr); am_drop(L.mlist);
    am_drop(L.r); am_drop(L.t); am_drop(L.bval32); am_drop(L.rep_brp); am_drop(L.rep_bcol); am_drop(L.rep_bval);
  }
  am_drop(s->lv_dev); am_drop(s->brp0); am_drop(s->arena); am_drop(s->act0); am_drop(s->act_global); am_drop(s->agg_global);
  for (DevBuf& b : s->work) am_drop(b);
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
  const int world = ctx->world, rank = ctx->rank;
  const bool dist = world > 1;
  if (dist && (world > MYC_MAX_WORLD || !ctx->comm || !ctx->node_offsets))
    MYC_FAIL(ctx, MYC_ERR_STATE, "amg_setup: multi-GPU context without communicator / partition (myc_dist_init, myc_dist_set_plan)");
  if (dist && (3 * ctx->node_offsets[rank] != row_offset || 3 * ctx->node_offsets[rank + 1] != row_offset + n_rows ||
               3 * ctx->node_offsets[world] != n_cols_global))
    MYC_FAIL(ctx, MYC_ERR_STATE, "amg_setup: the row block does not match the installed partition");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  // The hierarchy is built on the symmetric 3x3 node-block view of K: needs the node-block structure.  On several
  // GPUs the verdict is collective (every rank must take the same path), so local findings are only recorded here.
  int64_t local_ok = (ctx->csr_block3 && n_rows % 3 == 0 && row_offset % 3 == 0 && (dist || n_rows > 0) &&
                      (n_rows == 0 || (((uintptr_t)d_col_idx | (uintptr_t)d_val) & 15u) == 0)) ? 1 : 0;
  if (!dist && !local_ok) return MYC_OK;
  ctx->plan_valid = false;                  // the sort buffers of the assembly plan are reused below
  MYC_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
  const int64_t n0 = n_rows / 3;
  int64_t* h_pin = (int64_t*)ctx->h_pinned;
  MYC_TRY(myc_ensure(ctx, ctx->misc, 512));
  int* bad = (int*)((char*)ctx->misc.p + 320);
  int64_t* d_total = (int64_t*)((char*)ctx->misc.p + 64);
  int32_t* d_range = (int32_t*)((char*)ctx->misc.p + 384);     // [2][MYC_MAX_WORLD] halo ranges of one level

  // ---- level 0: symmetric block view of K (also what the solver sweeps), block row pointer, activity
  int64_t nb0 = 0;
  if (local_ok) {
    int32_t h_nnz = 0;
    MYC_CUDA(ctx, cudaMemcpyAsync(&h_nnz, d_row_ptr + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    nb0 = h_nnz / 9;
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
    if (h_bad[0] || h_bad[1]) local_ok = 0;   // K not blockwise symmetric, or a node with a partial Dirichlet set
  }
  int64_t h_all[32 * MYC_MAX_WORLD];
  MYC_TRY(myc_dist_allgather_host_i64(ctx, &local_ok, 1, h_all, st));
  for (int q = 0; q < world; ++q) if (!h_all[q]) return MYC_OK;
  ctx->sym_owner = 1;                       // sym_val / sym_col now belong to this hierarchy
  S->world = world;

  // partition of the level being coarsened: rank q owns the nodes [lvl_off[q], lvl_off[q+1]) of the level
  int64_t lvl_off[MYC_MAX_WORLD + 1] = {0};
  if (dist) for (int q = 0; q <= world; ++q) lvl_off[q] = ctx->node_offsets[q];
  else lvl_off[1] = n0;
  const uint8_t* act_global = nullptr;
  if (dist) {                               // activity of the remote ends of the cut blocks (keep rule)
    const int64_t ng = n_cols_global / 3;
    MYC_TRY(myc_ensure(ctx, S->act_global, (size_t)ng + 16));
    if (n0 > 0)
      MYC_CUDA(ctx, cudaMemcpyAsync((uint8_t*)S->act_global.p + lvl_off[rank], S->act0.p, (size_t)n0, cudaMemcpyDeviceToDevice, st));
    MYC_TRY(myc_dist_allgatherv(ctx, S->act_global.p, lvl_off, st));
    act_global = (const uint8_t*)S->act_global.p;
  }

  // level 0 views (not owned by the level: K's block view lives in ctx->sym_*)
  AmgLevelHost* L = &S->lv[0];
  L->n = n0; L->nb = nb0; L->n_global = n_cols_global / 3; L->node_off = row_offset / 3;
  L->replicated = 0; L->own_lo = L->own_n = 0; L->agg_shift = 0; L->r_off = -1;
  const int32_t* brp = (const int32_t*)S->brp0.p;
  const int32_t* bcol = (const int32_t*)ctx->sym_col.p;
  const double* bval = (const double*)ctx->sym_val.p;
  const uint8_t* act = (const uint8_t*)S->act0.p;
  int lv = 0;
  for (;;) {
    L = &S->lv[lv];
    const int64_t n = L->n, nb = L->nb;
    const bool part = dist && !L->replicated;        // this level's rows are partitioned over the ranks
    const int g_n = grid_for(ctx, ceil_div64(n, AM_THREADS), 8);
    for (int q = 0; q < MYC_MAX_WORLD; ++q) L->need_lo[q] = L->need_hi[q] = L->give_lo[q] = L->give_hi[q] = 0;
    MYC_TRY(myc_ensure(ctx, L->dinv, (size_t)(n + 1) * 6 * sizeof(double)));
    am_dinv_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, brp, bcol, bval, act, reg, (double*)L->dinv.p);
    MYC_LAUNCHED(ctx);
    if (part) {
      // ---- halo plan of the level: which of q's rows do my blocks gather, and (transposed) who gathers mine
      AmOffsets o;
      for (int q = 0; q <= MYC_MAX_WORLD; ++q) o.off[q] = q <= world ? lvl_off[q] : lvl_off[world];
      o.world = world; o.rank = rank;
      int32_t h_range[2 * MYC_MAX_WORLD];
      for (int q = 0; q < MYC_MAX_WORLD; ++q) { h_range[q] = 0x7fffffff; h_range[MYC_MAX_WORLD + q] = 0; }
      MYC_CUDA(ctx, cudaMemcpyAsync(d_range, h_range, sizeof(h_range), cudaMemcpyHostToDevice, st));
      if (nb > 0) {
        am_halo_range_kernel<<<grid_for(ctx, ceil_div64(nb, AM_THREADS), 8), AM_THREADS, 0, st>>>(nb, bcol, o, d_range,
                                                                                                  d_range + MYC_MAX_WORLD);
        MYC_LAUNCHED(ctx);
      }
      MYC_CUDA(ctx, cudaMemcpyAsync(h_range, d_range, sizeof(h_range), cudaMemcpyDeviceToHost, st));
      MYC_CUDA(ctx, cudaStreamSynchronize(st));
      int64_t mine[2 * MYC_MAX_WORLD];
      for (int q = 0; q < MYC_MAX_WORLD; ++q) {
        const bool any = q < world && q != rank && h_range[MYC_MAX_WORLD + q] > 0;
        mine[q] = L->need_lo[q] = any ? h_range[q] : 0;
        mine[MYC_MAX_WORLD + q] = L->need_hi[q] = any ? h_range[MYC_MAX_WORLD + q] : 0;
      }
      MYC_TRY(myc_dist_allgather_host_i64(ctx, mine, 2 * MYC_MAX_WORLD, h_all, st));
      for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        L->give_lo[q] = h_all[q * 2 * MYC_MAX_WORLD + rank];
        L->give_hi[q] = h_all[q * 2 * MYC_MAX_WORLD + MYC_MAX_WORLD + rank];
      }
    }
    if (lv + 1 >= AMG_MAX_LEVELS || L->n_global <= AMG_MIN_NODES) break;
    // ---- aggregates (rank-local on a partitioned level)
    for (int k = 0; k < 4; ++k) MYC_TRY(myc_ensure(ctx, S->work[k], (size_t)(n + 2) * sizeof(int32_t)));
    MYC_TRY(myc_ensure(ctx, S->work[4], (size_t)((nb > n ? nb : n) + 2) * sizeof(int32_t)));
    MYC_TRY(myc_ensure(ctx, S->work[5], (size_t)((nb > n ? nb : n) + 2) * sizeof(int32_t)));
    int32_t* best = (int32_t*)S->work[0].p;
    int32_t* paired = (int32_t*)S->work[1].p;
    int32_t* root = (int32_t*)S->work[2].p;
    int32_t* keep = (int32_t*)S->work[3].p;
    int32_t* flag = (int32_t*)S->work[4].p;
    int32_t* uidx = (int32_t*)S->work[5].p;
    am_propose_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, brp, bcol, bval, act, best);
    MYC_LAUNCHED(ctx);
    am_accept_kernel<<<g_n, AM_THREADS, 0, st>>>(n, best, paired);
    MYC_LAUNCHED(ctx);
    am_root_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, brp, bcol, bval, act, paired, root);
    MYC_LAUNCHED(ctx);
    MYC_CUDA(ctx, cudaMemsetAsync(keep, 0, (size_t)(n + 1) * sizeof(int32_t), st));
    am_keep_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, brp, bcol, act, lv == 0 ? act_global : nullptr, root, keep);
    MYC_LAUNCHED(ctx);
    am_lead_kernel<<<g_n, AM_THREADS, 0, st>>>(n, act, root, keep, flag);
    MYC_LAUNCHED(ctx);
    MYC_TRY(myc_exclusive_scan_i32(ctx, flag, flag, n, false, d_total, st));
    MYC_CUDA(ctx, cudaMemcpyAsync(h_pin, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t n_c = h_pin[0];
    // numbering over the ranks: rank-major, so the coarse level inherits the contiguous partition
    int64_t c_off[MYC_MAX_WORLD + 1] = {0};
    int64_t n_c_global = n_c, cnode_off = 0;
    if (part) {
      MYC_TRY(myc_dist_allgather_host_i64(ctx, &n_c, 1, h_all, st));
      for (int q = 0; q < world; ++q) c_off[q + 1] = c_off[q] + h_all[q];
      n_c_global = c_off[world];
      cnode_off = c_off[rank];
    } else {
      c_off[1] = n_c;
    }
    if (n_c_global == 0 || (double)n_c_global > AMG_MAX_RATIO * (double)L->n_global) break;
    if (n_c_global >= ((int64_t)1 << 30)) MYC_FAIL(ctx, MYC_ERR_CAPACITY, "amg_setup: too many aggregates for 32-bit ids");
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
    // ---- Galerkin operator of the aggregates (columns: GLOBAL aggregate ids)
    const int32_t* agg_global = nullptr;
    if (part) {
      MYC_TRY(myc_ensure(ctx, S->agg_global, (size_t)(L->n_global + 4) * sizeof(int32_t)));
      am_agg_global_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, (const int32_t*)L->agg.p, (int32_t)cnode_off,
                                                       (int32_t*)S->agg_global.p);
      MYC_LAUNCHED(ctx);
      int64_t boff[MYC_MAX_WORLD + 1];
      for (int q = 0; q <= world; ++q) boff[q] = 4 * lvl_off[q];
      MYC_TRY(myc_dist_allgatherv(ctx, S->agg_global.p, boff, st));
      agg_global = (const int32_t*)S->agg_global.p;
    }
    const int cbits = am_bits_for(n_c_global + 1);
    const int rbits = am_bits_for(n_c + 1);
    am_coarse_emit_kernel<<<g_n, AM_THREADS, 0, st>>>(n, L->node_off, brp, bcol, act, (const int32_t*)L->agg.p, agg_global,
                                                      cnode_off, (int32_t)n_c, cbits, (uint64_t*)ctx->sort_keys[0].p,
                                                      (uint32_t*)ctx->sort_vals[0].p);
    MYC_LAUNCHED(ctx);
    MYC_TRY(myc_radix_sort_pairs(ctx, nb, 0, cbits + rbits, &sorted, st));
    const uint64_t* skeys = (const uint64_t*)ctx->sort_keys[sorted].p;
    const uint32_t* svals = (const uint32_t*)ctx->sort_vals[sorted].p;
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
    if (nb == 0) MYC_CUDA(ctx, cudaMemsetAsync(C->brp.p, 0, (size_t)(n_c + 2) * sizeof(int32_t), st));
    am_coarse_fill_kernel<<<g_b, AM_THREADS, 0, st>>>(nb, (int32_t)n_c, cbits, skeys, svals, flag, uidx, bval, (int32_t)n_cb,
                                                      (int32_t*)C->brp.p, (int32_t*)C->bcol.p, (double*)C->bval.p);
    MYC_LAUNCHED(ctx);
    C->n = n_c; C->nb = n_cb; C->n_global = n_c_global; C->node_off = cnode_off;
    C->replicated = L->replicated ? 2 : 0;
    C->own_lo = C->own_n = 0; C->agg_shift = 0; C->r_off = -1;
    L->agg_shift = 0;
    if (part && n_c_global <= ctx->amg_replicate_nodes) {
      // ---- the seam: gather the coarse operator onto every rank; from here down everything is redundant
      int64_t cb_off[MYC_MAX_WORLD + 1] = {0};
      MYC_TRY(myc_dist_allgather_host_i64(ctx, &n_cb, 1, h_all, st));
      for (int q = 0; q < world; ++q) cb_off[q + 1] = cb_off[q] + h_all[q];
      const int64_t nb_g = cb_off[world];
      if (nb_g >= ((int64_t)1 << 31)) MYC_FAIL(ctx, MYC_ERR_CAPACITY, "amg_setup: replicated level too large");
      DevBuf& f_brp = C->rep_brp;
      DevBuf& f_bcol = C->rep_bcol;
      DevBuf& f_bval = C->rep_bval;
      MYC_TRY(myc_ensure(ctx, f_brp, (size_t)(n_c_global + 2) * sizeof(int32_t)));
      MYC_TRY(myc_ensure(ctx, f_bcol, (size_t)(nb_g + 4) * sizeof(int32_t)));
      MYC_TRY(myc_ensure(ctx, f_bval, (size_t)(nb_g + 4) * 6 * sizeof(double)));
      am_brp_section_kernel<<<grid_for(ctx, ceil_div64(n_c, AM_THREADS), 8), AM_THREADS, 0, st>>>(
          n_c, cnode_off, (const int32_t*)C->brp.p, (int32_t)cb_off[rank], n_c_global, (int32_t)nb_g, (int32_t*)f_brp.p);
      MYC_LAUNCHED(ctx);
      if (n_cb > 0) {
        MYC_CUDA(ctx, cudaMemcpyAsync((int32_t*)f_bcol.p + cb_off[rank], C->bcol.p, (size_t)n_cb * sizeof(int32_t),
                                      cudaMemcpyDeviceToDevice, st));
        MYC_CUDA(ctx, cudaMemcpyAsync((double*)f_bval.p + 6 * cb_off[rank], C->bval.p, (size_t)n_cb * 6 * sizeof(double),
                                      cudaMemcpyDeviceToDevice, st));
      }
      int64_t boff[MYC_MAX_WORLD + 1];
      for (int q = 0; q <= world; ++q) boff[q] = 4 * c_off[q];
      MYC_TRY(myc_dist_allgatherv(ctx, f_brp.p, boff, st));
      for (int q = 0; q <= world; ++q) boff[q] = 4 * cb_off[q];
      MYC_TRY(myc_dist_allgatherv(ctx, f_bcol.p, boff, st));
      for (int q = 0; q <= world; ++q) boff[q] = 48 * cb_off[q];
      MYC_TRY(myc_dist_allgatherv(ctx, f_bval.p, boff, st));
      C->n = n_c_global; C->nb = nb_g; C->node_off = 0;
      C->replicated = 1; C->own_lo = cnode_off; C->own_n = n_c;
      // agg[] of the level above indexes the coarse level relative to ITS node_off, which is now 0
      if (n > 0 && cnode_off > 0) {
        am_shift_kernel<<<g_n, AM_THREADS, 0, st>>>(n, (int32_t*)L->agg.p, (int32_t)cnode_off);
        MYC_LAUNCHED(ctx);
      }
      L->agg_shift = cnode_off;
      c_off[0] = 0; c_off[1] = n_c_global;
    }
    for (int q = 0; q <= world; ++q) lvl_off[q] = (dist && !C->replicated) ? c_off[q] : (q == 0 ? 0 : C->n);
    brp = (const int32_t*)(C->replicated == 1 ? C->rep_brp.p : C->brp.p);
    bcol = (const int32_t*)(C->replicated == 1 ? C->rep_bcol.p : C->bcol.p);
    bval = (const double*)(C->replicated == 1 ? C->rep_bval.p : C->bval.p);
    act = nullptr;
    ++lv;
  }
  S->n_levels = lv + 1;
  // ---- FP32 copies of the level operators for the V-cycle's sweeps
  S->f32 = !ctx->amg_fp64;
  if (S->f32)
    for (int l = 0; l < S->n_levels; ++l) {
      AmgLevelHost& H = S->lv[l];
      const double* src = l == 0 ? (const double*)ctx->sym_val.p
                                 : (const double*)(H.replicated == 1 ? H.rep_bval.p : H.bval.p);
      MYC_TRY(myc_ensure(ctx, H.bval32, (size_t)(H.nb + 4) * 6 * sizeof(float)));
      if (H.nb > 0) {
        am_to_f32_kernel<<<grid_for(ctx, ceil_div64(6 * H.nb, AM_THREADS), 8), AM_THREADS, 0, st>>>(6 * H.nb, src, (float*)H.bval32.p);
        MYC_LAUNCHED(ctx);
      }
    }
  // ---- vectors: r, t per level (level 0's r is the CG residual); arena of the gathered correction vectors
  int64_t off = 0;
  for (int l = 0; l < S->n_levels; ++l) {
    AmgLevelHost& H = S->lv[l];
    if (l > 0 && H.replicated != 1) MYC_TRY(myc_ensure(ctx, H.r, (size_t)(3 * H.n + 4) * sizeof(double)));
    MYC_TRY(myc_ensure(ctx, H.t, (size_t)(3 * H.n + 4) * sizeof(double)));
    for (int k = 0; k < 2; ++k) {
      H.e_off[k] = off;
      off += (3 * H.n_global + 15) / 16 * 16;
    }
    H.r_off = -1;
    if (H.replicated == 1) {                 // the seam level's right-hand side is assembled from every rank's part
      H.r_off = off;
      off += (3 * H.n_global + 15) / 16 * 16;
    }
  }
  S->arena_doubles = off;
  if (dist) {
    int ok = 1;
    MYC_TRY(am_peer_ensure(ctx, off + 16, st, &ok));
    if (!ok) return MYC_OK;                  // no P2P path between some pair of GPUs: collective fallback
  } else {
    MYC_TRY(myc_ensure(ctx, S->arena, (size_t)(off + 16) * sizeof(double) + 512 + myc_amg_peer_tail_bytes()));
  }
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
    const double mat64 = 52.0 * nb + 4.0 * rows / 3.0;        // FP64 block view + block row pointer
    const double mat = (S->f32 ? 28.0 : 52.0) * nb + 4.0 * rows / 3.0;   // what the V-cycle's sweeps stream
    const bool coarsest = l == S->n_levels - 1;
    // sweeps: matrix + gathered e (8) + r (8) + own e (8) + write (8) per row; smoothing sweeps also read dinv (16)
    const double sweeps = coarsest ? (AMG_COARSE_SWEEPS - 1) : 2;
    bytes += sweeps * (mat + 32.0 * rows) + (coarsest ? sweeps : 1.0) * 16.0 * rows;
    if (l == 0) {
      bytes += mat64 + 32.0 * rows;                            // w = A u: u gathered, u own, r read, w written
      bytes += (13.0 * 8.0) * rows;                            // D0: u w p s x r mask dinv(2) read, p s x r e written
    } else {
      bytes += (8.0 + 8.0 + 16.0 + 8.0) * rows + 4.0 * rows;   // restriction: t of the members, r and e written, dinv, lists
    }
    if (!coarsest) bytes += (8.0 + 8.0 + 8.0) * rows + 4.0 * rows / 3.0;   // prolongation: e read + written, coarse e, agg
  }
  return bytes;
}

// Introspection for tests and reports.  out[0] = nodes and out[1] = blocks of `level` held by this rank, out[2] =
// levels, out[3] = setup time in microseconds, out[4] = global id of the first held node, out[5] = nodes of the
// level over all ranks, out[6] = 0 partitioned / 1 first replicated level / 2 replicated, out[7] = what has been
// added to the level's aggregate map to make it index the next level from that level's first held node.
// d_out_agg (may be NULL): the level's aggregate map, out[0] int32.
extern "C" int myc_amg_level_info(myc_ctx* ctx, int level, int64_t* h_out8, int32_t* d_out_agg, void* stream) {
  if (!ctx || !h_out8) return MYC_ERR_BAD_ARG;
  AmgState* S = ctx->amg;
  if (!S || !S->valid || level < 0 || level >= S->n_levels) MYC_FAIL(ctx, MYC_ERR_STATE, "amg_level_info: no such level");
  const AmgLevelHost& H = S->lv[level];
  h_out8[0] = H.n;
  h_out8[1] = H.nb;
  h_out8[2] = S->n_levels;
  h_out8[3] = (int64_t)(S->setup_ms * 1e3);
  h_out8[4] = H.node_off;
  h_out8[5] = H.n_global;
  h_out8[6] = H.replicated;
  h_out8[7] = H.agg_shift;
  if (d_out_agg) {
    if (level + 1 >= S->n_levels) MYC_FAIL(ctx, MYC_ERR_STATE, "amg_level_info: the coarsest level has no aggregates");
    MYC_CUDA(ctx, cudaMemcpyAsync(d_out_agg, H.agg.p, (size_t)H.n * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                  (cudaStream_t)stream));
  }
  return MYC_OK;
}
