// Shared internals of libmycelium_fea_b200.so (sm_100a only).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mycelium_fea.h"

#define MYC_SM_COUNT_FALLBACK 148
#define MYC_MAX_WORLD 8          // ranks the NVLink peer-memory PCG supports (one HGX board)
#define MYC_MAXIT_LIMIT 4000000  // PCG iterations one solve may run (32-bit barrier epochs of the persistent kernels)

// Growable device scratch buffer owned by the context.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct NcclApi;   // dist.cu
struct AmgState;  // amg.cuh

struct PeerRange {  // half-open DOF range [lo, hi) of a global-length vector
  int64_t lo = 0, hi = 0;
};

struct myc_ctx {
  int device = 0;
  int sm_count = MYC_SM_COUNT_FALLBACK;
  char err[512] = {0};
  int64_t launches = 0;
  bool force_plain_spmv = false;   // MYC_FORCE_PLAIN_SPMV=1: use the non-TMA CSR-stream kernel
  bool no_fused_pcg = false;       // MYC_NO_FUSED_PCG=1: always use the multi-kernel PCG
  bool no_block3_spmv = false;     // MYC_NO_BLOCK3_SPMV=1: ignore the node-block hint
  bool no_sym3 = false;            // MYC_NO_SYM3=1: the fused PCG streams the CSR, not the symmetric block view
  bool no_halo_overlap = true;     // MYC_HALO_OVERLAP=1 enables the gated sweep (halo waits inside the sweep)
  bool asm_full_sort = false;      // MYC_ASM_FULL_SORT=1: radix passes over the whole (source, destination) key
  bool asm_short_sort = false;     // MYC_ASM_SHORT_SORT=1: radix passes over the source bits + per-node neighbour ordering
                                   // (default: no sort pass at all -- atomic placement into per-node segments + ordering)
  bool asm_direct_fill = false;    // MYC_ASM_DIRECT_FILL=1: numeric assembly stores rows straight to global memory (no staging)
  bool csr_block3 = false;         // caller's hint: the CSR it passes has the 3x3 node-block structure

  // ---- scratch arenas (grown on demand, never shrunk)
  DevBuf scan_tmp;              // block sums of the exclusive scan (all levels)
  DevBuf sort_keys[2];          // radix sort ping-pong
  DevBuf sort_vals[2];
  DevBuf sort_table;            // per-tile digit histograms
  DevBuf edge_cnt;              // per-element emitted-edge count / offsets
  DevBuf node_deg;              // per-owned-node incident edge count -> edge_start
  DevBuf node_bc;               // per-owned-node block count -> block_start
  DevBuf partials;              // per-block partial sums of the fused dots
  DevBuf scalars;               // PcgScalars + counters
  DevBuf vec[6];                // PCG work vectors (r, p(global), Ap, ...)
  DevBuf misc;                  // small temporaries (flags, gather-sum output ...)
  DevBuf xchg;                  // staging of the small host-value all-gathers (dist.cu)
  DevBuf lc[14];                // device copies owned by myc_load_case_host
  DevBuf sym_val, sym_col;      // symmetric 3x3 block view of K for the persistent solver kernels (spmv_sym3.cuh)
  int sym_owner = 0;            // 1: sym_val / sym_col are level 0 of the multigrid hierarchy below
  AmgState* amg = nullptr;      // aggregation-multigrid hierarchy of the last myc_amg_setup (amg.cuh)
  void* h_pinned = nullptr;     // 4 KB pinned staging for host scalars
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [4], [5]: multigrid setup timing

  // ---- sampled per-launch timing of the fused SpMV (bench.py roofline; off by default)
  static constexpr int PROF_PAIRS = 128;
  bool prof_on = false;
  cudaEvent_t prof_ev[2 * PROF_PAIRS] = {nullptr};
  double prof_ms = 0.0, prof_bytes = 0.0;
  int64_t prof_samples = 0, prof_launches = 0;
  int prof_op = 0;               // operator the last profiled fused solve streamed (0/1 CSR, 2 sym3)

  // ---- per-DEVICE one-time kernel setup (cudaFuncSetAttribute is per device; a process may hold one
  // context per GPU, so these flags live here and not in function-local statics)
  int fused_max_blocks_per_sm = -1;       // pcg_fused.cu: opt-in shared memory set + occupancy queried
  int amg_max_blocks_per_sm = -1;         // pcg_amg.cu

  // ---- assembly plan retained between symbolic and numeric
  bool plan_valid = false;
  int64_t plan_n_elem = 0, plan_n_nodes = 0, plan_node_begin = 0, plan_node_end = 0;
  int64_t plan_n_edges = 0, plan_nnz = 0;
  int plan_sorted_buf = 0;      // which ping-pong buffer holds the sorted edges
  int plan_dst_bits = 0;

  // ---- distributed state
  NcclApi* nccl = nullptr;
  void* comm = nullptr;         // ncclComm_t
  int rank = 0, world = 1;
  int64_t* node_offsets = nullptr;   // world+1
  PeerRange* recv_from = nullptr;    // [world] DOF ranges this rank receives from peer q
  PeerRange* send_to = nullptr;      // [world] DOF ranges this rank sends to peer q
  // NVLink peer memory of the fused multi-GPU PCG (pcg_fused.cu): one IPC-shared buffer per rank
  // holding the gathered vector u (global length) followed by the PeerSync slots/flags
  void* peer_own = nullptr;
  void* peer_base[MYC_MAX_WORLD] = {nullptr};
  int64_t peer_cap = 0;              // capacity of the u vector in doubles
  bool peer_ok = false;
  unsigned peer_epoch_red = 0, peer_epoch_halo = 0;
  unsigned amg_epoch_seam = 0;
  unsigned amg_epoch_red = 0, amg_epoch_halo = 0;     // the same for the multigrid solver kernel's own flag block
  // multigrid on several GPUs (amg_setup.cu / pcg_amg.cu): one IPC-shared buffer per rank holding the gathered
  // correction vectors of all levels followed by the AgPeerSync slots/flags
  void* amg_peer_own = nullptr;
  void* amg_peer_base[MYC_MAX_WORLD] = {nullptr};
  int64_t amg_peer_cap = 0;          // capacity of the vector arena in doubles
  bool amg_fp64 = false;             // MYC_AMG_FP64=1: the V-cycle streams the FP64 level operators (default: FP32 copies)
  int64_t amg_replicate_nodes = 65536;   // a level with at most this many nodes over all ranks is held by every rank
                                         // in full (MYC_AMG_REPLICATE_NODES overrides; tests use small values)
};

#define MYC_FAIL(ctx, code, ...)                                   \
  do {                                                             \
    if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
    return (code);                                                 \
  } while (0)

#define MYC_CUDA(ctx, call)                                                                  \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      MYC_FAIL(ctx, MYC_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,               \
               cudaGetErrorString(e__));                                                     \
  } while (0)

#define MYC_TRY(call)            \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != MYC_OK) return rc__; \
  } while (0)

// Check the launch that was just issued.
#define MYC_LAUNCHED(ctx)                                                                     \
  do {                                                                                        \
    (ctx)->launches++;                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess)                                                                   \
      MYC_FAIL(ctx, MYC_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,            \
               cudaGetErrorString(e__));                                                      \
  } while (0)

static inline int myc_ensure(myc_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return MYC_OK;
  if (b.p) MYC_CUDA(ctx, cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  MYC_CUDA(ctx, cudaMalloc(&b.p, want));
  b.cap = want;
  return MYC_OK;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Doubles per block of myc_block_inverse_packed.  R = 12: symmetric-packed upper triangle (78).  R = 6: all
// six rows in full (36), so that the solver reads a DOF's row with three contiguous 128-bit loads and no
// index arithmetic (measured 25.66 -> 24.61 us per iteration at 512^2 against the packed triangle, which
// -DMYC_BLOCK6_PACKED restores: 21 doubles).
#ifdef MYC_BLOCK6_PACKED
#define MYC_B6_FULL 0
#else
#define MYC_B6_FULL 1
#endif
__host__ __device__ __forceinline__ constexpr int myc_block_inverse_stride(int R) {
  return (R == 6 && MYC_B6_FULL) ? 36 : R * (R + 1) / 2;
}

// Symmetric-packed (upper triangle, row-major) position of entry (i, j) of an R x R block:
// the layout of myc_block_inverse_packed, read back by the fused PCG kernel.
__host__ __device__ __forceinline__ constexpr int myc_sympack(int R, int i, int j) {
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  return lo * R - lo * (lo - 1) / 2 + (hi - lo);
}

// Shared-memory carve-out preference (percent) for a kernel that runs `blocks_per_sm` blocks of `dynamic_bytes` +
// `static_bytes` shared memory on an SM: the smallest configuration of the SM's 256 KB array that holds them, so that
// the rest is L1.  Gathers and streamed operands in flight each hold an L1 line, so a kernel with many loads in flight
// per SM is throttled by a small L1 (multigrid solver kernel at 2048^2: 162.8 ms with 28 KB of L1, 145.7 ms with
// 60 KB; profiles/r2_ab_l1_carveout.md).  The driver rounds a preference UP to the next configuration, hence the
// floor; left to itself it picks the larger L1 most of the time, not always.  MYC_CARVEOUT=<percent> overrides (A/B).
// [host-test-begin carveout]  (tests/test_kernel_logic_host.py compiles this function for the host)
static inline int myc_carveout_percent(size_t dynamic_bytes, size_t static_bytes, int blocks_per_sm) {
  if (const char* e = getenv("MYC_CARVEOUT")) return atoi(e);
  const size_t need = (size_t)blocks_per_sm * (dynamic_bytes + static_bytes + 1024);   // + what the system reserves per block
  static const int config_kb[] = {8, 16, 32, 64, 100, 132, 164, 196, 228};
  for (int kb : config_kb)
    if (need <= (size_t)kb * 1024) return kb * 100 / 228;
  return 100;
}
// [host-test-end carveout]

// Grid for grid-stride kernels: whole waves of the SM count, never more than the work.
static inline int grid_for(const myc_ctx* ctx, int64_t n_tiles, int blocks_per_sm) {
  int64_t full = (int64_t)ctx->sm_count * blocks_per_sm;
  if (n_tiles < 1) n_tiles = 1;
  return (int)(n_tiles < full ? n_tiles : full);
}

// ---- internal cross-file entry points ----------------------------------------------------
// scan.cu: exclusive prefix sum of int32 counts, in place allowed; total (64-bit) to *d_total
// (device) if non-null.  n+1-th element (the total) is also written at d_out[n] when
// write_total_at_end is set (CSR-style offsets), truncated to int32 (caller checks d_total).
int myc_exclusive_scan_i32(myc_ctx* ctx, const int32_t* d_in, int32_t* d_out, int64_t n,
                           bool write_total_at_end, int64_t* d_total, cudaStream_t st);
// radix_sort.cu: stable LSD sort of (key64, val32) pairs on key bits [start_bit, key_bits).
// Returns the index (0/1) of the ping-pong buffer holding the result.
int myc_radix_sort_pairs(myc_ctx* ctx, int64_t n, int start_bit, int key_bits, int* out_buf, cudaStream_t st);

// dist.cu (collective; no-ops / copies on a single rank)
int myc_dist_allgatherv(myc_ctx* ctx, void* d_buf, const int64_t* h_off_bytes, cudaStream_t st);
int myc_dist_allgather_host_i64(myc_ctx* ctx, const int64_t* h_mine, int k, int64_t* h_all, cudaStream_t st);
int myc_dist_allgather_host_64b(myc_ctx* ctx, const void* h_mine64, void* h_all, cudaStream_t st);

// spmv.cu
int myc_launch_spmv(myc_ctx* ctx, int64_t n_rows, const int32_t* rp, const int32_t* ci,
                    const double* v, const double* x, double* y, cudaStream_t st);
