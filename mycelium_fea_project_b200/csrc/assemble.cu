// K2+K3: global stiffness assembly straight to CSR (replaces src/fea_solver.py:74-106).
//
// The reference emits 36 COO triplets per element and lets scipy sort / merge them.  Every
// triplet of an element lives in one of four 3x3 node blocks, and the two diagonal blocks
// exist for any node with an incident active element, so the CSR *structure* is fixed by
// the directed node pairs (n1->n2, n2->n1) alone.  Pipeline (all on the caller's stream):
//
//   symbolic  per-node degree (integer atomics) -> exclusive scan = the node's segment of the pair stream ->
//             every element drops its directed pairs (key = src_local<<dst_bits | dst, value = element id)
//             into the segments of its end nodes (slot by integer atomic: arrival order is arbitrary) ->
//             each node orders its few pairs by (destination, element id) -- a TOTAL order, so the stream is
//             bit-reproducible and equals what a stable sort of the element-ordered pairs by (source,
//             destination) gives, i.e. scipy's COO->CSR order -> per node: unique neighbours (+1 for the
//             diagonal block) -> exclusive scan -> row_ptr.  No sort pass over the stream at all; a mesh
//             with a hub node (> 64 incident pairs) takes the radix-sort route instead (radix_sort.cu).
//   numeric   one thread per owned node walks its sorted pair segment, evaluates S_e in
//             registers (ke.cuh -- the COO stream and K_e are never materialised), sums
//             duplicates in element order and the diagonal in (neighbour, element) order
//             (fixed order => deterministic, "segmented scatter-add" without atomics), and
//             writes the node's three CSR rows: col_idx ascending, explicit zeros kept.  A warp
//             stages the contiguous window of its 32 nodes in shared memory and stores it coalesced.
//
// An element with n1 == n2 contributes S - S - S + S = 0 to its node's diagonal block: it
// only makes the block exist, exactly as in the reference.
#include "common.cuh"
#include "ke.cuh"

namespace {

// [host-test-begin assembly_kernels]  (tests/test_kernel_logic_host.py compiles this text with g++)
constexpr int AS_THREADS = 256;

__device__ __forceinline__ bool owned(int32_t node, int64_t nb, int64_t ne) {
  return node >= nb && node < ne;
}

__global__ void __launch_bounds__(AS_THREADS)
edge_count_kernel(const int32_t* __restrict__ n1, const int32_t* __restrict__ n2,
                  const uint8_t* __restrict__ active, int64_t n_elem, int64_t n_nodes, int64_t nb,
                  int64_t ne, int32_t* __restrict__ cnt, int* __restrict__ bad_flag) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem;
       e += (int64_t)gridDim.x * blockDim.x) {
    int c = 0;
    if (!active || active[e]) {
      const int32_t a = n1[e], b = n2[e];
      if (a < 0 || b < 0 || a >= n_nodes || b >= n_nodes) {
        *bad_flag = 1;
      } else if (a == b) {
        c = owned(a, nb, ne) ? 1 : 0;
      } else {
        c = (owned(a, nb, ne) ? 1 : 0) + (owned(b, nb, ne) ? 1 : 0);
      }
    }
    cnt[e] = c;
  }
}

__global__ void __launch_bounds__(AS_THREADS)
edge_emit_kernel(const int32_t* __restrict__ n1, const int32_t* __restrict__ n2,
                 const uint8_t* __restrict__ active, int64_t n_elem, int64_t nb, int64_t ne,
                 int dst_bits, const int32_t* __restrict__ offs, uint64_t* __restrict__ keys,
                 uint32_t* __restrict__ vals, int32_t* __restrict__ deg) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem;
       e += (int64_t)gridDim.x * blockDim.x) {
    if (active && !active[e]) continue;
    const int32_t a = n1[e], b = n2[e];
    int64_t o = offs[e];
    if (owned(a, nb, ne)) {
      keys[o] = ((uint64_t)(a - nb) << dst_bits) | (uint64_t)b;
      vals[o] = (uint32_t)e;
      atomicAdd(&deg[a - nb], 1);
      ++o;
    }
    if (b != a && owned(b, nb, ne)) {
      keys[o] = ((uint64_t)(b - nb) << dst_bits) | (uint64_t)a;
      vals[o] = (uint32_t)e;
      atomicAdd(&deg[b - nb], 1);
    }
  }
}

// Sort-free route (default): degrees first ...
__global__ void __launch_bounds__(AS_THREADS)
edge_degree_kernel(const int32_t* __restrict__ n1, const int32_t* __restrict__ n2,
                   const uint8_t* __restrict__ active, int64_t n_elem, int64_t n_nodes, int64_t nb,
                   int64_t ne, int32_t* __restrict__ deg, int* __restrict__ bad_flag) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem;
       e += (int64_t)gridDim.x * blockDim.x) {
    if (active && !active[e]) continue;
    const int32_t a = n1[e], b = n2[e];
    if (a < 0 || b < 0 || a >= n_nodes || b >= n_nodes) { *bad_flag = 1; continue; }
    if (owned(a, nb, ne)) atomicAdd(&deg[a - nb], 1);
    if (b != a && owned(b, nb, ne)) atomicAdd(&deg[b - nb], 1);
  }
}

// ... then every directed pair takes the next free slot of its source node's segment.  Which slot is a race;
// segment_order_kernel removes every trace of it.
__global__ void __launch_bounds__(AS_THREADS)
edge_place_kernel(const int32_t* __restrict__ n1, const int32_t* __restrict__ n2,
                  const uint8_t* __restrict__ active, int64_t n_elem, int64_t n_nodes, int64_t nb, int64_t ne,
                  int dst_bits, const int32_t* __restrict__ edge_start, int32_t* __restrict__ fill,
                  uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem;
       e += (int64_t)gridDim.x * blockDim.x) {
    if (active && !active[e]) continue;
    const int32_t a = n1[e], b = n2[e];
    if (a < 0 || b < 0 || a >= n_nodes || b >= n_nodes) continue;
    if (owned(a, nb, ne)) {
      const int32_t o = edge_start[a - nb] + atomicAdd(&fill[a - nb], 1);
      keys[o] = ((uint64_t)(a - nb) << dst_bits) | (uint64_t)b;
      vals[o] = (uint32_t)e;
    }
    if (b != a && owned(b, nb, ne)) {
      const int32_t o = edge_start[b - nb] + atomicAdd(&fill[b - nb], 1);
      keys[o] = ((uint64_t)(b - nb) << dst_bits) | (uint64_t)a;
      vals[o] = (uint32_t)e;
    }
  }
}

// Short sort (MYC_ASM_SHORT_SORT=1; MYC_ASM_FULL_SORT=1 forces the full-key sort): the radix passes cover the source-node
// bits only, so a node's segment arrives in emission (= element) order; this kernel orders it by destination
// node (ties: element id) with an insertion sort -- the same permutation the full-key sort produces, at
// O(degree^2) per node (degree <= ~6 on hyphal networks); the sort-free route uses it on segments that arrive
// in arbitrary order.  A hub of degree d would cost d^2/2 moves in one thread, so a node
// with more than AS_SHORT_SORT_MAX_DEG incident pairs raises *too_long and the host redoes the full sort.
// Measured at 2048^2 (profiles/r2_ab_gpu1.md): assembly 1.93 -> 1.46 ms, CSR bit-identical.
constexpr int AS_SHORT_SORT_MAX_DEG = 64;
__global__ void __launch_bounds__(AS_THREADS)
segment_order_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, const int32_t* __restrict__ edge_start,
                     int64_t n_local, int dst_bits, int* __restrict__ too_long) {
  const uint64_t dmask = ((uint64_t)1 << dst_bits) - 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_local;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t es = edge_start[i], ee = edge_start[i + 1];
    if (ee - es > AS_SHORT_SORT_MAX_DEG) { *too_long = 1; continue; }
    for (int32_t k = es + 1; k < ee; ++k) {
      const uint64_t kk = keys[k];
      const uint32_t vv = vals[k];
      const uint64_t d = kk & dmask;
      int32_t j = k - 1;
      // (destination, element id): equal destinations end up in element order whatever the order of arrival
      while (j >= es && ((keys[j] & dmask) > d || ((keys[j] & dmask) == d && vals[j] > vv))) {
        keys[j + 1] = keys[j];
        vals[j + 1] = vals[j];
        --j;
      }
      keys[j + 1] = kk;
      vals[j + 1] = vv;
    }
  }
}

__global__ void __launch_bounds__(AS_THREADS)
block_count_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ edge_start,
                   int64_t n_local, int64_t nb, int dst_bits, int32_t* __restrict__ bc) {
  const uint64_t dmask = ((uint64_t)1 << dst_bits) - 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_local;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t es = edge_start[i], ee = edge_start[i + 1];
    int c = 0;
    if (ee > es) {
      c = 1;  // diagonal block
      const int64_t self = nb + i;
      int64_t prev = -1;
      for (int32_t k = es; k < ee; ++k) {
        const int64_t dst = (int64_t)(keys[k] & dmask);
        if (dst != prev && dst != self) ++c;
        prev = dst;
      }
    }
    bc[i] = c;
  }
}

__global__ void __launch_bounds__(AS_THREADS)
row_ptr_kernel(const int32_t* __restrict__ block_start, int64_t n_local, int32_t* __restrict__ row_ptr) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_local;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t bs = block_start[i];
    if (i == n_local) {
      row_ptr[3 * i] = 9 * bs;
    } else {
      const int32_t w = 3 * (block_start[i + 1] - bs);
      row_ptr[3 * i] = 9 * bs;
      row_ptr[3 * i + 1] = 9 * bs + w;
      row_ptr[3 * i + 2] = 9 * bs + 2 * w;
    }
  }
}

__device__ __forceinline__ void write_block(int32_t* col_idx, double* val, int64_t row0, int32_t w, int b,
                                            int64_t col_node, const Sym3& s, double sign) {
  const double m[3][3] = {{s.xx, s.xy, s.xz}, {s.xy, s.yy, s.yz}, {s.xz, s.yz, s.zz}};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int64_t p = row0 + (int64_t)a * w + 3 * b;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      col_idx[p + c] = (int32_t)(3 * col_node + c);
      val[p + c] = sign * m[a][c];
    }
  }
}

// The three CSR rows of node i (what one thread of the numeric phase produces): walk the node's
// sorted pair segment, evaluate S_e in registers, sum duplicate pairs and the diagonal block in
// (neighbour, element) order.  `out_col`/`out_val` + `row0` address either the global arrays
// (row0 = 9 * block_start[i]) or a warp's staging window in shared memory.
//
// S_e is evaluated from the coordinates of (this node, destination node).  The destination is in the pair's
// key, so neither the element id nor n1 / n2 is read here, and the element's own orientation does not matter:
// myc_bar_block(p, q) and myc_bar_block(q, p) are BITWISE equal (p - q = -(q - p) exactly, and every later
// operation is a product of two sign-flipped factors or sign-symmetric) -- checked on the CPU in
// tests/test_kernel_logic_host.py.  The first four destinations and their coordinates are requested together
// (degree <= 4 covers a lattice network): two dependent load latencies per node instead of four per pair.
__device__ __forceinline__ void fill_node(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ evals,
                                          int32_t es, int32_t ee, int32_t w, int64_t self, uint64_t dmask,
                                          const double* __restrict__ coords, const int32_t* __restrict__ n1,
                                          const int32_t* __restrict__ n2, const BarConsts& bc, int32_t* out_col,
                                          double* out_val, int64_t row0) {
  (void)evals; (void)n1; (void)n2;
  constexpr int PRE = 4;
  int64_t dpre[PRE];
  double qx[PRE], qy[PRE], qz[PRE];
#pragma unroll
  for (int u = 0; u < PRE; ++u) dpre[u] = es + u < ee ? (int64_t)(keys[es + u] & dmask) : -1;
  const double px = coords[3 * self], py = coords[3 * self + 1], pz = coords[3 * self + 2];
#pragma unroll
  for (int u = 0; u < PRE; ++u) {
    const int64_t d = dpre[u] >= 0 ? dpre[u] : self;
    qx[u] = coords[3 * d]; qy[u] = coords[3 * d + 1]; qz[u] = coords[3 * d + 2];
  }
  Sym3 diag = {0, 0, 0, 0, 0, 0}, acc = {0, 0, 0, 0, 0, 0};
  int b = 0, diag_pos = -1;
  int64_t run_dst = -1;
  bool open = false;
  auto flush = [&]() {
    if (open) {
      write_block(out_col, out_val, row0, w, b, run_dst, acc, -1.0);
      ++b;
      open = false;
    }
  };
  auto take = [&](int64_t dst, double x, double y, double z) {
    if (dst == self) return;                           // n1 == n2: contributes exact zeros
    if (!open || dst != run_dst) {
      flush();
      if (diag_pos < 0 && dst > self) diag_pos = b++;
      run_dst = dst;
      acc = Sym3{0, 0, 0, 0, 0, 0};
      open = true;
    }
    double L;
    const Sym3 s = myc_bar_block(px, py, pz, x, y, z, bc, &L);
    acc.xx += s.xx; acc.xy += s.xy; acc.xz += s.xz; acc.yy += s.yy; acc.yz += s.yz; acc.zz += s.zz;
    diag.xx += s.xx; diag.xy += s.xy; diag.xz += s.xz; diag.yy += s.yy; diag.yz += s.yz; diag.zz += s.zz;
  };
#pragma unroll
  for (int u = 0; u < PRE; ++u)
    if (dpre[u] >= 0) take(dpre[u], qx[u], qy[u], qz[u]);
  for (int32_t k = es + PRE; k < ee; ++k) {
    const int64_t dst = (int64_t)(keys[k] & dmask);
    take(dst, coords[3 * dst], coords[3 * dst + 1], coords[3 * dst + 2]);
  }
  flush();
  if (diag_pos < 0) diag_pos = b;
  write_block(out_col, out_val, row0, w, diag_pos, self, diag, 1.0);
}

// Direct form (MYC_ASM_DIRECT_FILL=1, and the fallback for windows that exceed the staging capacity):
// every thread stores its node's rows straight to global memory (strided 8-byte stores).
__global__ void __launch_bounds__(AS_THREADS)
fill_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ evals,
            const int32_t* __restrict__ edge_start, const int32_t* __restrict__ block_start,
            int64_t n_local, int64_t nb, int dst_bits, const double* __restrict__ coords,
            const int32_t* __restrict__ n1, const int32_t* __restrict__ n2, double E, double A,
            double I, int32_t* __restrict__ col_idx, double* __restrict__ val) {
  const uint64_t dmask = ((uint64_t)1 << dst_bits) - 1;
  const BarConsts bc = myc_bar_consts(E, A, I);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_local;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t es = edge_start[i], ee = edge_start[i + 1];
    if (ee == es) continue;
    const int32_t bs = block_start[i];
    const int32_t w = 3 * (block_start[i + 1] - bs);     // entries per row
    fill_node(keys, evals, es, ee, w, nb + i, dmask, coords, n1, n2, bc, col_idx, val, 9 * (int64_t)bs);
  }
}

// [host-test-end assembly_kernels]

// Staged form (default).  The rows of 32 consecutive nodes occupy ONE contiguous window of col_idx / val
// (9 * block_start[first] .. 9 * block_start[last + 1]), so a warp builds the window in shared memory
// -- each lane its own node, same arithmetic and order as above -- and then copies it out with
// consecutive lanes on consecutive entries: full-line stores instead of 8-byte stores strided by the
// row length (ncu, 2048^2: the direct form wrote 1.94 GB to DRAM for 1.10 GB of CSR).
constexpr int FS_WARPS = 4;
constexpr int FS_THREADS = 32 * FS_WARPS;
constexpr int FS_CAP = 1440;      // entries per warp window: 32 nodes x 5 node blocks x 9 (any mesh of degree <= 4 fits)
constexpr size_t FS_SMEM = (size_t)FS_WARPS * FS_CAP * (sizeof(double) + sizeof(int32_t));

__global__ void __launch_bounds__(FS_THREADS)
fill_staged_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ evals,
                   const int32_t* __restrict__ edge_start, const int32_t* __restrict__ block_start,
                   int64_t n_local, int64_t nb, int dst_bits, const double* __restrict__ coords,
                   const int32_t* __restrict__ n1, const int32_t* __restrict__ n2, double E, double A,
                   double I, int32_t* __restrict__ col_idx, double* __restrict__ val) {
  extern __shared__ __align__(16) unsigned char fs_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const sv = (double*)fs_smem + (size_t)warp * FS_CAP;
  int32_t* const sc = (int32_t*)(fs_smem + (size_t)FS_WARPS * FS_CAP * sizeof(double)) + (size_t)warp * FS_CAP;
  const uint64_t dmask = ((uint64_t)1 << dst_bits) - 1;
  const BarConsts bc = myc_bar_consts(E, A, I);
  const int64_t n_tiles = (n_local + 31) / 32;
  for (int64_t tile = (int64_t)blockIdx.x * FS_WARPS + warp; tile < n_tiles; tile += (int64_t)gridDim.x * FS_WARPS) {
    const int64_t i_first = tile * 32;
    const int64_t i_end = i_first + 32 < n_local ? i_first + 32 : n_local;
    const int64_t win0 = 9 * (int64_t)block_start[i_first];
    const int64_t win_len = 9 * (int64_t)block_start[i_end] - win0;
    const bool staged = win_len <= FS_CAP;               // warp-uniform
    const int64_t i = i_first + lane;
    if (i < n_local) {
      const int32_t es = edge_start[i], ee = edge_start[i + 1];
      if (ee > es) {
        const int32_t bs = block_start[i];
        const int32_t w = 3 * (block_start[i + 1] - bs);
        const int64_t row0 = 9 * (int64_t)bs;
        if (staged) fill_node(keys, evals, es, ee, w, nb + i, dmask, coords, n1, n2, bc, sc, sv, row0 - win0);
        else fill_node(keys, evals, es, ee, w, nb + i, dmask, coords, n1, n2, bc, col_idx, val, row0);
      }
    }
    __syncwarp();
    if (staged) {
      const int len = (int)win_len;
      for (int q = lane; q < len; q += 32) {
        val[win0 + q] = sv[q];
        col_idx[win0 + q] = sc[q];
      }
      __syncwarp();                                      // the window is reused by the next tile
    }
  }
}

int bits_for(int64_t n) {  // bits needed to represent values in [0, n)
  int b = 1;
  while (((int64_t)1 << b) < n) ++b;
  return b;
}

}  // namespace

extern "C" int myc_assemble_symbolic(myc_ctx* ctx, const int32_t* d_n1, const int32_t* d_n2,
                                     const uint8_t* d_active, int64_t n_elem, int64_t n_nodes,
                                     int64_t node_begin, int64_t node_end, int32_t* d_out_row_ptr,
                                     int64_t* h_out_nnz, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  ctx->plan_valid = false;
  if (n_elem < 0 || n_nodes < 0 || node_begin < 0 || node_end < node_begin || node_end > n_nodes ||
      !d_out_row_ptr || !h_out_nnz || (n_elem > 0 && (!d_n1 || !d_n2)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "assemble_symbolic: bad argument");
  if (n_nodes >= ((int64_t)1 << 31) / 3 || n_elem >= ((int64_t)1 << 31))
    MYC_FAIL(ctx, MYC_ERR_CAPACITY, "assemble_symbolic: mesh exceeds int32 DOF / element indices");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_local = node_end - node_begin;
  int64_t* h_pin = (int64_t*)ctx->h_pinned;

  const int64_t cnt_items = (n_elem > n_local ? n_elem : n_local) + 1;
  MYC_TRY(myc_ensure(ctx, ctx->edge_cnt, (size_t)cnt_items * sizeof(int32_t)));
  MYC_TRY(myc_ensure(ctx, ctx->node_deg, (size_t)(n_local + 1) * sizeof(int32_t)));
  MYC_TRY(myc_ensure(ctx, ctx->node_bc, (size_t)(n_local + 1) * sizeof(int32_t)));
  MYC_TRY(myc_ensure(ctx, ctx->misc, 256));
  int32_t* cnt = (int32_t*)ctx->edge_cnt.p;            // per-element pair count (sort routes) / per-node fill (place route)
  int32_t* deg = (int32_t*)ctx->node_deg.p;
  int32_t* nbc = (int32_t*)ctx->node_bc.p;
  int* bad_flag = (int*)ctx->misc.p;
  int* too_long = bad_flag + 8;
  int64_t* d_total = (int64_t*)((char*)ctx->misc.p + 64);
  const int g_elem = grid_for(ctx, ceil_div64(n_elem, AS_THREADS), 8);
  const int g_node = grid_for(ctx, ceil_div64(n_local + 1, AS_THREADS), 8);
  const int dst_bits = bits_for(n_nodes);
  const int key_bits = dst_bits + bits_for(n_local);
  int64_t n_edges = 0;
  int sorted = 0;
  auto ensure_pairs = [&](int64_t n) -> int {
    for (int k = 0; k < 2; ++k) {
      MYC_TRY(myc_ensure(ctx, ctx->sort_keys[k], (size_t)(n + 1) * sizeof(uint64_t)));
      MYC_TRY(myc_ensure(ctx, ctx->sort_vals[k], (size_t)(n + 1) * sizeof(uint32_t)));
    }
    return MYC_OK;
  };
  // Routes, tried in this order (a later one only if an earlier one meets a hub node; MYC_ASM_SHORT_SORT=1 /
  // MYC_ASM_FULL_SORT=1 start further down):
  //   0  sort-free: degrees -> segments -> atomic placement -> per-node ordering by (destination, element)
  //   1  short sort: element-ordered emission, radix passes over the source bits, per-node ordering
  //   2  full sort: radix passes over (source, destination)
  // Every route runs to the end speculatively; the hub flag is read together with nnz, so the common case pays
  // no extra host round trip.  All three produce the same pair stream, hence the same CSR, bit for bit.
  for (int route = ctx->asm_full_sort ? 2 : ctx->asm_short_sort ? 1 : 0; route < 3; ++route) {
    MYC_CUDA(ctx, cudaMemsetAsync(bad_flag, 0, 128, st));
    MYC_CUDA(ctx, cudaMemsetAsync(deg, 0, (size_t)(n_local + 1) * sizeof(int32_t), st));
    sorted = 0;
    if (route == 0) {
      MYC_CUDA(ctx, cudaMemsetAsync(cnt, 0, (size_t)(n_local + 1) * sizeof(int32_t), st));
      if (n_elem > 0) {
        edge_degree_kernel<<<g_elem, AS_THREADS, 0, st>>>(d_n1, d_n2, d_active, n_elem, n_nodes, node_begin, node_end, deg,
                                                          bad_flag);
        MYC_LAUNCHED(ctx);
      }
      MYC_TRY(myc_exclusive_scan_i32(ctx, deg, deg, n_local, true, d_total, st));
    } else {
      if (n_elem > 0) {
        edge_count_kernel<<<g_elem, AS_THREADS, 0, st>>>(d_n1, d_n2, d_active, n_elem, n_nodes,
                                                         node_begin, node_end, cnt, bad_flag);
        MYC_LAUNCHED(ctx);
      }
      MYC_TRY(myc_exclusive_scan_i32(ctx, cnt, cnt, n_elem, false, d_total, st));
    }
    MYC_CUDA(ctx, cudaMemcpyAsync(h_pin, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaMemcpyAsync(h_pin + 1, bad_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    n_edges = h_pin[0];
    if (*(int*)(h_pin + 1)) MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "assemble_symbolic: element end node outside [0, n_nodes)");
    if (n_edges >= ((int64_t)1 << 31)) MYC_FAIL(ctx, MYC_ERR_CAPACITY, "assemble_symbolic: too many directed pairs for one device");
    MYC_TRY(ensure_pairs(n_edges));
    if (route == 0) {
      if (n_edges > 0) {
        edge_place_kernel<<<g_elem, AS_THREADS, 0, st>>>(d_n1, d_n2, d_active, n_elem, n_nodes, node_begin, node_end, dst_bits,
                                                         deg, cnt, (uint64_t*)ctx->sort_keys[0].p, (uint32_t*)ctx->sort_vals[0].p);
        MYC_LAUNCHED(ctx);
      }
    } else {
      if (n_edges > 0) {
        edge_emit_kernel<<<g_elem, AS_THREADS, 0, st>>>(d_n1, d_n2, d_active, n_elem, node_begin, node_end,
                                                        dst_bits, cnt, (uint64_t*)ctx->sort_keys[0].p,
                                                        (uint32_t*)ctx->sort_vals[0].p, deg);
        MYC_LAUNCHED(ctx);
        MYC_TRY(myc_radix_sort_pairs(ctx, n_edges, route == 1 ? dst_bits : 0, key_bits, &sorted, st));
      }
      MYC_TRY(myc_exclusive_scan_i32(ctx, deg, deg, n_local, true, nullptr, st));     // per-node segments of the sorted stream
    }
    if (route < 2 && n_edges > 0 && n_local > 0) {
      segment_order_kernel<<<g_node, AS_THREADS, 0, st>>>((uint64_t*)ctx->sort_keys[sorted].p,
                                                          (uint32_t*)ctx->sort_vals[sorted].p, deg, n_local, dst_bits,
                                                          too_long);
      MYC_LAUNCHED(ctx);
    }
    if (n_local > 0) {
      block_count_kernel<<<g_node, AS_THREADS, 0, st>>>((const uint64_t*)ctx->sort_keys[sorted].p, deg,
                                                        n_local, node_begin, dst_bits, nbc);
      MYC_LAUNCHED(ctx);
    }
    MYC_TRY(myc_exclusive_scan_i32(ctx, nbc, nbc, n_local, true, d_total, st));
    row_ptr_kernel<<<g_node, AS_THREADS, 0, st>>>(nbc, n_local, d_out_row_ptr);
    MYC_LAUNCHED(ctx);
    MYC_CUDA(ctx, cudaMemcpyAsync(h_pin, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaMemcpyAsync(h_pin + 2, too_long, sizeof(int), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    if (route == 2 || !*(int*)(h_pin + 2)) break;      // no hub: this route's pair stream stands
  }
  const int64_t nnz = 9 * h_pin[0];
  if (nnz >= ((int64_t)1 << 31)) MYC_FAIL(ctx, MYC_ERR_CAPACITY, "assemble_symbolic: nnz %lld exceeds int32 row_ptr", (long long)nnz);
  *h_out_nnz = nnz;
  ctx->plan_valid = true;
  ctx->plan_n_elem = n_elem;
  ctx->plan_n_nodes = n_nodes;
  ctx->plan_node_begin = node_begin;
  ctx->plan_node_end = node_end;
  ctx->plan_n_edges = n_edges;
  ctx->plan_nnz = nnz;
  ctx->plan_sorted_buf = sorted;
  ctx->plan_dst_bits = dst_bits;
  return MYC_OK;
}

extern "C" int myc_assemble_numeric(myc_ctx* ctx, const double* d_coords, const int32_t* d_n1,
                                    const int32_t* d_n2, double E, double A, double I,
                                    int64_t nnz_capacity, const int32_t* d_row_ptr,
                                    int32_t* d_out_col_idx, double* d_out_val, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (!ctx->plan_valid) MYC_FAIL(ctx, MYC_ERR_STATE, "assemble_numeric: call myc_assemble_symbolic first");
  if (nnz_capacity < ctx->plan_nnz) MYC_FAIL(ctx, MYC_ERR_CAPACITY, "assemble_numeric: capacity %lld < nnz %lld", (long long)nnz_capacity, (long long)ctx->plan_nnz);
  if (ctx->plan_nnz == 0) return MYC_OK;
  if (!d_coords || !d_n1 || !d_n2 || !d_out_col_idx || !d_out_val || !d_row_ptr)
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "assemble_numeric: null pointer");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n_local = ctx->plan_node_end - ctx->plan_node_begin;
  const uint64_t* keys = (const uint64_t*)ctx->sort_keys[ctx->plan_sorted_buf].p;
  const uint32_t* evals = (const uint32_t*)ctx->sort_vals[ctx->plan_sorted_buf].p;
  const int32_t* edge_start = (const int32_t*)ctx->node_deg.p;
  const int32_t* block_start = (const int32_t*)ctx->node_bc.p;
  if (ctx->asm_direct_fill) {
    const int g_node = grid_for(ctx, ceil_div64(n_local, AS_THREADS), 8);
    fill_kernel<<<g_node, AS_THREADS, 0, (cudaStream_t)stream>>>(
        keys, evals, edge_start, block_start, n_local, ctx->plan_node_begin, ctx->plan_dst_bits, d_coords, d_n1,
        d_n2, E, A, I, d_out_col_idx, d_out_val);
  } else {
    MYC_CUDA(ctx, cudaFuncSetAttribute(fill_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS_SMEM));
    const int g_tile = grid_for(ctx, ceil_div64(ceil_div64(n_local, 32), FS_WARPS), 3);
    fill_staged_kernel<<<g_tile, FS_THREADS, FS_SMEM, (cudaStream_t)stream>>>(
        keys, evals, edge_start, block_start, n_local, ctx->plan_node_begin, ctx->plan_dst_bits, d_coords, d_n1,
        d_n2, E, A, I, d_out_col_idx, d_out_val);
  }
  MYC_LAUNCHED(ctx);
  return MYC_OK;
}
