// The multigrid solver kernel's sweep over one level's symmetric 3x3 block view (pcg_amg.cu): the per-warp TMA
// ring of spmv_sym3.cuh, generalised in two directions.
//
//   * S stages instead of two: the bulk copy of tile j + S - 1 is issued while tile j is multiplied, so S - 1 tiles
//     per warp are in flight.  ncu on the two-stage ring (profiles/r2_amg_ncu.md): DRAM traffic = algorithmic bytes,
//     56 % of the stall samples on the global-load scoreboard, 2.6 TB/s -- the sweep is bound by the bytes it keeps
//     in flight, not by bandwidth.
//   * the value type of the staged blocks: double for the operator of the CG iteration (w = A u), FLOAT for the
//     sweeps inside the V-cycle (residual, smoothing).  The preconditioner only steers convergence -- the answer is
//     decided by the FP64 residual recurrence -- so its level operators are stored rounded to FP32: 28 B instead of
//     52 B per block, and twice the stages in the same shared memory.  Products and sums stay FP64.
//
// Tile = NODES consecutive nodes (AgTileT); lane l < 3 NODES owns row l of the tile in the epilogue.  Stage layout per
// warp: S x (CAPB x 6 values), then S x (CAPB column indices), then S x (16 block row pointers: the tile's own window of
// brp travels with the same bulk-copy transaction, so the consumer never waits on a global load for it).  After the multiply the three row partials of a
// block are parked in the block's own value slot (48 B or 24 B: three doubles fit either way).
#pragma once
#include <type_traits>

#include "spmv_sym3.cuh"

#ifndef AG_TILE_NODES
#define AG_TILE_NODES 10              // nodes per warp tile of the FP32 sweeps inside the V-cycle
#endif
#ifndef AG_TILE_NODES_F64
#define AG_TILE_NODES_F64 AG_TILE_NODES   // ... of the FP64 sweeps (w = A u; the V-cycle's sweeps under MYC_AMG_FP64=1)
#endif
template <int N>
struct AgTileT {
  static constexpr int NODES = N;                                  // nodes per warp tile (<= 10: one lane per row)
  static constexpr int ROWS = 3 * NODES;
  static constexpr int CAPB = (NODES * 5 + 3 + 3) / 4 * 4;         // blocks per stage: 5 per node (lattice maximum) + alignment slack
  static constexpr int NPW = 16;                                   // block-row-pointer window per stage: the tile's NODES + 1
                                                                   // entries from a 16-byte-aligned start (<= 3 + 11 ints)
  static_assert(N >= 4 && N <= 10, "one lane per row of a tile");
};
// the tile is chosen by the value type of the staged blocks
template <class VT> struct AgTileOf { using type = AgTileT<AG_TILE_NODES_F64>; };
template <> struct AgTileOf<float> { using type = AgTileT<AG_TILE_NODES>; };
// Stage counts.  Shared memory is not only the ring: what the ring leaves of the SM's 256 KB array is L1, and every
// gather / epilogue operand in flight holds an L1 line -- with 32 warps x (a tile's gathers + epilogue operands) in
// flight a 28 KB L1 throttles the sweeps.  Three FP32 stages instead of four bring the block under 196 KB, i.e. a
// 60 KB L1: 162.8 -> 145.7 ms per 2048^2 solve (profiles/r2_ab_l1_carveout.md); the fourth stage bought nothing.
#ifndef AG_STAGES_F64
#define AG_STAGES_F64 2
#endif
#ifndef AG_STAGES_F32
#define AG_STAGES_F32 3
#endif
constexpr int AG_MAX_STAGES = AG_STAGES_F64 > AG_STAGES_F32 ? AG_STAGES_F64 : AG_STAGES_F32;
template <class VT, int S>
__host__ __device__ constexpr size_t ag_ring_bytes_of() {
  using T = typename AgTileOf<VT>::type;
  return (size_t)S * (T::CAPB * (6 * sizeof(VT) + sizeof(int32_t)) + T::NPW * sizeof(int32_t));
}
__host__ __device__ constexpr size_t ag_ring_bytes_per_warp() {
  constexpr size_t f64 = ag_ring_bytes_of<double, AG_STAGES_F64>();
  constexpr size_t f32 = ag_ring_bytes_of<float, AG_STAGES_F32>();
  return f64 > f32 ? f64 : f32;
}
static_assert(ag_ring_bytes_per_warp() % 16 == 0, "bulk-copy destinations are 16-byte aligned");
__host__ __device__ constexpr size_t ag_smem_bytes(int warps) {
  return warps * ag_ring_bytes_per_warp() + (size_t)warps * AG_MAX_STAGES * sizeof(uint64_t) + 128;
}

struct AgPipe {
  unsigned char* ring;     // this warp's stage memory
  uint64_t* bars;          // [AG_MAX_STAGES]
  uint32_t phase_bits;     // (the L2 policies -- evict-first for the big streams, evict-last for the levels that fit L2 and
                           // are re-read every iteration -- are created where a copy is issued: no registers held)
};

__device__ __forceinline__ void ag_pipe_init(AgPipe& pp, unsigned char* smem_base, int warps_per_block, int warp, int lane) {
  pp.ring = smem_base + (size_t)warp * ag_ring_bytes_per_warp();
  pp.bars = reinterpret_cast<uint64_t*>(smem_base + (size_t)warps_per_block * ag_ring_bytes_per_warp()) + warp * AG_MAX_STAGES;
  pp.phase_bits = 0;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < AG_MAX_STAGES; ++s) tm_mbar_init(&pp.bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}

#ifndef AG_SWEEP_INLINE
#define AG_SWEEP_INLINE __forceinline__
#endif
// One sweep of warp gw over its tiles (gw, gw + n_warps, ...).  brp: block row pointer of the level (n_nodes + 1),
// bval: 6 VT per block, bcol: DOF column of the block, x: the gathered vector (coherent loads: other SMs / GPUs
// rewrite it between phases).  Rows are handed to epi.row / epi.row_warp like tm_sym3_sweep does.
template <class Epi, class VT, int S>
__device__ AG_SWEEP_INLINE void ag_sweep(AgPipe& pp, int64_t n_rows, const int32_t* __restrict__ brp,
                                         const VT* __restrict__ bval, const int32_t* __restrict__ bcol, const double* x,
                                         const Epi& epi, double (&acc)[Epi::NACC == 0 ? 1 : Epi::NACC], int64_t gw,
                                         int64_t n_warps, int lane, int32_t nb_total, bool keep_in_l2) {
  using Tile = typename AgTileOf<VT>::type;
  constexpr int NODES = Tile::NODES, ROWS = Tile::ROWS, CAPB = Tile::CAPB, D = S - 1;
  constexpr bool F32 = std::is_same<VT, float>::value;
  static_assert(S >= 2 && S <= AG_MAX_STAGES, "stage count");
  // 32-bit tile arithmetic (a rank holds < 2^31 rows): the kernel runs at the 64-register limit of a 1024-thread block
  const int n_nodes = (int)(n_rows / 3);
  const int n_tiles = (n_nodes + NODES - 1) / NODES;
  const int gw32 = (int)gw, nw32 = (int)n_warps;
  const int t_count = gw32 < n_tiles ? (n_tiles - gw32 + nw32 - 1) / nw32 : 0;
  if (t_count == 0) return;
  const int32_t nb4 = nb_total & ~3;
  VT* const s_val = reinterpret_cast<VT*>(pp.ring);
  int32_t* const s_col = reinterpret_cast<int32_t*>(pp.ring + (size_t)S * CAPB * 6 * sizeof(VT));
  int32_t* const s_np = s_col + (size_t)S * CAPB;
  constexpr int NPW = Tile::NPW;
  uint64_t* const bars = pp.bars;
  auto tile_of = [&](int j) -> int { return gw32 + j * nw32; };
  auto node_ptr = [&](int nd) -> int32_t { return brp[nd < n_nodes ? nd : n_nodes]; };
  // every tile is one bulk-copy transaction: its window of brp (always; the arrays carry >= 256 B of slack behind
  // their last entry, common.cuh myc_ensure) plus, if the tile fits a stage, its values and column indices
  auto issue = [&](int s, int t, int32_t lo, int32_t hi) {
    int32_t a0, a1;
    const bool staged = tm_tile_staged(lo, hi, nb4, CAPB, a0, a1);
    const uint32_t n = staged ? (uint32_t)(a1 - a0) : 0u;
    uint64_t policy;
    if (keep_in_l2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    tm_mbar_expect_tx(&bars[s], n * (uint32_t)(6 * sizeof(VT) + 4) + (uint32_t)(NPW * sizeof(int32_t)));
    tm_bulk_load(s_np + (size_t)s * NPW, brp + ((t * NODES) & ~3), (uint32_t)(NPW * sizeof(int32_t)), &bars[s], policy);
    if (staged) {
      tm_bulk_load(s_val + (size_t)s * CAPB * 6, bval + (size_t)a0 * 6, n * (uint32_t)(6 * sizeof(VT)), &bars[s], policy);
      tm_bulk_load(s_col + (size_t)s * CAPB, bcol + a0, n * 4u, &bars[s], policy);
    }
  };

  // ---- prologue: tiles 0 .. D-1 are requested, the block range of tile D is fetched for the first iteration
  int32_t pre_lo = 0, pre_hi = 0;          // lane 0: block range of tile j + D, fetched one iteration ahead
  if (lane == 0) {
    int32_t lo[D], hi[D];
#pragma unroll
    for (int d = 0; d < D; ++d)
      if (d < t_count) { lo[d] = node_ptr(tile_of(d) * NODES); hi[d] = node_ptr(tile_of(d) * NODES + NODES); }
    if (D < t_count) { pre_lo = node_ptr(tile_of(D) * NODES); pre_hi = node_ptr(tile_of(D) * NODES + NODES); }
#pragma unroll
    for (int d = 0; d < D; ++d)
      if (d < t_count) issue(d, tile_of(d), lo[d], hi[d]);
  }

  for (int j = 0; j < t_count; ++j) {
    const int s = j % S;
    const int r0 = tile_of(j) * ROWS;
    if (lane == 0 && j + D < t_count) {
      issue((j + D) % S, tile_of(j + D), pre_lo, pre_hi);      // the stage tile j - 1 has just left
      if (j + D + 1 < t_count) {
        pre_lo = node_ptr(tile_of(j + D + 1) * NODES);
        pre_hi = node_ptr(tile_of(j + D + 1) * NODES + NODES);
      }
    }
    // the tile's transaction has landed: its block row pointers are in the stage (lane l < NODES: node NODES*t + l)
    tm_mbar_wait(&bars[s], (pp.phase_bits >> s) & 1u);
    pp.phase_bits ^= (1u << s);
    int32_t cur_lo, cur_hi;
    {
      const int t0 = tile_of(j) * NODES, w0 = t0 & ~3;
      const int nd = t0 + (lane < NODES ? lane : NODES - 1);
      const int32_t* np = s_np + (size_t)s * NPW - w0;
      cur_lo = np[nd < n_nodes ? nd : n_nodes];
      cur_hi = np[nd + 1 < n_nodes ? nd + 1 : n_nodes];
    }
    const int32_t lo = __shfl_sync(0xffffffffu, cur_lo, 0);
    const int32_t hi = __shfl_sync(0xffffffffu, cur_hi, NODES - 1);
    const int q = lane / 3, comp = lane - 3 * q;
    const int32_t my_lo = __shfl_sync(0xffffffffu, cur_lo, q < NODES ? q : NODES - 1);
    const int32_t my_hi = __shfl_sync(0xffffffffu, cur_hi, q < NODES ? q : NODES - 1);
    const bool row_ok = lane < ROWS && (r0 + lane) < (int)n_rows;
    typename Epi::Pre pre{};
    if (row_ok) pre = epi.load(r0 + lane);
    double sum = 0.0;
    if (hi > lo) {
      int32_t a0, a1;
      tm_tile_staged(lo, hi, nb4, CAPB, a0, a1);
      if (hi - a0 <= CAPB) {
        VT* sv = s_val + (size_t)s * CAPB * 6;
        int32_t* sc = s_col + (size_t)s * CAPB;
        double* sp = reinterpret_cast<double*>(sv);          // parked partials: 3 doubles per block (F32) / 6-stride (F64)
        constexpr int PSTRIDE = F32 ? 3 : 6;
        const int first = lo - a0, last = hi - a0, staged = a1 > a0 ? a1 - a0 : 0;
        if (staged < last) {                      // ragged end of the whole array (< 4 blocks)
          const int k = (staged > first ? staged : first) + lane;
          if (k < last) {
#pragma unroll
            for (int c = 0; c < 6; ++c) sv[6 * k + c] = bval[6 * (size_t)(a0 + k) + c];
            sc[k] = bcol[a0 + k];
          }
          __syncwarp();
        }
#pragma unroll
        for (int u = 0; u < (CAPB + 31) / 32; ++u) {          // one lane per block
          const int k = first + lane + 32 * u;
          if (k < last) {
            const int32_t c = sc[k];
            const double x0 = x[c], x1 = x[c + 1], x2 = x[c + 2];
            double xx, xy, xz, yy, yz, zz;
            if constexpr (F32) {
              const float2* f2 = reinterpret_cast<const float2*>(sv + 6 * k);
              const float2 v01 = f2[0], v23 = f2[1], v45 = f2[2];
              xx = v01.x; xy = v01.y; xz = v23.x; yy = v23.y; yz = v45.x; zz = v45.y;
            } else {
              const double2* b2 = reinterpret_cast<const double2*>(sv + 6 * k);
              const double2 v01 = b2[0], v23 = b2[1], v45 = b2[2];
              xx = v01.x; xy = v01.y; xz = v23.x; yy = v23.y; yz = v45.x; zz = v45.y;
            }
            const double p0 = fma(xz, x2, fma(xy, x1, xx * x0));
            const double p1 = fma(yz, x2, fma(yy, x1, xy * x0));
            const double p2 = fma(zz, x2, fma(yz, x1, xz * x0));
            sp[PSTRIDE * k] = p0; sp[PSTRIDE * k + 1] = p1; sp[PSTRIDE * k + 2] = p2;
          }
        }
        __syncwarp();
        if (row_ok) {
          // block order; the first five partials (a lattice node has at most five blocks) are requested together
          const int b0 = my_lo - a0, e = my_hi - a0;
          const double* q = sp + PSTRIDE * b0 + comp;
          const double s0 = b0 < e ? q[0] : 0.0, s1 = b0 + 1 < e ? q[PSTRIDE] : 0.0, s2 = b0 + 2 < e ? q[2 * PSTRIDE] : 0.0;
          const double s3 = b0 + 3 < e ? q[3 * PSTRIDE] : 0.0, s4 = b0 + 4 < e ? q[4 * PSTRIDE] : 0.0;
          sum = (((s0 + s1) + s2) + s3) + s4;
          for (int b = b0 + 5; b < e; ++b) sum += sp[PSTRIDE * b + comp];
        }
        tm_fence_proxy_async();
        __syncwarp();
      } else if (row_ok) {
        // oversize tile (a node with very many neighbours): straight from global memory
        for (int32_t b = my_lo; b < my_hi; ++b) {
          const VT* m = bval + 6 * (size_t)b;
          const int32_t c = bcol[b];
          const double x0 = tm_ld_cg(x + c), x1 = tm_ld_cg(x + c + 1), x2 = tm_ld_cg(x + c + 2);
          const double m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5];
          const double p = comp == 0 ? fma(m2, x2, fma(m1, x1, m0 * x0))
                         : comp == 1 ? fma(m4, x2, fma(m3, x1, m1 * x0))
                                     : fma(m5, x2, fma(m4, x1, m2 * x0));
          sum += p;
        }
      }
    }
    if constexpr (tm_epi_warp_uniform<Epi>::value) epi.row_warp(r0 + lane, row_ok, sum, pre, acc, lane);
    else if (row_ok) epi.row(r0 + lane, sum, pre, acc);
  }
}
