// Symmetric 3x3 node-block operator for the PCG sweep (sm_100a), derived from the CSR K.
//
// Every stiffness matrix of this problem is a sum of element blocks [[S,-S],[-S,S]] with S a
// SYMMETRIC 3x3 matrix (src/fea_solver.py:30-68), so in the assembled CSR (csrc/assemble.cu) the
// three rows of a node carry the same block columns and every 3x3 block is bitwise symmetric.
// The solver therefore does not have to stream 9 values + 9 column indices per block (108 B): six
// values and one column index (52 B) say the same.  `myc_sym3_convert_kernel` builds that view once
// per solve (one pass over the CSR, checking the symmetry it relies on); `tm_sym3_sweep` is the
// TMA-pipelined sweep over it, structured like tm_warp_sweep (spmv_tma.cuh):
//
//   tile   10 consecutive nodes = 30 rows; its blocks are one contiguous window of bval/bcol
//   lane 0 two bulk copies (values, columns) into a 2-stage ring, mbarrier completion
//   lanes  stride the window one BLOCK each: 3 x LDS.128 + 1 LDS.32, three contiguous x gathers,
//          nine FMAs; the three row partials are parked in the block's own slot
//   lanes  0..29 then add their row's partials in block order (fixed order -> reproducible)
//
// Per 9 matrix entries: 52 B streamed instead of 108 B, 6 shared-memory accesses instead of 16
// (node-block CSR scheme) or 36 (generic scheme).  The CSR stays the product's deliverable and the
// operand of myc_spmv; this is the solver's private copy.
#pragma once
#include "common.cuh"
#include "spmv.cuh"
#include "spmv_tma.cuh"

struct TmCfgSym {
  static constexpr int NODES = 10;          // nodes per warp tile
  static constexpr int ROWS = 30;           // one lane per row in the sum phase
#ifndef TM_SYM_CAPB
#define TM_SYM_CAPB 56
#endif
  static constexpr int CAPB = TM_SYM_CAPB;  // blocks a stage window may hold (10 nodes x 5 + alignment slack)
};
__host__ __device__ constexpr size_t tm_sym_smem_per_warp() {
  return (size_t)TM_STAGES * TmCfgSym::CAPB * (6 * sizeof(double) + sizeof(int32_t));
}
__host__ __device__ constexpr size_t tm_sym_smem_bytes(int warps) {
  return warps * tm_sym_smem_per_warp() + warps * TM_STAGES * sizeof(uint64_t) + 128;
}

// [host-test-begin sym3_convert]  (tests/test_kernel_logic_host.py compiles this text with g++)
// One thread per owned node: pack the node's blocks.  rp/ci/v: node-block-structured CSR (local rows).
// bval: 6 doubles per block (xx xy xz yy yz zz), bcol: DOF column of the block's first entry.
// *bad is raised if a block is not bitwise symmetric (then the caller keeps using the CSR sweep).
static __global__ void __launch_bounds__(256)
myc_sym3_convert_kernel(int64_t n_nodes, const int32_t* __restrict__ rp, const int32_t* __restrict__ ci,
                        const double* __restrict__ v, double* __restrict__ bval, int32_t* __restrict__ bcol,
                        int* __restrict__ bad) {
  for (int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nd < n_nodes; nd += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s0 = rp[3 * nd];
    const int32_t w = rp[3 * nd + 1] - s0;
    const int32_t b0 = s0 / 9;
    for (int32_t k = 0; k < w; k += 3) {
      const double* r0 = v + s0 + k;
      const double* r1 = r0 + w;
      const double* r2 = r1 + w;
      const double xx = r0[0], xy = r0[1], xz = r0[2], yy = r1[1], yz = r1[2], zz = r2[2];
      if (r1[0] != xy || r2[0] != xz || r2[1] != yz) *bad = 1;
      double* o = bval + 6 * (size_t)(b0 + k / 3);
      o[0] = xx; o[1] = xy; o[2] = xz; o[3] = yy; o[4] = yz; o[5] = zz;
      bcol[b0 + k / 3] = ci[s0 + k];
    }
  }
}

// [host-test-end sym3_convert]

// Per-warp pipeline state for the sym3 ring (same shape as TmPipe).
struct TmSymPipe {
  double* s_val;      // [stage][CAPB*6]
  int32_t* s_col;     // [stage][CAPB]
  uint64_t* bars;
  uint64_t l2_stream;
  uint32_t phase_bits;
  bool head_in_flight;
};

__device__ __forceinline__ void tm_sym_pipe_init(TmSymPipe& pp, unsigned char* smem_base, int warps_per_block, int warp,
                                                 int lane) {
  constexpr int CAPB = TmCfgSym::CAPB;
  pp.s_val = reinterpret_cast<double*>(smem_base) + (size_t)warp * TM_STAGES * CAPB * 6;
  pp.s_col = reinterpret_cast<int32_t*>(smem_base + (size_t)warps_per_block * TM_STAGES * CAPB * 6 * sizeof(double)) +
             (size_t)warp * TM_STAGES * CAPB;
  pp.bars = reinterpret_cast<uint64_t*>(smem_base + (size_t)warps_per_block * tm_sym_smem_per_warp()) + warp * TM_STAGES;
  pp.phase_bits = 0;
  pp.head_in_flight = false;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pp.l2_stream));
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < TM_STAGES; ++s) tm_mbar_init(&pp.bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}

// Multi-GPU gate: entries of x outside [own_lo, own_hi) are stored by peer GPUs (halo).  Instead of
// stalling every block at the barrier until the neighbours' halo flags arrive, only a warp whose
// tile actually gathers remote entries waits for them -- and the sweep visits its first round of
// tiles (which holds the low-boundary tiles) last, so the NVLink latency hides behind interior work.
// Remote entries are read with ld.global.cg: an L1 line that straddles the own/halo boundary may have
// been fetched (for an own entry) before the neighbour's store landed.
struct TmHaloGate {
  int64_t own_lo, own_hi;
  const unsigned* flags;       // this rank's PeerSync.flag_halo[], written by the peers
  unsigned epoch;
  int world;
  unsigned recv_mask;          // bit q: this rank gathers rows of peer q
};
__device__ __forceinline__ double tm_ld_cg(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned tm_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// block offset of local node `nd` (clamped): the scalar CSR row pointer of its first row / 9, or -- DIRECT,
// the multigrid levels of pcg_amg.cu -- an explicit block row pointer
template <bool DIRECT = false>
__device__ __forceinline__ int32_t tm_sym_node_ptr(const int32_t* __restrict__ rp, int64_t nd, int64_t n_nodes) {
  if constexpr (DIRECT) return rp[nd < n_nodes ? nd : n_nodes];
  else return rp[3 * (nd < n_nodes ? nd : n_nodes)] / 9;
}

// Epilogues that need a whole node's rows at once (3x3 block smoothers) declare
// `static constexpr bool WARP_UNIFORM = true` and provide row_warp(row, row_ok, sum, pre, acc, lane), which
// ALL 32 lanes call (lane = 3 * node_in_tile + component for lane < 30), so it may shuffle.
template <class E, class = void>
struct tm_epi_warp_uniform { static constexpr bool value = false; };
template <class E>
struct tm_epi_warp_uniform<E, decltype((void)E::WARP_UNIFORM)> { static constexpr bool value = E::WARP_UNIFORM; };

// One sweep of warp gw over its tiles.  x is gathered coherently (persistent solver kernel).
template <class Epi, bool PREFETCH_NEXT, bool GATED, bool DIRECT = false>
__device__ __forceinline__ void tm_sym3_sweep(TmSymPipe& pp, int64_t n_rows, const int32_t* __restrict__ rp,
                                              const double* __restrict__ bval, const int32_t* __restrict__ bcol,
                                              const double* x, const Epi& epi,
                                              double (&acc)[Epi::NACC == 0 ? 1 : Epi::NACC], int64_t gw,
                                              int64_t n_warps, int lane, int32_t nb_total, const TmHaloGate& gate) {
  constexpr int NODES = TmCfgSym::NODES, ROWS = TmCfgSym::ROWS, CAPB = TmCfgSym::CAPB;
  const int64_t n_nodes = n_rows / 3;
  const int64_t n_tiles = (n_nodes + NODES - 1) / NODES;
  const int64_t t_count = gw < n_tiles ? (n_tiles - gw + n_warps - 1) / n_warps : 0;
  const int32_t nb4 = nb_total & ~3;
  double* const s_val = pp.s_val;
  int32_t* const s_col = pp.s_col;
  uint64_t* const bars = pp.bars;
  // visiting order: round (j + rot) % t_count; with the gate the first round comes last
  const int64_t rot = (GATED && t_count > 1) ? 1 : 0;
  auto tile_of = [&](int64_t j) -> int64_t { return gw + ((j + rot) % t_count) * n_warps; };
  bool halo_ready = false;

  // lane l < NODES holds the block range [lo_l, hi_l) of node NODES*t + l
  auto load_np = [&](int64_t t, int32_t& lo_l, int32_t& hi_l) {
    const int64_t nd = t * NODES + (lane < NODES ? lane : NODES - 1);
    lo_l = tm_sym_node_ptr<DIRECT>(rp, nd, n_nodes);
    hi_l = tm_sym_node_ptr<DIRECT>(rp, nd + 1, n_nodes);
  };
  auto issue = [&](int s, int32_t lo, int32_t hi) {
    int32_t a0, a1;
    if (tm_tile_staged(lo, hi, nb4, CAPB, a0, a1)) {
      const int32_t n = a1 - a0;
      tm_mbar_expect_tx(&bars[s], (uint32_t)n * 52u);
      tm_bulk_load(s_val + (size_t)s * CAPB * 6, bval + (size_t)a0 * 6, (uint32_t)n * 48u, &bars[s], pp.l2_stream);
      tm_bulk_load(s_col + (size_t)s * CAPB, bcol + a0, (uint32_t)n * 4u, &bars[s], pp.l2_stream);
    }
  };

  int32_t cur_lo = 0, cur_hi = 0, nxt_lo = 0, nxt_hi = 0, head_lo = 0, head_hi = 0;
  if (t_count > 0) {
    load_np(tile_of(0), cur_lo, cur_hi);
    if (t_count > 1) load_np(tile_of(1), nxt_lo, nxt_hi);
    head_lo = __shfl_sync(0xffffffffu, cur_lo, 0);
    head_hi = __shfl_sync(0xffffffffu, cur_hi, NODES - 1);
    if (!pp.head_in_flight && lane == 0) issue(0, head_lo, head_hi);
  }

  for (int64_t j = 0; j < t_count; ++j) {
    const int s = (int)(j % TM_STAGES);
    const int64_t t = tile_of(j);
    const int64_t r0 = t * ROWS;
    if (j + 1 < t_count) {
      const int32_t lo1 = __shfl_sync(0xffffffffu, nxt_lo, 0), hi1 = __shfl_sync(0xffffffffu, nxt_hi, NODES - 1);
      if (lane == 0) issue((int)((j + 1) % TM_STAGES), lo1, hi1);
    }
    int32_t nn_lo = 0, nn_hi = 0;
    if (j + 2 < t_count) load_np(tile_of(j + 2), nn_lo, nn_hi);
    const int32_t lo = __shfl_sync(0xffffffffu, cur_lo, 0);
    const int32_t hi = __shfl_sync(0xffffffffu, cur_hi, NODES - 1);
    // row (lane) -> its node's block range
    const int q = lane / 3, comp = lane - 3 * q;
    const int32_t my_lo = __shfl_sync(0xffffffffu, cur_lo, q < NODES ? q : NODES - 1);
    const int32_t my_hi = __shfl_sync(0xffffffffu, cur_hi, q < NODES ? q : NODES - 1);
    const bool row_ok = lane < ROWS && (r0 + lane) < n_rows;
    typename Epi::Pre pre{};
    if (row_ok) pre = epi.load(r0 + lane);
    double sum = 0.0;
    if (hi > lo) {
      int32_t a0, a1;
      const bool staged_tile = tm_tile_staged(lo, hi, nb4, CAPB, a0, a1);
      if (hi - a0 <= CAPB) {
        double* sv = s_val + (size_t)s * CAPB * 6;
        int32_t* sc = s_col + (size_t)s * CAPB;
        if (staged_tile) {
          tm_mbar_wait(&bars[s], (pp.phase_bits >> s) & 1u);
          pp.phase_bits ^= (1u << s);
        }
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
        // one lane per block; CAPB <= 64 -> at most two rounds
#pragma unroll
        for (int u = 0; u < (CAPB + 31) / 32; ++u) {
          const int k = first + lane + 32 * u;
          const int32_t c = k < last ? sc[k] : 0;
          bool remote = false;
          if constexpr (GATED) {
            remote = k < last && ((int64_t)c < gate.own_lo || (int64_t)c >= gate.own_hi);
            if (!halo_ready && __any_sync(0xffffffffu, remote)) {
              // this tile gathers halo entries: make sure the neighbours' stores have landed
              if (lane < gate.world && ((gate.recv_mask >> lane) & 1u)) {
                unsigned spins = 0;
                while (tm_ld_acquire_sys(&gate.flags[lane]) < gate.epoch)
                  if (++spins > (1u << 28)) __trap();
              }
              __syncwarp();
              halo_ready = true;
            }
          }
          if (k < last) {
            double x0, x1, x2;
            if (GATED && remote) { x0 = tm_ld_cg(x + c); x1 = tm_ld_cg(x + c + 1); x2 = tm_ld_cg(x + c + 2); }
            else { x0 = x[c]; x1 = x[c + 1]; x2 = x[c + 2]; }
            double2* b2 = reinterpret_cast<double2*>(sv + 6 * k);
            const double2 v01 = b2[0], v23 = b2[1], v45 = b2[2];        // xx xy | xz yy | yz zz
            const double p0 = fma(v23.x, x2, fma(v01.y, x1, v01.x * x0));
            const double p1 = fma(v45.x, x2, fma(v23.y, x1, v01.y * x0));
            const double p2 = fma(v45.y, x2, fma(v45.x, x1, v23.x * x0));
            b2[0] = make_double2(p0, p1);
            sv[6 * k + 2] = p2;
          }
        }
        __syncwarp();
        if (row_ok) {
          const int e = my_hi - a0;
          for (int b = my_lo - a0; b < e; ++b) sum += sv[6 * b + comp];     // block order
        }
        tm_fence_proxy_async();
        __syncwarp();
      } else if (row_ok) {
        // oversize tile (a node with very many neighbours): straight from global memory
        if constexpr (GATED) {
          if (!halo_ready) {     // rare path: wait unconditionally
            if (lane < gate.world && ((gate.recv_mask >> lane) & 1u)) {
              unsigned spins = 0;
              while (tm_ld_acquire_sys(&gate.flags[lane]) < gate.epoch)
                if (++spins > (1u << 28)) __trap();
            }
            halo_ready = true;
          }
        }
        for (int32_t b = my_lo; b < my_hi; ++b) {
          const double* m = bval + 6 * (size_t)b;
          const int32_t c = bcol[b];
          const double x0 = tm_ld_cg(x + c), x1 = tm_ld_cg(x + c + 1), x2 = tm_ld_cg(x + c + 2);
          const double p = comp == 0 ? fma(m[2], x2, fma(m[1], x1, m[0] * x0))
                         : comp == 1 ? fma(m[4], x2, fma(m[3], x1, m[1] * x0))
                                     : fma(m[5], x2, fma(m[4], x1, m[2] * x0));
          sum += p;
        }
      }
    }
    if constexpr (tm_epi_warp_uniform<Epi>::value) epi.row_warp(r0 + lane, row_ok, sum, pre, acc, lane);
    else if (row_ok) epi.row(r0 + lane, sum, pre, acc);
    cur_lo = nxt_lo; cur_hi = nxt_hi;
    nxt_lo = nn_lo; nxt_hi = nn_hi;
  }
  if constexpr (PREFETCH_NEXT) {
    if (t_count > 0) {
      if (lane == 0) issue(0, head_lo, head_hi);
      pp.head_in_flight = true;
    }
  } else {
    pp.head_in_flight = false;
  }
}

template <bool GATED>
__device__ __forceinline__ void tm_sym_pipe_drain(TmSymPipe& pp, int64_t n_rows, const int32_t* rp, int64_t gw,
                                                  int64_t n_warps, int32_t nb_total) {
  if (!pp.head_in_flight) return;
  const int64_t n_nodes = n_rows / 3;
  const int64_t n_tiles = (n_nodes + TmCfgSym::NODES - 1) / TmCfgSym::NODES;
  if (gw < n_tiles) {
    const int64_t t_count = (n_tiles - gw + n_warps - 1) / n_warps;
    const int64_t t0 = gw + ((GATED && t_count > 1) ? 1 : 0) * n_warps;     // first tile the sweep visits
    const int32_t lo = tm_sym_node_ptr(rp, t0 * TmCfgSym::NODES, n_nodes);
    const int32_t hi = tm_sym_node_ptr(rp, t0 * TmCfgSym::NODES + TmCfgSym::NODES, n_nodes);
    int32_t a0, a1;
    if (tm_tile_staged(lo, hi, nb_total & ~3, TmCfgSym::CAPB, a0, a1)) tm_mbar_wait(&pp.bars[0], pp.phase_bits & 1u);
  }
  pp.head_in_flight = false;
}
