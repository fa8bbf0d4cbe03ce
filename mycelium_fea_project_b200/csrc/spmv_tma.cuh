// TMA-pipelined CSR SpMV (sm_100a).  The plain CSR-stream kernel (spmv.cuh) was measured at
// 3.6 TB/s on the 2048^2 operator with 73 % of warp stalls on the global-load scoreboard
// (profiles/r1_spmv2048_ncu.md): not enough bytes in flight.  Here every WARP runs its own
// asynchronous pipeline over small row tiles dealt round-robin over the grid's warps:
//
//   lane 0:  cp.async.bulk (1-D TMA, SASS UBLKCP, L2 evict-first) of the tile's value and column
//            windows into a 2-stage shared-memory ring, completion on a per-stage mbarrier
//   warp  :  waits the stage, multiplies, __syncwarp, one lane per row sums, epilogue, refill.
//
// There is no __syncthreads in the loop: warps drift apart freely, every warp always has its next
// tile in flight while it consumes the current one (32 pipelines per SM), and the streamed matrix
// bypasses L1 and the register file.  Windows are widened to 16-byte boundaries as cp.async.bulk
// requires; the <=3 leading elements belong to the previous tile and are ignored, the ragged end
// of the whole array is fetched with plain loads.  Tiles whose window exceeds the stage (rows much
// denser than this problem's 3*(neighbours+1)) are computed straight from global memory.
//
// Two multiply/sum schemes share the pipeline (template parameter Cfg):
//   TmCfgGeneric  any CSR.  16-row tiles; lanes stride the nnz window and form val*x[col] in
//                 place (4 shared-memory accesses and one gather per nnz), one lane per row sums
//                 its products left to right -- bit-identical to the plain kernel.
//   TmCfgBlock3   CSR with the 3x3 node-block structure every stiffness matrix of this problem
//                 has (3 DOF per node: the three rows of a node have equal length and their
//                 columns come in triples 3c,3c+1,3c+2).  18-row tiles (6 nodes); one lane per
//                 (node, block): ONE column index and three contiguous x values serve nine
//                 products, the three per-row partials are parked in the slots the values came
//                 from and one lane per row adds its <= ~5 partials in block order.  1.8 instead
//                 of 4 shared-memory accesses per nnz, a third of the gather instructions.
//                 ncu on the generic scheme showed the LSU/shared data pipe at 74 % and issue
//                 slots at 65 % as the limiters once DRAM latency was hidden.
#pragma once
#include "common.cuh"
#include "spmv.cuh"

struct TmCfgGeneric {
  static constexpr int ROWS = 16;     // rows per warp tile (one lane per row in the sum phase)
  static constexpr int CAP = 256;     // elements a stage window may hold (16*15 + alignment slack)
  static constexpr bool B3 = false;
};
#ifndef TM_B3_ROWS
#define TM_B3_ROWS 18                 // 6 nodes per tile
#endif
#ifndef TM_B3_CAP
#define TM_B3_CAP 288                 // 18*15 + alignment slack, multiple of 32
#endif
struct TmCfgBlock3 {
  static constexpr int ROWS = TM_B3_ROWS;
  static constexpr int CAP = TM_B3_CAP;
  static constexpr bool B3 = true;
};
static_assert(TM_B3_ROWS % 3 == 0, "node-block tiles hold whole nodes");
constexpr int TM_WARPS = 8;
constexpr int TM_THREADS = 32 * TM_WARPS;
constexpr int TM_STAGES = 2;
#ifndef TM_BLOCKS_PER_SM_N
#define TM_BLOCKS_PER_SM_N 4
#endif
constexpr int TM_BLOCKS_PER_SM = TM_BLOCKS_PER_SM_N;   // x TM_WARPS tile pipelines per SM
// shared memory of one block of `warps` tile pipelines whose stages hold `cap` elements: the smaller
// it is, the more of the 256 KB SM array is left to L1 for the x gathers
__host__ __device__ constexpr size_t tm_smem_per_warp(int cap) { return (size_t)TM_STAGES * cap * (sizeof(double) + sizeof(int32_t)); }
__host__ __device__ constexpr size_t tm_smem_bytes(int warps, int cap) {
  return warps * tm_smem_per_warp(cap) + warps * TM_STAGES * sizeof(uint64_t) + 128;
}

__device__ __forceinline__ uint32_t tm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tm_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tm_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tm_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                             uint64_t l2_policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          tm_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(tm_smem_u32(bar)), "l"(l2_policy)
      : "memory");
}
__device__ __forceinline__ void tm_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    if (++spins > (1u << 22)) __trap();   // a lost TMA completion must fail loudly, not hang the GPU
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(tm_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tm_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Per-warp pipeline state.  It outlives one sweep over the matrix, so a persistent kernel (the
// fused PCG, pcg_fused.cu) can keep the matrix stream running across solver iterations: at the
// end of a sweep the first tile of the NEXT sweep is already requested (the matrix does not
// change between iterations, only the gathered vector does).
struct TmPipe {
  double* s_val;
  int32_t* s_col;
  uint64_t* bars;
  uint64_t l2_stream;
  uint32_t phase_bits;
  bool head_in_flight;     // tile 0 of the coming sweep has been issued into stage 0
};

__device__ __forceinline__ void tm_pipe_init(TmPipe& pp, unsigned char* smem_base, int warps_per_block, int warp,
                                             int lane, int cap) {
  pp.s_val = reinterpret_cast<double*>(smem_base) + (size_t)warp * TM_STAGES * cap;
  pp.s_col = reinterpret_cast<int32_t*>(smem_base + (size_t)warps_per_block * TM_STAGES * cap * sizeof(double)) +
             (size_t)warp * TM_STAGES * cap;
  pp.bars = reinterpret_cast<uint64_t*>(smem_base + (size_t)warps_per_block * tm_smem_per_warp(cap)) + warp * TM_STAGES;
  pp.phase_bits = 0;
  pp.head_in_flight = false;
  // the matrix is read once per sweep: evict-first, keep L2 for the gathered vector
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pp.l2_stream));
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < TM_STAGES; ++s) tm_mbar_init(&pp.bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}

// Does the stage hold the tile [lo, hi) (and was a bulk copy issued for it)?  Shared by the issuing
// and the consuming side so that every expect_tx is matched by exactly one wait.
__device__ __forceinline__ bool tm_tile_staged(int32_t lo, int32_t hi, int32_t nnz4, int cap, int32_t& a0, int32_t& a1) {
  a0 = lo & ~3;
  a1 = (hi + 3) & ~3;
  if (a1 > nnz4) a1 = nnz4;                      // never read past the arrays
  return hi > lo && hi - a0 <= cap && a1 > a0;
}

// One sweep of warp `gw` (of n_warps) over its tiles: y-rows are handed to epi.row(row, sum, pre, acc).
// COHERENT_X: the gathered vector is written by other SMs between sweeps of the same kernel, so
// it must not be read through the non-coherent (ld.global.nc) path.
// PREFETCH_NEXT: request tile 0 of the next sweep before returning.
template <class Cfg, class Epi, bool COHERENT_X, bool PREFETCH_NEXT>
__device__ __forceinline__ void tm_warp_sweep(TmPipe& pp, int64_t n_rows, const int32_t* __restrict__ rp,
                                              const int32_t* __restrict__ ci, const double* __restrict__ v,
                                              const double* x, const Epi& epi,
                                              double (&acc)[Epi::NACC == 0 ? 1 : Epi::NACC], int64_t gw,
                                              int64_t n_warps, int lane, int32_t nnz_total) {
  static_assert(TM_STAGES == 2, "the sweep is written for a 2-stage ring");
  static_assert(Cfg::ROWS <= 32 && Cfg::CAP % 32 == 0, "tile shape");
  constexpr int ROWS = Cfg::ROWS;
  // Tiles are dealt round-robin over all warps of the grid (tile = gw + j * n_warps): at any time
  // the whole chip works inside one moving window of ~n_warps*ROWS rows, so the x entries gathered
  // by neighbouring tiles are shared in L2/L1 instead of being spread over the whole vector.
  const int64_t n_tiles = (n_rows + ROWS - 1) / ROWS;
  const int64_t t_count = gw < n_tiles ? (n_tiles - gw + n_warps - 1) / n_warps : 0;
  const int32_t nnz4 = nnz_total & ~3;
  double* const s_val = pp.s_val;
  int32_t* const s_col = pp.s_col;
  uint64_t* const bars = pp.bars;
  auto ldx = [&](int32_t c) -> double {
    if constexpr (COHERENT_X) return x[c];
    else return __ldg(x + c);
  };

  // row pointers of tile t: lane l holds rp[ROWS*t + l] and rp[ROWS*t + l + 1] (clamped)
  auto load_rp = [&](int64_t t, int32_t& lo_l, int32_t& hi_l) {
    const int64_t r = t * ROWS + (lane < ROWS ? lane : ROWS - 1);
    lo_l = rp[r < n_rows ? r : n_rows];
    hi_l = rp[r + 1 < n_rows ? r + 1 : n_rows];
  };
  // issue the TMA loads of one tile [lo, hi) into stage s (lane 0 only)
  auto issue = [&](int s, int32_t lo, int32_t hi) {
    int32_t a0, a1;
    if (tm_tile_staged(lo, hi, nnz4, Cfg::CAP, a0, a1)) {
      const int32_t n = a1 - a0;
      tm_mbar_expect_tx(&bars[s], (uint32_t)n * 12u);
      tm_bulk_load(s_val + (size_t)s * Cfg::CAP, v + a0, (uint32_t)n * 8u, &bars[s], pp.l2_stream);
      tm_bulk_load(s_col + (size_t)s * Cfg::CAP, ci + a0, (uint32_t)n * 4u, &bars[s], pp.l2_stream);
    }
  };

  // Register pipeline of row pointers: tile j (cur), j+1 (nxt); tile j+2 is requested at the top
  // of iteration j, so no global-load latency sits on the critical path.
  int32_t cur_lo = 0, cur_hi = 0, nxt_lo = 0, nxt_hi = 0;
  int32_t head_lo = 0, head_hi = 0;          // tile 0 again, for PREFETCH_NEXT
  if (t_count > 0) {
    load_rp(gw, cur_lo, cur_hi);
    if (t_count > 1) load_rp(gw + n_warps, nxt_lo, nxt_hi);
    head_lo = __shfl_sync(0xffffffffu, cur_lo, 0);
    head_hi = __shfl_sync(0xffffffffu, cur_hi, ROWS - 1);
    if (!pp.head_in_flight && lane == 0) issue(0, head_lo, head_hi);
  }

  for (int64_t j = 0; j < t_count; ++j) {
    const int s = (int)(j % TM_STAGES);
    const int64_t t = gw + j * n_warps;
    const int64_t r0 = t * ROWS;
    // refill the stage tile j-1 has just released with tile j+1
    if (j + 1 < t_count) {
      const int32_t lo1 = __shfl_sync(0xffffffffu, nxt_lo, 0), hi1 = __shfl_sync(0xffffffffu, nxt_hi, ROWS - 1);
      if (lane == 0) issue((int)((j + 1) % TM_STAGES), lo1, hi1);
    }
    int32_t nn_lo = 0, nn_hi = 0;
    if (j + 2 < t_count) load_rp(t + 2 * n_warps, nn_lo, nn_hi);
    const int32_t lo = __shfl_sync(0xffffffffu, cur_lo, 0);
    const int32_t hi = __shfl_sync(0xffffffffu, cur_hi, ROWS - 1);
    const int32_t my_lo = cur_lo, my_hi = cur_hi;
    const bool row_ok = lane < ROWS && (r0 + lane) < n_rows;
    typename Epi::Pre pre{};
    if (row_ok) pre = epi.load(r0 + lane);       // epilogue operands: in flight during the whole tile
    double sum = 0.0;
    if (hi > lo) {
      int32_t a0, a1;
      const bool staged_tile = tm_tile_staged(lo, hi, nnz4, Cfg::CAP, a0, a1);
      if (hi - a0 <= Cfg::CAP) {
        double* sv = s_val + (size_t)s * Cfg::CAP;
        int32_t* sc = s_col + (size_t)s * Cfg::CAP;
        if (staged_tile) {
          tm_mbar_wait(&bars[s], (pp.phase_bits >> s) & 1u);
          pp.phase_bits ^= (1u << s);
        }
        const int first = lo - a0, last = hi - a0, staged = a1 > a0 ? a1 - a0 : 0;
        if (staged < last) {                      // ragged end of the whole array (< 4 elements)
          const int k = (staged > first ? staged : first) + lane;
          if (k < last) { sv[k] = v[a0 + k]; sc[k] = ci[a0 + k]; }
          __syncwarp();
        }
        if constexpr (Cfg::B3) {
          // ---- node-block scheme: one lane per (node, block)
          constexpr int NODES = ROWS / 3;
          int nlo[NODES], nbc[NODES];            // stage offset of each node's first row, blocks per row
          int n_units = 0;
#pragma unroll
          for (int q = 0; q < NODES; ++q) {
            const int32_t l = __shfl_sync(0xffffffffu, my_lo, 3 * q), h = __shfl_sync(0xffffffffu, my_hi, 3 * q);
            nlo[q] = l - a0;
            nbc[q] = (h - l) / 3;
            n_units += nbc[q];
          }
          for (int ub = 0; ub < n_units; ub += 32) {
            int unit = ub + lane, base = 0, w = 0;
            const bool ok = unit < n_units;
#pragma unroll
            for (int q = 0; q < NODES; ++q) {    // locate this lane's node: unit index -> (node, block)
              if (ok && w == 0) {
                if (unit < nbc[q]) { base = nlo[q] + 3 * unit; w = 3 * nbc[q]; }
                else unit -= nbc[q];
              }
            }
            if (ok) {
              const int32_t c = sc[base];
              const double x0 = ldx(c), x1 = ldx(c + 1), x2 = ldx(c + 2);
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                double* q = sv + base + a * w;
                q[0] = fma(q[2], x2, fma(q[1], x1, q[0] * x0));   // partial of (row a, this block)
              }
            }
          }
          __syncwarp();
          if (row_ok) {
            const int e = my_hi - a0;
            for (int q = my_lo - a0; q < e; q += 3) sum += sv[q];   // block order
          }
        } else {
          // ---- generic scheme: products in place, fully unrolled so that every x gather of the tile
          // is in flight at once
          constexpr int TM_MAXIT = Cfg::CAP / 32;
          int32_t c[TM_MAXIT];
          double xv[TM_MAXIT];
#pragma unroll
          for (int u = 0; u < TM_MAXIT; ++u) {
            const int k = first + lane + 32 * u;
            c[u] = k < last ? sc[k] : -1;
          }
#pragma unroll
          for (int u = 0; u < TM_MAXIT; ++u) xv[u] = c[u] >= 0 ? ldx(c[u]) : 0.0;
#pragma unroll
          for (int u = 0; u < TM_MAXIT; ++u) {
            const int k = first + lane + 32 * u;
            if (c[u] >= 0) sv[k] = sv[k] * xv[u];
          }
          __syncwarp();
          if (row_ok) {
            int q = my_lo - a0;
            const int e = my_hi - a0;
            for (; q + 3 < e; q += 4) {           // 4 independent loads, adds kept in row order
              const double p0 = sv[q], p1 = sv[q + 1], p2 = sv[q + 2], p3 = sv[q + 3];
              sum = (((sum + p0) + p1) + p2) + p3;
            }
            for (; q < e; ++q) sum += sv[q];
          }
        }
        // every lane orders its generic-proxy accesses to this stage before the async-proxy
        // refill that lane 0 issues after the warp barrier
        tm_fence_proxy_async();
        __syncwarp();
      } else if (row_ok) {
        // oversize tile: lane-per-row straight from global memory (left-to-right, deterministic;
        // product rounded before the add like the staged generic path)
        for (int32_t q = my_lo; q < my_hi; ++q) sum = __dadd_rn(sum, __dmul_rn(v[q], ldx(ci[q])));
      }
    }
    if (row_ok) epi.row(r0 + lane, sum, pre, acc);
    cur_lo = nxt_lo; cur_hi = nxt_hi;
    nxt_lo = nn_lo; nxt_hi = nn_hi;
  }
  // Tile j uses stage j % 2 and every sweep starts with tile 0 in stage 0.  Stage 0 is free here:
  // with an odd tile count the last tile (even index) has just released it, with an even count it
  // was released one tile earlier.
  if constexpr (PREFETCH_NEXT) {
    if (t_count > 0) {
      if (lane == 0) issue(0, head_lo, head_hi);
      pp.head_in_flight = true;
    }
  } else {
    pp.head_in_flight = false;
  }
}

// Wait for a prefetched head tile (persistent kernels call this before exiting so that no bulk copy
// is in flight when the block retires).
template <class Cfg>
__device__ __forceinline__ void tm_pipe_drain(TmPipe& pp, int64_t n_rows, const int32_t* rp, int64_t gw,
                                              int32_t nnz_total) {
  if (!pp.head_in_flight) return;
  const int64_t n_tiles = (n_rows + Cfg::ROWS - 1) / Cfg::ROWS;
  if (gw < n_tiles) {
    const int64_t r0 = gw * Cfg::ROWS;
    const int64_t re = r0 + Cfg::ROWS < n_rows ? r0 + Cfg::ROWS : n_rows;
    int32_t a0, a1;
    if (tm_tile_staged(rp[r0], rp[re], nnz_total & ~3, Cfg::CAP, a0, a1)) tm_mbar_wait(&pp.bars[0], pp.phase_bits & 1u);
  }
  pp.head_in_flight = false;
}

template <class Cfg, class Epi>
__global__ void __launch_bounds__(TM_THREADS, TM_BLOCKS_PER_SM)
myc_spmv_tma_kernel(int64_t n_rows, const int32_t* __restrict__ rp, const int32_t* __restrict__ ci,
                    const double* __restrict__ v, const double* __restrict__ x, Epi epi, double* partials,
                    unsigned* counter, double* out, const int* done) {
  extern __shared__ __align__(128) unsigned char tm_smem[];
  __shared__ double s_warp[TM_THREADS / 32];
  if (done && *done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TmPipe pp;
  tm_pipe_init(pp, tm_smem, TM_WARPS, warp, lane, Cfg::CAP);
  double acc[Epi::NACC == 0 ? 1 : Epi::NACC];
#pragma unroll
  for (int j = 0; j < (Epi::NACC == 0 ? 1 : Epi::NACC); ++j) acc[j] = 0.0;
  tm_warp_sweep<Cfg, Epi, false, false>(pp, n_rows, rp, ci, v, x, epi, acc, (int64_t)blockIdx.x * TM_WARPS + warp,
                                        (int64_t)gridDim.x * TM_WARPS, lane, rp[n_rows]);
  if constexpr (Epi::NACC > 0) {
#pragma unroll
    for (int j = 0; j < Epi::NACC; ++j) acc[j] = myc_block_reduce(acc[j], s_warp);
    myc_finalize_partials<Epi::NACC>(partials, acc, counter, out, s_warp);
  }
}

template <class Cfg, class Epi>
static inline int myc_launch_spmv_tma(myc_ctx* ctx, int64_t n_rows, const int32_t* rp, const int32_t* ci,
                                      const double* v, const double* x, const Epi& epi, double* partials,
                                      unsigned* counter, double* out, const int* done, cudaStream_t st) {
  // the opt-in is per device and costs ~1 us: set it on every launch (a process may drive several GPUs)
  MYC_CUDA(ctx, cudaFuncSetAttribute(myc_spmv_tma_kernel<Cfg, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tm_smem_bytes(TM_WARPS, Cfg::CAP)));
  MYC_CUDA(ctx, cudaFuncSetAttribute(myc_spmv_tma_kernel<Cfg, Epi>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     myc_carveout_percent(tm_smem_bytes(TM_WARPS, Cfg::CAP), sizeof(double) * (TM_THREADS / 32) + 64,
                                                          TM_BLOCKS_PER_SM)));
  const int64_t n_tiles = ceil_div64(n_rows, Cfg::ROWS);
  const int grid = grid_for(ctx, ceil_div64(n_tiles, TM_WARPS), TM_BLOCKS_PER_SM);
  myc_spmv_tma_kernel<Cfg, Epi><<<grid, TM_THREADS, tm_smem_bytes(TM_WARPS, Cfg::CAP), st>>>(n_rows, rp, ci, v, x, epi, partials, counter,
                                                                        out, done);
  MYC_LAUNCHED(ctx);
  return MYC_OK;
}

// Launch helper: TMA kernel when the arrays are 16-byte aligned (always true for torch / cudaMalloc
// buffers) -- node-block scheme if the caller declared the structure (myc_set_csr_hint) -- and the
// plain CSR-stream kernel otherwise.
template <class Epi>
static inline int myc_launch_spmv_epi(myc_ctx* ctx, int64_t n_rows, const int32_t* rp, const int32_t* ci,
                                      const double* v, const double* x, const Epi& epi, double* partials,
                                      unsigned* counter, double* out, const int* done, cudaStream_t st) {
  if (n_rows == 0) return MYC_OK;
  const bool aligned = (((uintptr_t)ci | (uintptr_t)v) & 15u) == 0;
  if (aligned && !ctx->force_plain_spmv) {
    if (ctx->csr_block3 && n_rows % 3 == 0 && !ctx->no_block3_spmv)
      return myc_launch_spmv_tma<TmCfgBlock3, Epi>(ctx, n_rows, rp, ci, v, x, epi, partials, counter, out, done, st);
    return myc_launch_spmv_tma<TmCfgGeneric, Epi>(ctx, n_rows, rp, ci, v, x, epi, partials, counter, out, done, st);
  }
  const int grid = grid_for(ctx, ceil_div64(n_rows, SP_ROWS), SP_BLOCKS_PER_SM);
  myc_spmv_kernel<Epi><<<grid, SP_THREADS, 0, st>>>(n_rows, rp, ci, v, x, epi, partials, counter, out, done);
  MYC_LAUNCHED(ctx);
  return MYC_OK;
}
