// K5, single-kernel form: the whole (block-)Jacobi-PCG solve is ONE persistent cooperative launch per GPU
// (one 1024-thread block per SM), on one GPU or on several GPUs that talk through NVLink peer
// memory -- no NCCL call, no host round trip and no kernel boundary inside the iteration loop.
//
// Per iteration (single-reduction recurrence of Chronopoulos & Gear):
//   A. w = K u + reg u   (TMA-pipelined sweep, spmv_tma.cuh) with gamma = r.u, delta = w.u and
//      r.r folded into the epilogue                                  -> barrier #1 (+ reduction)
//   B. beta = gamma/gamma_old, alpha = gamma/(delta - beta gamma/alpha_old)
//      p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;  u = M^-1 r
//      -- rows that a neighbouring GPU gathers are ALSO stored straight into that GPU's u buffer
//      (same global offset, P2P store over NVLink)                   -> barrier #2 (halo ready)
//
// Barriers are "arrive / leader / release": every block arrives on a local counter, the last
// arriver's warp 0 does the cross-GPU part, then releases the local blocks.
//   #1: the leader sums the per-block partials in a fixed order, stores the three local totals
//       into every rank's PeerSync slot, raises its reduction flag there, waits for all ranks'
//       flags and adds the slots in rank order -> every GPU obtains bit-identical global sums and
//       therefore takes identical decisions (convergence, breakdown, maxit).
//   #2: the leader raises its halo flag at the neighbours that read its rows and waits for the
//       neighbours it reads from.
// The all-rank barrier #1 sits between a sweep (which reads the halo) and the pushes of the next
// values, so the halo needs no double buffering; the reduction slots are double-buffered by
// parity because only neighbours are synchronised by #2.  Flags are monotone epochs that continue
// across solves.  Every spin is bounded and traps instead of hanging the GPU.
// The matrix stream never stops: each warp requests the first tile of the next sweep before it
// enters phase B.  u is read with ordinary loads (never ld.global.nc); the acquiring
// __threadfence() after each barrier invalidates the SM's L1.
//
// Measured motivation (profiles/r1_pcg_fused.md): on the L2-resident 512^2 benchmark operator the
// three-kernel PCG costs ~47 us per iteration, about half of it launch gaps, per-kernel prologues
// and last-block reductions; with NCCL (halo + 2 all-reduces) ~78 us at 2 GPUs.
#include "common.cuh"
#include "spmv.cuh"
#include "spmv_tma.cuh"
#include "spmv_sym3.cuh"
#include "amg.cuh"

#include <type_traits>

namespace {

constexpr int FU_WARPS = 32;
constexpr int FU_THREADS = 32 * FU_WARPS;

struct PeerSync {                       // lives behind the u vector in each rank's IPC-shared buffer
  double sums[2][MYC_MAX_WORLD][4];     // [parity][writer rank][gamma, delta, r.r, -]
  unsigned flag_red[MYC_MAX_WORLD];     // written by rank q: reductions q has published
  unsigned flag_halo[MYC_MAX_WORLD];    // written by rank q: phase-B passes q has completed
};

struct FusedArgs {
  int64_t n_rows, row_offset;
  const int32_t* rp;
  const int32_t* ci;
  const double* v;
  const double* bval;      // symmetric 3x3 node-block view of K (spmv_sym3.cuh), or null
  const int32_t* bcol;
  int32_t nb_total;
  const double* dinv;
  const double* binv;      // block-Jacobi inverses: (n_rows/3, 9) for PC 1, symmetric-packed R x R blocks for PC 2 / 3
                           // (myc_block_inverse_packed), null for point Jacobi
  double* x;
  double* r;
  double* w;
  double* p;
  double* s;
  double reg;
  long long maxit;
  double* partials;        // [gridDim][3]
  unsigned* bar;           // [0] arrive counter, [1] release epoch (both zeroed before the launch)
  double* gsum;            // [2][4] global sums published by the leader
  PcgScalars* sc;          // in: tol2 ; out: iters, rr_final, done, breakdown, final epochs
  int world, rank;
  unsigned epoch_red0, epoch_halo0;            // epochs reached by previous solves
  double* peer_u[MYC_MAX_WORLD];               // u buffers (global length); [rank] is the local one
  PeerSync* peer_sync[MYC_MAX_WORLD];
  int64_t give_lo[MYC_MAX_WORLD], give_hi[MYC_MAX_WORLD];   // DOF ranges of MY rows that peer q gathers
  unsigned char recv_any[MYC_MAX_WORLD];       // I gather rows of peer q
  unsigned recv_mask;                          // the same as a bit mask
  int halo_overlap;                            // gated sweep: halo waits move from barrier #2 into the sweep
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

constexpr unsigned FU_SPIN_LIMIT = 1u << 28;

// arrive / leader / release barrier.  `leader_work(lane)` runs on warp 0 of the last-arriving block.
// `sys_release`: this block stored into a peer GPU's memory since the last barrier, so its release
// fence must have system scope (an NVLink round trip); blocks that only wrote local memory use the
// cheaper gpu-scope fence.
template <class F>
__device__ __forceinline__ void fused_barrier(const FusedArgs& a, unsigned& epoch, int* s_leader, bool sys_release,
                                              F&& leader_work) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    if (sys_release) __threadfence_system(); else __threadfence();      // release this block's stores
    const unsigned old = atomicAdd(&a.bar[0], 1u);
    *s_leader = (old == epoch * gridDim.x - 1u);
  }
  __syncthreads();
  if (*s_leader && warp == 0) {
    __threadfence();
    __syncwarp();          // the other lanes' work is ordered after lane 0's observation of the arrivals
    leader_work(lane);
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      st_release_gpu(&a.bar[1], epoch);
    }
  }
  if (threadIdx.x == 0) {
    unsigned spins = 0;
    while (ld_acquire_gpu(&a.bar[1]) < epoch)
      if (++spins > FU_SPIN_LIMIT) __trap();
    __threadfence();                                                     // acquire + L1 invalidate
  }
  __syncthreads();
}

// Single-GPU barrier: every block spins on the arrive counter itself (no leader hop).
// The arrive is one release-atomic (MEMBAR.ALL.GPU + REDG) and the wait one acquire-load spin
// (LDG.STRONG.GPU + CCTL.IVALL).  The release covers the whole block's stores through the preceding
// bar.sync (cumulativity); the acquire-load's CCTL.IVALL invalidates the SM's L1 for every warp.
// -DMYC_HEAVY_BARRIER restores the first form, __threadfence() + atomicAdd ... spin + __threadfence(),
// whose fences each compile to MEMBAR.ALL.CTA + MEMBAR.SC.GPU + ERRBAR + CCTL.IVALL: measured
// 24.39 -> 23.57 us per iteration at 512^2 (profiles/r1_block_jacobi_groups.md).
__device__ __forceinline__ void local_barrier(const FusedArgs& a, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    const unsigned target = epoch * gridDim.x;
    unsigned spins = 0;
#ifndef MYC_HEAVY_BARRIER
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(&a.bar[0]), "r"(1u) : "memory");
    while (ld_acquire_gpu(&a.bar[0]) < target)
      if (++spins > FU_SPIN_LIMIT) __trap();
#else
    __threadfence();
    atomicAdd(&a.bar[0], 1u);
#ifdef MYC_BARRIER_BACKOFF
    while (ld_acquire_gpu(&a.bar[0]) < target) {
      __nanosleep(MYC_BARRIER_BACKOFF);
      if (++spins > FU_SPIN_LIMIT) __trap();
    }
#else
    while (ld_acquire_gpu(&a.bar[0]) < target)
      if (++spins > FU_SPIN_LIMIT) __trap();
#endif
    __threadfence();
#endif
  }
  __syncthreads();
}

struct EpiFused {   // w = K u + reg u ; acc = {r.u, w.u, r.r}
  static constexpr int NACC = 3;
  double* w;
  const double* u_own;   // u + row_offset
  const double* r;
  double reg;
  struct Pre { double ui, ri; };
  __device__ __forceinline__ Pre load(int64_t i) const { return Pre{u_own[i], r[i]}; }
  __device__ __forceinline__ void row(int64_t i, double sum, const Pre& pre, double (&acc)[3]) const {
    const double ui = pre.ui, ri = pre.ri;
    const double wi = sum + reg * ui;
    w[i] = wi;
    acc[0] += ri * ui;
    acc[1] += wi * ui;
    acc[2] += ri * ri;
  }
};

#ifdef MYC_FUSED_TIMING
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define FT_MARK(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long n_ = gtimer(); tacc[k] += n_ - tlast; tlast = n_; } } while (0)
#else
#define FT_MARK(k) do { } while (0)
#endif

// PC: 0 = point Jacobi, 1 = 3x3 node blocks, 2 / 3 = aligned blocks of 2 / 4 nodes (6 / 12 rows, single GPU)
// OP: 0 = generic CSR sweep, 1 = node-block CSR sweep, 2 = symmetric 3x3 block operator
template <int PC, int OP, bool DIST>
__global__ void __launch_bounds__(FU_THREADS, 1) pcg_fused_kernel(FusedArgs a) {
  constexpr bool BLOCK3 = PC == 1;
  constexpr int GR = PC == 2 ? 6 : 12;                 // rows per block of the node-group preconditioners
  constexpr int GLW = (32 / GR) * GR;                  // lanes of a warp that carry rows: 30 (GR 6), 24 (GR 12)
  static_assert(PC >= 0 && PC <= 3 && !(DIST && PC == 3), "12x12 blocks are single-GPU");
  using Cfg = std::conditional_t<OP == 1, TmCfgBlock3, TmCfgGeneric>;
  using Pipe = std::conditional_t<OP == 2, TmSymPipe, TmPipe>;
  extern __shared__ __align__(128) unsigned char fu_smem[];
  __shared__ double s_red[FU_WARPS][3];
  __shared__ double s_tot[3];
  __shared__ int s_leader;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n = a.n_rows;
  const int64_t gtid = (int64_t)blockIdx.x * FU_THREADS + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * FU_THREADS;
  const double tol2 = a.sc->tol2;
  unsigned epoch = 0;                                  // local barrier epoch
  unsigned ep_red = a.epoch_red0, ep_halo = a.epoch_halo0;
  double* const u = a.peer_u[a.rank];
  PeerSync* const my_sync = a.peer_sync[a.rank];

  Pipe pp;
  if constexpr (OP == 2) tm_sym_pipe_init(pp, fu_smem, FU_WARPS, warp, lane);
  else tm_pipe_init(pp, fu_smem, FU_WARPS, warp, lane, Cfg::CAP);
  const int32_t nnz_total = a.rp[n];
  // Global warp number: tiles (and the vector pass's row chunks) are dealt round-robin over it.  Block-major
  // numbering would give the last, partial round to the first blocks only (at 512^2: 17,474 tiles over 4,736
  // warps = 4 tiles per warp in blocks 0..101 and 3 in blocks 103..147); numbering the warps block-fastest
  // gives every SM the same share of the partial round: 157.1 -> 153.8 ms per 512^2 solve, same iteration
  // counts, whole GPU suite green (profiles/r2_ab_gpu1.md).
  const int64_t gw = (int64_t)warp * gridDim.x + blockIdx.x;
  const int64_t n_warps = (int64_t)gridDim.x * FU_WARPS;

  // store one entry of u locally and into every GPU that gathers it
  bool pushed = false;            // this thread stored into a peer since the last halo barrier
  auto put_u = [&](int64_t i, double val) {
    const int64_t g = a.row_offset + i;
    u[g] = val;
    if constexpr (DIST) {
#pragma unroll 1
      for (int q = 0; q < a.world; ++q)
        if (g >= a.give_lo[q] && g < a.give_hi[q]) { a.peer_u[q][g] = val; pushed = true; }
    }
  };
  // barrier #2: phase-B results (and the pushed halo) are complete everywhere they are needed
  auto halo_barrier = [&]() {
    if constexpr (!DIST) {
      local_barrier(a, epoch);
    } else {
      ++ep_halo;
      const unsigned e = ep_halo;
      const bool block_pushed = __syncthreads_or(pushed ? 1 : 0) != 0;
      pushed = false;
      fused_barrier(a, epoch, &s_leader, block_pushed, [&](int ln) {
        // one lane per peer: all NVLink round trips overlap
        if (ln < a.world && ln != a.rank) {
          __threadfence_system();
          if (a.give_hi[ln] > a.give_lo[ln]) st_relaxed_sys(&a.peer_sync[ln]->flag_halo[a.rank], e);
          // with the gated sweep (OP == 2) only the warps that gather halo entries wait for them
          if (!(OP == 2 && a.halo_overlap) && a.recv_any[ln]) {
            unsigned spins = 0;
            while (ld_acquire_sys(&my_sync->flag_halo[ln]) < e)
              if (++spins > FU_SPIN_LIMIT) __trap();
          }
        }
      });
    }
  };

  // Block-Jacobi passes use 30 lanes per warp = 10 whole nodes, rows dealt coalesced; the lane of
  // row (node, c) gets its siblings' residuals by shuffle and applies row c of the 3x3 inverse.
  const int64_t b3_stride = n_warps * 30;
  const int b3_sib = lane - lane % 3;                     // first lane of this lane's node
  auto block3_z = [&](int64_t i, double ri) -> double {   // all 32 lanes call; lanes >= 30 / i >= n idle
    const double r0 = __shfl_sync(0xffffffffu, ri, b3_sib < 30 ? b3_sib : 0);
    const double r1 = __shfl_sync(0xffffffffu, ri, b3_sib < 30 ? b3_sib + 1 : 0);
    const double r2 = __shfl_sync(0xffffffffu, ri, b3_sib < 30 ? b3_sib + 2 : 0);
    if (lane >= 30 || i >= n) return 0.0;
    const double* m = a.binv + 3 * i;                     // row c of node i/3: binv[9*(i/3) + 3*c .. +2]
    return m[0] * r0 + m[1] * r1 + m[2] * r2;
  };
  // Node-group passes: GLW lanes per warp = whole blocks of GR rows; the lane of row c of a block gets
  // its siblings' residuals by shuffle and applies row c of the symmetric-packed inverse.
  const int64_t g_stride = n_warps * GLW;
  const int g_c = lane % GR, g_first = lane - lane % GR;
  auto group_z = [&](int64_t i, double ri) -> double {    // all 32 lanes call; lanes >= GLW / i >= n idle
    const bool ok = lane < GLW && i < n;
    double z = 0.0;
    if constexpr (GR == 6 && MYC_B6_FULL) {
      // full rows: row i of its block is the six doubles at binv + 6 i (48-byte rows, 16-byte aligned)
      double2 m01 = make_double2(0.0, 0.0), m23 = m01, m45 = m01;
      if (ok) {
        const double2* m = reinterpret_cast<const double2*>(a.binv + 6 * i);
        m01 = m[0]; m23 = m[1]; m45 = m[2];
      }
      const int src = lane < GLW ? g_first : 0;
      const double r0 = __shfl_sync(0xffffffffu, ri, src), r1 = __shfl_sync(0xffffffffu, ri, src + 1);
      const double r2 = __shfl_sync(0xffffffffu, ri, src + 2), r3 = __shfl_sync(0xffffffffu, ri, src + 3);
      const double r4 = __shfl_sync(0xffffffffu, ri, src + 4), r5 = __shfl_sync(0xffffffffu, ri, src + 5);
      z = m01.x * r0;
      z += m01.y * r1;
      z += m23.x * r2;
      z += m23.y * r3;
      z += m45.x * r4;
      z += m45.y * r5;
    } else {
      const double* m = a.binv + (ok ? i / GR : 0) * myc_block_inverse_stride(GR);
#pragma unroll
      for (int j = 0; j < GR; ++j) {
        const double rj = __shfl_sync(0xffffffffu, ri, lane < GLW ? g_first + j : 0);
        if (ok) z += m[myc_sympack(GR, g_c, j)] * rj;
      }
    }
    return z;
  };
  // init: u = M^-1 r, p = s = 0
  if constexpr (PC >= 2) {
    for (int64_t base = gw * GLW; base < n; base += g_stride) {
      const int64_t i = base + lane;
      const bool ok = lane < GLW && i < n;
      const double ri = ok ? a.r[i] : 0.0;
      const double z = group_z(i, ri);
      if (ok) {
        put_u(i, z);
        a.p[i] = 0.0;
        a.s[i] = 0.0;
      }
    }
  } else if constexpr (BLOCK3) {
    for (int64_t base = gw * 30; base < n; base += b3_stride) {
      const int64_t i = base + lane;
      const bool ok = lane < 30 && i < n;
      const double ri = ok ? a.r[i] : 0.0;
      const double z = block3_z(i, ri);
      if (ok) {
        put_u(i, z);
        a.p[i] = 0.0;
        a.s[i] = 0.0;
      }
    }
  } else {
    for (int64_t i = gtid; i < n; i += gstride) {
      put_u(i, a.dinv[i] * a.r[i]);
      a.p[i] = 0.0;
      a.s[i] = 0.0;
    }
  }
  halo_barrier();

#ifdef MYC_FUSED_TIMING
  unsigned long long tacc[4] = {0, 0, 0, 0}, tlast = gtimer();
#endif
  double gamma_old = 1.0, alpha_old = 1.0, rr = 0.0;
  long long it = 0;
  int status = 0;   // 1 converged, 2 breakdown, 0 maxit
  EpiFused epi{a.w, u + a.row_offset, a.r, a.reg};
  for (;;) {
    // ---- phase A: w = A u, partial dots
    double acc[3] = {0.0, 0.0, 0.0};
    if constexpr (OP == 2) {
      if constexpr (DIST) {
        if (a.halo_overlap) {
          TmHaloGate gate;
          gate.own_lo = a.row_offset;
          gate.own_hi = a.row_offset + n;
          gate.flags = my_sync->flag_halo;
          gate.epoch = ep_halo;
          gate.world = a.world;
          gate.recv_mask = a.recv_mask;
          tm_sym3_sweep<EpiFused, true, true>(pp, n, a.rp, a.bval, a.bcol, u, epi, acc, gw, n_warps, lane, a.nb_total, gate);
        } else {
          tm_sym3_sweep<EpiFused, true, false>(pp, n, a.rp, a.bval, a.bcol, u, epi, acc, gw, n_warps, lane, a.nb_total, TmHaloGate{});
        }
      } else {
        tm_sym3_sweep<EpiFused, true, false>(pp, n, a.rp, a.bval, a.bcol, u, epi, acc, gw, n_warps, lane, a.nb_total, TmHaloGate{});
      }
    }
    else
      tm_warp_sweep<Cfg, EpiFused, true, true>(pp, n, a.rp, a.ci, a.v, u, epi, acc, gw, n_warps, lane, nnz_total);
    FT_MARK(0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double t = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (lane == 0) s_red[warp][j] = t;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = 0.0;
      for (int wq = 0; wq < FU_WARPS; ++wq) t += s_red[wq][threadIdx.x];
      a.partials[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
    // ---- barrier #1 with the reduction
    if constexpr (!DIST) {
      local_barrier(a, epoch);
      if (warp == 0) {              // every block sums the per-block partials in the same fixed order
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double t = 0.0;
          for (unsigned b = lane; b < gridDim.x; b += 32) t += __ldcg(&a.partials[(size_t)b * 3 + j]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) s_tot[j] = t;
        }
      }
    } else {
      // the leader (last-arriving block) reduces locally, exchanges with the peers, publishes
      ++ep_red;
      const unsigned er = ep_red;
      const int par = (int)(er & 1u);
      fused_barrier(a, epoch, &s_leader, false, [&](int ln) {
        double tot[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double t = 0.0;
          for (unsigned b = ln; b < gridDim.x; b += 32) t += __ldcg(&a.partials[(size_t)b * 3 + j]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          tot[j] = t;
        }
        // one lane per peer: publish the local totals everywhere, then collect everybody's
        if (ln < a.world) {
          double* slot = a.peer_sync[ln]->sums[par][a.rank];
          slot[0] = tot[0]; slot[1] = tot[1]; slot[2] = tot[2];
          __threadfence_system();
          st_relaxed_sys(&a.peer_sync[ln]->flag_red[a.rank], er);
          unsigned spins = 0;
          while (ld_acquire_sys(&my_sync->flag_red[ln]) < er)
            if (++spins > FU_SPIN_LIMIT) __trap();
        }
        __syncwarp();
        if (ln < 3) {
          double t = 0.0;
          for (int q = 0; q < a.world; ++q) t += ld_volatile_f64(&my_sync->sums[par][q][ln]);   // rank order
          a.gsum[par * 4 + ln] = t;
        }
      });
      if (threadIdx.x < 3) s_tot[threadIdx.x] = __ldcg(&a.gsum[par * 4 + threadIdx.x]);
    }
    __syncthreads();
    FT_MARK(1);
    const double gamma = s_tot[0], delta = s_tot[1];
    rr = s_tot[2];
    if (!isfinite(rr)) { status = 2; break; }            // NaN / inf in K, b or x0: breakdown, not convergence
    if (!(rr > tol2)) { status = 1; break; }             // converged (x, r are consistent)
    if (it >= a.maxit) { status = 0; break; }
    const double beta = (it == 0) ? 0.0 : gamma / gamma_old;
    const double denom = (it == 0) ? delta : delta - beta * gamma / alpha_old;
    if (!(denom > 0.0) || !isfinite(gamma)) { status = 2; break; }
    const double alpha = gamma / denom;
    // ---- phase B: all vector recurrences in one pass
    if constexpr (PC >= 2) {
      for (int64_t base = gw * GLW; base < n; base += g_stride) {
        const int64_t i = base + lane;
        const bool ok = lane < GLW && i < n;
        double ri = 0.0;
        if (ok) {
          const double pi = u[a.row_offset + i] + beta * a.p[i];
          const double si = a.w[i] + beta * a.s[i];
          a.p[i] = pi;
          a.s[i] = si;
          a.x[i] += alpha * pi;
          ri = a.dinv[i] != 0.0 ? a.r[i] - alpha * si : 0.0;
          a.r[i] = ri;
        }
        const double z = group_z(i, ri);
        if (ok) put_u(i, z);
      }
    } else if constexpr (BLOCK3) {
      for (int64_t base = gw * 30; base < n; base += b3_stride) {
        const int64_t i = base + lane;
        const bool ok = lane < 30 && i < n;
        double ri = 0.0;
        if (ok) {
          const double pi = u[a.row_offset + i] + beta * a.p[i];
          const double si = a.w[i] + beta * a.s[i];
          a.p[i] = pi;
          a.s[i] = si;
          a.x[i] += alpha * pi;
          ri = a.dinv[i] != 0.0 ? a.r[i] - alpha * si : 0.0;
          a.r[i] = ri;
        }
        const double z = block3_z(i, ri);
        if (ok) put_u(i, z);
      }
    } else {
      for (int64_t i = gtid; i < n; i += gstride) {
        const double d = a.dinv[i];
        const double pi = u[a.row_offset + i] + beta * a.p[i];
        const double si = a.w[i] + beta * a.s[i];
        a.p[i] = pi;
        a.s[i] = si;
        a.x[i] += alpha * pi;
        const double ri = d != 0.0 ? a.r[i] - alpha * si : 0.0;
        a.r[i] = ri;
        put_u(i, d * ri);
      }
    }
    gamma_old = gamma;
    alpha_old = alpha;
    ++it;
    FT_MARK(2);
    halo_barrier();
    FT_MARK(3);
  }
  // no bulk copy may be in flight when the block exits
  if constexpr (OP == 2) {
    if (DIST && a.halo_overlap) tm_sym_pipe_drain<true>(pp, n, a.rp, gw, n_warps, a.nb_total);
    else tm_sym_pipe_drain<false>(pp, n, a.rp, gw, n_warps, a.nb_total);
  }
  else tm_pipe_drain<Cfg>(pp, n, a.rp, gw, nnz_total);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.sc->iters = it;
    a.sc->rr_final = rr;
    a.sc->red[1] = rr;
    a.sc->done = (status == 1);
    a.sc->breakdown = (status == 2);
    a.sc->pAp = (double)ep_red;          // final epochs, carried into the next solve by the host
    a.sc->rz_old = (double)ep_halo;
#ifdef MYC_FUSED_TIMING
    for (int k = 0; k < 4; ++k) a.sc->out[k] = (double)tacc[k] / (double)(it > 0 ? it : 1);   // ns per iteration
#endif
  }
}

size_t peer_buffer_bytes(int64_t cap) { return ((size_t)cap * sizeof(double) + 255) / 256 * 256 + sizeof(PeerSync); }
PeerSync* peer_sync_of(void* base, int64_t cap) {
  return (PeerSync*)((char*)base + ((size_t)cap * sizeof(double) + 255) / 256 * 256);
}

}  // namespace

// ---- peer-memory setup (multi-GPU) -----------------------------------------------------------
extern "C" int myc_dist_peer_alloc(myc_ctx* ctx, int64_t n_cols_capacity, uint8_t* h_out_handle64) {
  if (!ctx || !h_out_handle64 || n_cols_capacity < 0) return MYC_ERR_BAD_ARG;
  if (ctx->world > MYC_MAX_WORLD) MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "peer path supports at most %d ranks", MYC_MAX_WORLD);
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  MYC_CUDA(ctx, cudaDeviceSynchronize());
  for (int q = 0; q < MYC_MAX_WORLD; ++q) {
    if (ctx->peer_base[q] && q != ctx->rank) cudaIpcCloseMemHandle(ctx->peer_base[q]);
    ctx->peer_base[q] = nullptr;
  }
  if (ctx->peer_own) MYC_CUDA(ctx, cudaFree(ctx->peer_own));
  ctx->peer_own = nullptr;
  ctx->peer_ok = false;
  const size_t bytes = peer_buffer_bytes(n_cols_capacity);
  MYC_CUDA(ctx, cudaMalloc(&ctx->peer_own, bytes));
  MYC_CUDA(ctx, cudaMemset(ctx->peer_own, 0, bytes));
  MYC_CUDA(ctx, cudaDeviceSynchronize());
  ctx->peer_cap = n_cols_capacity;
  cudaIpcMemHandle_t h;
  MYC_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->peer_own));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(h_out_handle64, &h, 64);
  return MYC_OK;
}

extern "C" int myc_dist_peer_open(myc_ctx* ctx, const uint8_t* h_handles) {
  if (!ctx || !h_handles) return MYC_ERR_BAD_ARG;
  if (!ctx->peer_own) MYC_FAIL(ctx, MYC_ERR_STATE, "peer_open: call myc_dist_peer_alloc first");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int q = 0; q < ctx->world; ++q) {
    if (q == ctx->rank) { ctx->peer_base[q] = ctx->peer_own; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handles + 64 * (size_t)q, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      MYC_FAIL(ctx, MYC_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) -> %s (no P2P path: the NCCL PCG is used instead)", q,
               cudaGetErrorString(e));
    }
    ctx->peer_base[q] = p;
  }
  ctx->peer_epoch_red = ctx->peer_epoch_halo = 0;
  ctx->peer_ok = true;
  return MYC_OK;
}

// Returns MYC_OK and fills *handled = 1 if the fused path ran; *handled = 0 means "not applicable
// here" (caller falls back to the multi-kernel PCG).  On entry r = b - A x0 is in ctx->vec[1] and
// sc->tol2 is set (pcg.cu does that for both paths).  *op_used: 0 generic CSR, 1 node-block CSR,
// 2 symmetric 3x3 block view.
namespace {
template <int PC, int OP, bool DIST>
const void* fused_fn() { return (const void*)pcg_fused_kernel<PC, OP, DIST>; }
constexpr int FU_VARIANTS = 21;
// index = dist*6 + op*2 + pc for pc 0 / 1;  12 + op*2 + (pc - 2) for the single-GPU node-group blocks;
// 18 + op for the 6x6 blocks on several GPUs (needs rank boundaries on even nodes: dist.py cuts there)
const void* fused_variant(int idx) {
  static const void* tab[FU_VARIANTS] = {
      fused_fn<0, 0, false>(), fused_fn<1, 0, false>(), fused_fn<0, 1, false>(), fused_fn<1, 1, false>(),
      fused_fn<0, 2, false>(), fused_fn<1, 2, false>(), fused_fn<0, 0, true>(),  fused_fn<1, 0, true>(),
      fused_fn<0, 1, true>(),  fused_fn<1, 1, true>(),  fused_fn<0, 2, true>(),  fused_fn<1, 2, true>(),
      fused_fn<2, 0, false>(), fused_fn<3, 0, false>(), fused_fn<2, 1, false>(), fused_fn<3, 1, false>(),
      fused_fn<2, 2, false>(), fused_fn<3, 2, false>(),
      fused_fn<2, 0, true>(),  fused_fn<2, 1, true>(),  fused_fn<2, 2, true>()};
  return tab[idx];
}
int fused_variant_op(int idx) { return idx >= 18 ? idx - 18 : (idx % 6) / 2; }
size_t fused_smem(int op) {
  return op == 2 ? tm_sym_smem_bytes(FU_WARPS) : tm_smem_bytes(FU_WARPS, op == 1 ? TmCfgBlock3::CAP : TmCfgGeneric::CAP);
}
}  // namespace

int myc_pcg_fused_try(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                      const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                      const double* d_dinv, const double* d_binv, int pc, double reg, int64_t maxit, double* d_x,
                      cudaStream_t st, int* handled, int* op_used) {
  *handled = 0;
  if (ctx->no_fused_pcg) return MYC_OK;
  const bool dist = ctx->world > 1;
  // node-group blocks: rows of a block must be rank-local; on several GPUs only the 6x6 blocks
  if (pc >= 2 && (row_offset % (pc == 2 ? 6 : 12) != 0 || ((uintptr_t)d_binv & 15u) != 0 || (dist && pc != 2)))
    return MYC_OK;
  if (dist && (!ctx->peer_ok || ctx->peer_cap < n_cols_global || ctx->world > MYC_MAX_WORLD)) return MYC_OK;
  if (!dist && n_rows == 0) return MYC_OK;
  if ((((uintptr_t)d_col_idx | (uintptr_t)d_val) & 15u) != 0) return MYC_OK;   // (same allocator on every rank)
  int& max_blocks_per_sm = ctx->fused_max_blocks_per_sm;      // per device (the attribute below is per device)
  if (max_blocks_per_sm < 0) {
    int mn = 1 << 30;
    for (int k = 0; k < FU_VARIANTS; ++k) {
      const size_t smem = fused_smem(fused_variant_op(k));
      int b = 0;
      MYC_CUDA(ctx, cudaFuncSetAttribute(fused_variant(k), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaFuncAttributes fa;
      MYC_CUDA(ctx, cudaFuncGetAttributes(&fa, fused_variant(k)));
      MYC_CUDA(ctx, cudaFuncSetAttribute(fused_variant(k), cudaFuncAttributePreferredSharedMemoryCarveout,
                                         myc_carveout_percent(smem, fa.sharedSizeBytes, 1)));     // (common.cuh: the rest is L1)
      MYC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fused_variant(k), FU_THREADS, smem));
      mn = b < mn ? b : mn;
    }
    max_blocks_per_sm = mn;
  }
  int coop = 0;
  MYC_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  if (max_blocks_per_sm < 1 || !coop) return MYC_OK;
  // work vectors: w, p, s (r is vec[1]); u is vec[0] on one GPU, the IPC-shared buffer otherwise
  if (!dist) MYC_TRY(myc_ensure(ctx, ctx->vec[0], (size_t)(n_cols_global + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->vec[2], (size_t)(n_rows + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->vec[3], (size_t)(n_rows + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->vec[5], (size_t)(n_rows + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->misc, 512));
  MYC_TRY(myc_ensure(ctx, ctx->partials, (size_t)ctx->sm_count * 16 * 4 * sizeof(double)));
  // ---- operator: node-block structured K -> symmetric 3x3 block view (52 B instead of 108 B per block)
  const bool b3 = ctx->csr_block3 && n_rows % 3 == 0 && !ctx->no_block3_spmv;
  int op = b3 ? 1 : 0;
  int32_t nb_total = 0;
  if (b3 && !ctx->no_sym3 && n_rows > 0) {
    int32_t h_nnz = 0;
    MYC_CUDA(ctx, cudaMemcpyAsync(&h_nnz, d_row_ptr + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    nb_total = h_nnz / 9;
    ctx->sym_owner = 0;                       // the block view is rebuilt for THIS matrix: a multigrid hierarchy
    if (ctx->amg) ctx->amg->valid = false;    // whose level 0 lived there is gone
    MYC_TRY(myc_ensure(ctx, ctx->sym_val, ((size_t)nb_total + 4) * 6 * sizeof(double)));
    MYC_TRY(myc_ensure(ctx, ctx->sym_col, ((size_t)nb_total + 4) * sizeof(int32_t)));
    int* bad = (int*)((char*)ctx->misc.p + 320);
    MYC_CUDA(ctx, cudaMemsetAsync(bad, 0, sizeof(int), st));
    myc_sym3_convert_kernel<<<grid_for(ctx, ceil_div64(n_rows / 3, 256), 8), 256, 0, st>>>(
        n_rows / 3, d_row_ptr, d_col_idx, d_val, (double*)ctx->sym_val.p, (int32_t*)ctx->sym_col.p, bad);
    MYC_LAUNCHED(ctx);
    int* h = (int*)ctx->h_pinned;
    MYC_CUDA(ctx, cudaMemcpyAsync(h, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    if (*h == 0) op = 2;       // every block bitwise symmetric: the solver streams the compact view
  }
  if (op_used) *op_used = op;
  const int64_t n_tiles = op == 2 ? ceil_div64(n_rows / 3, TmCfgSym::NODES)
                                  : ceil_div64(n_rows, op == 1 ? TmCfgBlock3::ROWS : TmCfgGeneric::ROWS);
  int grid = ctx->sm_count;
  if (ceil_div64(n_tiles, FU_WARPS) < grid) grid = (int)ceil_div64(n_tiles, FU_WARPS);
  if (grid < 1) grid = 1;
  unsigned* bar = (unsigned*)((char*)ctx->misc.p + 224);          // [0] counter [1] release
  double* gsum = (double*)((char*)ctx->misc.p + 256);             // [2][4]
  MYC_CUDA(ctx, cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), st));
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.n_rows = n_rows;
  a.row_offset = row_offset;
  a.rp = d_row_ptr;
  a.ci = d_col_idx;
  a.v = d_val;
  a.bval = op == 2 ? (const double*)ctx->sym_val.p : nullptr;
  a.bcol = op == 2 ? (const int32_t*)ctx->sym_col.p : nullptr;
  a.nb_total = nb_total;
  a.dinv = d_dinv;
  a.binv = d_binv;
  a.x = d_x;
  a.r = (double*)ctx->vec[1].p;
  a.w = (double*)ctx->vec[2].p;
  a.p = (double*)ctx->vec[3].p;
  a.s = (double*)ctx->vec[5].p;
  a.reg = reg;
  a.maxit = (long long)maxit;
  a.partials = (double*)ctx->partials.p;
  a.bar = bar;
  a.gsum = gsum;
  a.sc = (PcgScalars*)ctx->scalars.p;
  a.world = ctx->world;
  a.rank = ctx->rank;
  if (dist) {
    a.epoch_red0 = ctx->peer_epoch_red;
    a.epoch_halo0 = ctx->peer_epoch_halo;
    for (int q = 0; q < ctx->world; ++q) {
      a.peer_u[q] = (double*)ctx->peer_base[q];
      a.peer_sync[q] = peer_sync_of(ctx->peer_base[q], ctx->peer_cap);
      a.give_lo[q] = q == ctx->rank ? 0 : ctx->send_to[q].lo;
      a.give_hi[q] = q == ctx->rank ? 0 : ctx->send_to[q].hi;
      a.recv_any[q] = (q != ctx->rank && ctx->recv_from[q].hi > ctx->recv_from[q].lo) ? 1 : 0;
      if (a.recv_any[q]) a.recv_mask |= 1u << q;
    }
    a.halo_overlap = ctx->no_halo_overlap ? 0 : 1;
  } else {
    a.peer_u[0] = (double*)ctx->vec[0].p;
    a.peer_sync[0] = nullptr;
  }
  void* params[] = {&a};
  const void* fn = pc >= 2 ? (dist ? fused_variant(18 + op) : fused_variant(12 + op * 2 + (pc - 2)))
                           : fused_variant((dist ? 6 : 0) + op * 2 + (pc == 1 ? 1 : 0));
  MYC_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(FU_THREADS), params, fused_smem(op), st));
  ctx->launches++;
  *handled = 1;
  return MYC_OK;
}
