// K5 with the multigrid preconditioner: the whole AMG-preconditioned CG solve is ONE persistent cooperative
// launch per GPU (one 1024-thread block per SM), like pcg_fused.cu, with the V-cycle inside the iteration loop.
//
// Per iteration (single-reduction CG of Chronopoulos & Gear; u = M^-1 r is the V-cycle of amg.cuh):
//   D0   p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s        (all CG recurrences, one pass)
//        fused with the pre-smoothing of level 0 from a zero guess: e_0 = omega D_0^-1 r
//   down for l = 0 .. L-2:   R_l: t_l = r_l - A_l e_l        (TMA-pipelined sweep over the level's block view)
//                            D_l+1: r_l+1 = P^T t_l (each coarse row gathers its members: no atomics),
//                                   e_l+1 = omega D^-1 r_l+1
//   coarsest: AMG_COARSE_SWEEPS - 1 further sweeps  e += omega D^-1 (r - A e)
//   up   for l = L-2 .. 0:   U1_l: e_l += AMG_SCALE * P e_l+1;   U2_l: e_l += omega D^-1 (r_l - A_l e_l)  (sweep)
//   CG   w = A_0 u + reg u with gamma = r.u, delta = w.u, r.r folded into the sweep's epilogue -> reduction
// Every phase ends in a grid barrier.  Several GPUs: on a PARTITIONED level, phases whose output another GPU
// gathers (the e vectors read by a sweep) also store the rows a neighbour needs straight into that neighbour's
// arena (P2P over NVLink) and end in a halo barrier (neighbour flags); restriction and prolongation are
// rank-local because aggregates never span ranks.  REPLICATED levels (amg.cuh) are processed by every GPU in
// full with local barriers only; at the seam every rank restricts onto its own aggregates, stores that part of
// the right-hand side into every arena, and one all-rank flag exchange completes it.  Vector passes deal rows in chunks of 30 per warp (10 whole nodes), the lane of row (node, c) gets its
// siblings' values by shuffle and applies row c of the symmetric 3x3 inverse -- the same row-per-lane layout as
// the sweep's epilogue, so all accesses are coalesced.
#include <stdlib.h>

#include "amg.cuh"
#include "amg_sweep.cuh"

namespace {

#ifndef AG_WARPS_N
#define AG_WARPS_N 32
#endif
constexpr int AG_WARPS = AG_WARPS_N;
constexpr int AG_THREADS = 32 * AG_WARPS;
constexpr unsigned AG_SPIN_LIMIT = 1u << 28;
static_assert(AMG_COARSE_SWEEPS >= 2 && AMG_COARSE_SWEEPS % 2 == 0, "the coarsest level's result must land in buffer 1");

struct AgPeerSync {                       // lives behind the vector arena in each rank's IPC-shared buffer
  double sums[2][MYC_MAX_WORLD][4];       // [parity][writer rank][gamma, delta, r.r, -]
  unsigned flag_red[MYC_MAX_WORLD];       // written by rank q: reductions q has published
  unsigned long long arrive[MYC_MAX_WORLD];   // arrive[q]: how many BLOCKS of rank q have arrived at cross-GPU phase barriers,
                                              // counted in THIS rank's memory by q's blocks themselves (red.release.sys)
};

struct AmgArgs {
  const AmgLevelDev* lv;
  int n_levels;
  double* arena[MYC_MAX_WORLD];           // vector arena of every rank ([rank] = own)
  AgPeerSync* sync[MYC_MAX_WORLD];
  const double* mask0;                    // level-0 Jacobi diagonal: 0 marks a known row
  double* x;
  double* w;
  double* p;
  double* s;
  double reg;
  long long maxit;
  double* partials;                       // [gridDim][3]
  unsigned* bar;                          // [0] arrive counter, [1] release epoch
  double* gsum;                           // [2][4]
  PcgScalars* sc;
  int world, rank;
  unsigned epoch_red0, epoch_halo0;       // barrier counts carried over from earlier solves (the flags / counters are monotone)
  unsigned epoch_seam0;
  unsigned recv_mask_all;                 // neighbours: peers this rank exchanges halo rows with on any level (symmetric)
  unsigned long long* timing;             // -DMYC_AMG_TIMING: ns per phase slot, accumulated by block 0 ([4 l + phase], [64] = CG)
};

__device__ __forceinline__ unsigned ag_ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ag_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ag_st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ag_ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier of one GPU: every block arrives with one release-atomic and spins on the counter with an
// acquire load (whose CCTL.IVALL invalidates the SM's L1) -- see local_barrier in pcg_fused.cu.
__device__ __forceinline__ void ag_grid_barrier(unsigned* bar, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    const unsigned target = epoch * gridDim.x;
    unsigned spins = 0;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
    while (ag_ld_acquire_gpu(bar) < target)
      if (++spins > AG_SPIN_LIMIT) __trap();
  }
  __syncthreads();
}

// z = (D^-1 v)_row for the row-per-lane layout: lanes 3q, 3q+1, 3q+2 (< 30) hold the rows of one node.
// All 32 lanes call; `ok` lanes get their row of the symmetric inverse (xx xy xz yy yz zz) applied.
// The row of the inverse is loaded separately (AgDinvRow, early: its latency then hides behind the rest of the
// phase) from the shuffle + multiply (ag_dinv_mul, when the residual is known).
struct AgDinvRow { double m0, m1, m2; };
__device__ __forceinline__ AgDinvRow ag_dinv_row(const double* __restrict__ dinv, int64_t row, bool ok, int lane) {
  if (!ok) return AgDinvRow{0.0, 0.0, 0.0};
  const int c = lane % 3;
  const double* d = dinv + 6 * (row / 3);
  // row c of [[d0 d1 d2] [d1 d3 d4] [d2 d4 d5]]
  return AgDinvRow{d[c], d[c == 0 ? 1 : c + 2], d[c + 2 + (c > 0)]};
}
__device__ __forceinline__ double ag_dinv_mul(const AgDinvRow& m, double v, int lane) {
  const int c = lane % 3;
  const int src = lane < 30 ? lane - c : 0;
  const double v0 = __shfl_sync(0xffffffffu, v, src);
  const double v1 = __shfl_sync(0xffffffffu, v, src + 1);
  const double v2 = __shfl_sync(0xffffffffu, v, src + 2);
  return m.m0 * v0 + m.m1 * v1 + m.m2 * v2;
}
__device__ __forceinline__ double ag_dinv_apply(const double* __restrict__ dinv, int64_t row, bool ok, double v, int lane) {
  const AgDinvRow m = ag_dinv_row(dinv, row, ok, lane);
  return ag_dinv_mul(m, v, lane);
}

// stores one row of a gathered correction vector: own arena, plus every peer that gathers that row
template <bool DIST>
struct AgPut {
  const AmgArgs* a;
  __device__ __forceinline__ void operator()(const AmgLevelDev& L, int k, int64_t row, double val) const {
    const int64_t g = 3 * (int64_t)L.node_off + row;
    a->arena[a->rank][L.e_off[k] + g] = val;
    if constexpr (DIST) {
      // rows a neighbour gathers lie in [0, zone_lo) or [zone_hi, 3n) of the own rows (strips: next to the cuts);
      // everything in between -- almost every row -- is done after the two comparisons
      if (row < L.zone_lo || row >= L.zone_hi) {
#pragma unroll 1
        for (int q = 0; q < a->world; ++q)
          if (g >= L.give_lo[q] && g < L.give_hi[q]) a->arena[q][L.e_off[k] + g] = val;
      }
    }
  }
};

// ---- sweep epilogues (row-per-lane; see tm_sym3_sweep) --------------------------------------------------
struct EpiAgResidual {      // t = r - (A e + reg e), 0 on known rows of level 0
  static constexpr int NACC = 0;
  double* t;
  const double* r;
  const double* e_own;
  const double* mask;       // null: every row is free
  double reg;
  struct Pre { double ri, ei, mi; };
  __device__ __forceinline__ Pre load(int64_t i) const { return Pre{r[i], e_own[i], mask ? mask[i] : 1.0}; }
  __device__ __forceinline__ void row(int64_t i, double sum, const Pre& pre, double (&)[1]) const {
    t[i] = pre.mi != 0.0 ? pre.ri - (sum + reg * pre.ei) : 0.0;
  }
};

template <bool DIST>
struct EpiAgSmooth {        // e_out = e + omega D^-1 (r - (A e + reg e))
  static constexpr int NACC = 0;
  static constexpr bool WARP_UNIFORM = true;
  const AmgLevelDev* L;
  int k_out;
  bool push;                // the output is gathered by another GPU's sweep
  const double* r;
  const double* e_own;
  const double* mask;
  double reg;
  AgPut<DIST> put;
  struct Pre { double ri, ei, mi; AgDinvRow m; };
  __device__ __forceinline__ Pre load(int64_t i) const {      // called by the lanes that own a row (lane = i % 30 + ...)
    return Pre{r[i], e_own[i], mask ? mask[i] : 1.0, ag_dinv_row(L->dinv, i, true, (int)(i % 3))};
  }
  __device__ __forceinline__ void row_warp(int64_t i, bool ok, double sum, const Pre& pre, double (&)[1], int lane) const {
    const double res = (ok && pre.mi != 0.0) ? pre.ri - (sum + reg * pre.ei) : 0.0;
    const double z = ag_dinv_mul(pre.m, res, lane);
    if (ok) {
      const double val = pre.ei + AMG_OMEGA * z;
      if (push) put(*L, k_out, i, val);
      else put.a->arena[put.a->rank][L->e_off[k_out] + 3 * (int64_t)L->node_off + i] = val;
    }
  }
};

struct EpiAgCg {            // w = A u + reg u ; acc = {r.u, w.u, r.r}
  static constexpr int NACC = 3;
  double* w;
  const double* u_own;
  const double* r;
  double reg;
  struct Pre { double ui, ri; };
  __device__ __forceinline__ Pre load(int64_t i) const { return Pre{u_own[i], r[i]}; }
  __device__ __forceinline__ void row(int64_t i, double sum, const Pre& pre, double (&acc)[3]) const {
    const double wi = sum + reg * pre.ui;
    w[i] = wi;
    acc[0] += pre.ri * pre.ui;
    acc[1] += wi * pre.ui;
    acc[2] += pre.ri * pre.ri;
  }
};

// F32: the sweeps inside the V-cycle stream the FP32 copies of the level operators (amg_sweep.cuh); the CG sweep
// w = A u always streams the FP64 operator.
template <bool DIST, bool F32>
__global__ void __launch_bounds__(AG_THREADS, 1) pcg_amg_kernel(AmgArgs a) {
  extern __shared__ __align__(128) unsigned char ag_smem[];
  __shared__ double s_red[AG_WARPS][3];
  __shared__ double s_tot[3];
  __shared__ int s_leader;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // 32-bit row / tile arithmetic throughout (a rank holds < 2^31 rows): the kernel runs at the 64-register limit
  const int gw = warp * (int)gridDim.x + (int)blockIdx.x;          // block-fastest: small levels spread over all SMs
  const int n_warps = (int)gridDim.x * AG_WARPS;
  const AmgLevelDev* const lv = a.lv;
  const int NL = a.n_levels;
  double* const arena = a.arena[a.rank];
  AgPeerSync* const my_sync = a.sync[a.rank];
  unsigned epoch = 0, ep_red = a.epoch_red0, ep_halo = a.epoch_halo0, ep_seam = a.epoch_seam0;
  const AgPut<DIST> put{&a};
  AgPipe pp;
  ag_pipe_init(pp, ag_smem, AG_WARPS, warp, lane);
  using PcVal = typename std::conditional<F32, float, double>::type;       // value type of the V-cycle's operators
  constexpr int PcStages = F32 ? AG_STAGES_F32 : AG_STAGES_F64;
  auto pc_val = [](const AmgLevelDev& L) -> const PcVal* {
    if constexpr (F32) return L.bval32; else return L.bval;
  };

#ifdef MYC_AMG_TIMING
  unsigned long long t_last = 0;
  auto tick = [&](int slot) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (slot >= 0) a.timing[slot] += t - t_last;
      t_last = t;
    }
  };
#else
  auto tick = [](int) {};
#endif
  // ---- barriers
  auto local_barrier = [&]() { ag_grid_barrier(&a.bar[0], epoch); };
  // arrive / leader / release with cross-GPU work done by warp 0 of the last-arriving block
  auto leader_barrier = [&](auto&& leader_work) {
    __syncthreads();
    if (threadIdx.x == 0) {
      ++epoch;
      __threadfence();
      const unsigned old = atomicAdd(&a.bar[0], 1u);
      s_leader = (old == epoch * gridDim.x - 1u);
    }
    __syncthreads();
    if (s_leader && warp == 0) {
      __threadfence();
      __syncwarp();
      leader_work(lane);
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&a.bar[1]), "r"(epoch) : "memory");
      }
    }
    if (threadIdx.x == 0) {
      unsigned spins = 0;
      while (ag_ld_acquire_gpu(&a.bar[1]) < epoch)
        if (++spins > AG_SPIN_LIMIT) __trap();
      __threadfence();
    }
    __syncthreads();
  };
  // Cross-GPU phase barrier, FLAT: no leader hop.  Every block announces itself directly -- one release-add on its own
  // GPU's counter and one on the counter each signalled peer keeps for this rank (a fire-and-forget NVLink atomic) --
  // and then waits until its own GPU and every awaited peer have been announced by ALL their blocks.  One hop instead
  // of arrive -> leader -> peer flags -> release (measured 10-14 us per barrier with the leader form on 2 GPUs).
  // The release covers the whole block's stores, P2P ones included (they are ordered before it by the __syncthreads).
  // Neighbours signal each other at every halo barrier, all ranks at the seam barrier; both sides count the same
  // barriers, so the expected value of a counter is (barriers so far) x (blocks per GPU) -- every rank of a multi-GPU
  // solve launches the same grid.  Counters are 64-bit and monotone across solves.
  auto flag_barrier = [&](bool seam) {
    if (seam) ++ep_seam; else ++ep_halo;
    __syncthreads();
    if (threadIdx.x < (unsigned)a.world) {
      const int q = (int)threadIdx.x;
      const bool nbr = (a.recv_mask_all >> q) & 1u;
      if (q == a.rank) {
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(&my_sync->arrive[a.rank]), "l"(1ull) : "memory");
      } else if (seam || nbr) {
        asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(&a.sync[q]->arrive[a.rank]), "l"(1ull) : "memory");
      }
      if (q == a.rank || seam || nbr) {
        const unsigned long long target =
            (unsigned long long)gridDim.x * ((q == a.rank || nbr) ? (unsigned long long)ep_halo + ep_seam : (unsigned long long)ep_seam);
        unsigned spins = 0;
        for (;;) {
          unsigned long long v;
          if (q == a.rank) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(&my_sync->arrive[q]) : "memory");
          else asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(&my_sync->arrive[q]) : "memory");
          if (v >= target) break;
          if (++spins > AG_SPIN_LIMIT) __trap();
        }
      }
    }
    __syncthreads();
  };
  // the level's gathered vector is complete everywhere it is read.  The wait covers the neighbours of ALL
  // partitioned levels (the same ranks on every level for contiguous strips), which also orders this rank's
  // next stores into a neighbour's vector behind that neighbour's last reads of it.
  auto halo_barrier = [&](const AmgLevelDev& L) {
    if constexpr (!DIST) {
      local_barrier();
    } else {
      if (L.replicated) local_barrier();
      else flag_barrier(false);
    }
  };

  // ---- phases
  const AmgLevelDev L0 = lv[0];
  const int n0 = 3 * L0.n;
  double* const r0 = L0.r;
  auto own = [&](const AmgLevelDev& L, int k) -> double* { return arena + L.e_off[k] + 3 * (int64_t)L.node_off; };
  const int fin_coarsest = (AMG_COARSE_SWEEPS - 1) & 1;
  auto fin = [&](int l) -> int { return l == NL - 1 ? fin_coarsest : 1; };

  // D0: CG recurrences (or their initialisation) + pre-smoothing of level 0
  auto phase_d0 = [&](bool first, double alpha, double beta) {
    const double* u = own(L0, 1);
    for (int base = gw * 30; base < n0; base += n_warps * 30) {
      const int i = base + lane;
      const bool ok = lane < 30 && i < n0;
      double ri = 0.0;
      // every load of the row is issued before the first store (the arrays may alias as far as the compiler knows)
      const AgDinvRow m = ag_dinv_row(L0.dinv, i, ok, lane);
      if (ok) {
        if (first) {
          ri = r0[i];
          a.p[i] = 0.0;
          a.s[i] = 0.0;
        } else {
          const double ui = u[i], pi0 = a.p[i], wi = a.w[i], si0 = a.s[i], xi = a.x[i], rr0 = r0[i], mk = a.mask0[i];
          const double pi = ui + beta * pi0;
          const double si = wi + beta * si0;
          ri = mk != 0.0 ? rr0 - alpha * si : 0.0;
          a.p[i] = pi;
          a.s[i] = si;
          a.x[i] = xi + alpha * pi;
          r0[i] = ri;
        }
      }
      const double z = ag_dinv_mul(m, ri, lane);
      if (ok) put(L0, 0, i, AMG_OMEGA * z);
    }
  };
  // R_l: t = r - A e   (e = buffer 0)
  auto phase_residual = [&](const AmgLevelDev& L, const double* mask) {
    double dummy[1] = {0.0};
    EpiAgResidual epi{L.t, L.r, own(L, 0), mask, a.reg};
    ag_sweep<EpiAgResidual, PcVal, PcStages>(pp, 3 * (int64_t)L.n, L.brp, pc_val(L), L.bcol, arena + L.e_off[0], epi, dummy,
                                             gw, n_warps, lane, L.nb, L.l2_keep != 0);
  };
  // seam (several GPUs): restriction onto THIS rank's aggregates of the first replicated level, stored into
  // every rank's copy of the level's right-hand side
  auto phase_restrict_seam = [&](const AmgLevelDev& Lf, const AmgLevelDev& L) {
    const int n = 3 * L.own_n;
    for (int base = gw * 30; base < n; base += n_warps * 30) {
      const int i = base + lane;
      if (lane < 30 && i < n) {
        const int nd = i / 3;
        const int c = (int)(i - 3 * nd);
        double sum = 0.0;
        const int32_t me = L.mptr[nd + 1];
        for (int32_t m = L.mptr[nd]; m < me; ++m) sum += Lf.t[3 * (int64_t)L.mlist[m] + c];
        const int64_t g = L.r_off + 3 * (int64_t)L.own_lo + i;
#pragma unroll 1
        for (int q = 0; q < a.world; ++q) a.arena[q][g] = sum;
      }
    }
  };
  // pre-smoothing from zero on a level whose right-hand side is complete: e = omega D^-1 r
  auto phase_presmooth = [&](const AmgLevelDev& L) {
    const int n = 3 * L.n;
    for (int base = gw * 30; base < n; base += n_warps * 30) {
      const int i = base + lane;
      const bool ok = lane < 30 && i < n;
      const double ri = ok ? L.r[i] : 0.0;
      const double z = ag_dinv_apply(L.dinv, i, ok, ri, lane);
      if (ok) put(L, 0, i, AMG_OMEGA * z);
    }
  };
  // D_l (l >= 1): restriction by member lists + pre-smoothing from zero
  auto phase_restrict = [&](const AmgLevelDev& Lf, const AmgLevelDev& L) {
    const int n = 3 * L.n;
    for (int base = gw * 30; base < n; base += n_warps * 30) {
      const int i = base + lane;
      const bool ok = lane < 30 && i < n;
      double sum = 0.0;
      const AgDinvRow dm = ag_dinv_row(L.dinv, i, ok, lane);
      if (ok) {
        const int nd = i / 3;
        const int c = i - 3 * nd;
        const int32_t mb = L.mptr[nd], me = L.mptr[nd + 1];
        // aggregates have 2..4 members as a rule: fetch the first four lists entries / values together
        const int32_t k0 = mb < me ? L.mlist[mb] : 0, k1 = mb + 1 < me ? L.mlist[mb + 1] : 0;
        const int32_t k2 = mb + 2 < me ? L.mlist[mb + 2] : 0, k3 = mb + 3 < me ? L.mlist[mb + 3] : 0;
        const double t0 = mb < me ? Lf.t[3 * (int64_t)k0 + c] : 0.0, t1 = mb + 1 < me ? Lf.t[3 * (int64_t)k1 + c] : 0.0;
        const double t2 = mb + 2 < me ? Lf.t[3 * (int64_t)k2 + c] : 0.0, t3 = mb + 3 < me ? Lf.t[3 * (int64_t)k3 + c] : 0.0;
        sum = ((t0 + t1) + t2) + t3;                         // member order, like the loop below
        for (int32_t m = mb + 4; m < me; ++m) sum += Lf.t[3 * (int64_t)L.mlist[m] + c];
        L.r[i] = sum;
      }
      const double z = ag_dinv_mul(dm, sum, lane);
      if (ok) put(L, 0, i, AMG_OMEGA * z);
    }
  };
  // smoothing sweep: e[k_out] = e[k_in] + omega D^-1 (r - A e[k_in])
  auto phase_smooth = [&](const AmgLevelDev& L, int k_in, int k_out, const double* mask, bool push) {
    double dummy[1] = {0.0};
    EpiAgSmooth<DIST> epi{&L, k_out, push, L.r, own(L, k_in), mask, a.reg, put};
    ag_sweep<EpiAgSmooth<DIST>, PcVal, PcStages>(pp, 3 * (int64_t)L.n, L.brp, pc_val(L), L.bcol, arena + L.e_off[k_in], epi,
                                                 dummy, gw, n_warps, lane, L.nb, L.l2_keep != 0);
  };
  // U1_l: e_l += SCALE * P e_l+1
  auto phase_prolong = [&](const AmgLevelDev& L, const AmgLevelDev& Lc, int kc) {
    const int n = 3 * L.n;
    const double* e = own(L, 0);
    const double* ec = own(Lc, kc);
    const int stride = n_warps * 30;
    for (int base = gw * 30; base < n; base += 2 * stride) {          // two chunks per round: twice the loads in flight
      const int i0 = base + lane, i1 = i0 + stride;
      const bool ok0 = lane < 30 && i0 < n, ok1 = lane < 30 && i1 < n;
      const int nd0 = i0 / 3, nd1 = i1 / 3;
      const int32_t ag0 = ok0 ? L.agg[nd0] : -1, ag1 = ok1 ? L.agg[nd1] : -1;
      const double e0 = ag0 >= 0 ? e[i0] : 0.0, e1 = ag1 >= 0 ? e[i1] : 0.0;
      const double c0 = ag0 >= 0 ? ec[3 * (int64_t)ag0 + (i0 - 3 * nd0)] : 0.0;
      const double c1 = ag1 >= 0 ? ec[3 * (int64_t)ag1 + (i1 - 3 * nd1)] : 0.0;
      if (ag0 >= 0) put(L, 0, i0, e0 + AMG_SCALE * c0);
      if (ag1 >= 0) put(L, 0, i1, e1 + AMG_SCALE * c1);
    }
  };

  // The CG scalars are block-uniform: they live in shared memory (thread 0 updates them after each reduction), not
  // in every thread's registers -- the kernel runs at the 64-register limit and these would be live in every phase.
  __shared__ double s_cg[5];          // gamma_old, alpha_old, rr, alpha, beta
  __shared__ long long s_it;
  __shared__ int s_status;            // -1 running, 1 converged, 2 breakdown, 0 maxit
  if (threadIdx.x == 0) {
    s_cg[0] = 1.0; s_cg[1] = 1.0; s_cg[2] = 0.0; s_cg[3] = 0.0; s_cg[4] = 0.0;
    s_it = 0;
    s_status = -1;
  }
  __syncthreads();
  bool first = true;
  for (;;) {
    // ---- u = M^-1 r : V-cycle, its first phase fused with the CG recurrences
    tick(-1);
    phase_d0(first, s_cg[3], s_cg[4]);
    first = false;
    halo_barrier(L0);
    tick(0);
    if (NL == 1) {
      for (int k = 1; k < AMG_COARSE_SWEEPS; ++k) {
        phase_smooth(lv[0], (k - 1) & 1, k & 1, a.mask0, true);
        halo_barrier(L0);
      }
    } else {
      for (int l = 0; l + 1 < NL; ++l) {
        phase_residual(lv[l], l == 0 ? a.mask0 : nullptr);
        local_barrier();
        tick(4 * l + 1);
        if constexpr (DIST) {
          if (lv[l + 1].replicated == 1) {
            phase_restrict_seam(lv[l], lv[l + 1]);
            flag_barrier(true);                       // every rank's part of r has landed
            phase_presmooth(lv[l + 1]);
            local_barrier();
            tick(4 * (l + 1));
            continue;
          }
        }
        phase_restrict(lv[l], lv[l + 1]);
        halo_barrier(lv[l + 1]);
        tick(4 * (l + 1));
      }
      for (int k = 1; k < AMG_COARSE_SWEEPS; ++k) {
        phase_smooth(lv[NL - 1], (k - 1) & 1, k & 1, nullptr, true);
        halo_barrier(lv[NL - 1]);
      }
      tick(4 * (NL - 1) + 1);
      for (int l = NL - 2; l >= 0; --l) {
        phase_prolong(lv[l], lv[l + 1], fin(l + 1));
        halo_barrier(lv[l]);
        tick(4 * l + 2);
        phase_smooth(lv[l], 0, 1, l == 0 ? a.mask0 : nullptr, l == 0);
        if (l == 0) halo_barrier(L0); else local_barrier();
        tick(4 * l + 3);
      }
    }
    // ---- w = A u, partial dots
    double acc[3] = {0.0, 0.0, 0.0};
    {
      EpiAgCg epi{a.w, own(L0, 1), r0, a.reg};
      ag_sweep<EpiAgCg, double, AG_STAGES_F64>(pp, n0, L0.brp, L0.bval, L0.bcol, arena + L0.e_off[1], epi, acc, gw, n_warps,
                                               lane, L0.nb, L0.l2_keep != 0);
    }
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
      for (int wq = 0; wq < AG_WARPS; ++wq) t += s_red[wq][threadIdx.x];
      a.partials[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
#ifdef MYC_AMG_TIMING
    local_barrier();            // timing builds only: separates the CG sweep (slot 64) from the reduction exchange (slot 65)
    tick(64);
#endif
    // ---- reduction barrier: every GPU obtains bit-identical global sums
    if constexpr (!DIST) {
      local_barrier();
      if (warp == 0) {
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
      ++ep_red;
      const unsigned er = ep_red;
      const int par = (int)(er & 1u);
      leader_barrier([&](int ln) {
        double tot[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double t = 0.0;
          for (unsigned b = ln; b < gridDim.x; b += 32) t += __ldcg(&a.partials[(size_t)b * 3 + j]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          tot[j] = t;
        }
        if (ln < a.world) {        // one lane per peer: publish the local totals everywhere, collect everybody's
          double* slot = a.sync[ln]->sums[par][a.rank];
          slot[0] = tot[0]; slot[1] = tot[1]; slot[2] = tot[2];
          __threadfence_system();
          ag_st_release_sys(&a.sync[ln]->flag_red[a.rank], er);
          unsigned spins = 0;
          while (ag_ld_acquire_sys(&my_sync->flag_red[ln]) < er)
            if (++spins > AG_SPIN_LIMIT) __trap();
        }
        __syncwarp();
        if (ln < 3) {
          double t = 0.0;
          for (int q = 0; q < a.world; ++q) t += ag_ld_volatile_f64(&my_sync->sums[par][q][ln]);   // rank order
          a.gsum[par * 4 + ln] = t;
        }
      });
      if (threadIdx.x < 3) s_tot[threadIdx.x] = __ldcg(&a.gsum[par * 4 + threadIdx.x]);
    }
    __syncthreads();
#ifdef MYC_AMG_TIMING
    tick(65);
#else
    tick(64);
#endif
    if (threadIdx.x == 0) {   // every block computes the same scalars from the same bit-identical sums
      const double gamma = s_tot[0], delta = s_tot[1], rr = s_tot[2];
      const long long it = s_it;
      s_cg[2] = rr;
      int status = -1;
      if (!isfinite(rr)) status = 2;
      else if (!(rr > a.sc->tol2)) status = 1;
      else if (it >= a.maxit) status = 0;
      else {
        const double beta = (it == 0) ? 0.0 : gamma / s_cg[0];
        const double denom = (it == 0) ? delta : delta - beta * gamma / s_cg[1];
        if (!(denom > 0.0) || !isfinite(gamma)) status = 2;
        else {
          const double alpha = gamma / denom;
          s_cg[0] = gamma; s_cg[1] = alpha; s_cg[3] = alpha; s_cg[4] = beta;
          s_it = it + 1;
        }
      }
      s_status = status;
    }
    __syncthreads();          // (also: s_tot is rewritten by the next reduction)
    if (s_status >= 0) break;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int status = s_status;
    a.sc->iters = s_it;
    a.sc->rr_final = s_cg[2];
    a.sc->red[1] = s_cg[2];
    a.sc->done = (status == 1);
    a.sc->breakdown = (status == 2);
    a.sc->pAp = (double)ep_red;          // final epochs, carried into the next solve by the host
    a.sc->rz_old = (double)ep_halo;
    a.sc->out[3] = (double)ep_seam;
  }
}

}  // namespace

size_t myc_amg_peer_tail_bytes() { return sizeof(AgPeerSync) + 256; }

int myc_pcg_amg_try(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset, const int32_t* d_row_ptr,
                    const double* d_dinv, double reg, int64_t maxit, double* d_x, cudaStream_t st, int* handled) {
  *handled = 0;
  AmgState* S = ctx->amg;
  if (!S || !S->valid || S->n_rows0 != n_rows || S->row_offset0 != row_offset || S->key_rp != d_row_ptr ||
      S->key_dinv != d_dinv || ctx->sym_owner != 1)
    MYC_FAIL(ctx, MYC_ERR_STATE, "pcg_solve(MYC_PC_AMG): call myc_amg_setup for this operator and Dirichlet set first");
  const bool dist = ctx->world > 1;
  if (S->world != ctx->world || (dist && (!ctx->amg_peer_own || ctx->amg_peer_cap < S->arena_doubles)))
    MYC_FAIL(ctx, MYC_ERR_STATE, "pcg_solve(MYC_PC_AMG): the hierarchy was built for another communicator");
  int coop = 0;
  MYC_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  const size_t smem = ag_smem_bytes(AG_WARPS);
  const bool f32 = S->f32;
  if (ctx->amg_max_blocks_per_sm < 0) {
    int mn = 1 << 30;
    const void* fns[4] = {(const void*)pcg_amg_kernel<false, false>, (const void*)pcg_amg_kernel<true, false>,
                          (const void*)pcg_amg_kernel<false, true>, (const void*)pcg_amg_kernel<true, true>};
    for (const void* fn : fns) {
      int b = 0;
      MYC_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaFuncAttributes fa;
      MYC_CUDA(ctx, cudaFuncGetAttributes(&fa, fn));
      MYC_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         myc_carveout_percent(smem, fa.sharedSizeBytes, 1)));
      MYC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fn, AG_THREADS, smem));
      mn = b < mn ? b : mn;
    }
    ctx->amg_max_blocks_per_sm = mn;
  }
  if (!coop || ctx->amg_max_blocks_per_sm < 1)
    MYC_FAIL(ctx, MYC_ERR_STATE, "pcg_solve(MYC_PC_AMG): cooperative launch unavailable on this device");
  MYC_TRY(myc_ensure(ctx, ctx->vec[2], (size_t)(n_rows + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->vec[3], (size_t)(n_rows + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->vec[5], (size_t)(n_rows + 1) * sizeof(double)));
  MYC_TRY(myc_ensure(ctx, ctx->misc, 512));
  MYC_TRY(myc_ensure(ctx, ctx->partials, (size_t)ctx->sm_count * 16 * 4 * sizeof(double)));
  // ---- level table
  AmgLevelDev h_lv[AMG_MAX_LEVELS];
  memset(h_lv, 0, sizeof(h_lv));
  unsigned recv_mask_all = 0;
  // small levels stay in L2 across iterations (126 MB): evict-last for them, evict-first for the big streams.
  // Counted from the coarsest level up: a level's operator <= 16 MB, 48 MB in total.
  bool l2_keep[AMG_MAX_LEVELS] = {false};
  {
    double total = 0.0;
    for (int l = S->n_levels - 1; l >= 0; --l) {
      const double need = (double)S->lv[l].nb * 52.0;
      if (need > 16e6 || total + need > 48e6) break;
      total += need;
      l2_keep[l] = true;
    }
  }
  for (int l = 0; l < S->n_levels; ++l) {
    AmgLevelHost& H = S->lv[l];
    AmgLevelDev& D = h_lv[l];
    D.n = (int32_t)H.n; D.node_off = (int32_t)H.node_off; D.n_global = (int32_t)H.n_global; D.nb = (int32_t)H.nb;
    const bool seam = H.replicated == 1;            // the seam level's operator is the all-gathered copy
    D.brp = l == 0 ? (const int32_t*)S->brp0.p : (const int32_t*)(seam ? H.rep_brp.p : H.brp.p);
    D.bcol = l == 0 ? (const int32_t*)ctx->sym_col.p : (const int32_t*)(seam ? H.rep_bcol.p : H.bcol.p);
    D.bval = l == 0 ? (const double*)ctx->sym_val.p : (const double*)(seam ? H.rep_bval.p : H.bval.p);
    D.bval32 = f32 ? (const float*)H.bval32.p : nullptr;
    D.dinv = (const double*)H.dinv.p;
    D.agg = l + 1 < S->n_levels ? (const int32_t*)H.agg.p : nullptr;
    D.mptr = l > 0 ? (const int32_t*)H.mptr.p : nullptr;
    D.mlist = l > 0 ? (const int32_t*)H.mlist.p : nullptr;
    D.r = l == 0 ? (double*)ctx->vec[1].p : (double*)H.r.p;
    D.t = (double*)H.t.p;
    D.e_off[0] = H.e_off[0]; D.e_off[1] = H.e_off[1];
    D.l2_keep = l2_keep[l] ? 1 : 0;
    D.replicated = H.replicated;
    D.own_lo = (int32_t)H.own_lo; D.own_n = (int32_t)H.own_n;
    D.r_off = H.r_off;
    if (H.replicated == 1) D.r = (double*)ctx->amg_peer_own + H.r_off;
    D.zone_lo = 0;
    D.zone_hi = 3 * (int64_t)H.n;                     // no row is pushed (single GPU, replicated level)
    if (dist && !H.replicated) {
      const int64_t own_lo = 3 * H.node_off;          // (every give range ends up inside one of the two zones by construction)
      for (int q = 0; q < ctx->world; ++q) {
        if (q == ctx->rank) continue;
        D.give_lo[q] = 3 * H.give_lo[q]; D.give_hi[q] = 3 * H.give_hi[q];
        if (H.need_hi[q] > H.need_lo[q] || H.give_hi[q] > H.give_lo[q]) D.recv_mask |= 1u << q;   // (symmetric by construction)
        if (D.give_hi[q] > D.give_lo[q]) {            // lower ranks read next to the low cut, higher ranks next to the high cut
          if (q < ctx->rank) D.zone_lo = D.give_hi[q] - own_lo > D.zone_lo ? D.give_hi[q] - own_lo : D.zone_lo;
          else D.zone_hi = D.give_lo[q] - own_lo < D.zone_hi ? D.give_lo[q] - own_lo : D.zone_hi;
        }
      }
    }
    recv_mask_all |= D.recv_mask;
  }
  MYC_CUDA(ctx, cudaMemcpyAsync(S->lv_dev.p, h_lv, sizeof(AmgLevelDev) * S->n_levels, cudaMemcpyHostToDevice, st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));      // h_lv is a stack array
  unsigned* bar = (unsigned*)((char*)ctx->misc.p + 224);
  double* gsum = (double*)((char*)ctx->misc.p + 256);
  MYC_CUDA(ctx, cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), st));
  AmgArgs a;
  memset(&a, 0, sizeof(a));
  a.lv = (const AmgLevelDev*)S->lv_dev.p;
  a.n_levels = S->n_levels;
  if (dist) {
    for (int q = 0; q < ctx->world; ++q) {
      a.arena[q] = (double*)ctx->amg_peer_base[q];
      a.sync[q] = (AgPeerSync*)myc_amg_peer_sync_of(ctx, q);
    }
    a.epoch_red0 = ctx->amg_epoch_red;
    a.epoch_halo0 = ctx->amg_epoch_halo;
    a.epoch_seam0 = ctx->amg_epoch_seam;
    a.recv_mask_all = recv_mask_all;
  } else {
    a.arena[0] = (double*)S->arena.p;
    a.sync[0] = (AgPeerSync*)((char*)S->arena.p + (((size_t)S->arena_doubles + 16) * sizeof(double) + 255) / 256 * 256);
  }
  a.mask0 = d_dinv;
  a.x = d_x;
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
#ifdef MYC_AMG_TIMING
  MYC_TRY(myc_ensure(ctx, ctx->xchg, 4096));
  a.timing = (unsigned long long*)ctx->xchg.p;
  MYC_CUDA(ctx, cudaMemsetAsync(a.timing, 0, 72 * sizeof(unsigned long long), st));
#endif
  const int64_t n_tiles = ceil_div64(n_rows / 3, AG_TILE_NODES < AG_TILE_NODES_F64 ? AG_TILE_NODES : AG_TILE_NODES_F64);
  int grid = ctx->sm_count;
  // (several GPUs: every rank launches the full grid -- the flat cross-GPU barrier counts blocks)
  if (!dist && ceil_div64(n_tiles, AG_WARPS) < grid) grid = (int)ceil_div64(n_tiles, AG_WARPS);
  if (grid < 1) grid = 1;
  void* params[] = {&a};
  // diagnostic: MYC_AMG_FORCE_DIST_KERNEL=1 runs the multi-GPU instantiation on one GPU (no peers: its flag barriers
  // degenerate to leader barriers), which separates code-generation effects from communication when profiling
  static const bool force_dist = getenv("MYC_AMG_FORCE_DIST_KERNEL") && getenv("MYC_AMG_FORCE_DIST_KERNEL")[0] == '1';
  if (force_dist && !dist) MYC_CUDA(ctx, cudaMemsetAsync(a.sync[0], 0, sizeof(AgPeerSync), st));
  const void* fn = (dist || force_dist) ? (f32 ? (const void*)pcg_amg_kernel<true, true> : (const void*)pcg_amg_kernel<true, false>)
                        : (f32 ? (const void*)pcg_amg_kernel<false, true> : (const void*)pcg_amg_kernel<false, false>);
  MYC_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(AG_THREADS), params, smem, st));
  ctx->launches++;
  *handled = 1;
#ifdef MYC_AMG_TIMING
  {
    unsigned long long h[72];
    MYC_CUDA(ctx, cudaMemcpyAsync(h, a.timing, sizeof(h), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    fprintf(stderr, "[amg timing, block 0, us summed over the solve] level: presmooth/restrict  residual(|coarsest sweeps)  prolong  postsmooth\n");
    for (int l = 0; l < S->n_levels; ++l)
      fprintf(stderr, "  L%-2d n=%-9lld %10.1f %10.1f %10.1f %10.1f\n", l, (long long)S->lv[l].n, h[4 * l] / 1e3, h[4 * l + 1] / 1e3,
              h[4 * l + 2] / 1e3, h[4 * l + 3] / 1e3);
    fprintf(stderr, "  CG sweep %10.1f   reduction %10.1f\n", h[64] / 1e3, h[65] / 1e3);
  }
#endif
  return MYC_OK;
}
