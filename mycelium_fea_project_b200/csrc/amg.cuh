// Aggregation multigrid for the PCG (shared declarations of amg_setup.cu and pcg_amg.cu).
//
// The reference's PETSc benchmark offers {jacobi, sor, ilu, icc, gamg} (src/fea_petsc_solverAndPC.cpp:330-331);
// this is the GPU path's counterpart of that menu's multigrid entry, designed for this operator: K is a sum of
// [[S,-S],[-S,S]] bar blocks with S symmetric 3x3 and full rank (axial + transverse spring), so its near-null
// space is the three translations and a piecewise-constant prolongation per component is the right coarse
// space.  Every level therefore keeps the fine level's format -- symmetric 3x3 node blocks, 6 values + 1 column
// index each (spmv_sym3.cuh) -- and is swept by the same TMA-pipelined kernel code.
//
// Level l:  n nodes owned by this rank (global ids node_off .. node_off+n), block row pointer brp, block columns
// bcol = 3 * GLOBAL node id, values bval (xx xy xz yy yz zz), dinv = symmetric inverse of (diagonal block + reg I).
// agg maps a node to its aggregate (index on level l+1 relative to that level's node_off, -1 = not represented
// there); mptr/mlist list the members of each node of level l on level l-1 (restriction gathers, so it needs no
// atomics).
//
// Several GPUs (one process per GPU, rows partitioned by contiguous node ranges -- the PETSc MPIAIJ layout of
// src/fea_petsc_parallel.cpp:236): aggregates are formed inside a rank, so level l+1 inherits the partition and
// restriction / prolongation need no communication; the correction vectors a sweep gathers are kept at the
// level's GLOBAL length in every rank's IPC-shared arena and the rows a neighbour reads are stored there
// directly (NVLink P2P).  A level with at most amg_replicate_nodes nodes over all ranks is REPLICATED instead:
// its operator is all-gathered once at setup, every rank then coarsens and smooths it in full, redundantly and
// bit-identically, and from there down the V-cycle needs no cross-GPU barrier at all (and aggregates are no
// longer confined to a rank, so the coarsest level ends up as small as on one GPU).  The one exchange at the
// seam: every rank restricts onto its own aggregates and stores that part of the right-hand side into
// everybody's arena.
// Algorithm and constants are restated in numpy in oracle/amg_oracle.py, which the tests compare with.
#pragma once
#include "common.cuh"

constexpr int AMG_MAX_LEVELS = 16;
constexpr int AMG_MIN_NODES = 200;       // a level with at most this many nodes (over all ranks) is the coarsest
constexpr double AMG_MAX_RATIO = 0.8;    // stop coarsening if a level does not shrink below this fraction
constexpr double AMG_OMEGA = 0.9;        // damping of the 3x3-block Jacobi smoother
constexpr double AMG_SCALE = 1.5;        // over-correction of the piecewise-constant coarse correction
constexpr int AMG_COARSE_SWEEPS = 8;     // smoother sweeps on the coarsest level

struct AmgLevelDev {                     // what the solver kernel reads (device array, one entry per level)
  int32_t n;                             // owned nodes
  int32_t node_off;                      // global id of the first owned node
  int32_t n_global;                      // nodes of this level over all ranks
  int32_t nb;                            // owned blocks
  const int32_t* brp;                    // [n+1]
  const int32_t* bcol;                   // [nb]   3 * global node id
  const double* bval;                    // [nb*6]
  const float* bval32;                   // [nb*6]  the same rounded to FP32: what the V-cycle's sweeps stream (null: FP64)
  const double* dinv;                    // [n*6]
  const int32_t* agg;                    // [n]    aggregate (local index on the next level) or -1; null on the coarsest
  const int32_t* mptr;                   // [n+1]  members on the previous level; null on level 0
  const int32_t* mlist;                  //        local node ids on the previous level
  double* r;                             // [3n]   right-hand side of this level (level 0: the CG residual)
  double* t;                             // [3n]   residual after pre-smoothing
  int64_t e_off[2];                      // offsets (doubles) of the two gathered correction vectors (3*n_global
                                         // each) inside every rank's vector arena; own entries at 3*node_off
  int64_t give_lo[MYC_MAX_WORLD];        // DOF ranges [lo, hi) (global, this level) of MY rows that peer q gathers
  int64_t give_hi[MYC_MAX_WORLD];
  int64_t zone_lo, zone_hi;              // own rows (local numbering) outside [zone_lo, zone_hi) may lie in a give range
  unsigned recv_mask;                    // bit q: this rank gathers rows of peer q on this level
  int32_t replicated;                    // 0: rows partitioned over the ranks (or one GPU); 1: first replicated level
                                         // (the seam: r is assembled from every rank's part); 2: replicated, deeper
  int32_t own_lo, own_n;                 // seam level: the aggregates of THIS rank (global ids own_lo .. own_lo+own_n);
                                         // mptr / mlist describe exactly those
  int32_t l2_keep;                       // the level's operator is small enough to stay in L2 between iterations
  int64_t r_off;                         // seam level: offset of r inside every rank's arena (r then points there)
};

struct AmgLevelHost {
  int64_t n = 0, nb = 0, n_global = 0, node_off = 0;
  int replicated = 0;
  int64_t own_lo = 0, own_n = 0;         // seam level: this rank's aggregates
  int64_t agg_shift = 0;                 // what was added to this level's agg[] so that it indexes level l+1 relative
                                         // to that level's node_off (non-zero only above the seam)
  DevBuf brp, bcol, bval, dinv, agg, mptr, mlist, r, t;
  DevBuf rep_brp, rep_bcol, rep_bval;    // seam level (replicated == 1): the all-gathered operator (brp / bcol / bval above
                                         // then hold only this rank's rows, the input of the all-gather).  Separate,
                                         // grow-only buffers: no cudaMalloc / cudaFree in the steady state of a load-case loop
  DevBuf bval32;                         // FP32 copy of the level's block values (level 0: of the context's sym_val)
  int64_t e_off[2] = {0, 0};
  int64_t r_off = -1;                    // seam level: r lives in the arena
  int64_t need_lo[MYC_MAX_WORLD] = {0}, need_hi[MYC_MAX_WORLD] = {0};     // node ranges (global, this level)
  int64_t give_lo[MYC_MAX_WORLD] = {0}, give_hi[MYC_MAX_WORLD] = {0};
};

struct AmgState {
  bool valid = false;
  int n_levels = 0;
  int64_t n_rows0 = 0, row_offset0 = 0;  // the fine operator the hierarchy belongs to
  const void* key_rp = nullptr;          // identity of that operator (pointers the caller passed)
  const void* key_val = nullptr;
  const void* key_dinv = nullptr;
  double reg = 0.0;
  AmgLevelHost lv[AMG_MAX_LEVELS];
  DevBuf lv_dev;                         // AmgLevelDev[n_levels]
  DevBuf brp0;                           // level 0 block row pointer (rp[3 i] / 9)
  DevBuf act0;                           // level 0 activity (1 = node with three free DOFs), uint8 per node
  DevBuf act_global;                     // several GPUs: level 0 activity of every node (global ids)
  DevBuf agg_global;                     // several GPUs: global aggregate id of every node of the level being coarsened
  int world = 1;                         // ranks the hierarchy was built for
  bool f32 = true;                       // the V-cycle streams FP32 copies of the level operators (MYC_AMG_FP64=1: no)
  DevBuf work[6];                        // best / paired / root / keep / flags / scan output (int32 x n)
  DevBuf arena;                          // single GPU: the gathered correction vectors of all levels
  int64_t arena_doubles = 0;             // doubles the e vectors of all levels need
  double setup_ms = 0.0;
};

// amg_setup.cu
int myc_amg_destroy(myc_ctx* ctx);
void* myc_amg_peer_sync_of(const myc_ctx* ctx, int q);   // the AgPeerSync block behind rank q's arena
// pcg_amg.cu: runs the AMG-PCG iteration loop as one persistent kernel; on entry r = b - A x0 is in ctx->vec[1]
// and sc->tol2 is set (pcg.cu).  *handled = 0: not applicable (no valid hierarchy, no cooperative launch).
int myc_pcg_amg_try(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset, const int32_t* d_row_ptr,
                    const double* d_dinv, double reg, int64_t maxit, double* d_x, cudaStream_t st, int* handled);
// Algorithmic bytes one AMG-PCG iteration has to stream on this rank (bench.py roofline): per level the block
// view is swept twice (residual, post-smoothing; the coarsest level AMG_COARSE_SWEEPS - 1 times), level 0 once
// more for w = A u, plus the vector traffic of every phase (see DESIGN.md section 4 for the per-phase table).
double myc_amg_bytes_per_iteration(const myc_ctx* ctx);
