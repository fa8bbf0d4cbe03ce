/*
 * mycelium_fea.h -- C-ABI of the B200-native FEA hot path (libmycelium_fea_b200.so).
 *
 * The reference (YiKwanwoo2/mycelium-fea-project) has no FFI/plugin interface: its boundary
 * is four Python functions and a CSV directory layout.  This header is
 * the native layer the Python drop-in (mycelium_fea_project_b200/fea_solver.py) binds with
 * ctypes; each entry point names the reference code it replaces.
 *
 * Conventions
 *   - every function returns int: MYC_OK (0) or a negative MYC_ERR_*; myc_last_error() gives
 *     the message.  No C++ exception crosses this boundary.
 *   - pointers named d_* are DEVICE pointers on the context's device (e.g. torch
 *     tensor.data_ptr()); pointers named h_* are HOST pointers.  The caller owns every
 *     input/output buffer and keeps it alive until the stream has drained; the library owns
 *     only scratch inside the context.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     NULL is the legacy default stream.  Calls are asynchronous on that stream except where a
 *     host scalar (h_*) is returned, which implies one stream synchronise.
 *   - one context per device per thread; contexts are not thread-safe.
 *   - float64 everywhere; DOF/column indices are int32 (scipy's CSR index type at these
 *     sizes), element end nodes are int32, DOF lists are int64 (numpy default).
 *   - DOF layout is the reference's: dof = 3*node + {0,1,2}  (src/fea_solver.py:96).
 */
#ifndef MYCELIUM_FEA_H
#define MYCELIUM_FEA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MYC_ABI_VERSION 1

enum {
  MYC_OK = 0,
  MYC_ERR_BAD_ARG = -1,        /* null pointer, negative size, node id out of range ...      */
  MYC_ERR_CUDA = -2,           /* a CUDA runtime call failed                                  */
  MYC_ERR_NCCL = -3,           /* NCCL missing or a collective failed                         */
  MYC_ERR_NOT_CONVERGED = -4,  /* PCG hit maxit (x holds the last iterate)                    */
  MYC_ERR_BREAKDOWN = -5,      /* PCG: p.Ap <= 0 or non-finite residual                       */
  MYC_ERR_CAPACITY = -6,       /* output buffer too small / int32 index overflow              */
  MYC_ERR_STATE = -7           /* call order violated (e.g. numeric before symbolic)          */
};

typedef struct myc_ctx myc_ctx;

/* preconditioners of myc_pcg_solve (the reference's PETSc menu has jacobi / bjacobi:
 * src/fea_petsc_solverAndPC.cpp:331, src/fea_petsc_parallel.cpp:339) */
enum {
  MYC_PC_JACOBI = 0,   /* point Jacobi                                                          */
  MYC_PC_BLOCK3 = 1,   /* 3x3 node blocks                                                       */
  MYC_PC_BLOCK6 = 2,   /* aligned blocks of 2 consecutive nodes (rows 6k .. 6k+5)               */
  MYC_PC_BLOCK12 = 3,  /* aligned blocks of 4 consecutive nodes (rows 12k .. 12k+11)            */
  MYC_PC_AMG = 4       /* aggregation-multigrid V-cycle (myc_amg_setup): the counterpart of the
                          reference's "gamg" entry, src/fea_petsc_solverAndPC.cpp:331              */
};

int myc_abi_version(void);

/* Create / destroy the per-device context (scratch arenas, partial-sum buffers). */
int myc_create(int device_ordinal, myc_ctx** out_ctx);
int myc_destroy(myc_ctx* ctx);
/* Message of the last failure on this context (ctx may be NULL: last create failure). */
const char* myc_last_error(const myc_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py "gpu_launches"). */
int64_t myc_launch_count(const myc_ctx* ctx);

/* Sampled device timing of the dominant kernel (the fused SpMV inside myc_pcg_solve), for the
 * live roofline figure of bench.py.  reset(enable): clears the counters and switches sampling
 * on/off (off by default; when on, every 32nd SpMV launch of a solve is bracketed by CUDA
 * events on the solve's stream).  get: out[0] = summed duration of the sampled launches (ms),
 * out[1] = number of sampled launches, out[2] = summed algorithmic bytes of the sampled
 * launches (12*nnz + 20*n_rows each), out[3] = total SpMV launches since reset. */
int myc_profile_reset(myc_ctx* ctx, int enable);
int myc_profile_get(myc_ctx* ctx, double* h_out4);

/* Structure hint for every CSR the caller passes from now on (sticky): non-zero = the matrix has
 * the 3x3 node-block structure of this problem (3 DOF per node: the three rows of a node have equal
 * length and their columns come in triples 3c,3c+1,3c+2 -- what myc_assemble_* emits), which lets
 * the SpMV inside myc_spmv / myc_apply_dirichlet / myc_pcg_solve / myc_true_residual use one column
 * index and one x gather per block.  A wrong hint gives wrong results; myc_csr_is_block3 verifies
 * an arbitrary CSR on the device (one pass over col_idx). */
int myc_set_csr_hint(myc_ctx* ctx, int node_block3);
int myc_csr_is_block3(myc_ctx* ctx, int64_t n_rows, const int32_t* d_row_ptr, const int32_t* d_col_idx,
                      int* h_out_is_block3, void* stream);

/* ------------------------------------------------------------------------------------------
 * K1  element stiffness.  Replaces bar_stiffness_bulk(p1s,p2s,E,A,I) -> (K(N,6,6), L(N,))
 *     src/fea_solver.py:30-68   (C++ twin: element_stiffness_6x6, src/fea_petsc.cpp:88-140)
 * d_p1s, d_p2s : (n,3) row-major end points.  d_out_ke : (n,6,6) row-major.  d_out_L : (n,)
 * (the UNclamped length, as the reference returns it).  Products are rounded before adds
 * (no FMA contraction) exactly like the numpy expression; L**3 is the correctly rounded cube.
 */
int myc_bar_stiffness_bulk(myc_ctx* ctx, const double* d_p1s, const double* d_p2s, int64_t n,
                           double E, double A, double I, double* d_out_ke, double* d_out_L,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * K2+K3  assembly.  Replaces assemble_global_stiffness(coords, elems, active) -> csr_matrix
 *     src/fea_solver.py:74-106  (PETSc twin: MatSetValue loop, src/fea_petsc.cpp:229-263)
 * Two phases so that the caller can size the output:
 *   symbolic: radix-sorts the directed node pairs of the active elements and writes the CSR
 *             row pointer of the rows this context owns; returns nnz through h_out_nnz.
 *   numeric : writes col_idx / val (K_e is evaluated inside, the COO stream never exists).
 * Rows owned = DOFs of nodes [node_begin, node_end) (whole mesh: 0, n_nodes); row_ptr is
 * local (starts at 0, 3*(node_end-node_begin)+1 entries), col_idx is GLOBAL.  The result is
 * bit-identical in row_ptr / col_idx to the reference's K.indptr / K.indices (sorted columns,
 * merged duplicates, explicit zeros kept); values are summed in a fixed order (deterministic).
 * d_active may be NULL (all active).  Element end nodes must lie in [0, n_nodes).
 */
int myc_assemble_symbolic(myc_ctx* ctx, const int32_t* d_n1, const int32_t* d_n2,
                          const uint8_t* d_active, int64_t n_elem, int64_t n_nodes,
                          int64_t node_begin, int64_t node_end, int32_t* d_out_row_ptr,
                          int64_t* h_out_nnz, void* stream);
int myc_assemble_numeric(myc_ctx* ctx, const double* d_coords, const int32_t* d_n1,
                         const int32_t* d_n2, double E, double A, double I, int64_t nnz_capacity,
                         const int32_t* d_row_ptr, int32_t* d_out_col_idx, double* d_out_val,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * K4  Dirichlet elimination.  Replaces the reduction half of solve_system
 *     src/fea_solver.py:113-125  (PETSc twin: MatZeroRowsColumnsIS, src/fea_petsc.cpp:309-325)
 * Works in the full index space instead of extracting K_ff: the operator the solver sees is
 *   A = P K P + reg*P  on free rows (P = projector on free DOFs), identity elsewhere,
 * which is K_ff + reg*I on the free DOFs.  Outputs (all local length n_rows unless noted):
 *   d_out_ubc   (n_cols_global) prescribed values scattered, 0 elsewhere
 *   d_out_rhs   b = -K_fk u_k on free rows, 0 on known rows          (fea_solver.py:121-122)
 *   d_out_dinv  Jacobi: 1/(K_ii + reg) on free rows, 0 on known rows; 0 marks "known"
 * Duplicate entries in d_known_dofs are NOT allowed (the reference passes dict keys).
 * row_offset = 3*node_begin of the owned block (0 on one GPU).
 */
int myc_apply_dirichlet(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                        const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                        const int64_t* d_known_dofs, const double* d_known_vals, int64_t n_known,
                        double reg, double* d_out_ubc, double* d_out_rhs, double* d_out_dinv,
                        void* stream);

/* 3x3 node-block Jacobi: inverse of the free-free part of each node's diagonal block of
 * K + reg*I (rows/cols of known DOFs replaced by identity).  d_out_binv: (n_rows/3, 9). */
int myc_block3_inverse(myc_ctx* ctx, int64_t n_rows, int64_t row_offset, const int32_t* d_row_ptr,
                       const int32_t* d_col_idx, const double* d_val, const double* d_dinv,
                       double reg, double* d_out_binv, void* stream);

/* Node-group block Jacobi (MYC_PC_BLOCK6 / MYC_PC_BLOCK12): inverse of the aligned diagonal blocks of
 * nodes_per_block (2 or 4) consecutive nodes -- R = 3*nodes_per_block rows -- of K + reg*I, rows/cols
 * of known DOFs (and the padding of a ragged last block) zeroed; a block that is singular in floating
 * point (coincident nodes coupled with ~1e28) keeps the inverses of its 3x3 node blocks only.  The layout
 * of d_out_pinv belongs to the library (6x6 blocks row by row, 36 doubles each; 12x12 blocks as
 * symmetric-packed upper triangles, entry (i <= j) at i*R - i(i-1)/2 + (j-i), 78 doubles each):
 * allocate myc_block_inverse_size(nodes_per_block, n_rows) doubles and pass the buffer to myc_pcg_solve
 * unchanged.  Single GPU (row_offset must be a multiple of R). */
int64_t myc_block_inverse_size(int nodes_per_block, int64_t n_rows);
int myc_block_inverse_packed(myc_ctx* ctx, int nodes_per_block, int64_t n_rows, int64_t row_offset,
                             const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                             const double* d_dinv, double reg, double* d_out_pinv, void* stream);

/* Aggregation multigrid for MYC_PC_AMG (the multigrid row of the reference's PETSc preconditioner menu,
 * src/fea_petsc_solverAndPC.cpp:330-331, restated for this operator; algorithm in oracle/amg_oracle.py).
 * Builds, inside the context, a hierarchy of Galerkin operators on node aggregates (pairwise matching by
 * coupling strength + joining of leftovers; piecewise-constant prolongation per displacement component;
 * 3x3-block Jacobi smoothing) for the operator A of myc_apply_dirichlet: call it after myc_apply_dirichlet
 * with the same CSR and the d_dinv it produced, then pass MYC_PC_AMG (d_binv = NULL) to myc_pcg_solve with
 * the same d_row_ptr / d_dinv.  The hierarchy stays valid for further right-hand sides on the same operator
 * and Dirichlet set.  *h_out_levels = 0 means "not applicable" (the CSR lacks the 3x3 node-block structure or
 * blockwise symmetry, or a node has only some of its DOFs prescribed): use a Jacobi-type preconditioner.
 * Several GPUs (after myc_dist_init + myc_dist_set_plan; the row block must be the installed partition's):
 * COLLECTIVE -- every rank calls it with its rows, and every rank gets the same *h_out_levels.  Aggregates are
 * formed inside a rank, levels stay row-partitioned with the correction vectors exchanged through NVLink peer
 * memory; a level with at most 65,536 nodes over all ranks (MYC_AMG_REPLICATE_NODES) is gathered onto every
 * rank and processed redundantly from there down.  This is the GPU path's form of the reference's PCBJACOBI /
 * GAMG under mpirun (src/fea_petsc_parallel.cpp:330-351).
 * myc_amg_level_info (this rank's part): out[0] nodes, out[1] blocks of `level`, out[2] number of levels,
 * out[3] setup time (us), out[4] global id of the first held node, out[5] nodes of the level over all ranks,
 * out[6] 0 = row-partitioned (or single GPU) / 1 = first replicated level / 2 = replicated, out[7] = offset
 * already added to the aggregate map so that it indexes the next level from this rank's first node there;
 * d_out_agg (may be NULL): the level's node -> aggregate map (int32, -1 = not represented on the next level). */
int myc_amg_setup(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                  const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                  const double* d_dinv, double reg, int* h_out_levels, void* stream);
int myc_amg_level_info(myc_ctx* ctx, int level, int64_t* h_out8, int32_t* d_out_agg, void* stream);

/* Structure parity aid: the explicit reduced matrix K[free][:,free] the reference builds
 * (src/fea_solver.py:118), as CSR over the compacted free numbering.  Two-phase like assembly:
 * call with d_out_col_idx == NULL to get row_ptr + nnz, then again with buffers. */
int myc_reduce_csr(myc_ctx* ctx, int64_t n_rows, const int32_t* d_row_ptr, const int32_t* d_col_idx,
                   const double* d_val, const double* d_dinv, int32_t* d_out_free_index,
                   int32_t* d_out_row_ptr, int32_t* d_out_col_idx, double* d_out_val,
                   int64_t* h_out_n_free, int64_t* h_out_nnz, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5a/K6  y = K x.  Replaces K @ U (src/fea_solver.py:257; MatMult, src/fea_petsc.cpp:363) and
 * is the SpMV inside the solver.  d_x has n_cols_global entries (global column space), d_y has
 * n_rows.  On a distributed context the halo of d_x is NOT refreshed here (see myc_halo_exchange).
 */
int myc_spmv(myc_ctx* ctx, int64_t n_rows, const int32_t* d_row_ptr, const int32_t* d_col_idx,
             const double* d_val, const double* d_x, double* d_y, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5  preconditioned CG.  Replaces spsolve(K_ff, F_f) (src/fea_solver.py:128) and
 *     KSPSolve with KSPCG (src/fea_petsc.cpp:323-341, src/fea_petsc_parallel.cpp:330-351).
 * Solves A x = b with A as defined under myc_apply_dirichlet.  d_x (n_rows) holds the initial
 * guess on entry (must be 0 on known rows) and the solution on return.  Converged when
 * ||r||2 <= max(rtol*||b||2, atol).  d_binv: NULL for MYC_PC_JACOBI, the output of myc_block3_inverse
 * for MYC_PC_BLOCK3, of myc_block_inverse_packed for MYC_PC_BLOCK6 / MYC_PC_BLOCK12 (those two run only
 * in the single-GPU persistent solver kernel and return MYC_ERR_STATE where it is unavailable).
 * On a distributed context (myc_dist_init) every rank calls this collectively; dot products
 * are NCCL all-reduces and the search direction's halo is exchanged every iteration.
 * h_out_iters / h_out_relres (||r||/||b|| from the recurrence) may be NULL.
 */
int myc_pcg_solve(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                  const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                  const double* d_rhs, const double* d_dinv, const double* d_binv, int precond,
                  double reg, double rtol, double atol, int64_t maxit, double* d_x,
                  int64_t* h_out_iters, double* h_out_relres, void* stream);

/* ||b - A x||2 / ||b||2 with A as above, recomputed from scratch (reported with every solve). */
int myc_true_residual(myc_ctx* ctx, int64_t n_rows, int64_t n_cols_global, int64_t row_offset,
                      const int32_t* d_row_ptr, const int32_t* d_col_idx, const double* d_val,
                      const double* d_rhs, const double* d_dinv, double reg, const double* d_x,
                      double* h_out_relres, void* stream);

/* U = x on free rows, prescribed value on known rows (src/fea_solver.py:131-133).
 * d_out_U is written at [row_offset, row_offset+n_rows) of a global-length vector. */
int myc_merge_solution(myc_ctx* ctx, int64_t n_rows, int64_t row_offset, const double* d_x,
                       const double* d_dinv, const double* d_ubc, double* d_out_U, void* stream);

/* sum_i d_v[d_idx[i]] in a fixed order (reaction sum over the top grip,
 * src/fea_solver.py:263-264).  idx are int64 positions into d_v. */
int myc_gather_sum(myc_ctx* ctx, const double* d_v, const int64_t* d_idx, int64_t n,
                   double* h_out_sum, void* stream);

/* ------------------------------------------------------------------------------------------
 * K7  strain / stress / failure.  Replaces the iterrows loop src/fea_solver.py:269-284
 *     (C++ twin src/fea_petsc.cpp:386-406; like the Python path, L is NOT clamped here).
 * d_U is the global displacement vector.  d_active is updated in place (|strain| > max_strain
 * switches the element off); d_out_stress gets E*strain (0 for inactive elements).
 * h_out_n_active (may be NULL) receives the number of active elements after the update.
 */
int myc_strain_update(myc_ctx* ctx, const double* d_coords, const int32_t* d_n1,
                      const int32_t* d_n2, int64_t n_elem, const double* d_U, double E,
                      double max_strain, uint8_t* d_active, double* d_out_stress,
                      int64_t* h_out_n_active, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU).  The matrix is row-partitioned by contiguous node ranges
 * (PETSc MPIAIJ row blocks: MatSetSizes(..PETSC_DECIDE..), src/fea_petsc_parallel.cpp:236).
 * myc_dist_init loads NCCL (h_nccl_path: path of libnccl.so.2, or NULL for the default search
 * path) and joins the communicator described by the 128-byte unique id created by
 * myc_dist_unique_id on rank 0 and distributed by the host (torch.distributed broadcast).
 * myc_dist_set_plan installs the partition of the mesh about to be solved (may be called any
 * number of times, no communication): h_node_offsets = world+1 node offsets; for THIS rank and
 * each peer q, [need_lo,need_hi) is the half-open global node range of q's nodes whose values
 * this rank reads and [give_lo,give_hi) the range of this rank's nodes that q reads (the
 * transpose, exchanged by the host); lo == hi means nothing.
 */
int myc_dist_unique_id(const char* h_nccl_path, uint8_t* h_out_id128);
int myc_dist_init(myc_ctx* ctx, const char* h_nccl_path, const uint8_t* h_id128, int rank, int world);
int myc_dist_set_plan(myc_ctx* ctx, const int64_t* h_node_offsets, const int64_t* h_need_lo,
                      const int64_t* h_need_hi, const int64_t* h_give_lo, const int64_t* h_give_hi);
/* NVLink peer-memory path of myc_pcg_solve (Jacobi): every rank allocates one IPC-shareable buffer
 * (gathered vector of n_cols_capacity doubles + flag/slot block), the host exchanges the 64-byte
 * cudaIpcMemHandle_t of all ranks (h_handles = world x 64 bytes, rank order) and every rank opens
 * its peers' buffers.  After that myc_pcg_solve runs as ONE persistent kernel per GPU: halo
 * values are stored straight into the neighbours' buffers and dot products are exchanged through
 * peer-written slots -- no NCCL call inside the iteration loop (csrc/pcg_fused.cu).  All three
 * calls are collective; myc_dist_peer_disable makes every rank use the NCCL loop again (call it
 * on all ranks if any rank failed to open a handle).  Growing the buffer later: every rank first calls
 * myc_dist_release_peers, then a rank barrier, then myc_dist_peer_alloc again (importers unmap before exporters free). */
int myc_dist_peer_alloc(myc_ctx* ctx, int64_t n_cols_capacity, uint8_t* h_out_handle64);
int myc_dist_peer_open(myc_ctx* ctx, const uint8_t* h_handles);
int myc_dist_peer_disable(myc_ctx* ctx);

/* Refresh the halo entries of a global-length vector from their owners (VecScatter of
 * MatMult, src/fea_petsc_parallel.cpp:402).  Collective. */
int myc_halo_exchange(myc_ctx* ctx, double* d_x_global, void* stream);
/* Sum h_inout[0..n) over ranks (n <= 8).  Collective; synchronises the stream. */
int myc_allreduce_sum(myc_ctx* ctx, double* h_inout, int n, void* stream);
/* Gather the owned slices of a global-length vector so that every rank holds all of it
 * (VecScatterCreateToZero + MPI_Bcast, src/fea_petsc_parallel.cpp:374-391).  Collective. */
int myc_allgather_owned(myc_ctx* ctx, double* d_x_global, void* stream);
/* Orderly shutdown of a multi-GPU job, step 1 of 2 (all ranks; put a rank barrier between this call and myc_destroy):
 * unmaps every peer's memory this context had mapped for the persistent solver kernels.  CUDA IPC requires the
 * importers to unmap before the exporter frees. */
int myc_dist_release_peers(myc_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * End-to-end load case on HOST buffers (what a caller without torch uses; bench.py "e2e"):
 * H2D of the mesh and BCs, assembly, Dirichlet elimination, PCG, reactions, D2H of U.
 * Equivalent to  K = assemble_global_stiffness(..); U = solve_system(K, known_dofs, known_vals);
 * F = K @ U  (src/fea_solver.py:220-257).  Single GPU.  h_out_U: n_dof doubles.  h_react_idx
 * (may be NULL/0): DOF indices whose reactions are summed into h_out_force.  precond = MYC_PC_AMG builds the
 * multigrid hierarchy inside the call (and uses MYC_PC_BLOCK6 where it is not applicable).  Host buffers may
 * be pageable or pinned (pinned: the copies run at full PCIe speed).  h_out_ms_assemble covers H2D + assembly,
 * h_out_ms_solve everything after it (Dirichlet, preconditioner setup, PCG, reactions, D2H).
 */
int myc_load_case_host(myc_ctx* ctx, const double* h_coords, const int32_t* h_n1, const int32_t* h_n2,
                       const uint8_t* h_active, int64_t n_elem, int64_t n_nodes, double E, double A,
                       double I, const int64_t* h_known_dofs, const double* h_known_vals,
                       int64_t n_known, double reg, int precond, double rtol, int64_t maxit,
                       const int64_t* h_react_idx, int64_t n_react, double* h_out_U,
                       double* h_out_force, int64_t* h_out_iters, double* h_out_relres,
                       int64_t* h_out_nnz, double* h_out_ms_assemble, double* h_out_ms_solve);

#ifdef __cplusplus
}
#endif
#endif /* MYCELIUM_FEA_H */
