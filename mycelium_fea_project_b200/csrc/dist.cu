// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.  The matrix is
// row-partitioned by contiguous node ranges (PETSc MPIAIJ row blocks,
// src/fea_petsc_parallel.cpp:236); vectors that are gathered through column indices (the CG
// search direction, the prescribed-displacement vector, U) are stored at GLOBAL length on every
// rank (a 134 M-DOF vector is 1 GB of 180 GB), so CSR column indices stay global, local rows are
// verbatim slices of the global matrix, and a halo refresh is a plain ncclSend/ncclRecv of
// contiguous DOF ranges straight between the two p buffers -- no pack/unpack kernels, no index
// translation.  Dot products are in-place ncclAllReduce on the device-resident scalars.
//
// NCCL is bound at run time with dlopen (the torch-bundled libnccl.so.2), so the library loads
// and the single-GPU path runs on machines without NCCL.
#include <dlfcn.h>
#include <stdlib.h>

#include "common.cuh"

// ---- minimal NCCL ABI (stable since NCCL 2.x) -----------------------------------------------
typedef struct { char internal[128]; } myc_ncclUniqueId;
typedef void* myc_ncclComm_t;
enum { MYC_NCCL_SUM = 0 };
enum { MYC_NCCL_UINT8 = 1, MYC_NCCL_INT64 = 4, MYC_NCCL_FLOAT64 = 8 };

struct NcclApi {
  void* handle;
  int (*GetUniqueId)(myc_ncclUniqueId*);
  int (*CommInitRank)(myc_ncclComm_t*, int, myc_ncclUniqueId, int);
  int (*CommDestroy)(myc_ncclComm_t);
  const char* (*GetErrorString)(int);
  int (*AllReduce)(const void*, void*, size_t, int, int, myc_ncclComm_t, cudaStream_t);
  int (*AllGather)(const void*, void*, size_t, int, myc_ncclComm_t, cudaStream_t);
  int (*Send)(const void*, size_t, int, int, myc_ncclComm_t, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, myc_ncclComm_t, cudaStream_t);
  int (*GroupStart)(void);
  int (*GroupEnd)(void);
};

static NcclApi* g_nccl = nullptr;
static char g_nccl_err[256] = {0};

static NcclApi* load_nccl(const char* path) {
  if (g_nccl) return g_nccl;
  const char* cands[3] = {path, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (int i = 0; i < 3 && !h; ++i)
    if (cands[i] && cands[i][0]) h = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    snprintf(g_nccl_err, sizeof(g_nccl_err), "dlopen(libnccl.so.2) failed: %s", dlerror());
    return nullptr;
  }
  NcclApi* a = (NcclApi*)calloc(1, sizeof(NcclApi));
  a->handle = h;
#define MYC_SYM(field, name)                                                         \
  *(void**)(&a->field) = dlsym(h, name);                                             \
  if (!a->field) {                                                                   \
    snprintf(g_nccl_err, sizeof(g_nccl_err), "NCCL symbol %s not found", name);      \
    free(a);                                                                         \
    return nullptr;                                                                  \
  }
  MYC_SYM(GetUniqueId, "ncclGetUniqueId")
  MYC_SYM(CommInitRank, "ncclCommInitRank")
  MYC_SYM(CommDestroy, "ncclCommDestroy")
  MYC_SYM(GetErrorString, "ncclGetErrorString")
  MYC_SYM(AllReduce, "ncclAllReduce")
  MYC_SYM(AllGather, "ncclAllGather")
  MYC_SYM(Send, "ncclSend")
  MYC_SYM(Recv, "ncclRecv")
  MYC_SYM(GroupStart, "ncclGroupStart")
  MYC_SYM(GroupEnd, "ncclGroupEnd")
#undef MYC_SYM
  g_nccl = a;
  return a;
}

#define MYC_NCCL(ctx, call)                                                                \
  do {                                                                                     \
    int r__ = (call);                                                                      \
    if (r__ != 0)                                                                          \
      MYC_FAIL(ctx, MYC_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
               (ctx)->nccl->GetErrorString(r__));                                          \
  } while (0)

extern "C" int myc_dist_unique_id(const char* h_nccl_path, uint8_t* h_out_id128) {
  if (!h_out_id128) return MYC_ERR_BAD_ARG;
  NcclApi* a = load_nccl(h_nccl_path);
  if (!a) return MYC_ERR_NCCL;
  myc_ncclUniqueId id;
  if (a->GetUniqueId(&id) != 0) return MYC_ERR_NCCL;
  memcpy(h_out_id128, id.internal, 128);
  return MYC_OK;
}

extern "C" int myc_dist_init(myc_ctx* ctx, const char* h_nccl_path, const uint8_t* h_id128, int rank,
                             int world) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (!h_id128 || world < 1 || rank < 0 || rank >= world) MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "dist_init: bad argument");
  if (ctx->comm) MYC_FAIL(ctx, MYC_ERR_STATE, "dist_init: context already has a communicator");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->nccl = load_nccl(h_nccl_path);
  if (!ctx->nccl) MYC_FAIL(ctx, MYC_ERR_NCCL, "%s", g_nccl_err);
  myc_ncclUniqueId id;
  memcpy(id.internal, h_id128, 128);
  myc_ncclComm_t comm = nullptr;
  MYC_NCCL(ctx, ctx->nccl->CommInitRank(&comm, world, id, rank));
  ctx->comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  ctx->node_offsets = (int64_t*)calloc(world + 1, sizeof(int64_t));
  ctx->recv_from = (PeerRange*)calloc(world, sizeof(PeerRange));
  ctx->send_to = (PeerRange*)calloc(world, sizeof(PeerRange));
  return MYC_OK;
}

extern "C" int myc_dist_set_plan(myc_ctx* ctx, const int64_t* h_node_offsets, const int64_t* h_need_lo,
                                 const int64_t* h_need_hi, const int64_t* h_give_lo, const int64_t* h_give_hi) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (!h_node_offsets || !h_need_lo || !h_need_hi || !h_give_lo || !h_give_hi)
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "dist_set_plan: null pointer");
  if (ctx->world <= 1) return MYC_OK;
  if (!ctx->comm) MYC_FAIL(ctx, MYC_ERR_STATE, "dist_set_plan: call myc_dist_init first");
  const int world = ctx->world, rank = ctx->rank;
  for (int q = 0; q <= world; ++q)
    if (h_node_offsets[q] < 0 || (q > 0 && h_node_offsets[q] < h_node_offsets[q - 1]))
      MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "dist_set_plan: node offsets must be non-decreasing");
  memcpy(ctx->node_offsets, h_node_offsets, sizeof(int64_t) * (world + 1));
  for (int q = 0; q < world; ++q) {
    int64_t nlo = h_need_lo[q], nhi = h_need_hi[q], glo = h_give_lo[q], ghi = h_give_hi[q];
    if (q == rank || nhi <= nlo) nlo = nhi = 0;
    if (q == rank || ghi <= glo) glo = ghi = 0;
    if (nhi > nlo && (nlo < h_node_offsets[q] || nhi > h_node_offsets[q + 1]))
      MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "dist_set_plan: needed range [%lld,%lld) is not owned by rank %d", (long long)nlo, (long long)nhi, q);
    if (ghi > glo && (glo < h_node_offsets[rank] || ghi > h_node_offsets[rank + 1]))
      MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "dist_set_plan: given range [%lld,%lld) is not owned by this rank", (long long)glo, (long long)ghi);
    ctx->recv_from[q].lo = 3 * nlo;
    ctx->recv_from[q].hi = 3 * nhi;
    ctx->send_to[q].lo = 3 * glo;
    ctx->send_to[q].hi = 3 * ghi;
  }
  return MYC_OK;
}

extern "C" int myc_dist_peer_disable(myc_ctx* ctx) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  ctx->peer_ok = false;
  return MYC_OK;
}

void myc_amg_close_imports(myc_ctx* ctx);      // amg_setup.cu

// Orderly multi-process shutdown, step 1 of 2: close every mapping of a PEER's memory (both solver kernels').  The
// caller puts a rank barrier between this call and myc_destroy / re-allocation, which free the exported buffers
// (CUDA IPC: importers unmap before the exporter frees).  Afterwards the peer-memory solvers are unavailable until
// the buffers are set up again.
extern "C" int myc_dist_release_peers(myc_ctx* ctx) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int q = 0; q < MYC_MAX_WORLD; ++q) {
    if (ctx->peer_base[q] && q != ctx->rank) cudaIpcCloseMemHandle(ctx->peer_base[q]);
    if (q != ctx->rank) ctx->peer_base[q] = nullptr;
  }
  ctx->peer_ok = false;
  myc_amg_close_imports(ctx);
  return MYC_OK;
}

int myc_dist_destroy(myc_ctx* ctx) {
  for (int q = 0; q < MYC_MAX_WORLD; ++q) {
    if (ctx->peer_base[q] && q != ctx->rank) cudaIpcCloseMemHandle(ctx->peer_base[q]);
    ctx->peer_base[q] = nullptr;
  }
  if (ctx->peer_own) cudaFree(ctx->peer_own);
  ctx->peer_own = nullptr;
  ctx->peer_ok = false;
  if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy((myc_ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  free(ctx->node_offsets);
  free(ctx->recv_from);
  free(ctx->send_to);
  ctx->node_offsets = nullptr;
  ctx->recv_from = ctx->send_to = nullptr;
  return MYC_OK;
}

int myc_dist_halo(myc_ctx* ctx, double* d_x_global, cudaStream_t st) {
  if (ctx->world <= 1) return MYC_OK;
  NcclApi* a = ctx->nccl;
  myc_ncclComm_t comm = (myc_ncclComm_t)ctx->comm;
  MYC_NCCL(ctx, a->GroupStart());
  for (int q = 0; q < ctx->world; ++q) {
    if (q == ctx->rank) continue;
    const PeerRange& s = ctx->send_to[q];
    const PeerRange& r = ctx->recv_from[q];
    if (s.hi > s.lo) MYC_NCCL(ctx, a->Send(d_x_global + s.lo, (size_t)(s.hi - s.lo), MYC_NCCL_FLOAT64, q, comm, st));
    if (r.hi > r.lo) MYC_NCCL(ctx, a->Recv(d_x_global + r.lo, (size_t)(r.hi - r.lo), MYC_NCCL_FLOAT64, q, comm, st));
  }
  MYC_NCCL(ctx, a->GroupEnd());
  return MYC_OK;
}

int myc_dist_allreduce_dev(myc_ctx* ctx, double* d_buf, int n, cudaStream_t st) {
  if (ctx->world <= 1) return MYC_OK;
  MYC_NCCL(ctx, ctx->nccl->AllReduce(d_buf, d_buf, (size_t)n, MYC_NCCL_FLOAT64, MYC_NCCL_SUM,
                                     (myc_ncclComm_t)ctx->comm, st));
  return MYC_OK;
}

// All-gather of contiguous, variably sized sections of one device array: rank q owns bytes
// [h_off_bytes[q], h_off_bytes[q+1]); on return every rank holds all sections (in place).
int myc_dist_allgatherv(myc_ctx* ctx, void* d_buf, const int64_t* h_off_bytes, cudaStream_t st) {
  if (ctx->world <= 1) return MYC_OK;
  NcclApi* a = ctx->nccl;
  myc_ncclComm_t comm = (myc_ncclComm_t)ctx->comm;
  char* base = (char*)d_buf;
  const int64_t my_lo = h_off_bytes[ctx->rank], my_hi = h_off_bytes[ctx->rank + 1];
  MYC_NCCL(ctx, a->GroupStart());
  for (int q = 0; q < ctx->world; ++q) {
    if (q == ctx->rank) continue;
    const int64_t lo = h_off_bytes[q], hi = h_off_bytes[q + 1];
    if (my_hi > my_lo) MYC_NCCL(ctx, a->Send(base + my_lo, (size_t)(my_hi - my_lo), MYC_NCCL_UINT8, q, comm, st));
    if (hi > lo) MYC_NCCL(ctx, a->Recv(base + lo, (size_t)(hi - lo), MYC_NCCL_UINT8, q, comm, st));
  }
  MYC_NCCL(ctx, a->GroupEnd());
  return MYC_OK;
}

// All-gather of k host int64 per rank (k <= 32): h_all[q * k + j] = value j of rank q.  Synchronises the stream.
int myc_dist_allgather_host_i64(myc_ctx* ctx, const int64_t* h_mine, int k, int64_t* h_all, cudaStream_t st) {
  if (k < 1 || k > 32) MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "allgather_host: bad count");
  if (ctx->world <= 1) {
    memcpy(h_all, h_mine, sizeof(int64_t) * k);
    return MYC_OK;
  }
  MYC_TRY(myc_ensure(ctx, ctx->xchg, (size_t)(ctx->world + 1) * 32 * sizeof(int64_t)));
  int64_t* d_mine = (int64_t*)ctx->xchg.p;
  int64_t* d_all = d_mine + 32;
  MYC_CUDA(ctx, cudaMemcpyAsync(d_mine, h_mine, sizeof(int64_t) * k, cudaMemcpyHostToDevice, st));
  MYC_NCCL(ctx, ctx->nccl->AllGather(d_mine, d_all, (size_t)k, MYC_NCCL_INT64, (myc_ncclComm_t)ctx->comm, st));
  MYC_CUDA(ctx, cudaMemcpyAsync(h_all, d_all, sizeof(int64_t) * k * ctx->world, cudaMemcpyDeviceToHost, st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  return MYC_OK;
}

// All-gather of 64 opaque bytes per rank (IPC memory handles).  Synchronises the stream.
int myc_dist_allgather_host_64b(myc_ctx* ctx, const void* h_mine64, void* h_all, cudaStream_t st) {
  int64_t mine[8], all[8 * MYC_MAX_WORLD];
  memcpy(mine, h_mine64, 64);
  MYC_TRY(myc_dist_allgather_host_i64(ctx, mine, 8, all, st));
  memcpy(h_all, all, (size_t)64 * ctx->world);
  return MYC_OK;
}

extern "C" int myc_halo_exchange(myc_ctx* ctx, double* d_x_global, void* stream) {
  if (!ctx || !d_x_global) return MYC_ERR_BAD_ARG;
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  return myc_dist_halo(ctx, d_x_global, (cudaStream_t)stream);
}

extern "C" int myc_allreduce_sum(myc_ctx* ctx, double* h_inout, int n, void* stream) {
  if (!ctx || !h_inout || n < 0 || n > 8) return MYC_ERR_BAD_ARG;
  if (ctx->world <= 1 || n == 0) return MYC_OK;
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  MYC_TRY(myc_ensure(ctx, ctx->misc, 256));
  double* d = (double*)((char*)ctx->misc.p + 128);
  double* h = (double*)ctx->h_pinned;
  memcpy(h, h_inout, sizeof(double) * n);
  MYC_CUDA(ctx, cudaMemcpyAsync(d, h, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  MYC_TRY(myc_dist_allreduce_dev(ctx, d, n, st));
  MYC_CUDA(ctx, cudaMemcpyAsync(h, d, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  MYC_CUDA(ctx, cudaStreamSynchronize(st));
  memcpy(h_inout, h, sizeof(double) * n);
  return MYC_OK;
}

extern "C" int myc_allgather_owned(myc_ctx* ctx, double* d_x_global, void* stream) {
  if (!ctx || !d_x_global) return MYC_ERR_BAD_ARG;
  if (ctx->world <= 1) return MYC_OK;
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  NcclApi* a = ctx->nccl;
  myc_ncclComm_t comm = (myc_ncclComm_t)ctx->comm;
  const int64_t my_lo = 3 * ctx->node_offsets[ctx->rank], my_hi = 3 * ctx->node_offsets[ctx->rank + 1];
  MYC_NCCL(ctx, a->GroupStart());
  for (int q = 0; q < ctx->world; ++q) {
    if (q == ctx->rank) continue;
    const int64_t lo = 3 * ctx->node_offsets[q], hi = 3 * ctx->node_offsets[q + 1];
    if (my_hi > my_lo) MYC_NCCL(ctx, a->Send(d_x_global + my_lo, (size_t)(my_hi - my_lo), MYC_NCCL_FLOAT64, q, comm, st));
    if (hi > lo) MYC_NCCL(ctx, a->Recv(d_x_global + lo, (size_t)(hi - lo), MYC_NCCL_FLOAT64, q, comm, st));
  }
  MYC_NCCL(ctx, a->GroupEnd());
  return MYC_OK;
}
