// K7: per-element axial strain / stress and failure switch-off, the device form of the
// iterrows loop src/fea_solver.py:269-284 (C++ twin src/fea_petsc.cpp:386-406).  One thread per
// element; U is gathered through the read-only path.  As in the Python path the length is NOT
// clamped (a zero-length active element gives nan strain, which never exceeds the threshold).
#include "common.cuh"
#include "spmv.cuh"

namespace {

// [host-test-begin strain_kernel]  (tests/test_kernel_logic_host.py compiles this text with g++)
constexpr int ST_THREADS = 256;

__global__ void __launch_bounds__(ST_THREADS)
strain_kernel(const double* __restrict__ coords, const int32_t* __restrict__ n1,
              const int32_t* __restrict__ n2, int64_t n_elem, const double* __restrict__ U, double E,
              double max_strain, uint8_t* __restrict__ active, double* __restrict__ stress,
              unsigned long long* __restrict__ n_active) {
  unsigned long long cnt = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem;
       e += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    if (active[e]) {
      const int64_t a = n1[e], b = n2[e];
      const double vx = __dsub_rn(coords[3 * b], coords[3 * a]);
      const double vy = __dsub_rn(coords[3 * b + 1], coords[3 * a + 1]);
      const double vz = __dsub_rn(coords[3 * b + 2], coords[3 * a + 2]);
      const double L = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
      const double nx = __ddiv_rn(vx, L), ny = __ddiv_rn(vy, L), nz = __ddiv_rn(vz, L);
      const double dx = __dsub_rn(U[3 * b], U[3 * a]);
      const double dy = __dsub_rn(U[3 * b + 1], U[3 * a + 1]);
      const double dz = __dsub_rn(U[3 * b + 2], U[3 * a + 2]);
      const double dot = __dadd_rn(__dadd_rn(__dmul_rn(nx, dx), __dmul_rn(ny, dy)), __dmul_rn(nz, dz));
      const double strain = __ddiv_rn(dot, L);
      s = __dmul_rn(E, strain);
      if (fabs(strain) > max_strain) active[e] = 0; else ++cnt;
    }
    stress[e] = s;
  }
  // integer count: order-independent
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_active, cnt);
}

// [host-test-end strain_kernel]

}  // namespace

extern "C" int myc_strain_update(myc_ctx* ctx, const double* d_coords, const int32_t* d_n1,
                                 const int32_t* d_n2, int64_t n_elem, const double* d_U, double E,
                                 double max_strain, uint8_t* d_active, double* d_out_stress,
                                 int64_t* h_out_n_active, void* stream) {
  if (!ctx) return MYC_ERR_BAD_ARG;
  if (n_elem < 0 || (n_elem > 0 && (!d_coords || !d_n1 || !d_n2 || !d_U || !d_active || !d_out_stress)))
    MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "strain_update: bad argument");
  MYC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  MYC_TRY(myc_ensure(ctx, ctx->misc, 256));
  unsigned long long* d_cnt = (unsigned long long*)((char*)ctx->misc.p + 192);
  MYC_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), st));
  if (n_elem > 0) {
    strain_kernel<<<grid_for(ctx, ceil_div64(n_elem, ST_THREADS), 8), ST_THREADS, 0, st>>>(
        d_coords, d_n1, d_n2, n_elem, d_U, E, max_strain, d_active, d_out_stress, d_cnt);
    MYC_LAUNCHED(ctx);
  }
  if (h_out_n_active) {
    unsigned long long* h = (unsigned long long*)ctx->h_pinned;
    MYC_CUDA(ctx, cudaMemcpyAsync(h, d_cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    MYC_CUDA(ctx, cudaStreamSynchronize(st));
    *h_out_n_active = (int64_t)*h;
  }
  return MYC_OK;
}
