// Stand-alone SpMV micro-benchmark for kernel-variant sweeps (compile with -DTM_CFG_* overrides):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I mycelium_fea_project_b200/csrc \
//        tools/spmv_bench.cu -o build/spmv_bench_X [-DTM_CFG_ROWS=16 ...]
//   build/spmv_bench_X /tmp/csr_2048.bin [reps]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#include "spmv_tma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "/tmp/csr_2048.bin";
  const int reps = argc > 2 ? atoi(argv[2]) : 20;
  FILE* f = fopen(path, "rb");
  if (!f) { printf("cannot open %s\n", path); return 1; }
  int64_t hdr[2];
  if (fread(hdr, 8, 2, f) != 2) return 1;
  const int64_t n = hdr[0], nnz = hdr[1];
  std::vector<int32_t> rp(n + 1), ci(nnz);
  std::vector<double> v(nnz), x(n);
  if (fread(rp.data(), 4, n + 1, f) != (size_t)(n + 1) || fread(ci.data(), 4, nnz, f) != (size_t)nnz ||
      fread(v.data(), 8, nnz, f) != (size_t)nnz) return 1;
  fclose(f);
  srand(1);
  for (auto& e : x) e = rand() / (double)RAND_MAX - 0.5;
  int32_t *d_rp, *d_ci; double *d_v, *d_x, *d_y, *d_y0; char* d_flush;
  CK(cudaMalloc(&d_rp, (n + 1) * 4)); CK(cudaMalloc(&d_ci, nnz * 4)); CK(cudaMalloc(&d_v, nnz * 8));
  CK(cudaMalloc(&d_x, n * 8)); CK(cudaMalloc(&d_y, n * 8)); CK(cudaMalloc(&d_y0, n * 8));
  CK(cudaMalloc(&d_flush, 256 << 20));
  CK(cudaMemcpy(d_rp, rp.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ci, ci.data(), nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_v, v.data(), nnz * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_x, x.data(), n * 8, cudaMemcpyHostToDevice));
  myc_ctx ctx;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  ctx.sm_count = prop.multiProcessorCount;
  // reference: plain CSR-stream kernel
  ctx.force_plain_spmv = true;
  EpiPlain e0{d_y0};
  myc_launch_spmv_epi<EpiPlain>(&ctx, n, d_rp, d_ci, d_v, d_x, e0, nullptr, nullptr, nullptr, nullptr, 0);
  CK(cudaDeviceSynchronize());
  const double bytes = 12.0 * nnz + 20.0 * n;
  for (int variant = 0; variant < 3; ++variant) {
    ctx.force_plain_spmv = variant == 0;
    ctx.csr_block3 = variant == 2;
    EpiPlain e1{d_y};
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    std::vector<float> ts;
    for (int r = 0; r < reps + 3; ++r) {
      CK(cudaMemsetAsync(d_flush, r, 256 << 20));
      CK(cudaEventRecord(a));
      int rc = myc_launch_spmv_epi<EpiPlain>(&ctx, n, d_rp, d_ci, d_v, d_x, e1, nullptr, nullptr, nullptr, nullptr, 0);
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      if (rc) { printf("launch failed: %s\n", ctx.err); return 1; }
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (r >= 3) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    std::vector<double> y(n), y0(n);
    CK(cudaMemcpy(y.data(), d_y, n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(y0.data(), d_y0, n * 8, cudaMemcpyDeviceToHost));
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) bad += y[i] != y0[i];
    const float med = ts[ts.size() / 2];
    double maxrel = 0.0;
    for (int64_t i = 0; i < n; ++i) { double d = fabs(y[i] - y0[i]); double sc = fabs(y0[i]) + 1e-300; if (d / sc > maxrel) maxrel = d / sc; }
    printf("%s : median %.4f ms  min %.4f ms  %.1f GB/s  bit-mismatches vs plain %lld  max rel diff %.2e\n",
           variant == 0 ? "plain CSR-stream   " : (variant == 1 ? "TMA generic (16/256)" : "TMA block3 (18/288) "), med, ts[0],
           bytes / med / 1e6, (long long)bad, maxrel);
  }
  return 0;
}
