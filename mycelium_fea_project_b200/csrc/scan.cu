// Exclusive prefix sums (reduce-then-scan, three kernels per level) used by the assembly:
// element -> edge offsets, node -> edge_start, node -> block_start, radix digit tables.
// Integer adds only, so the result is independent of scheduling.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t* total, int64_t* smem_warp) {
  // returns the exclusive prefix of v over the block's threads (thread order); *total = block sum
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int64_t w = lane < (SCAN_THREADS / 32) ? smem_warp[lane] : 0;
    int64_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < (SCAN_THREADS / 32)) smem_warp[lane] = winc - w;   // exclusive warp base
    if (lane == (SCAN_THREADS / 32) - 1) smem_warp[SCAN_THREADS / 32] = winc;
  }
  __syncthreads();
  *total = smem_warp[SCAN_THREADS / 32];
  return smem_warp[warp] + inc - v;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums_kernel(const T* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sums) {
  __shared__ int64_t sw[SCAN_THREADS / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) s += (int64_t)in[base + k];
  int64_t total;
  block_exclusive_scan(s, &total, sw);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const TIn* in, TOut* out, int64_t n,   // in may alias out
                  const int64_t* __restrict__ tile_base, bool write_total_at_end,
                  int64_t* __restrict__ d_total) {
  __shared__ int64_t sw[SCAN_THREADS / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t v[SCAN_ITEMS];
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? (int64_t)in[base + k] : 0;
    s += v[k];
  }
  int64_t total;
  int64_t pre = block_exclusive_scan(s, &total, sw) + (tile_base ? tile_base[blockIdx.x] : 0);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = (TOut)pre;
    pre += v[k];
  }
  // the thread owning the last element knows the grand total
  const int64_t last = n - 1;
  if (last >= base && last < base + SCAN_ITEMS) {
    if (write_total_at_end) out[n] = (TOut)pre;
    if (d_total) *d_total = pre;
  }
}

__global__ void scan_empty_kernel(int32_t* out, bool write_total_at_end, int64_t* d_total) {
  if (write_total_at_end) out[0] = 0;
  if (d_total) *d_total = 0;
}

// scans `n` int64 values in place; scratch must hold the tile sums of every further level
int scan_i64_inplace(myc_ctx* ctx, int64_t* d, int64_t n, int64_t* scratch, cudaStream_t st) {
  const int64_t tiles = ceil_div64(n, SCAN_TILE);
  if (tiles > 1) {
    scan_tile_sums_kernel<int64_t><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(d, n, scratch);
    MYC_LAUNCHED(ctx);
    MYC_TRY(scan_i64_inplace(ctx, scratch, tiles, scratch + tiles, st));
  }
  scan_apply_kernel<int64_t, int64_t><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(
      d, d, n, tiles > 1 ? scratch : nullptr, false, nullptr);
  MYC_LAUNCHED(ctx);
  return MYC_OK;
}

}  // namespace

int myc_exclusive_scan_i32(myc_ctx* ctx, const int32_t* d_in, int32_t* d_out, int64_t n,
                           bool write_total_at_end, int64_t* d_total, cudaStream_t st) {
  if (n < 0) MYC_FAIL(ctx, MYC_ERR_BAD_ARG, "scan: negative length");
  if (n == 0) {
    scan_empty_kernel<<<1, 1, 0, st>>>(d_out, write_total_at_end, d_total);
    MYC_LAUNCHED(ctx);
    return MYC_OK;
  }
  const int64_t tiles = ceil_div64(n, SCAN_TILE);
  // scratch for all levels: tiles + tiles/2048 + ... < tiles + tiles/1024 + 8
  size_t scratch_elems = (size_t)tiles + (size_t)tiles / 1024 + 64;
  MYC_TRY(myc_ensure(ctx, ctx->scan_tmp, scratch_elems * sizeof(int64_t)));
  int64_t* scratch = (int64_t*)ctx->scan_tmp.p;
  if (tiles > 1) {
    scan_tile_sums_kernel<int32_t><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(d_in, n, scratch);
    MYC_LAUNCHED(ctx);
    MYC_TRY(scan_i64_inplace(ctx, scratch, tiles, scratch + tiles, st));
  }
  scan_apply_kernel<int32_t, int32_t><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(
      d_in, d_out, n, tiles > 1 ? scratch : nullptr, write_total_at_end, d_total);
  MYC_LAUNCHED(ctx);
  return MYC_OK;
}
