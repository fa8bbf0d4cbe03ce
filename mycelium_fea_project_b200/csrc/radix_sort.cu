// Stable LSD radix sort of (uint64 key, uint32 value) pairs, 8 bits per pass, written for the
// assembly's directed node-pair stream (key = src_local << dst_bits | dst, value = element id).
// Three kernels per pass: per-tile digit histogram -> exclusive scan of the digit-major table
// (scan.cu) -> stable scatter.  Stability inside a tile comes from ranking items in
// (warp, round, lane) order with __match_any_sync; there are no floating-point operations and no
// order-dependent atomics, so the permutation is deterministic.
#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                       // rounds per warp
constexpr int RS_WARP_CHUNK = 32 * RS_ITEMS;       // 512 consecutive items per warp
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;     // 4096 items per block
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int64_t n_tiles,
               int32_t* __restrict__ table) {
  __shared__ int hist[RS_RADIX];
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = tile * RS_TILE + (int64_t)(threadIdx.x >> 5) * RS_WARP_CHUNK + (threadIdx.x & 31);
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const int64_t idx = base + r * 32;
      if (idx < n) atomicAdd(&hist[(int)((keys[idx] >> shift) & (RS_RADIX - 1))], 1);
    }
    __syncthreads();
    table[(int64_t)threadIdx.x * n_tiles + tile] = hist[threadIdx.x];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                  int shift, int64_t n_tiles, const int32_t* __restrict__ table_scanned) {
  __shared__ int wc[RS_WARPS][RS_RADIX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) wc[w][threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = tile * RS_TILE + (int64_t)warp * RS_WARP_CHUNK + lane;
    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS];
    int rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const int64_t idx = base + r * 32;
      const bool valid = idx < n;
      key[r] = valid ? keys_in[idx] : 0;
      val[r] = valid ? vals_in[idx] : 0;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const bool valid = (base + r * 32) < n;
      const int digit = (int)((key[r] >> shift) & (RS_RADIX - 1));
      const unsigned peers = __match_any_sync(0xffffffffu, valid ? digit : RS_RADIX + lane);
      int c = 0;
      if (valid) c = wc[warp][digit];
      __syncwarp();
      rank[r] = c + __popc(peers & lt_mask);
      if (valid && (peers & lt_mask) == 0) wc[warp][digit] = c + __popc(peers);   // group leader
      __syncwarp();
    }
    __syncthreads();
    {  // exclusive prefix over warps for digit = threadIdx.x, plus this tile's global offset
      int run = table_scanned[(int64_t)threadIdx.x * n_tiles + tile];
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) {
        const int t = wc[w][threadIdx.x];
        wc[w][threadIdx.x] = run;
        run += t;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      if ((base + r * 32) < n) {
        const int digit = (int)((key[r] >> shift) & (RS_RADIX - 1));
        const int64_t pos = (int64_t)wc[warp][digit] + rank[r];
        keys_out[pos] = key[r];
        vals_out[pos] = val[r];
      }
    }
    __syncthreads();
  }
}

}  // namespace

int myc_radix_sort_pairs(myc_ctx* ctx, int64_t n, int start_bit, int key_bits, int* out_buf, cudaStream_t st) {
  // input pairs are in sort_keys[0] / sort_vals[0]; both ping-pong buffers must hold n items.
  // Sorts on key bits [start_bit, key_bits); stable, so the order on the lower bits is the input order.
  *out_buf = 0;
  if (n <= 1 || key_bits <= start_bit) return MYC_OK;
  if (n >= (int64_t)1 << 31) MYC_FAIL(ctx, MYC_ERR_CAPACITY, "radix sort: %lld items exceed int32 offsets", (long long)n);
  const int64_t n_tiles = ceil_div64(n, RS_TILE);
  MYC_TRY(myc_ensure(ctx, ctx->sort_table, (size_t)(n_tiles * RS_RADIX + 1) * sizeof(int32_t)));
  int32_t* table = (int32_t*)ctx->sort_table.p;
  const int grid = grid_for(ctx, n_tiles, 4);
  int cur = 0;
  for (int shift = start_bit; shift < key_bits; shift += 8) {
    const uint64_t* kin = (const uint64_t*)ctx->sort_keys[cur].p;
    const uint32_t* vin = (const uint32_t*)ctx->sort_vals[cur].p;
    uint64_t* kout = (uint64_t*)ctx->sort_keys[cur ^ 1].p;
    uint32_t* vout = (uint32_t*)ctx->sort_vals[cur ^ 1].p;
    rs_hist_kernel<<<grid, RS_THREADS, 0, st>>>(kin, n, shift, n_tiles, table);
    MYC_LAUNCHED(ctx);
    MYC_TRY(myc_exclusive_scan_i32(ctx, table, table, n_tiles * RS_RADIX, false, nullptr, st));
    rs_scatter_kernel<<<grid, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, n_tiles, table);
    MYC_LAUNCHED(ctx);
    cur ^= 1;
  }
  *out_buf = cur;
  return MYC_OK;
}
