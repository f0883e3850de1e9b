// scan.cuh -- exclusive prefix sums used by the index build, the query sort and
// the CSR offset computation.  Three launches: per-tile sums, one-block scan of
// the tile sums, per-tile rescan with the tile offset.  out[n] receives the total.
#pragma once
#include "common.cuh"

namespace rrtqx {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename TOut>
__device__ __forceinline__ TOut block_exclusive_scan(TOut v, TOut *smem /*>=33*/, TOut *total) {
  // warp scan then scan of warp sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TOut incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    TOut t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    TOut w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : (TOut)0;
    TOut wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      TOut t = __shfl_up_sync(FULL, wi, o);
      if (lane >= o) wi += t;
    }
    smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  TOut r = incl - v + smem[warp];
  if (total) *total = smem[32];
  __syncthreads();
  return r;
}

template <typename TIn, typename TOut>
__global__ void scan_tile_sums_kernel(const TIn *__restrict__ in, int64_t n, TOut *__restrict__ tile_sums) {
  __shared__ TOut sm[33];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  TOut s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += (TOut)in[i];
  }
  TOut tot;
  block_exclusive_scan<TOut>(s, sm, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

template <typename TOut>
__global__ void scan_sums_kernel(TOut *__restrict__ tile_sums, int64_t n_tiles) {
  __shared__ TOut sm[33];
  __shared__ TOut carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_tiles; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    TOut v = (i < n_tiles) ? tile_sums[i] : (TOut)0;
    TOut tot;
    TOut ex = block_exclusive_scan<TOut>(v, sm, &tot);
    TOut carry = carry_s;
    if (i < n_tiles) tile_sums[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[n_tiles] = carry_s;  // grand total
}

template <typename TIn, typename TOut>
__global__ void scan_apply_kernel(const TIn *__restrict__ in, int64_t n, const TOut *__restrict__ tile_sums,
                                  TOut *__restrict__ out) {
  __shared__ TOut sm[33];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  TOut v[SCAN_ITEMS];
  TOut s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    v[k] = (i < n) ? (TOut)in[i] : (TOut)0;
    s += v[k];
  }
  TOut ex = block_exclusive_scan<TOut>(s, sm, (TOut *)nullptr) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = tile_sums[gridDim.x];
}

// Exclusive scan of in[0..n) into out[0..n], out[n] = total.  `tmp` must hold
// at least n/SCAN_TILE + 2 elements.  in and out may alias only if TIn == TOut.
template <typename TIn, typename TOut>
void exclusive_scan(rrtqx_ctx *ctx, const TIn *in, int64_t n, TOut *out, DevBuf<TOut> &tmp) {
  if (n <= 0) {
    RQ_CUDA(cudaMemsetAsync(out, 0, sizeof(TOut), ctx->stream));
    return;
  }
  int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  tmp.ensure((size_t)tiles + 2, ctx->stream);
  scan_tile_sums_kernel<TIn, TOut><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, tmp.p);
  scan_sums_kernel<TOut><<<1, 1024, 0, ctx->stream>>>(tmp.p, tiles);
  scan_apply_kernel<TIn, TOut><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, tmp.p, out);
  post_launch(ctx, 3);
}

}  // namespace rrtqx
