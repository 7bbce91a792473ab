// Device-wide inclusive/exclusive scan (reduce-then-scan, three launches) for any associative op.
// Hand-written so the pileup builder carries no library dependency; tiles of 2048 elements,
// 256 threads x 8 items, warp-shuffle scans inside the tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace biodb {

struct OpAdd {
  template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct OpMax {
  template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; }
};

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T shfl_up_t(T v, int d) {
  if constexpr (sizeof(T) == 8) {
    unsigned long long x = (unsigned long long)v;
    x = __shfl_up_sync(0xffffffffu, x, d);
    return (T)x;
  } else {
    return (T)__shfl_up_sync(0xffffffffu, v, d);
  }
}

// block-wide inclusive scan of one value per thread; returns the inclusive value, *total = block aggregate
template <typename T, typename Op>
__device__ __forceinline__ T block_scan_incl(T v, Op op, T identity, T* total) {
  __shared__ T warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = shfl_up_t(v, d);
    if (lane >= d) v = op(o, v);
  }
  if (lane == 31) warp_tot[wid] = v;
  __syncthreads();
  T pre = identity;
  T tot = identity;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    T x = warp_tot[w];
    if (w < wid) pre = op(pre, x);
    tot = op(tot, x);
  }
  __syncthreads();
  *total = tot;
  return op(pre, v);
}

template <typename TIn, typename T, typename Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const TIn* __restrict__ in, uint64_t n, T* __restrict__ tile_agg,
                                                                   Op op, T identity) {
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
  T acc = identity;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) acc = op(acc, (T)in[i]);
  }
  T tot;
  block_scan_incl(acc, op, identity, &tot);
  if (threadIdx.x == 0) tile_agg[blockIdx.x] = tot;
}

// exclusive scan of the tile aggregates, single block; also writes the grand total to tile_agg[n_tiles]
template <typename T, typename Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(T* tile_agg, uint32_t n_tiles, Op op, T identity) {
  __shared__ T carry;
  if (threadIdx.x == 0) carry = identity;
  __syncthreads();
  for (uint32_t base = 0; base < n_tiles; base += SCAN_THREADS) {
    uint32_t i = base + threadIdx.x;
    T v = i < n_tiles ? tile_agg[i] : identity;
    T tot;
    T incl = block_scan_incl(v, op, identity, &tot);
    T c = carry;
    // exclusive = carry (+) (inclusive without own value): recompute via shuffle-free trick
    T excl_in_block = identity;
    {
      __shared__ T tmp[SCAN_THREADS];
      tmp[threadIdx.x] = incl;
      __syncthreads();
      if (threadIdx.x > 0) excl_in_block = tmp[threadIdx.x - 1];
      __syncthreads();
    }
    if (i < n_tiles) tile_agg[i] = op(c, excl_in_block);
    __syncthreads();
    if (threadIdx.x == 0) carry = op(c, tot);
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_agg[n_tiles] = carry;
}

template <typename TIn, typename T, typename Op, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const TIn* __restrict__ in, T* __restrict__ out, uint64_t n,
                                                                  const T* __restrict__ tile_agg, Op op, T identity) {
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  T v[SCAN_ITEMS];
  T acc = identity;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    uint64_t i = base + k;
    v[k] = i < n ? (T)in[i] : identity;
    acc = op(acc, v[k]);
  }
  T tot;
  T incl = block_scan_incl(acc, op, identity, &tot);
  // exclusive prefix of this thread = tile prefix (+) inclusive-of-previous-thread
  __shared__ T tmp[SCAN_THREADS];
  tmp[threadIdx.x] = incl;
  __syncthreads();
  T pre = tile_agg[blockIdx.x];
  if (threadIdx.x > 0) pre = op(pre, tmp[threadIdx.x - 1]);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    uint64_t i = base + k;
    T nxt = op(pre, v[k]);
    if (i < n) out[i] = INCLUSIVE ? nxt : pre;
    pre = nxt;
  }
}

inline size_t scan_temp_elems(uint64_t n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 2; }

// out may alias in.  tile_tmp needs scan_temp_elems(n) elements of T; after the call
// tile_tmp[n_tiles] holds the grand total (device side).
template <bool INCLUSIVE, typename TIn, typename T, typename Op>
inline void device_scan(const TIn* in, T* out, uint64_t n, T* tile_tmp, Op op, T identity, cudaStream_t st) {
  uint32_t n_tiles = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
  g_kernel_launches += n_tiles ? 3 : 1;
  if (n_tiles == 0) {
    scan_tiles_kernel<T, Op><<<1, SCAN_THREADS, 0, st>>>(tile_tmp, 0, op, identity);
    return;
  }
  scan_reduce_kernel<TIn, T, Op><<<n_tiles, SCAN_THREADS, 0, st>>>(in, n, tile_tmp, op, identity);
  scan_tiles_kernel<T, Op><<<1, SCAN_THREADS, 0, st>>>(tile_tmp, n_tiles, op, identity);
  scan_apply_kernel<TIn, T, Op, INCLUSIVE><<<n_tiles, SCAN_THREADS, 0, st>>>(in, out, n, tile_tmp, op, identity);
}

}  // namespace biodb
