// Exact order statistics on the device by radix selection: the value of a given rank among n floats, four 8-bit passes over the
// order-preserving integer image of the floats with a shared-memory histogram.  One CTA; every thread of it calls select_rank and
// receives the result.  Used by estimate_channel_noise (noise.cu: two medians) and estimate_white_balance (pointwise.cu: the two
// ranks torch.quantile interpolates between).
#pragma once

#include "tdb_common.cuh"

namespace tdb {
namespace sel {

__device__ __forceinline__ uint32_t ordered_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// value of rank `rank` (0-based, ascending) among { value(i) : i < n, keep(i) }; hist = 256 words, pick = 2 words of shared memory.
// rank must be smaller than the number of kept items.
template <class V, class P>
__device__ float select_rank(int64_t n, int64_t rank, V value, P keep, uint32_t *hist, uint32_t *pick) {
  uint32_t prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      if (!keep(i)) continue;
      const uint32_t k = ordered_key(value(i));
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t acc = 0, b = 0;
      for (; b < 255; b++) {
        if (acc + hist[b] > rank) break;
        acc += hist[b];
      }
      pick[0] = b, pick[1] = acc;
    }
    __syncthreads();
    prefix |= pick[0] << shift, mask |= 255u << shift;
    rank -= pick[1];
    __syncthreads();
  }
  return key_value(prefix);
}

}  // namespace sel
}  // namespace tdb
