// Device-resident statistics state of the fused frame pipeline (one per ImageProcessor), see include/tdb200.h "Fused frame pipeline".
//
// The reference computes the green-equilibration ratio, the joint bounds of an image set and its tone-mapping metrics with separate
// reduction kernels and reads them back on the host (postprocess.cu:362-366, image_processor.py:288-294).  Here the kernels that
// already touch the pixels leave per-CTA partials behind and the LAST CTA to finish (ticket counter) reduces them in a fixed
// order, merges them into the state of the image set and -- on the last frame of a set -- applies the moving average.  The
// tickets reset themselves, so the state only needs to be zero-filled once.
#pragma once

#include <cfloat>

#include "tdb_common.cuh"

namespace tdb {

constexpr int kMaxStatCtas = 148 * 16;

struct FrameState {
  unsigned int ticket[2];  // [0] bounds, [1] metrics
  float set_lo, set_hi;    // bounds of the image set so far
  float set_sums[6];       // metric sums of the image set so far (log-gray, gray, r, g, b, count)
  float pad[6];
  float partials[kMaxStatCtas * 6];
};

// true in every thread of the last CTA of the grid to get here; `slot` is a __shared__ word
__device__ __forceinline__ bool last_cta(unsigned int *ticket, unsigned int total, unsigned int *slot) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *slot = atomicAdd(ticket, 1u);
  __syncthreads();
  const bool last = *slot == total - 1;
  if (last) __threadfence();
  return last;
}

__device__ __forceinline__ float ema(const float *prev, int k, float value, float t) {
  if (!prev) return value;
  const float a = prev[k];
  return a + (value - a) * t;  // pipeline/util.py:4 lerp
}

}  // namespace tdb
