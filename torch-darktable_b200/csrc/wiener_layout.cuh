// Layout of the Wiener scratch buffer (tdb_wiener_scratch_bytes), shared with the fused frame pipeline:
// [64 words: job counters of the K = 32 kernels][accumulator: H*W*C floats][one extra plane: the log-luminance composite]
#pragma once

#include <stddef.h>

namespace tdb {

struct WienerScratch {
  unsigned int *counters;
  float *acc, *lum;
};
inline WienerScratch wiener_scratch(void *scratch, int width, int height, int channels) {
  float *base = static_cast<float *>(scratch);
  return WienerScratch{reinterpret_cast<unsigned int *>(base), base + 64, base + 64 + (size_t)width * height * channels};
}

}  // namespace tdb
