// Internal launch entry points shared between translation units.
#pragma once
#include "cfa_tile.cuh"

namespace tdb {
int launch_bilinear(const CfaSource &src, float *rgb, int width, int height, uint32_t filters, cudaStream_t s);
int launch_ppg(const CfaSource &src, float *rgb, int width, int height, uint32_t filters, float median_threshold, cudaStream_t s);
int launch_rcd(const CfaSource &src, float *rgb, int width, int height, uint32_t filters, cudaStream_t s);
}  // namespace tdb
