// Pipe-throughput probes: the denominators of the NON-HBM rooflines ("Wiener: report FLOP/s ... both").
//
// MEASURED_PEAKS.json holds a copy bandwidth and a tensor-core GEMM rate; neither bounds a kernel that is limited by the FP32 FMA
// pipe (the Wiener tile transforms) or by the MUFU unit (pow / exp / log of the tone curves, Lab conversions).  These two kernels
// keep nothing but that one pipe busy -- every resident warp issues a long run of independent FFMA (or MUFU.EX2) instructions on
// registers -- so bench.py can time them with CUDA events on the same GPU, in the same process, under the same clocks, and quote a
// kernel's achieved FLOP/s (or MUFU op/s) against a peak that was measured, not taken from a data sheet.
//   FP32: 128 lanes x 2 FLOP x 148 SMs x 1.965 GHz = 74.4 TFLOP/s nominal
//   MUFU:  16 lanes x 148 SMs x 1.965 GHz = 4.65 Tops/s nominal
#include "tdb_common.cuh"

namespace tdb {
namespace {

constexpr int kThreads = 256;
constexpr int kIlp = 16;  // independent dependency chains per thread: hides the 4-cycle FFMA latency at any occupancy

__global__ void __launch_bounds__(kThreads) fp32_probe_kernel(float *__restrict__ sink, int iters, float a, float b) {
  float v[kIlp];
#pragma unroll
  for (int k = 0; k < kIlp; k++) v[k] = (float)(threadIdx.x + k) * 1e-3f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < kIlp; k++) v[k] = fmaf(v[k], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < kIlp; k++) s += v[k];
  if (s == 123.456f) sink[0] = s;  // never true for the arguments used; keeps the chains alive
}

__global__ void __launch_bounds__(kThreads) mufu_probe_kernel(float *__restrict__ sink, int iters) {
  float v[kIlp];
#pragma unroll
  for (int k = 0; k < kIlp; k++) v[k] = (float)(threadIdx.x + k) * 1e-3f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < kIlp; k++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[k]));
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < kIlp; k++) s += v[k];
  if (s == 123.456f) sink[0] = s;
}

}  // namespace
}  // namespace tdb

using namespace tdb;

extern "C" {

// Launches the FP32 FMA probe; *flops = floating-point operations it executes (2 per FFMA).  sink: any device float.
int tdb_probe_fp32(float *sink, int iters, double *flops, tdb_stream_t stream) {
  TDB_REQUIRE(sink && iters > 0, "probe_fp32: bad argument");
  const int grid = kNumSMs * 8;
  fp32_probe_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(sink, iters, 0.999f, 1e-3f);
  if (flops) *flops = 2.0 * kIlp * (double)iters * kThreads * grid;
  return check_launch("probe_fp32");
}

// Launches the MUFU probe (ex2.approx); *ops = MUFU operations it executes.
int tdb_probe_mufu(float *sink, int iters, double *ops, tdb_stream_t stream) {
  TDB_REQUIRE(sink && iters > 0, "probe_mufu: bad argument");
  const int grid = kNumSMs * 8;
  mufu_probe_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(sink, iters);
  if (ops) *ops = (double)kIlp * iters * kThreads * grid;
  return check_launch("probe_mufu");
}

}  // extern "C"
