// estimate_channel_noise (reference denoise.py:131-158): per-channel sigma = MAD / 0.6745 of a 4-neighbour Laplacian response sampled
// every `stride` pixels.  The reference builds it from library calls: a grouped 3x3 conv2d over the WHOLE image (12 B/px read, 12 B/px
// written), a strided slice, and two torch.median (sorts).  Here:
//   laplacian_samples_kernel  evaluates the response only AT the sampled pixels (zero padding outside the image, like conv2d padding=1):
//                             5 loads per sample and channel, 1/stride^2 of the image touched, nothing else written;
//   mad_select_kernel         one CTA per channel: exact median by radix selection on the order-preserving integer image of the floats
//                             (four 8-bit passes over the samples, shared-memory histogram), then the same selection on |r - median|.
// torch.median returns the LOWER of the two middle values for an even count: rank (n - 1) / 2 of the sorted samples; so does this.
// The result stays on the device (3 floats) and can feed Wiener.process as its noise tensor without a host round trip.
#include "select.cuh"

namespace tdb {
namespace {

constexpr int kThreads = 256;
constexpr int kSelectThreads = 1024;

__global__ void __launch_bounds__(kThreads) laplacian_samples_kernel(const float *__restrict__ rgb, float *__restrict__ resp, int width, int height,
                                                                     int stride, int sw, int sh) {
  const int64_t n = (int64_t)sw * sh;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const int sy = (int)(i / sw), sx = (int)(i - (int64_t)sy * sw);
    const int y = sy * stride, x = sx * stride;
    const float *c = rgb + 3 * ((int64_t)y * width + x);
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      const float up = y > 0 ? __ldg(c - 3 * (int64_t)width + ch) : 0.0f, dn = y + 1 < height ? __ldg(c + 3 * (int64_t)width + ch) : 0.0f;
      const float lf = x > 0 ? __ldg(c - 3 + ch) : 0.0f, rt = x + 1 < width ? __ldg(c + 3 + ch) : 0.0f;
      // cross-correlation with [[0,-1,0],[-1,4,-1],[0,-1,0]], taps in row-major order like a direct convolution loop
      resp[ch * n + i] = ((((0.0f - up) - lf) + 4.0f * __ldg(c + ch)) - rt) - dn;
    }
  }
}

__global__ void __launch_bounds__(kSelectThreads) mad_select_kernel(const float *__restrict__ resp, int64_t n, float *__restrict__ sigma) {
  __shared__ uint32_t hist[256], pick[2];
  const float *v = resp + (int64_t)blockIdx.x * n;
  const int64_t rank = (n - 1) / 2;
  auto all = [](int64_t) { return true; };
  const float med = sel::select_rank(n, rank, [v](int64_t i) { return v[i]; }, all, hist, pick);
  const float mad = sel::select_rank(n, rank, [v, med](int64_t i) { return fabsf(v[i] - med); }, all, hist, pick);
  if (threadIdx.x == 0) sigma[blockIdx.x] = __fdiv_rn(mad, 0.6745f);  // IEEE division like torch's (the library is built with fast-math)
}

}  // namespace
}  // namespace tdb

using namespace tdb;

extern "C" {

size_t tdb_channel_noise_scratch_bytes(int width, int height, int stride) {
  if (width <= 0 || height <= 0 || stride <= 0) return 0;
  const int64_t sw = (width + stride - 1) / stride, sh = (height + stride - 1) / stride;
  return (size_t)(3 * sw * sh) * sizeof(float);
}

int tdb_channel_noise(const float *rgb, int width, int height, int stride, void *scratch, float *sigma, tdb_stream_t stream) {
  TDB_REQUIRE(rgb && scratch && sigma, "channel_noise: null pointer");
  TDB_REQUIRE(width > 0 && height > 0 && stride > 0, "channel_noise: invalid size or stride");
  cudaStream_t s = as_stream(stream);
  const int sw = (width + stride - 1) / stride, sh = (height + stride - 1) / stride;  // rows / columns 0, stride, 2 stride, ... (python [::stride])
  const int64_t n = (int64_t)sw * sh;
  float *resp = static_cast<float *>(scratch);
  const int grid = (int)((n + kThreads - 1) / kThreads < (int64_t)kNumSMs * 8 ? (n + kThreads - 1) / kThreads : (int64_t)kNumSMs * 8);
  laplacian_samples_kernel<<<grid, kThreads, 0, s>>>(rgb, resp, width, height, stride, sw, sh);
  if (int e = check_launch("noise_laplacian_samples")) return e;
  mad_select_kernel<<<3, kSelectThreads, 0, s>>>(resp, n, sigma);
  return check_launch("noise_mad_select");
}

}  // extern "C"
