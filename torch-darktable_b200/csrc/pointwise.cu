// Pointwise family: white balance, colour conversions, luminance extract / replace, normalisation, image statistics.
//
// Replaces csrc/white_balance.cu, csrc/color_conversions.cu and csrc/tonemap/color_adaption.cu of the reference, which
// launch one thread per pixel with scalar 4-byte accesses to interleaved RGB on the legacy default stream and device-
// synchronise after every op.  Here a thread owns four pixels (48 bytes = three 128-bit accesses), kernels are
// grid-stride over a machine-sized grid, and every scalar the reference reads back to the host stays on the device.
#include "select.cuh"
#include <cfloat>

#include "color_math.cuh"

namespace tdb {
namespace {

constexpr int kThreads = 256;

inline int flat_grid(int64_t items) {
  const int64_t want = (items + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- generic RGB -> RGB map over 4-pixel groups ------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(kThreads) map_rgb_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t ngroups,
                                                           int64_t npixels, F f) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t t0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  for (int64_t g = t0; g < ngroups; g += stride) {
    const float4 *src = reinterpret_cast<const float4 *>(in) + 3 * g;
    rgb_t p[4];
    unpack4(__ldg(src), __ldg(src + 1), __ldg(src + 2), p);
#pragma unroll
    for (int k = 0; k < 4; k++) p[k] = f(p[k]);
    float4 a, b, c;
    pack4(p, a, b, c);
    float4 *dst = reinterpret_cast<float4 *>(out) + 3 * g;
    dst[0] = a, dst[1] = b, dst[2] = c;
  }
  for (int64_t i = ngroups * 4 + t0; i < npixels; i += stride) {
    const rgb_t r = f(rgb_t{in[3 * i], in[3 * i + 1], in[3 * i + 2]});
    out[3 * i] = r.x, out[3 * i + 1] = r.y, out[3 * i + 2] = r.z;
  }
}

// RGB -> scalar
template <class F>
__global__ void __launch_bounds__(kThreads) map_rgb_scalar_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                                  int64_t ngroups, int64_t npixels, F f) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t t0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  for (int64_t g = t0; g < ngroups; g += stride) {
    const float4 *src = reinterpret_cast<const float4 *>(in) + 3 * g;
    rgb_t p[4];
    unpack4(__ldg(src), __ldg(src + 1), __ldg(src + 2), p);
    reinterpret_cast<float4 *>(out)[g] = make_float4(f(p[0]), f(p[1]), f(p[2]), f(p[3]));
  }
  for (int64_t i = ngroups * 4 + t0; i < npixels; i += stride) out[i] = f(rgb_t{in[3 * i], in[3 * i + 1], in[3 * i + 2]});
}

// (RGB, scalar) -> RGB
template <class F>
__global__ void __launch_bounds__(kThreads) map_rgb_with_scalar_kernel(const float *__restrict__ in, const float *__restrict__ s,
                                                                       float *__restrict__ out, int64_t ngroups, int64_t npixels, F f) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t t0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  for (int64_t g = t0; g < ngroups; g += stride) {
    const float4 *src = reinterpret_cast<const float4 *>(in) + 3 * g;
    rgb_t p[4];
    unpack4(__ldg(src), __ldg(src + 1), __ldg(src + 2), p);
    const float4 sv = __ldg(reinterpret_cast<const float4 *>(s) + g);
    p[0] = f(p[0], sv.x), p[1] = f(p[1], sv.y), p[2] = f(p[2], sv.z), p[3] = f(p[3], sv.w);
    float4 a, b, c;
    pack4(p, a, b, c);
    float4 *dst = reinterpret_cast<float4 *>(out) + 3 * g;
    dst[0] = a, dst[1] = b, dst[2] = c;
  }
  for (int64_t i = ngroups * 4 + t0; i < npixels; i += stride) {
    const rgb_t r = f(rgb_t{in[3 * i], in[3 * i + 1], in[3 * i + 2]}, s[i]);
    out[3 * i] = r.x, out[3 * i + 1] = r.y, out[3 * i + 2] = r.z;
  }
}

struct OpRgbToXyz { __device__ rgb_t operator()(rgb_t c) const { return pub::rgb_to_xyz(c); } };
struct OpXyzToLab { __device__ rgb_t operator()(rgb_t c) const { return pub::xyz_to_lab(c); } };
struct OpLabToXyz { __device__ rgb_t operator()(rgb_t c) const { return pub::lab_to_xyz(c); } };
struct OpXyzToRgb { __device__ rgb_t operator()(rgb_t c) const { return pub::xyz_to_rgb(c); } };
struct OpRgbToLab { __device__ rgb_t operator()(rgb_t c) const { return pub::rgb_to_lab(c); } };
struct OpLabToRgb { __device__ rgb_t operator()(rgb_t c) const { return pub::lab_to_rgb(c); } };
struct OpHsl {
  float h, s, l;
  __device__ rgb_t operator()(rgb_t c) const { return pub::modify_hsl(c, h, s, l); }
};
struct OpVibrance {
  float amount;
  __device__ rgb_t operator()(rgb_t c) const { return pub::modify_vibrance(c, amount); }
};
struct OpMatrix {
  const float *m;  // device pointer, 9 floats
  __device__ rgb_t operator()(rgb_t c) const {
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; i++) k[i] = __ldg(m + i);
    return clip01(mat3(k, c));
  }
};
struct OpLum { __device__ float operator()(rgb_t c) const { return pub::luminance(c); } };
struct OpLogLum {
  float eps;
  __device__ float operator()(rgb_t c) const { return logf(fmaxf(eps, pub::luminance(c))); }
};
struct OpSetLum { __device__ rgb_t operator()(rgb_t c, float l) const { return pub::with_luminance(c, l); } };
struct OpSetLogLum { __device__ rgb_t operator()(rgb_t c, float l) const { return pub::with_luminance(c, expf(l)); } };

template <class F>
int launch_map_rgb(const float *in, float *out, int64_t npixels, F f, tdb_stream_t stream, const char *name) {
  TDB_REQUIRE(in && out, "%s: null pointer", name);
  if (npixels <= 0) return TDB_OK;
  const int64_t ngroups = (aligned16(in) && aligned16(out)) ? npixels / 4 : 0;
  map_rgb_kernel<F><<<flat_grid(ngroups ? ngroups : npixels), kThreads, 0, as_stream(stream)>>>(in, out, ngroups, npixels, f);
  return check_launch(name);
}

// ---- flat float ops ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) normalize_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t nvec,
                                                             int64_t n, const float *__restrict__ bounds) {
  const float b0 = __ldg(bounds), range = __ldg(bounds + 1) - b0;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t t0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  for (int64_t i = t0; i < nvec; i += stride) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(in) + i);
    v.x = (v.x - b0) / range, v.y = (v.y - b0) / range, v.z = (v.z - b0) / range, v.w = (v.w - b0) / range;
    reinterpret_cast<float4 *>(out)[i] = v;
  }
  for (int64_t i = nvec * 4 + t0; i < n; i += stride) out[i] = (in[i] - b0) / range;
}

__global__ void lerp_kernel(const float *a, const float *b, float t, float *out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b[i] - a[i]) * t;
}

// ---- white balance -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) white_balance_kernel(const float *__restrict__ in, float *__restrict__ out, int width,
                                                                 int height, uint32_t filters, const float *__restrict__ gains,
                                                                 int vec) {
  const float gr = __ldg(gains), gg = __ldg(gains + 1), gb = __ldg(gains + 2);
  const int64_t n = (int64_t)width * height;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t t0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  auto gain_at = [&](int y, int x) {
    const int c = fc(y, x, filters);
    return c == 0 ? gr : (c == 2 ? gb : gg);
  };
  if (vec) {  // width % 4 == 0: a float4 never straddles a row and starts on an even column
    const int wq = width >> 2;
    for (int64_t q = t0; q < n / 4; q += stride) {
      const int y = (int)(q / wq);
      const float g0 = gain_at(y, 0), g1 = gain_at(y, 1);
      float4 v = ld_stream(reinterpret_cast<const float4 *>(in) + q);
      v.x = clip01(v.x * g0), v.y = clip01(v.y * g1), v.z = clip01(v.z * g0), v.w = clip01(v.w * g1);
      reinterpret_cast<float4 *>(out)[q] = v;
    }
  } else {
    for (int64_t i = t0; i < n; i += stride) {
      const int y = (int)(i / width), x = (int)(i - (int64_t)y * width);
      out[i] = clip01(in[i] * gain_at(y, x));
    }
  }
}

// 2x2 patch -> chromaticity / intensity / unsaturated flag (reference white_balance.cu:57-82; patch origin is pos*2)
__global__ void wb_collect_kernel(const float *__restrict__ cfa, int width, int height, uint32_t filters, int stride,
                                  float *__restrict__ chroma, float *__restrict__ intensity, uint8_t *__restrict__ valid) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
  const int sw = width / stride, sh = height / stride;
  if (px >= sw || py >= sh) return;
  const int idx = py * sw + px;
  if (px + 1 >= sw || py + 1 >= sh) {  // the reference leaves these uninitialised; define them as invalid
    chroma[2 * idx] = chroma[2 * idx + 1] = 0.0f;
    intensity[idx] = 0.0f;
    valid[idx] = 0;
    return;
  }
  const int x = px * 2, y = py * 2;
  const float p00 = cfa[(int64_t)y * width + x], p01 = cfa[(int64_t)y * width + x + 1];
  const float p10 = cfa[(int64_t)(y + 1) * width + x], p11 = cfa[(int64_t)(y + 1) * width + x + 1];
  float r, g, b;
  switch (filters) {
    case TDB_FILTERS_RGGB: r = p00, g = (p01 + p10) * 0.5f, b = p11; break;
    case TDB_FILTERS_BGGR: r = p11, g = (p01 + p10) * 0.5f, b = p00; break;
    case TDB_FILTERS_GRBG: r = p01, g = (p00 + p11) * 0.5f, b = p10; break;
    default: r = p10, g = (p00 + p11) * 0.5f, b = p01; break;
  }
  const float s = r + g + b;
  chroma[2 * idx] = r / s, chroma[2 * idx + 1] = g / s;
  intensity[idx] = s;
  valid[idx] = fmaxf(fmaxf(p00, p01), fmaxf(p10, p11)) < 1.0f;
}

// The rest of estimate_white_balance (reference white_balance.cu:135-161: boolean-mask gathers, torch::quantile -- a full sort --,
// a second masked gather and a mean, about ten library launches and two device-side compactions) in ONE single-CTA kernel over the
// sample arrays: count the valid samples, find the two order statistics torch.quantile interpolates between by exact radix
// selection (select.cuh), and average the chromaticity of the samples at or above the threshold.
//   torch.quantile(v, q), linear: rank = q * (n - 1) in the dtype of v (float32), lo = floor(rank), hi = ceil(rank),
//   result = lerp(sorted[lo], sorted[hi], rank - lo) with ATen's lerp (w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)).
// gains = (mean_r / mean_g, 1, (1 - mean_r - mean_g) / mean_g), or (1, 1, 1) when nothing is valid.  IEEE operations throughout
// (the library is built with fast-math; the reference's ATen kernels are not).
constexpr int kWbThreads = 1024;
__global__ void __launch_bounds__(kWbThreads) wb_finish_kernel(const float *__restrict__ chroma, const float *__restrict__ intensity,
                                                                const uint8_t *__restrict__ valid, int64_t n, float quantile,
                                                                float *__restrict__ gains) {
  __shared__ uint32_t hist[256], pick[2];
  __shared__ double red[3][kWbThreads / 32];
  __shared__ double total[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto block_sum3 = [&](double a, double b, double c) {  // -> total[0..2], visible to every thread
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o), b += __shfl_xor_sync(0xffffffffu, b, o), c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    __syncthreads();  // the previous round's readers of total[] are done
    if (lane == 0) red[0][warp] = a, red[1][warp] = b, red[2][warp] = c;
    __syncthreads();
    if (tid < 3) {
      double s = 0.0;
      for (int w = 0; w < kWbThreads / 32; w++) s += red[tid][w];
      total[tid] = s;
    }
    __syncthreads();
  };
  double cnt = 0.0;
  for (int64_t i = tid; i < n; i += kWbThreads) cnt += valid[i] ? 1.0 : 0.0;
  block_sum3(cnt, 0.0, 0.0);
  const int64_t nv = (int64_t)total[0];
  if (nv == 0) {
    if (tid < 3) gains[tid] = 1.0f;
    return;
  }
  const float rank = __fmul_rn(quantile, (float)(nv - 1));
  const float below = floorf(rank), above = ceilf(rank);
  const int64_t lo = min((int64_t)below, nv - 1), hi = min((int64_t)above, nv - 1);
  auto value = [intensity](int64_t i) { return intensity[i]; };
  auto keep = [valid](int64_t i) { return valid[i] != 0; };
  const float a = sel::select_rank(n, lo, value, keep, hist, pick);
  const float b = hi == lo ? a : sel::select_rank(n, hi, value, keep, hist, pick);
  const float w = __fsub_rn(rank, below), d = __fsub_rn(b, a);
  const float thr = w < 0.5f ? __fadd_rn(a, __fmul_rn(w, d)) : __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, w)));
  double sr = 0.0, sg = 0.0, m = 0.0;
  for (int64_t i = tid; i < n; i += kWbThreads) {
    if (valid[i] && intensity[i] >= thr) sr += (double)chroma[2 * i], sg += (double)chroma[2 * i + 1], m += 1.0;
  }
  block_sum3(sr, sg, m);
  if (tid == 0) {
    if (total[2] == 0.0) {
      gains[0] = gains[1] = gains[2] = 1.0f;
    } else {
      const float mr = (float)(total[0] / total[2]), mg = (float)(total[1] / total[2]);
      gains[0] = __fdiv_rn(mr, mg), gains[1] = 1.0f, gains[2] = __fdiv_rn(__fsub_rn(__fsub_rn(1.0f, mr), mg), mg);
    }
  }
}

// ---- image statistics ----------------------------------------------------------------------------------------------
__global__ void init2_kernel(float *p, float a, float b) { p[0] = a, p[1] = b; }
__global__ void zero_kernel(float *p, int n) {
  if (threadIdx.x < n) p[threadIdx.x] = 0.0f;
}

__global__ void __launch_bounds__(kThreads) bounds_kernel(const float *__restrict__ rgb, int width, int height, int stride, int sw,
                                                          int64_t nsamples, float *__restrict__ bounds) {
  float lo = FLT_MAX, hi = -FLT_MAX;
  const int64_t step = (int64_t)gridDim.x * kThreads;
  for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < nsamples; s += step) {
    const int sy = (int)(s / sw), sx = (int)(s - (int64_t)sy * sw);
    const float *p = rgb + 3 * ((int64_t)sy * stride * width + (int64_t)sx * stride);
    const float r = __ldg(p), g = __ldg(p + 1), b = __ldg(p + 2);
    lo = fminf(lo, fminf(fminf(r, g), b));
    hi = fmaxf(hi, fmaxf(fmaxf(r, g), b));
  }
  lo = warp_min(lo), hi = warp_max(hi);
  __shared__ float slo[kThreads / 32], shi[kThreads / 32];
  if ((threadIdx.x & 31) == 0) slo[threadIdx.x >> 5] = lo, shi[threadIdx.x >> 5] = hi;
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = threadIdx.x < kThreads / 32 ? slo[threadIdx.x] : FLT_MAX;
    hi = threadIdx.x < kThreads / 32 ? shi[threadIdx.x] : -FLT_MAX;
    lo = warp_min(lo), hi = warp_max(hi);
    if (threadIdx.x == 0) atomic_min_float(bounds, lo), atomic_max_float(bounds + 1, hi);
  }
}

__global__ void __launch_bounds__(kThreads) metrics_kernel(const float *__restrict__ rgb, int width, int height, int stride, int sw,
                                                           int64_t nsamples, float min_gray, const float *__restrict__ bounds,
                                                           float *__restrict__ sums) {
  const float b0 = bounds ? __ldg(bounds) : 0.0f, b1 = bounds ? __ldg(bounds + 1) : 1.0f;
  const float range = b1 - b0 + 1e-6f;
  float acc[6] = {0, 0, 0, 0, 0, 0};
  const int64_t step = (int64_t)gridDim.x * kThreads;
  for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < nsamples; s += step) {
    const int sy = (int)(s / sw), sx = (int)(s - (int64_t)sy * sw);
    const float *p = rgb + 3 * ((int64_t)sy * stride * width + (int64_t)sx * stride);
    const float r = (__ldg(p) - b0) / range, g = (__ldg(p + 1) - b0) / range, b = (__ldg(p + 2) - b0) / range;
    const float mask = (r >= 0.99f || g >= 0.99f || b >= 0.99f) ? 0.0f : 1.0f;
    const float gray = r * 0.299f + g * 0.587f + b * 0.114f;
    acc[0] += logf(fmaxf(gray, min_gray)) * mask;
    acc[1] += gray * mask, acc[2] += r * mask, acc[3] += g * mask, acc[4] += b * mask, acc[5] += mask;
  }
  __shared__ float sh[6][kThreads / 32];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) v += sh[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, v);
  }
}

__global__ void metrics_finalize_kernel(const float *sums, float *metrics) {
  if (threadIdx.x < 5) metrics[threadIdx.x] = sums[threadIdx.x] * (1.0f / fmaxf(sums[5], 1.0f));
}

}  // namespace
}  // namespace tdb

using namespace tdb;

template <class F>
static int launch_scalar(const float *rgb, float *out, int64_t npixels, F f, tdb_stream_t stream, const char *name) {
  TDB_REQUIRE(rgb && out, "%s: null pointer", name);
  if (npixels <= 0) return TDB_OK;
  const int64_t ngroups = (aligned16(rgb) && aligned16(out)) ? npixels / 4 : 0;
  map_rgb_scalar_kernel<F><<<flat_grid(ngroups ? ngroups : npixels), kThreads, 0, as_stream(stream)>>>(rgb, out, ngroups, npixels, f);
  return check_launch(name);
}

template <class F>
static int launch_with_scalar(const float *rgb, const float *s, float *out, int64_t npixels, F f, tdb_stream_t stream, const char *name) {
  TDB_REQUIRE(rgb && s && out, "%s: null pointer", name);
  if (npixels <= 0) return TDB_OK;
  const int64_t ngroups = (aligned16(rgb) && aligned16(out) && aligned16(s)) ? npixels / 4 : 0;
  map_rgb_with_scalar_kernel<F><<<flat_grid(ngroups ? ngroups : npixels), kThreads, 0, as_stream(stream)>>>(rgb, s, out, ngroups, npixels, f);
  return check_launch(name);
}

extern "C" {

int tdb_color_convert(const float *in, float *out, int64_t npixels, int op, float p0, float p1, float p2, tdb_stream_t stream) {
  switch (op) {
    case TDB_RGB_TO_XYZ: return launch_map_rgb(in, out, npixels, OpRgbToXyz{}, stream, "rgb_to_xyz");
    case TDB_XYZ_TO_LAB: return launch_map_rgb(in, out, npixels, OpXyzToLab{}, stream, "xyz_to_lab");
    case TDB_LAB_TO_XYZ: return launch_map_rgb(in, out, npixels, OpLabToXyz{}, stream, "lab_to_xyz");
    case TDB_XYZ_TO_RGB: return launch_map_rgb(in, out, npixels, OpXyzToRgb{}, stream, "xyz_to_rgb");
    case TDB_RGB_TO_LAB: return launch_map_rgb(in, out, npixels, OpRgbToLab{}, stream, "rgb_to_lab");
    case TDB_LAB_TO_RGB: return launch_map_rgb(in, out, npixels, OpLabToRgb{}, stream, "lab_to_rgb");
    case TDB_MODIFY_HSL: return launch_map_rgb(in, out, npixels, OpHsl{p0, p1, p2}, stream, "modify_hsl");
    case TDB_MODIFY_VIBRANCE: return launch_map_rgb(in, out, npixels, OpVibrance{p0}, stream, "modify_vibrance");
  }
  set_error("color_convert: unknown op %d", op);
  return TDB_EINVAL;
}

int tdb_color_transform_3x3(const float *in, float *out, int64_t npixels, const float *matrix, tdb_stream_t stream) {
  TDB_REQUIRE(matrix, "color_transform_3x3: null matrix");
  return launch_map_rgb(in, out, npixels, OpMatrix{matrix}, stream, "color_transform_3x3");
}

int tdb_compute_luminance(const float *rgb, float *lum, int64_t npixels, tdb_stream_t stream) {
  return launch_scalar(rgb, lum, npixels, OpLum{}, stream, "compute_luminance");
}
int tdb_compute_log_luminance(const float *rgb, float *loglum, int64_t npixels, float eps, tdb_stream_t stream) {
  TDB_REQUIRE(eps > 0.0f, "Epsilon must be positive");
  return launch_scalar(rgb, loglum, npixels, OpLogLum{eps}, stream, "compute_log_luminance");
}
int tdb_modify_luminance(const float *rgb, const float *lum, float *out, int64_t npixels, tdb_stream_t stream) {
  return launch_with_scalar(rgb, lum, out, npixels, OpSetLum{}, stream, "modify_luminance");
}
int tdb_modify_log_luminance(const float *rgb, const float *loglum, float *out, int64_t npixels, float eps, tdb_stream_t stream) {
  TDB_REQUIRE(eps > 0.0f, "Epsilon must be positive");
  return launch_with_scalar(rgb, loglum, out, npixels, OpSetLogLum{}, stream, "modify_log_luminance");
}

int tdb_normalize(const float *in, float *out, int64_t nvalues, const float *bounds, tdb_stream_t stream) {
  TDB_REQUIRE(in && out && bounds, "normalize: null pointer");
  if (nvalues <= 0) return TDB_OK;
  const int64_t nvec = (aligned16(in) && aligned16(out)) ? nvalues / 4 : 0;
  normalize_kernel<<<flat_grid(nvec ? nvec : nvalues), kThreads, 0, as_stream(stream)>>>(in, out, nvec, nvalues, bounds);
  return check_launch("normalize");
}

int tdb_lerp(const float *a, const float *b, float t, float *out, int n, tdb_stream_t stream) {
  TDB_REQUIRE(a && b && out && n > 0, "lerp: bad arguments");
  lerp_kernel<<<div_up(n, 64), 64, 0, as_stream(stream)>>>(a, b, t, out, n);
  return check_launch("lerp");
}

int tdb_white_balance(const float *in, float *out, int width, int height, uint32_t filters, const float *gains, tdb_stream_t stream) {
  TDB_REQUIRE(in && out && gains, "apply_white_balance: null pointer");
  TDB_REQUIRE(width > 0 && height > 0, "apply_white_balance: empty image");
  const int vec = (width % 4 == 0) && aligned16(in) && aligned16(out);
  const int64_t items = vec ? (int64_t)width * height / 4 : (int64_t)width * height;
  white_balance_kernel<<<flat_grid(items), kThreads, 0, as_stream(stream)>>>(in, out, width, height, filters, gains, vec);
  return check_launch("apply_white_balance");
}

int tdb_wb_collect_samples(const float *cfa, int width, int height, uint32_t filters, int stride, float *chroma, float *intensity,
                           uint8_t *valid, tdb_stream_t stream) {
  TDB_REQUIRE(cfa && chroma && intensity && valid && stride > 0, "estimate_white_balance: bad arguments");
  const int sw = width / stride, sh = height / stride;
  if (sw <= 0 || sh <= 0) return TDB_OK;
  dim3 block(16, 16), grid(div_up(sw, 16), div_up(sh, 16));
  wb_collect_kernel<<<grid, block, 0, as_stream(stream)>>>(cfa, width, height, filters, stride, chroma, intensity, valid);
  return check_launch("wb_collect_samples");
}

int tdb_wb_estimate_gains(const float *chroma, const float *intensity, const uint8_t *valid, int64_t n, float quantile, float *gains,
                          tdb_stream_t stream) {
  TDB_REQUIRE(chroma && intensity && valid && gains, "estimate_white_balance: null pointer");
  TDB_REQUIRE(n >= 0, "estimate_white_balance: negative sample count");
  wb_finish_kernel<<<1, kWbThreads, 0, as_stream(stream)>>>(chroma, intensity, valid, n, quantile, gains);
  return check_launch("wb_estimate_gains");
}

int tdb_bounds_init(float *bounds, tdb_stream_t stream) {
  TDB_REQUIRE(bounds, "bounds_init: null pointer");
  init2_kernel<<<1, 1, 0, as_stream(stream)>>>(bounds, FLT_MAX, -FLT_MAX);
  return check_launch("bounds_init");
}

int tdb_bounds_accumulate(const float *rgb, int width, int height, int stride, float *bounds, tdb_stream_t stream) {
  TDB_REQUIRE(rgb && bounds && stride > 0 && width > 0 && height > 0, "compute_image_bounds: bad arguments");
  const int sw = (width + stride - 1) / stride, sh = (height + stride - 1) / stride;
  const int64_t n = (int64_t)sw * sh;
  bounds_kernel<<<flat_grid(n), kThreads, 0, as_stream(stream)>>>(rgb, width, height, stride, sw, n, bounds);
  return check_launch("compute_image_bounds");
}

int tdb_metrics_init(float *sums, tdb_stream_t stream) {
  TDB_REQUIRE(sums, "metrics_init: null pointer");
  zero_kernel<<<1, 32, 0, as_stream(stream)>>>(sums, 6);
  return check_launch("metrics_init");
}

int tdb_metrics_accumulate(const float *rgb, int width, int height, int stride, float min_gray, const float *bounds, float *sums,
                           tdb_stream_t stream) {
  TDB_REQUIRE(rgb && sums && stride > 0 && width > 0 && height > 0, "compute_image_metrics: bad arguments");
  const int sw = (width + stride - 1) / stride, sh = (height + stride - 1) / stride;
  const int64_t n = (int64_t)sw * sh;
  metrics_kernel<<<flat_grid(n), kThreads, 0, as_stream(stream)>>>(rgb, width, height, stride, sw, n, min_gray, bounds, sums);
  return check_launch("compute_image_metrics");
}

int tdb_metrics_finalize(const float *sums, float *metrics, tdb_stream_t stream) {
  TDB_REQUIRE(sums && metrics, "metrics_finalize: null pointer");
  metrics_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(sums, metrics);
  return check_launch("metrics_finalize");
}

}  // extern "C"
