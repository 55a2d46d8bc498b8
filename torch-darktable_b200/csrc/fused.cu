// Pointwise glue of the fused frame pipeline (include/tdb200.h, "Fused frame pipeline").
//
//   frame_prepare  : global green equilibration (the ratio comes from the smoothing kernel's statistics) + normalisation with the
//                    image-set bounds + log-luminance for the Wiener tiles + clearing of the Wiener accumulator, in one pass.
//                    Replaces green_eq_kernel -> normalize_kernel -> loglum_kernel -> memset (24 + 24 + 16 + 4 B/px, four launches);
//                    the arithmetic per pixel is the same, term by term.  When the Wiener stage follows, the normalised colour
//                    leaves as its Lab (a, b) pair instead of RGB (12 B read, 8 + 4 + 4 B written per pixel): the write-back of the
//                    denoised luminance (modify_log_luminance) needs exactly rgb_to_lab(colour).a/b, and computing it here shares
//                    the linearisation with the log-luminance, which saves the write-back pass six of its fifteen pow().
//   metrics_sliced : compute_image_metrics (color_adaption.cu:121-166) of the image the bilateral slice WOULD produce, evaluated
//                    only at the sampled pixels (every stride-th), so that the slice itself can be fused into the tone-map kernel
//                    and the locally contrasted image never exists in HBM.  The last CTA merges the sums into the image set and
//                    applies the moving average (image_processor.py:292-294).
#include <cfloat>

#include "bilateral.cuh"
#include "frame_state.cuh"
#include "wiener_layout.cuh"

namespace tdb {
namespace {

constexpr int kThreads = 256;

inline int flat_grid(int64_t items) {
  const int64_t want = (items + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

struct PrepareArgs {
  const float *in;
  float *out;
  float *loglum;        // or null; with it, `out` receives Lab (a, b) pairs instead of RGB
  float *zero;          // plane to clear (the Wiener accumulator), or null
  unsigned int *counters;  // two job counters of the Wiener tile kernels to clear, or null
  const float *ratio;   // device float[1], or null = no equilibration
  const float *bounds;  // device float[2]
  int width, height;
  uint32_t filters;
  float eps;
};

__device__ __forceinline__ rgb_t prepare_pixel(rgb_t c, bool g1, bool eq, float ratio, float b0, float range) {
  if (eq) {  // green_eq_kernel (global part): G1 sites take the ratio, everything is clamped at zero
    c.y *= g1 ? ratio : 1.0f;
    c.x = fmaxf(c.x, 0.0f), c.y = fmaxf(c.y, 0.0f), c.z = fmaxf(c.z, 0.0f);
  }
  return rgb_t{(c.x - b0) / range, (c.y - b0) / range, (c.z - b0) / range};  // normalize_kernel
}
// (a, b) of rgb_to_lab(c) and log(max(eps, compute_luminance(c))).  compute_luminance clips the colour first; inside [0,1]^3 the clip
// is the identity and its L is the L of the same Lab conversion (same expressions as pub::rgb_to_lab / pub::luminance)
__device__ __forceinline__ float2 lab_ab_and_loglum(rgb_t c, float eps, float &loglum) {
  const bool inside = c.x >= 0.0f && c.x <= 1.0f && c.y >= 0.0f && c.y <= 1.0f && c.z >= 0.0f && c.z <= 1.0f;
  const rgb_t v = pub::rgb_to_xyz(c);
  const float fx = pub::lab_f(v.x / 0.95047f), fy = pub::lab_f(v.y / 1.0f), fz = pub::lab_f(v.z / 1.08883f);
  float L = fmaxf(0.0f, (116.0f / 100.0f) * fy - (16.0f / 100.0f));
  if (!inside) L = pub::luminance(c);
  loglum = logf(fmaxf(eps, L));
  return make_float2((500.0f / 128.0f) * (fx - fy), (200.0f / 128.0f) * (fy - fz));
}

template <bool kVec>
__global__ void __launch_bounds__(kThreads) prepare_kernel(const PrepareArgs a) {
  const bool eq = a.ratio != nullptr;
  const float ratio = eq ? __ldg(a.ratio) : 1.0f;
  const float b0 = __ldg(a.bounds), range = __ldg(a.bounds + 1) - b0;
  const int64_t n = (int64_t)a.width * a.height;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t t0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  // G1 sites: green sites of even rows.  (row 0, col 0) green -> even columns, else odd columns
  const int g1_col = fc(0, 0, a.filters) == 1 ? 0 : 1;
  if (t0 < 2 && a.counters) a.counters[t0] = 0u;
  if (kVec) {  // width % 4 == 0: a 4-pixel group stays inside one row and starts on an even column
    // the row of a group is kept incrementally (one division per thread, not a 64-bit one per group); n / 4 < 2^31
    const unsigned wq = (unsigned)a.width >> 2, ngroups = (unsigned)(n / 4);
    const unsigned step = (unsigned)stride, dy = step / wq, dq = step - dy * wq;
    unsigned y = (unsigned)t0 / wq, q = (unsigned)t0 - y * wq;
    for (unsigned g = (unsigned)t0; g < ngroups; g += step, q += dq, y += dy) {
      if (q >= wq) q -= wq, y++;
      const bool even_row = !(y & 1);
      const float4 *src = reinterpret_cast<const float4 *>(a.in) + 3 * (size_t)g;
      rgb_t p[4];
      unpack4(ld_stream(src), ld_stream(src + 1), ld_stream(src + 2), p);
#pragma unroll
      for (int k = 0; k < 4; k++) p[k] = prepare_pixel(p[k], even_row && ((k & 1) == g1_col), eq, ratio, b0, range);
      if (a.loglum) {
        float ll[4];
        float2 ab[4];
#pragma unroll
        for (int k = 0; k < 4; k++) ab[k] = lab_ab_and_loglum(p[k], a.eps, ll[k]);
        float4 *dst = reinterpret_cast<float4 *>(a.out) + 2 * (size_t)g;
        dst[0] = make_float4(ab[0].x, ab[0].y, ab[1].x, ab[1].y), dst[1] = make_float4(ab[2].x, ab[2].y, ab[3].x, ab[3].y);
        reinterpret_cast<float4 *>(a.loglum)[g] = make_float4(ll[0], ll[1], ll[2], ll[3]);
      } else {
        float4 o0, o1, o2;
        pack4(p, o0, o1, o2);
        float4 *dst = reinterpret_cast<float4 *>(a.out) + 3 * (size_t)g;
        dst[0] = o0, dst[1] = o1, dst[2] = o2;
      }
      if (a.zero) reinterpret_cast<float4 *>(a.zero)[g] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
  } else {
    for (int64_t i = t0; i < n; i += stride) {
      const int y = (int)(i / a.width), x = (int)(i - (int64_t)y * a.width);
      const rgb_t p = prepare_pixel(rgb_t{a.in[3 * i], a.in[3 * i + 1], a.in[3 * i + 2]}, !(y & 1) && ((x & 1) == g1_col), eq, ratio, b0, range);
      if (a.loglum) {
        float ll;
        const float2 ab = lab_ab_and_loglum(p, a.eps, ll);
        a.out[2 * i] = ab.x, a.out[2 * i + 1] = ab.y, a.loglum[i] = ll;
      } else {
        a.out[3 * i] = p.x, a.out[3 * i + 1] = p.y, a.out[3 * i + 2] = p.z;
      }
      if (a.zero) a.zero[i] = 0.0f;
    }
  }
}

struct MetricsArgs {
  const float *rgb;
  const float *grid;  // blurred bilateral grid, or null = measure rgb as it is
  int lab_input;      // the image holds Lab pixels (fused Wiener write-back) instead of RGB
  bil::GridDims g;
  float sigma_s, sigma_r, detail;
  int width, height, stride, sw;
  int64_t nsamples;
  float min_gray;
  FrameState *state;
  int first_in_set, last_in_set;
  const float *prev_metrics;  // EMA state (device float[5]) or null
  float moving_average;
  float *metrics_out;         // device float[5], written on the last frame of a set
  int row_lo, row_hi;         // only the samples of these image rows count (one frame split into row bands)
  float *raw_out;             // device float[6] or null: leave the raw sums of this call here, no set merge, no finalisation
};

__global__ void __launch_bounds__(kThreads) metrics_sliced_kernel(const MetricsArgs a) {
  // same arithmetic as metrics_kernel (pointwise.cu) with bounds = (0, 1)
  const float b0 = 0.0f, range = 1.0f - b0 + 1e-6f;
  float acc[6] = {0, 0, 0, 0, 0, 0};
  const int64_t step = (int64_t)gridDim.x * kThreads;
  for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < a.nsamples; s += step) {
    const int sy = (int)(s / a.sw), sx = (int)(s - (int64_t)sy * a.sw);
    const int x = sx * a.stride, y = sy * a.stride;
    if (y < a.row_lo || y >= a.row_hi) continue;
    const float *p = a.rgb + 3 * ((int64_t)y * a.width + x);
    rgb_t c{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
    if (a.grid) c = a.lab_input ? bil::slice_lab(a.grid, x, y, c, a.g, a.sigma_s, a.sigma_r, a.detail)
                                : bil::slice_rgb(a.grid, x, y, c, a.g, a.sigma_s, a.sigma_r, a.detail);
    const float r = (c.x - b0) / range, g = (c.y - b0) / range, b = (c.z - b0) / range;
    const float mask = (r >= 0.99f || g >= 0.99f || b >= 0.99f) ? 0.0f : 1.0f;
    const float gray = r * 0.299f + g * 0.587f + b * 0.114f;
    acc[0] += logf(fmaxf(gray, a.min_gray)) * mask;
    acc[1] += gray * mask, acc[2] += r * mask, acc[3] += g * mask, acc[4] += b * mask, acc[5] += mask;
  }
  __shared__ float sh[6][kThreads / 32];
  __shared__ unsigned int slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) sh[k][warp] = v;
  }
  __syncthreads();
  FrameState *fs = a.state;
  if (threadIdx.x < 6) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) v += sh[threadIdx.x][w];
    fs->partials[6 * blockIdx.x + threadIdx.x] = v;
  }
  if (!last_cta(&fs->ticket[1], gridDim.x, &slot)) return;
  // fixed-order reduction of the per-CTA sums: warp k sums component k
  if (warp < 6) {
    float v = 0.0f;
    for (unsigned int i = lane; i < gridDim.x; i += 32) v += __ldcg(fs->partials + 6 * i + warp);
    v = warp_sum(v);
    if (lane == 0) sh[warp][0] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0 && a.raw_out) {
    for (int k = 0; k < 6; k++) a.raw_out[k] = sh[k][0];
    fs->ticket[1] = 0;
  } else if (threadIdx.x == 0) {
    float sums[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
      sums[k] = sh[k][0] + (a.first_in_set ? 0.0f : fs->set_sums[k]);
      fs->set_sums[k] = sums[k];
    }
    if (a.last_in_set) {
      const float inv = 1.0f / fmaxf(sums[5], 1.0f);  // metrics_finalize_kernel
      for (int k = 0; k < 5; k++) a.metrics_out[k] = ema(a.prev_metrics, k, sums[k] * inv, a.moving_average);
    }
    fs->ticket[1] = 0;
  }
}

}  // namespace
}  // namespace tdb

namespace tdb {
namespace {
// ---- one frame split into row bands (pipeline/tiled.py): the two places where the ranks' partial statistics meet.  Each used to be a
// dozen tiny torch kernels (sum / min / max over the gathered rows, where, maximum, stack, lerp ...), i.e. ~0.15 ms of launch latency
// per frame on every rank; here each is ONE single-thread kernel.  The formulas are frame_stats_kernel's (csrc/postprocess.cu) and
// metrics_finalize + lerp (csrc/pointwise.cu, pipeline/util.py:4), the sums run over the ranks in rank order (identical on every rank).
__global__ void band_stats_finish_kernel(const float *__restrict__ gathered, int world, const float *prev_bounds, float ma, float *bounds_out,
                                         float *ratio_out) {
  if (threadIdx.x != 0) return;
  float s1 = 0.0f, s2 = 0.0f, g_lo = FLT_MAX, g_hi = -FLT_MAX, o_lo = FLT_MAX, o_hi = -FLT_MAX;
  for (int r = 0; r < world; r++) {
    const float *p = gathered + 6 * r;  // G1 sum, G2 sum, min / max of the sampled G1 greens, min / max of every other sampled value
    s1 += p[0], s2 += p[1];
    g_lo = fminf(g_lo, p[2]), g_hi = fmaxf(g_hi, p[3]), o_lo = fminf(o_lo, p[4]), o_hi = fmaxf(o_hi, p[5]);
  }
  const float ratio = (s1 > 0.0f && s2 > 0.0f) ? __fdiv_rn(s2, s1) : 1.0f;
  *ratio_out = ratio;
  float lo = o_lo, hi = o_hi;
  if (g_lo <= g_hi) lo = fminf(lo, fmaxf(g_lo * ratio, 0.0f)), hi = fmaxf(hi, fmaxf(g_hi * ratio, 0.0f));
  const float p0 = prev_bounds ? prev_bounds[0] : lo, p1 = prev_bounds ? prev_bounds[1] : hi;
  bounds_out[0] = p0 + (lo - p0) * ma, bounds_out[1] = p1 + (hi - p1) * ma;
}
__global__ void band_metrics_finish_kernel(const float *__restrict__ sums, const float *prev_metrics, float ma, float *metrics_out) {
  if (threadIdx.x >= 5) return;
  const float v = sums[threadIdx.x] * (1.0f / fmaxf(sums[5], 1.0f));
  const float p = prev_metrics ? prev_metrics[threadIdx.x] : v;
  metrics_out[threadIdx.x] = p + (v - p) * ma;
}
}  // namespace
}  // namespace tdb

using namespace tdb;

extern "C" {

int tdb_frame_prepare(const float *rgb, float *out, void *wiener_scratch_buf, int width, int height, uint32_t filters, const float *ratio,
                      const float *bounds, float eps, tdb_stream_t stream) {
  TDB_REQUIRE(rgb && out && bounds, "frame_prepare: null pointer");
  TDB_REQUIRE(width > 0 && height > 0, "frame_prepare: empty image");
  TDB_REQUIRE((int64_t)width * height < ((int64_t)1 << 32), "frame_prepare: image too large (2^32 pixels)");
  TDB_REQUIRE(!wiener_scratch_buf || eps > 0.0f, "Epsilon must be positive");
  PrepareArgs a{rgb, out, nullptr, nullptr, nullptr, ratio, bounds, width, height, filters, eps};
  if (wiener_scratch_buf) {
    const WienerScratch ws = wiener_scratch(wiener_scratch_buf, width, height, 1);
    a.loglum = ws.lum, a.zero = ws.acc, a.counters = ws.counters;
  }
  auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = (width % 4 == 0) && al(rgb) && al(out) && al(a.loglum) && al(a.zero);
  const int64_t items = (int64_t)width * height / (vec ? 4 : 1);
  if (vec) prepare_kernel<true><<<flat_grid(items), kThreads, 0, as_stream(stream)>>>(a);
  else prepare_kernel<false><<<flat_grid(items), kThreads, 0, as_stream(stream)>>>(a);
  return check_launch("frame_prepare");
}

static int run_metrics_sliced(const float *rgb, int lab_input, const void *bilateral_scratch, int width, int height, float sigma_s,
                              float sigma_r, float detail, int stride, float min_gray, void *frame_state, int first_in_set, int last_in_set,
                              const float *prev_metrics, float moving_average, float *metrics_out, int row_lo, int row_hi, float *raw_out,
                              tdb_stream_t stream) {
  TDB_REQUIRE(rgb && frame_state && (metrics_out || raw_out), "metrics_sliced: null pointer");
  TDB_REQUIRE(width > 0 && height > 0 && stride > 0, "metrics_sliced: bad arguments");
  MetricsArgs a{};
  a.rgb = rgb, a.lab_input = lab_input;
  TDB_REQUIRE(!lab_input || bilateral_scratch, "metrics_sliced: a Lab image needs the bilateral grid");
  if (bilateral_scratch) {
    TDB_REQUIRE(sigma_r > 0.0f && sigma_s > 0.0f, "metrics_sliced: invalid sigmas");
    a.g = bil::grid_dims(width, height, sigma_s, sigma_r);
    a.grid = static_cast<const float *>(bilateral_scratch) + (size_t)a.g.x * a.g.y * a.g.z;  // [splatted grid][blurred grid]
  }
  a.sigma_s = sigma_s, a.sigma_r = sigma_r, a.detail = detail;
  a.width = width, a.height = height, a.stride = stride;
  a.sw = (width + stride - 1) / stride;
  a.nsamples = (int64_t)a.sw * ((height + stride - 1) / stride);
  a.min_gray = min_gray;
  a.state = static_cast<FrameState *>(frame_state);
  a.first_in_set = first_in_set, a.last_in_set = last_in_set, a.prev_metrics = prev_metrics, a.moving_average = moving_average;
  a.metrics_out = metrics_out, a.row_lo = row_lo, a.row_hi = row_hi, a.raw_out = raw_out;
  metrics_sliced_kernel<<<flat_grid(a.nsamples), kThreads, 0, as_stream(stream)>>>(a);
  return check_launch("metrics_sliced");
}

int tdb_metrics_sliced(const float *rgb, int lab_input, const void *bilateral_scratch, int width, int height, float sigma_s, float sigma_r, float detail,
                       int stride, float min_gray, void *frame_state, int first_in_set, int last_in_set, const float *prev_metrics,
                       float moving_average, float *metrics_out, tdb_stream_t stream) {
  return run_metrics_sliced(rgb, lab_input, bilateral_scratch, width, height, sigma_s, sigma_r, detail, stride, min_gray, frame_state,
                            first_in_set, last_in_set, prev_metrics, moving_average, metrics_out, 0, height, nullptr, stream);
}

int tdb_metrics_sliced_band(const float *rgb, int lab_input, const void *bilateral_scratch, int width, int height, float sigma_s,
                            float sigma_r, float detail, int stride, float min_gray, void *frame_state, int row_lo, int row_hi,
                            float *raw_sums, tdb_stream_t stream) {
  TDB_REQUIRE(raw_sums && row_lo >= 0 && row_lo < row_hi && row_hi <= height, "metrics_sliced_band: bad arguments");
  return run_metrics_sliced(rgb, lab_input, bilateral_scratch, width, height, sigma_s, sigma_r, detail, stride, min_gray, frame_state, 1, 1,
                            nullptr, 1.0f, nullptr, row_lo, row_hi, raw_sums, stream);
}


int tdb_band_stats_finish(const float *gathered, int world, const float *prev_bounds, float moving_average, float *bounds_out, float *ratio_out,
                          tdb_stream_t stream) {
  TDB_REQUIRE(gathered && bounds_out && ratio_out && world > 0, "band_stats_finish: bad arguments");
  band_stats_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(gathered, world, prev_bounds, moving_average, bounds_out, ratio_out);
  return check_launch("band_stats_finish");
}

int tdb_band_metrics_finish(const float *sums, const float *prev_metrics, float moving_average, float *metrics_out, tdb_stream_t stream) {
  TDB_REQUIRE(sums && metrics_out, "band_metrics_finish: null pointer");
  band_metrics_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(sums, prev_metrics, moving_average, metrics_out);
  return check_launch("band_metrics_finish");
}

}  // extern "C"
