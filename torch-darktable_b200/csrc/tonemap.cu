// Tone mapping epilogue: [3x3 colour matrix] -> Reinhard / ACES / adaptive ACES / linear -> gamma -> vibrance -> uint8,
// with the camera's rotate/flip transform folded into the store.
//
// Replaces csrc/tonemap/{reinhard,aces,linear}.cu (one thread per pixel, byte stores, device sync after the launch) and
// the separate torch.rot90/flip + .contiguous() pass of pipeline/transform.py:39-56.  A CTA owns a 32x32 pixel tile:
// rows are read coalesced, results are staged as packed RGB bytes in shared memory and written out as rows of the
// TRANSFORMED image, so that transposing transforms still store contiguous 96-byte runs.
// Algorithmic traffic: 12 B read + 3 B written per pixel.  The arithmetic (about 36 MUFU ops per pixel for the
// pow/cbrt chains) rather than HBM bounds this kernel; see DESIGN.md.
#include "bilateral.cuh"

namespace tdb {
namespace {

constexpr int kTile = 32;
constexpr int kThreads = 256;

struct TonemapArgs {
  float gamma, intensity, light_adapt, vibrance;
  const float *metrics;  // device float[5] or null
  const float *matrix;   // device float[9] or null
  int op, transform;
};

// kSlice (1: the image is RGB, 2: the image is Lab): the pixel is produced by the bilateral slice (Bilateral.process_rgb's last step) instead of being read as is, so the
// locally contrasted image is never written to HBM: 12 B read + 3 B written per pixel for slice + tone map together.
struct SliceArgs {
  const float *grid;  // blurred bilateral grid
  bil::GridDims g;
  float sigma_s, sigma_r, detail;
  bil::GridLimits lim;  // the grid limits as floats (bil::make_sample's float form: no int <-> float conversions per pixel)
};

// destination coordinates of source pixel (x, y); (ow, oh) = transformed size.  torch.rot90(k) is counter-clockwise.
__device__ __forceinline__ void map_xy(int tf, int x, int y, int w, int h, int &ox, int &oy) {
  switch (tf) {
    case TDB_TF_ROTATE_90: ox = y, oy = w - 1 - x; break;
    case TDB_TF_ROTATE_180: ox = w - 1 - x, oy = h - 1 - y; break;
    case TDB_TF_ROTATE_270: ox = h - 1 - y, oy = x; break;
    case TDB_TF_TRANSPOSE: ox = y, oy = x; break;
    case TDB_TF_FLIP_HORIZ: ox = w - 1 - x, oy = y; break;
    case TDB_TF_FLIP_VERT: ox = x, oy = h - 1 - y; break;
    case TDB_TF_TRANSVERSE: ox = w - 1 - x, oy = h - 1 - y; break;  // torch.flip(image, (0, 1)) in the reference
    default: ox = x, oy = y; break;
  }
}

// per-launch constants of the tone curve (reference color_adaption.h:17-44)
struct ToneConsts {
  float m[9];
  bool has_matrix;
  float map_key, exposure, mean[3], aces_gain, inv_gamma, light_adapt, vibrance;
};
template <int kOp>
__device__ __forceinline__ ToneConsts tone_consts(const TonemapArgs &a) {
  ToneConsts k{};
  k.has_matrix = a.matrix != nullptr;
  if (k.has_matrix) {
#pragma unroll
    for (int i = 0; i < 9; i++) k.m[i] = __ldg(a.matrix + i);
  }
  k.map_key = 0.0f, k.exposure = 1.0f;
  if (kOp != TDB_TM_ACES) {
    const float normalized = fmaxf(0.0f, fminf(1.0f, (-__ldg(a.metrics)) / 9.21034f));
    k.map_key = 0.3f + 0.7f * powf(normalized, 1.4f);
    k.exposure = expf(a.intensity);
    k.mean[0] = __ldg(a.metrics + 2), k.mean[1] = __ldg(a.metrics + 3), k.mean[2] = __ldg(a.metrics + 4);
  }
  k.aces_gain = powf(2.0f, a.intensity);
  k.inv_gamma = 1.0f / a.gamma;
  k.light_adapt = a.light_adapt, k.vibrance = a.vibrance;
  return k;
}
// one pixel: [slice] -> [3x3] -> tone curve -> gamma -> vibrance -> 0x00BBGGRR
template <int kOp, int kSlice>
__device__ __forceinline__ uint32_t tone_pixel(rgb_t c, float x, float y, const ToneConsts &k, const SliceArgs &sl) {
  if (kSlice == 1) c = bil::slice_rgb(sl.grid, x, y, c, sl.g, sl.lim, sl.sigma_s, sl.sigma_r, sl.detail);
  if (kSlice == 2) c = bil::slice_lab(sl.grid, x, y, c, sl.g, sl.lim, sl.sigma_s, sl.sigma_r, sl.detail);  // the input pixel is Lab
  if (k.has_matrix) c = mat3(k.m, c);
  rgb_t t;
  if (kOp == TDB_TM_ACES) {
    t = tm::aces_fit(rgb_t{c.x * k.aces_gain, c.y * k.aces_gain, c.z * k.aces_gain});
  } else {
    // lerp(light_adapt, global_mean, pixel) / exposure, to the power map_key  (color_adaption.h:46-76)
    const float ax = powf((k.mean[0] + k.light_adapt * (c.x - k.mean[0])) / k.exposure, k.map_key);
    const float ay = powf((k.mean[1] + k.light_adapt * (c.y - k.mean[1])) / k.exposure, k.map_key);
    const float az = powf((k.mean[2] + k.light_adapt * (c.z - k.mean[2])) / k.exposure, k.map_key);
    if (kOp == TDB_TM_REINHARD) t = rgb_t{c.x / (ax + c.x), c.y / (ay + c.y), c.z / (az + c.z)};
    else if (kOp == TDB_TM_LINEAR) t = rgb_t{c.x / ax, c.y / ay, c.z / az};
    else t = tm::aces_fit(rgb_t{c.x / ax, c.y / ay, c.z / az});
  }
  const rgb_t g{powf(fmaxf(t.x, 0.0f), k.inv_gamma), powf(fmaxf(t.y, 0.0f), k.inv_gamma), powf(fmaxf(t.z, 0.0f), k.inv_gamma)};
  const rgb_t v = tm::vibrance(g, k.vibrance);
  return tm::pack_u8(v);
}

// Transforms that keep rows as rows (none, flips, rotate_180): a thread owns four consecutive pixels of a row -- three 128-bit
// loads, three 32-bit stores of the twelve result bytes -- and no shared-memory tile is needed.  width % 4 == 0.
template <int kOp, int kSlice>
__global__ void __launch_bounds__(kThreads) tonemap_rows_kernel(const float *__restrict__ rgb, uint8_t *__restrict__ out, int width,
                                                                int height, TonemapArgs a, SliceArgs sl) {
  const ToneConsts k = tone_consts<kOp>(a);
  const bool flip_x = a.transform == TDB_TF_FLIP_HORIZ || a.transform == TDB_TF_ROTATE_180 || a.transform == TDB_TF_TRANSVERSE;
  const bool flip_y = a.transform == TDB_TF_FLIP_VERT || a.transform == TDB_TF_ROTATE_180 || a.transform == TDB_TF_TRANSVERSE;
  const int wq = width >> 2;
  const int64_t ngroups = (int64_t)wq * height;
  // row / column of a group kept incrementally: one 64-bit division per thread instead of one per group
  const int64_t step = (int64_t)gridDim.x * kThreads, g0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int dy = (int)(step / wq), dq = (int)(step - (int64_t)dy * wq);
  int y = (int)(g0 / wq), q = (int)(g0 - (int64_t)y * wq);
  for (int64_t g = g0; g < ngroups; g += step, q += dq, y += dy) {
    if (q >= wq) q -= wq, y++;
    const int x = 4 * q;
    const float4 *src = reinterpret_cast<const float4 *>(rgb) + 3 * g;
    rgb_t p[4];
    unpack4(__ldg(src), __ldg(src + 1), __ldg(src + 2), p);
    uint32_t v[4];
    const float xf = (float)x, yf = (float)y;  // pixel coordinates < 2^24: xf + i is exact
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = tone_pixel<kOp, kSlice>(p[i], xf + (float)i, yf, k, sl);
    if (flip_x) {
      const uint32_t t0 = v[0], t1 = v[1];
      v[0] = v[3], v[1] = v[2], v[2] = t1, v[3] = t0;
    }
    const int oy = flip_y ? height - 1 - y : y, ox = flip_x ? width - 4 - x : x;
    uint32_t *o = reinterpret_cast<uint32_t *>(out + 3 * ((int64_t)oy * width + ox));  // 12-byte groups of a 4-byte aligned buffer
    o[0] = v[0] | (v[1] << 24), o[1] = (v[1] >> 8) | (v[2] << 16), o[2] = (v[2] >> 16) | (v[3] << 8);
  }
}

template <int kOp, int kSlice>
__global__ void __launch_bounds__(kThreads) tonemap_kernel(const float *__restrict__ rgb, uint8_t *__restrict__ out, int width,
                                                           int height, TonemapArgs a, SliceArgs sl) {
  __shared__ uint32_t tile[kTile][kTile + 1];  // 0x00BBGGRR per pixel, indexed [y][x] in SOURCE tile coordinates

  const ToneConsts kc = tone_consts<kOp>(a);

  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const int tw = min(kTile, width - x0), th = min(kTile, height - y0);  // valid source extent of this tile
  if ((width & 3) == 0 && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0) {
    // a thread owns four consecutive pixels of a tile row: three 128-bit loads (tw is a multiple of 4 here)
    const int q = threadIdx.x & 7, ly = threadIdx.x >> 3;
    const int x = x0 + 4 * q, y = y0 + ly;
    if (4 * q < tw && ly < th) {
      const float4 *src = reinterpret_cast<const float4 *>(rgb + 3 * ((int64_t)y * width + x));
      rgb_t p[4];
      unpack4(__ldg(src), __ldg(src + 1), __ldg(src + 2), p);
      const float xf = (float)x, yf = (float)y;  // pixel coordinates < 2^24: xf + i is exact
#pragma unroll
      for (int i = 0; i < 4; i++) tile[ly][4 * q + i] = tone_pixel<kOp, kSlice>(p[i], xf + (float)i, yf, kc, sl);
    }
  } else {
    const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kTile / 8; k++) {
      const int ly = ly0 + 8 * k, x = x0 + lx, y = y0 + ly;
      if (x < width && y < height) {
        const float *p = rgb + 3 * ((int64_t)y * width + x);
        tile[ly][lx] = tone_pixel<kOp, kSlice>(rgb_t{__ldg(p), __ldg(p + 1), __ldg(p + 2)}, (float)x, (float)y, kc, sl);
      }
    }
  }
  __syncthreads();

  // store: walk the tile in destination order so that consecutive threads write consecutive destination bytes
  const int tf = a.transform;
  const bool swap = (tf == TDB_TF_ROTATE_90 || tf == TDB_TF_ROTATE_270 || tf == TDB_TF_TRANSPOSE);
  const int ow = swap ? height : width;
  const int dw = swap ? th : tw, dh = swap ? tw : th;                   // destination extent
  int ox0, oy0, ox1, oy1;
  map_xy(tf, x0, y0, width, height, ox0, oy0);
  map_xy(tf, x0 + tw - 1, y0 + th - 1, width, height, ox1, oy1);
  const int dx0 = min(ox0, ox1), dy0 = min(oy0, oy1);
  // source pixel that lands on destination (ox, oy)
  auto source = [&](int ox, int oy) -> uint32_t {
    int sx, sy;
    switch (tf) {
      case TDB_TF_ROTATE_90: sx = width - 1 - oy, sy = ox; break;
      case TDB_TF_ROTATE_180: sx = width - 1 - ox, sy = height - 1 - oy; break;
      case TDB_TF_ROTATE_270: sx = oy, sy = height - 1 - ox; break;
      case TDB_TF_TRANSPOSE: sx = oy, sy = ox; break;
      case TDB_TF_FLIP_HORIZ: sx = width - 1 - ox, sy = oy; break;
      case TDB_TF_FLIP_VERT: sx = ox, sy = height - 1 - oy; break;
      case TDB_TF_TRANSVERSE: sx = width - 1 - ox, sy = height - 1 - oy; break;
      default: sx = ox, sy = oy; break;
    }
    return tile[sy - y0][sx - x0];
  };
  if ((dw & 3) == 0 && (dx0 & 3) == 0 && (ow & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
    // 32-bit stores: a destination row of the tile is dw * 3 / 4 words, each made of two neighbouring pixels
    auto store_words = [&](const int wpr) {
      for (int i = threadIdx.x; i < wpr * dh; i += kThreads) {
        const int dy = i / wpr, wq = i - dy * wpr;
        const int g = wq / 3, ph = wq - 3 * g;  // pixel group of four, word inside its twelve bytes
        const int ox = dx0 + 4 * g + ph, oy = dy0 + dy;
        const uint32_t va = source(ox, oy), vb = source(ox + 1, oy);
        const uint32_t word = ph == 0 ? (va | (vb << 24)) : (ph == 1 ? ((va >> 8) | (vb << 16)) : ((va >> 16) | (vb << 8)));
        reinterpret_cast<uint32_t *>(out + 3 * ((int64_t)oy * ow + dx0))[wq] = word;
      }
    };
    // full-width tiles (all but the last tile column): the words per row are a compile-time constant, so the division by it is a
    // multiply-high instead of the I2F / MUFU.RCP / F2I sequence of a run-time divisor (ncu: 2 of them per stored word)
    if (dw == kTile) store_words(kTile * 3 / 4);
    else store_words(dw * 3 / 4);
  } else {
    for (int i = threadIdx.x; i < dw * dh; i += kThreads) {
      const int dy = i / dw, dx = i - dy * dw;
      const int ox = dx0 + dx, oy = dy0 + dy;
      const uint32_t v = source(ox, oy);
      uint8_t *o = out + 3 * ((int64_t)oy * ow + ox);
      o[0] = (uint8_t)v, o[1] = (uint8_t)(v >> 8), o[2] = (uint8_t)(v >> 16);
    }
  }
}

}  // namespace
}  // namespace tdb

using namespace tdb;

template <int kSlice>
static int launch_tonemap(const float *rgb, uint8_t *out, int width, int height, const TonemapArgs &a, const SliceArgs &sl, cudaStream_t s,
                          const char *name) {
  const bool rows = a.transform == TDB_TF_NONE || a.transform == TDB_TF_FLIP_HORIZ || a.transform == TDB_TF_FLIP_VERT ||
                    a.transform == TDB_TF_ROTATE_180 || a.transform == TDB_TF_TRANSVERSE;
  if (rows && width % 4 == 0 && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
    const int64_t groups = (int64_t)(width / 4) * height;
    const int g1 = (int)((groups + kThreads - 1) / kThreads < (int64_t)kNumSMs * 16 ? (groups + kThreads - 1) / kThreads : kNumSMs * 16);
    switch (a.op) {
      case TDB_TM_REINHARD: tonemap_rows_kernel<TDB_TM_REINHARD, kSlice><<<g1, kThreads, 0, s>>>(rgb, out, width, height, a, sl); break;
      case TDB_TM_ACES: tonemap_rows_kernel<TDB_TM_ACES, kSlice><<<g1, kThreads, 0, s>>>(rgb, out, width, height, a, sl); break;
      case TDB_TM_ADAPTIVE_ACES: tonemap_rows_kernel<TDB_TM_ADAPTIVE_ACES, kSlice><<<g1, kThreads, 0, s>>>(rgb, out, width, height, a, sl); break;
      case TDB_TM_LINEAR: tonemap_rows_kernel<TDB_TM_LINEAR, kSlice><<<g1, kThreads, 0, s>>>(rgb, out, width, height, a, sl); break;
      default: set_error("tonemap: unknown op %d", a.op); return TDB_EINVAL;
    }
    return check_launch(name);
  }
  dim3 grid(div_up(width, kTile), div_up(height, kTile));
  switch (a.op) {
    case TDB_TM_REINHARD: tonemap_kernel<TDB_TM_REINHARD, kSlice><<<grid, kThreads, 0, s>>>(rgb, out, width, height, a, sl); break;
    case TDB_TM_ACES: tonemap_kernel<TDB_TM_ACES, kSlice><<<grid, kThreads, 0, s>>>(rgb, out, width, height, a, sl); break;
    case TDB_TM_ADAPTIVE_ACES: tonemap_kernel<TDB_TM_ADAPTIVE_ACES, kSlice><<<grid, kThreads, 0, s>>>(rgb, out, width, height, a, sl); break;
    case TDB_TM_LINEAR: tonemap_kernel<TDB_TM_LINEAR, kSlice><<<grid, kThreads, 0, s>>>(rgb, out, width, height, a, sl); break;
    default: set_error("tonemap: unknown op %d", a.op); return TDB_EINVAL;
  }
  return check_launch(name);
}

extern "C" {

int tdb_tonemap(const float *rgb, uint8_t *out, int width, int height, int op, const float *metrics, float gamma, float intensity,
                float light_adapt, float vibrance, const float *matrix, int transform, tdb_stream_t stream) {
  TDB_REQUIRE(rgb && out, "tonemap: null pointer");
  TDB_REQUIRE(width > 0 && height > 0, "tonemap: empty image");
  TDB_REQUIRE(op == TDB_TM_ACES || metrics, "tonemap: metrics required");
  TDB_REQUIRE(transform >= TDB_TF_NONE && transform <= TDB_TF_TRANSVERSE, "tonemap: bad transform %d", transform);
  TonemapArgs a{gamma, intensity, light_adapt, vibrance, metrics, matrix, op, transform};
  return launch_tonemap<0>(rgb, out, width, height, a, SliceArgs{}, as_stream(stream), "tonemap");
}

int tdb_bilateral_slice_tonemap(const float *rgb, int lab_input, const void *bilateral_scratch, uint8_t *out, int width, int height,
                                float sigma_s, float sigma_r, float detail, int op, const float *metrics, float gamma, float intensity,
                                float light_adapt, float vibrance, const float *matrix, int transform, tdb_stream_t stream) {
  TDB_REQUIRE(rgb && out && bilateral_scratch, "slice_tonemap: null pointer");
  TDB_REQUIRE(width > 0 && height > 0 && sigma_r > 0.0f && sigma_s > 0.0f, "slice_tonemap: invalid dimensions or sigmas");
  TDB_REQUIRE(op == TDB_TM_ACES || metrics, "tonemap: metrics required");
  TDB_REQUIRE(transform >= TDB_TF_NONE && transform <= TDB_TF_TRANSVERSE, "tonemap: bad transform %d", transform);
  TonemapArgs a{gamma, intensity, light_adapt, vibrance, metrics, matrix, op, transform};
  const bil::GridDims g = bil::grid_dims(width, height, sigma_s, sigma_r);
  // scratch layout of bilateral.cu: [splatted grid][blurred grid]
  const float *blurred = static_cast<const float *>(bilateral_scratch) + (size_t)g.x * g.y * g.z;
  const SliceArgs sl{blurred, g, sigma_s, sigma_r, detail, bil::grid_limits(g)};
  if (lab_input) return launch_tonemap<2>(rgb, out, width, height, a, sl, as_stream(stream), "bilateral_slice_tonemap");
  return launch_tonemap<1>(rgb, out, width, height, a, sl, as_stream(stream), "bilateral_slice_tonemap");
}

}  // extern "C"
