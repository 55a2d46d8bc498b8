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

// kSlice: the pixel is produced by the bilateral slice (Bilateral.process_rgb's last step) instead of being read as is, so the
// locally contrasted image is never written to HBM: 12 B read + 3 B written per pixel for slice + tone map together.
struct SliceArgs {
  const float *grid;  // blurred bilateral grid
  bil::GridDims g;
  float sigma_s, sigma_r, detail;
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

template <int kOp, bool kSlice>
__global__ void __launch_bounds__(kThreads) tonemap_kernel(const float *__restrict__ rgb, uint8_t *__restrict__ out, int width,
                                                           int height, TonemapArgs a, SliceArgs sl) {
  __shared__ uint32_t tile[kTile][kTile + 1];  // 0x00BBGGRR per pixel, indexed [y][x] in SOURCE tile coordinates

  float m[9];
  const bool has_matrix = a.matrix != nullptr;
  if (has_matrix) {
#pragma unroll
    for (int i = 0; i < 9; i++) m[i] = __ldg(a.matrix + i);
  }
  float map_key = 0.0f, exposure = 1.0f, mean[3] = {0, 0, 0};
  if (kOp != TDB_TM_ACES) {  // reference color_adaption.h:17-44
    const float normalized = fmaxf(0.0f, fminf(1.0f, (-__ldg(a.metrics)) / 9.21034f));
    map_key = 0.3f + 0.7f * powf(normalized, 1.4f);
    exposure = expf(a.intensity);
    mean[0] = __ldg(a.metrics + 2), mean[1] = __ldg(a.metrics + 3), mean[2] = __ldg(a.metrics + 4);
  }
  const float aces_gain = powf(2.0f, a.intensity);
  const float inv_gamma = 1.0f / a.gamma;

  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kTile / 8; k++) {
    const int ly = ly0 + 8 * k, x = x0 + lx, y = y0 + ly;
    if (x < width && y < height) {
      const float *p = rgb + 3 * ((int64_t)y * width + x);
      rgb_t c{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
      if (kSlice) c = bil::slice_rgb(sl.grid, x, y, c, sl.g, sl.sigma_s, sl.sigma_r, sl.detail);
      if (has_matrix) c = mat3(m, c);
      rgb_t t;
      if (kOp == TDB_TM_ACES) {
        t = tm::aces_fit(rgb_t{c.x * aces_gain, c.y * aces_gain, c.z * aces_gain});
      } else {
        // lerp(light_adapt, global_mean, pixel) / exposure, to the power map_key  (color_adaption.h:46-76)
        const float ax = powf((mean[0] + a.light_adapt * (c.x - mean[0])) / exposure, map_key);
        const float ay = powf((mean[1] + a.light_adapt * (c.y - mean[1])) / exposure, map_key);
        const float az = powf((mean[2] + a.light_adapt * (c.z - mean[2])) / exposure, map_key);
        if (kOp == TDB_TM_REINHARD) t = rgb_t{c.x / (ax + c.x), c.y / (ay + c.y), c.z / (az + c.z)};
        else if (kOp == TDB_TM_LINEAR) t = rgb_t{c.x / ax, c.y / ay, c.z / az};
        else t = tm::aces_fit(rgb_t{c.x / ax, c.y / ay, c.z / az});
      }
      const rgb_t g{powf(fmaxf(t.x, 0.0f), inv_gamma), powf(fmaxf(t.y, 0.0f), inv_gamma), powf(fmaxf(t.z, 0.0f), inv_gamma)};
      const rgb_t v = tm::vibrance(g, a.vibrance);
      tile[ly][lx] = tm::to_u8(v.x) | (tm::to_u8(v.y) << 8) | (tm::to_u8(v.z) << 16);
    }
  }
  __syncthreads();

  // store: walk the tile in destination order so that consecutive threads write consecutive destination pixels
  const int tf = a.transform;
  const bool swap = (tf == TDB_TF_ROTATE_90 || tf == TDB_TF_ROTATE_270 || tf == TDB_TF_TRANSPOSE);
  const int ow = swap ? height : width;
  const int tw = min(kTile, width - x0), th = min(kTile, height - y0);  // valid source extent of this tile
  const int dw = swap ? th : tw, dh = swap ? tw : th;                   // destination extent
  int ox0, oy0, ox1, oy1;
  map_xy(tf, x0, y0, width, height, ox0, oy0);
  map_xy(tf, x0 + tw - 1, y0 + th - 1, width, height, ox1, oy1);
  const int dx0 = min(ox0, ox1), dy0 = min(oy0, oy1);
  for (int i = threadIdx.x; i < dw * dh; i += kThreads) {
    const int dy = i / dw, dx = i - dy * dw;
    const int ox = dx0 + dx, oy = dy0 + dy;
    int sx, sy;  // inverse map: which source pixel lands on (ox, oy)
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
    const uint32_t v = tile[sy - y0][sx - x0];
    uint8_t *o = out + 3 * ((int64_t)oy * ow + ox);
    o[0] = (uint8_t)v, o[1] = (uint8_t)(v >> 8), o[2] = (uint8_t)(v >> 16);
  }
}

}  // namespace
}  // namespace tdb

using namespace tdb;

template <bool kSlice>
static int launch_tonemap(const float *rgb, uint8_t *out, int width, int height, const TonemapArgs &a, const SliceArgs &sl, cudaStream_t s,
                          const char *name) {
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
  return launch_tonemap<false>(rgb, out, width, height, a, SliceArgs{}, as_stream(stream), "tonemap");
}

int tdb_bilateral_slice_tonemap(const float *rgb, const void *bilateral_scratch, uint8_t *out, int width, int height, float sigma_s,
                                float sigma_r, float detail, int op, const float *metrics, float gamma, float intensity,
                                float light_adapt, float vibrance, const float *matrix, int transform, tdb_stream_t stream) {
  TDB_REQUIRE(rgb && out && bilateral_scratch, "slice_tonemap: null pointer");
  TDB_REQUIRE(width > 0 && height > 0 && sigma_r > 0.0f && sigma_s > 0.0f, "slice_tonemap: invalid dimensions or sigmas");
  TDB_REQUIRE(op == TDB_TM_ACES || metrics, "tonemap: metrics required");
  TDB_REQUIRE(transform >= TDB_TF_NONE && transform <= TDB_TF_TRANSVERSE, "tonemap: bad transform %d", transform);
  TonemapArgs a{gamma, intensity, light_adapt, vibrance, metrics, matrix, op, transform};
  const bil::GridDims g = bil::grid_dims(width, height, sigma_s, sigma_r);
  // scratch layout of bilateral.cu: [splatted grid][blurred grid]
  const float *blurred = static_cast<const float *>(bilateral_scratch) + (size_t)g.x * g.y * g.z;
  return launch_tonemap<true>(rgb, out, width, height, a, SliceArgs{blurred, g, sigma_s, sigma_r, detail}, as_stream(stream),
                              "bilateral_slice_tonemap");
}

}  // extern "C"
