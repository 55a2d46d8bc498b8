// Bilateral-grid local contrast (darktable "bilateral" / local contrast), splat -> blur -> slice.
//
// Reference: csrc/local_contrast/bilateral.cu:358-385 = zero_ + splat (8 global float atomics per pixel) + two 5-tap blur
// passes whose thread-x walks the z axis (uncoalesced) + z-derivative pass + slice; Bilateral.process_rgb adds a
// compute_luminance pass before and a modify_luminance pass after (local_contrast.py:110-114), about 116 B/px in total.
// Here:
//   splat : a CTA privatises the grid cells under its 64x64 pixel tile in shared memory, then flushes one atomic per
//           cell (about 1.2-1.6 global atomics per pixel instead of 8); the RGB variant computes Lab L on the fly;
//   blur  : x, y (1-4-6-4-1) and the z derivative filter fused into one shared-memory pass over the grid (8 B/cell
//           instead of 24 B/cell), thread-x along the contiguous x axis;
//   slice : trilinear gather; the RGB variant recomputes L and applies modify_luminance in the same pass.
// Algorithmic traffic: lum->lum 12 B/px, rgb->rgb 36 B/px, plus 2 x grid.
#include "bilateral.cuh"

namespace tdb {
namespace {

using namespace bil;

constexpr int kThreads = 256;
constexpr int TP = 64;  // splat pixel tile edge

template <bool kRgb>
__device__ __forceinline__ float load_lum(const float *__restrict__ in, int64_t idx) {
  if (kRgb) return pub::luminance(rgb_t{__ldg(in + 3 * idx), __ldg(in + 3 * idx + 1), __ldg(in + 3 * idx + 2)});
  return __ldg(in + idx);
}

template <bool kRgb>
__global__ void __launch_bounds__(kThreads) splat_kernel(const float *__restrict__ in, float *__restrict__ grid, int width, int height,
                                                         GridDims g, float sigma_s, float sigma_r) {
  // (a CTA-private copy of the cells in shared memory was tried: shared-memory float atomicAdd compiles to a CAS spin loop on
  // sm_100a, ATOMS.CAST.SPIN, and loses to native red.global.add.f32 on an L2-resident grid)
  const int x0 = blockIdx.x * TP, y0 = blockIdx.y * TP;
  const int x1 = min(x0 + TP, width), y1 = min(y0 + TP, height);
  const int tw = x1 - x0, th = y1 - y0;
  for (int i = threadIdx.x; i < tw * th; i += kThreads) {
    const int ly = i / tw, lx = i - ly * tw;
    const int x = x0 + lx, y = y0 + ly;
    splat_pixel(grid, x, y, load_lum<kRgb>(in, (int64_t)y * width + x), g, sigma_s, sigma_r);
  }
}

// fused x / y gaussian (1-4-6-4-1)/16 and z derivative (-2,-4,0,4,2)/16, zero beyond the grid (bilateral.cu:132-203).
// Thread (tx, ty) of a 32 x 8 CTA owns grid column (x, y): the x taps come straight from global memory (five coalesced,
// L1-resident loads), the x-blurred rows of the 32 x 12 patch go through shared memory once, and the y blur feeds a
// five-deep register window along z that produces the derivative: one barrier, no index arithmetic in the loops.
constexpr int BX = 32, BY = 8, PYB = BY + 4;
__global__ void __launch_bounds__(kThreads) blur_kernel(const float *__restrict__ in, float *__restrict__ out, GridDims g) {
  extern __shared__ float sm[];  // [z][PYB][BX] after the x pass
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = blockIdx.x * BX + tx, y0 = blockIdx.y * BY;
  const int64_t plane = (int64_t)g.x * g.y;
  const float w0 = 6.0f / 16.0f, w1 = 4.0f / 16.0f, w2 = 1.0f / 16.0f;
  const bool xm2 = x - 2 >= 0 && x - 2 < g.x, xm1 = x - 1 >= 0 && x - 1 < g.x, xc = x < g.x, xp1 = x + 1 < g.x, xp2 = x + 2 < g.x;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int ly = ty + 8 * k;  // patch rows ty and ty + 8 (the latter only for ty < 4)
    if (ly >= PYB) break;
    const int y = y0 - 2 + ly;
    const bool row = y >= 0 && y < g.y;
    const float *p = in + (int64_t)y * g.x + x;
    float *d = sm + ly * BX + tx;
    for (int z = 0; z < g.z; z++, p += plane, d += PYB * BX) {
      float v = 0.0f;
      if (row) {
        const float a = xm2 ? __ldg(p - 2) : 0.0f, b = xm1 ? __ldg(p - 1) : 0.0f, c = xc ? __ldg(p) : 0.0f;
        const float e = xp1 ? __ldg(p + 1) : 0.0f, f = xp2 ? __ldg(p + 2) : 0.0f;
        v = c * w0 + w1 * (e + b) + w2 * (f + a);
      }
      *d = v;
    }
  }
  __syncthreads();
  const int y = y0 + ty;
  if (x >= g.x || y >= g.y) return;
  const float d1 = 4.0f / 16.0f, d2 = 2.0f / 16.0f;
  const float *b = sm + (ty + 2) * BX + tx;
  float *o = out + (int64_t)y * g.x + x;
  // register window over z: m2, m1, c0, p1 hold the xy-blurred values at z-4 .. z-1 when plane z is read
  float m2 = 0.0f, m1 = 0.0f, c0 = 0.0f, p1 = 0.0f;
  for (int z = 0; z < g.z + 2; z++, b += PYB * BX) {
    float p2 = 0.0f;
    if (z < g.z) p2 = b[0] * w0 + w1 * (b[BX] + b[-BX]) + w2 * (b[2 * BX] + b[-2 * BX]);
    // output plane z - 2: neighbours z-1 (p1), z-3 (m1), z (p2), z-4 (m2)
    if (z >= 2) o[plane * (z - 2)] = d1 * (p1 - m1) + d2 * (p2 - m2);
    m2 = m1, m1 = c0, c0 = p1, p1 = p2;
  }
}

template <bool kRgb>
__global__ void __launch_bounds__(kThreads) slice_kernel(const float *__restrict__ in, const float *__restrict__ grid, float *__restrict__ out,
                                                         int width, int height, GridDims g, float sigma_s, float sigma_r, float detail) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= width || y >= height) return;
  const int64_t idx = (int64_t)y * width + x;
  if (kRgb) {
    const rgb_t r = slice_rgb(grid, x, y, rgb_t{__ldg(in + 3 * idx), __ldg(in + 3 * idx + 1), __ldg(in + 3 * idx + 2)}, g, sigma_s, sigma_r, detail);
    out[3 * idx] = r.x, out[3 * idx + 1] = r.y, out[3 * idx + 2] = r.z;
  } else {
    out[idx] = slice_luminance(grid, x, y, __ldg(in + idx), g, sigma_s, sigma_r, detail);
  }
}

template <bool kRgb>
int run_bilateral(const float *in, float *out, void *scratch, int width, int height, float sigma_s, float sigma_r, float detail,
                  cudaStream_t s) {
  const GridDims g = grid_dims(width, height, sigma_s, sigma_r);
  const size_t cells = (size_t)g.x * g.y * g.z;
  float *grid = static_cast<float *>(scratch), *blurred = grid + cells;
  if (int e = bilateral_zero_grid(scratch, g, s)) return e;
  dim3 sgrid(div_up(width, TP), div_up(height, TP));
  splat_kernel<kRgb><<<sgrid, kThreads, 0, s>>>(in, grid, width, height, g, sigma_s, sigma_r);
  if (int e = check_launch("bilateral_splat")) return e;
  if (int e = bilateral_blur(scratch, g, s)) return e;
  dim3 pgrid(div_up(width, 32), div_up(height, 8));
  slice_kernel<kRgb><<<pgrid, kThreads, 0, s>>>(in, blurred, out, width, height, g, sigma_s, sigma_r, detail);
  return check_launch("bilateral_slice");
}

}  // namespace

int bilateral_zero_grid(void *scratch, bil::GridDims g, cudaStream_t s) {
  cudaMemsetAsync(scratch, 0, (size_t)g.x * g.y * g.z * sizeof(float), s);
  return check_launch("bilateral_zero_grid");
}

int bilateral_blur(void *scratch, bil::GridDims g, cudaStream_t s) {
  static bool attr = false;
  const size_t cells = (size_t)g.x * g.y * g.z;
  float *grid = static_cast<float *>(scratch), *blurred = grid + cells;
  const size_t bytes = (size_t)g.z * PYB * BX * sizeof(float);
  if (!attr) {
    cudaFuncSetAttribute(blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 51 * PYB * BX * 4);
    attr = true;
  }
  dim3 bgrid(div_up(g.x, BX), div_up(g.y, BY));
  blur_kernel<<<bgrid, kThreads, bytes, s>>>(grid, blurred, g);
  return check_launch("bilateral_blur");
}

}  // namespace tdb

using namespace tdb;

extern "C" {

int tdb_bilateral_grid_size(int width, int height, float sigma_s, float sigma_r, int size[3]) {
  TDB_REQUIRE(width > 0 && height > 0 && sigma_r > 0.0f && sigma_s > 0.0f, "Bilateral: invalid dimensions or sigmas");
  const GridDims g = grid_dims(width, height, sigma_s, sigma_r);
  size[0] = g.x, size[1] = g.y, size[2] = g.z;
  return TDB_OK;
}

size_t tdb_bilateral_scratch_bytes(int width, int height, float sigma_s, float sigma_r) {
  if (width <= 0 || height <= 0 || !(sigma_s > 0.0f) || !(sigma_r > 0.0f)) return 0;
  const GridDims g = grid_dims(width, height, sigma_s, sigma_r);
  return 2 * (size_t)g.x * g.y * g.z * sizeof(float);
}

int tdb_bilateral(const float *lum, float *out, void *scratch, int width, int height, float sigma_s, float sigma_r, float detail,
                  tdb_stream_t stream) {
  TDB_REQUIRE(lum && out && scratch, "Bilateral: null pointer");
  TDB_REQUIRE(width > 0 && height > 0 && sigma_r > 0.0f && sigma_s > 0.0f, "Bilateral: invalid dimensions or sigmas");
  return run_bilateral<false>(lum, out, scratch, width, height, sigma_s, sigma_r, detail, as_stream(stream));
}

int tdb_bilateral_rgb(const float *rgb, float *out, void *scratch, int width, int height, float sigma_s, float sigma_r, float detail,
                      tdb_stream_t stream) {
  TDB_REQUIRE(rgb && out && scratch, "Bilateral: null pointer");
  TDB_REQUIRE(width > 0 && height > 0 && sigma_r > 0.0f && sigma_s > 0.0f, "Bilateral: invalid dimensions or sigmas");
  return run_bilateral<true>(rgb, out, scratch, width, height, sigma_s, sigma_r, detail, as_stream(stream));
}

int tdb_bilateral_grid_rgb(const float *rgb, void *scratch, int width, int height, float sigma_s, float sigma_r, tdb_stream_t stream) {
  TDB_REQUIRE(rgb && scratch, "Bilateral: null pointer");
  TDB_REQUIRE(width > 0 && height > 0 && sigma_r > 0.0f && sigma_s > 0.0f, "Bilateral: invalid dimensions or sigmas");
  cudaStream_t s = as_stream(stream);
  const bil::GridDims g = bil::grid_dims(width, height, sigma_s, sigma_r);
  if (int e = bilateral_zero_grid(scratch, g, s)) return e;
  dim3 sgrid(div_up(width, TP), div_up(height, TP));
  splat_kernel<true><<<sgrid, kThreads, 0, s>>>(rgb, static_cast<float *>(scratch), width, height, g, sigma_s, sigma_r);
  if (int e = check_launch("bilateral_splat")) return e;
  return bilateral_blur(scratch, g, s);
}

}  // extern "C"
