// Bilateral-grid geometry and the trilinear slice, shared by bilateral.cu and the fused slice + tone-map kernel (tonemap.cu).
#pragma once

#include <cmath>

#include "color_math.cuh"

namespace tdb {
namespace bil {

struct GridDims {
  int x, y, z;
};

inline float clampf_host(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// reference bilateral.cu:273-299
inline GridDims grid_dims(int width, int height, float sigma_s, float sigma_r) {
  float ss = sigma_s;
  if (ss < 0.5f) ss = 0.5f;
  const float gx = clampf_host(roundf(width / ss), 4.0f, 3000.0f);
  const float gy = clampf_host(roundf(height / ss), 4.0f, 3000.0f);
  const float gz = clampf_host(roundf(1.0f / sigma_r), 4.0f, 50.0f);
  const float eff_s = fmaxf(height / gy, width / gx), eff_r = 1.0f / gz;
  return GridDims{(int)ceilf(width / eff_s) + 1, (int)ceilf(height / eff_s) + 1, (int)ceilf(1.0f / eff_r) + 1};
}

struct Sample {
  int ix, iy, iz;
  float fx, fy, fz;
};

// reference bilateral.cu:71-86: coordinates use the RAW sigmas and saturate at the last cell
__device__ __forceinline__ Sample make_sample(int x, int y, float L, GridDims g, float sigma_s, float sigma_r) {
  const float gx = fminf(fmaxf(x / sigma_s, 0.0f), (float)(g.x - 1));
  const float gy = fminf(fmaxf(y / sigma_s, 0.0f), (float)(g.y - 1));
  const float gz = fminf(fmaxf(L / sigma_r, 0.0f), (float)(g.z - 1));
  Sample s;
  s.ix = min((int)gx, g.x - 2), s.iy = min((int)gy, g.y - 2), s.iz = min((int)gz, g.z - 2);
  s.fx = gx - (float)s.ix, s.fy = gy - (float)s.iy, s.fz = gz - (float)s.iz;
  return s;
}
// The same sample for the slice of the frame pipeline's epilogue, which is bound by the XU pipe (MUFU and every int <-> float
// conversion): pixel coordinates arrive as floats, the grid limits as floats in the kernel arguments (constant bank), and
// floor() of the non-negative coordinate is taken with an addition of 2^23 rounded toward zero -- its low mantissa bits are the
// integer, the sum minus 2^23 is the float -- instead of the F2I / I2F pair.  Bit-identical to make_sample for 0 <= v < 2^22.
struct GridLimits {
  float x1, y1, z1;  // (float)(n - 1)
  float x2, y2, z2;  // (float)(n - 2)
};
inline GridLimits grid_limits(GridDims g) {
  return GridLimits{(float)(g.x - 1), (float)(g.y - 1), (float)(g.z - 1), (float)(g.x - 2), (float)(g.y - 2), (float)(g.z - 2)};
}
__device__ __forceinline__ void floor_split(float v, float limit2, int &i, float &frac) {
  const float t = __fadd_rz(v, 8388608.0f);
  const float fl = fminf(t - 8388608.0f, limit2);  // min((int)v, n - 2) as a float
  i = (int)(__float_as_uint(__fadd_rz(fl, 8388608.0f)) & 0x7fffffu);
  frac = v - fl;
}
__device__ __forceinline__ Sample make_sample(float xf, float yf, float L, const GridLimits &lim, float sigma_s, float sigma_r) {
  const float gx = fminf(fmaxf(xf / sigma_s, 0.0f), lim.x1);
  const float gy = fminf(fmaxf(yf / sigma_s, 0.0f), lim.y1);
  const float gz = fminf(fmaxf(L / sigma_r, 0.0f), lim.z1);
  Sample s;
  floor_split(gx, lim.x2, s.ix, s.fx), floor_split(gy, lim.y2, s.iy, s.fy), floor_split(gz, lim.z2, s.iz, s.fz);
  return s;
}
__device__ __forceinline__ int cell_of(int p, float sigma_s, int n) { return min((int)fminf(fmaxf(p / sigma_s, 0.0f), (float)(n - 1)), n - 2); }


// trilinear gather of the blurred grid at (x, y, L) and the contrast step (reference bilateral.cu:206-249):
// L' = max(0, L - detail * sigma_r * 4 * d)
__device__ __forceinline__ float slice_at(const float *__restrict__ grid, const Sample &s, float L, GridDims g, float sigma_r, float detail);
__device__ __forceinline__ float slice_luminance(const float *__restrict__ grid, int x, int y, float L, GridDims g, float sigma_s,
                                                 float sigma_r, float detail) {
  return slice_at(grid, make_sample(x, y, L, g, sigma_s, sigma_r), L, g, sigma_r, detail);
}
__device__ __forceinline__ float slice_at(const float *__restrict__ grid, const Sample &s, float L, GridDims g, float sigma_r, float detail) {
  const float ax = 1.0f - s.fx, ay = 1.0f - s.fy, az = 1.0f - s.fz;
  // 32-bit cell indices: a grid has at most 3001 x 3001 x 51 cells (grid_dims)
  const unsigned oy = (unsigned)g.x, oz = (unsigned)g.x * (unsigned)g.y;
  const unsigned i00 = (unsigned)s.ix + oy * ((unsigned)s.iy + (unsigned)g.y * (unsigned)s.iz), i10 = i00 + oy, i01 = i00 + oz, i11 = i01 + oy;
  const float *p00 = grid + i00, *p10 = grid + i10, *p01 = grid + i01, *p11 = grid + i11;
  const float d = __ldg(p00) * ax * ay * az + __ldg(p00 + 1) * s.fx * ay * az + __ldg(p10) * ax * s.fy * az +
                  __ldg(p10 + 1) * s.fx * s.fy * az + __ldg(p01) * ax * ay * s.fz + __ldg(p01 + 1) * s.fx * ay * s.fz +
                  __ldg(p11) * ax * s.fy * s.fz + __ldg(p11 + 1) * s.fx * s.fy * s.fz;
  const float norm = -detail * sigma_r * 4.0f;
  return fmaxf(0.0f, L + norm * d);
}

// Bilateral.process_rgb for one pixel: Lab L of the colour -> slice -> modify_luminance (local_contrast.py:110-114).
// compute_luminance clips the colour first; for a colour inside [0,1]^3 (always, after the Wiener write-back) the clip is the
// identity and its L is the L of the rgb_to_lab that modify_luminance needs anyway, so the linearisation and lab_f(Y) -- four of the
// thirteen pow() of this step -- are evaluated once.  Colours outside the cube take the literal path.
__device__ __forceinline__ rgb_t slice_rgb(const float *__restrict__ grid, int x, int y, rgb_t c, GridDims g, float sigma_s, float sigma_r,
                                           float detail) {
  const bool inside = c.x >= 0.0f && c.x <= 1.0f && c.y >= 0.0f && c.y <= 1.0f && c.z >= 0.0f && c.z <= 1.0f;
  const rgb_t v = pub::rgb_to_xyz(c);
  const float fx = pub::lab_f(v.x / 0.95047f), fy = pub::lab_f(v.y / 1.0f), fz = pub::lab_f(v.z / 1.08883f);
  float L = fmaxf(0.0f, (116.0f / 100.0f) * fy - (16.0f / 100.0f));
  if (!inside) L = pub::luminance(c);
  const float Lout = slice_luminance(grid, x, y, L, g, sigma_s, sigma_r, detail);
  const rgb_t lab{fmaxf(0.0f, fminf(1.0f, Lout)), (500.0f / 128.0f) * (fx - fy), (200.0f / 128.0f) * (fy - fz)};
  return clip01(pub::lab_to_rgb(lab));
}

// The same step for a colour that is already in Lab (what the fused Wiener write-back leaves behind for a clipped colour r:
// rgb_to_lab(r), whose L is compute_luminance(r) because r is inside [0,1]^3): no forward conversion at all.
__device__ __forceinline__ rgb_t slice_lab(const float *__restrict__ grid, int x, int y, rgb_t lab, GridDims g, float sigma_s, float sigma_r,
                                           float detail) {
  const float Lout = slice_luminance(grid, x, y, fmaxf(0.0f, lab.x), g, sigma_s, sigma_r, detail);
  return clip01(pub::lab_to_rgb(rgb_t{fmaxf(0.0f, fminf(1.0f, Lout)), lab.y, lab.z}));
}

// the two steps above with float pixel coordinates and float grid limits (see GridLimits)
__device__ __forceinline__ rgb_t slice_rgb(const float *__restrict__ grid, float xf, float yf, rgb_t c, GridDims g, const GridLimits &lim,
                                           float sigma_s, float sigma_r, float detail) {
  const bool inside = c.x >= 0.0f && c.x <= 1.0f && c.y >= 0.0f && c.y <= 1.0f && c.z >= 0.0f && c.z <= 1.0f;
  const rgb_t v = pub::rgb_to_xyz(c);
  const float fx = pub::lab_f(v.x / 0.95047f), fy = pub::lab_f(v.y / 1.0f), fz = pub::lab_f(v.z / 1.08883f);
  float L = fmaxf(0.0f, (116.0f / 100.0f) * fy - (16.0f / 100.0f));
  if (!inside) L = pub::luminance(c);
  const float Lout = slice_at(grid, make_sample(xf, yf, L, lim, sigma_s, sigma_r), L, g, sigma_r, detail);
  const rgb_t lab{fmaxf(0.0f, fminf(1.0f, Lout)), (500.0f / 128.0f) * (fx - fy), (200.0f / 128.0f) * (fy - fz)};
  return clip01(pub::lab_to_rgb(lab));
}
__device__ __forceinline__ rgb_t slice_lab(const float *__restrict__ grid, float xf, float yf, rgb_t lab, GridDims g, const GridLimits &lim,
                                           float sigma_s, float sigma_r, float detail) {
  const float L = fmaxf(0.0f, lab.x);
  const float Lout = slice_at(grid, make_sample(xf, yf, L, lim, sigma_s, sigma_r), L, g, sigma_r, detail);
  return clip01(pub::lab_to_rgb(rgb_t{fmaxf(0.0f, fminf(1.0f, Lout)), lab.y, lab.z}));
}

// splat of one pixel into the grid with native red.global.add.f32 (reference bilateral.cu:89-129).  A zero weight leaves the cell
// unchanged, so its atomic is skipped: for integer sigma_s the fractions are multiples of 1/sigma_s and on average only 4.5 of the 8
// corners carry weight at sigma_s = 2
__device__ __forceinline__ void splat_pixel(float *__restrict__ grid, int x, int y, float L, GridDims g, float sigma_s, float sigma_r) {
  const float contrib = 1.0f / (sigma_s * sigma_s);
  const Sample s = make_sample(x, y, L, g, sigma_s, sigma_r);
  const float ax = 1.0f - s.fx, ay = 1.0f - s.fy, az = 1.0f - s.fz;
  const int ox = 1, oy = g.x;
  const int64_t oz = (int64_t)g.x * g.y;
  float *base = grid + s.ix + (int64_t)g.x * (s.iy + (int64_t)g.y * s.iz);
  const float w000 = ax * ay * az * contrib, w100 = s.fx * ay * az * contrib, w010 = ax * s.fy * az * contrib,
              w110 = s.fx * s.fy * az * contrib, w001 = ax * ay * s.fz * contrib, w101 = s.fx * ay * s.fz * contrib,
              w011 = ax * s.fy * s.fz * contrib, w111 = s.fx * s.fy * s.fz * contrib;
  if (w000 != 0.0f) atomicAdd(base, w000);
  if (w100 != 0.0f) atomicAdd(base + ox, w100);
  if (w010 != 0.0f) atomicAdd(base + oy, w010);
  if (w110 != 0.0f) atomicAdd(base + oy + ox, w110);
  if (w001 != 0.0f) atomicAdd(base + oz, w001);
  if (w101 != 0.0f) atomicAdd(base + oz + ox, w101);
  if (w011 != 0.0f) atomicAdd(base + oz + oy, w011);
  if (w111 != 0.0f) atomicAdd(base + oz + oy + ox, w111);
}

}  // namespace bil

// host-side pieces of bilateral.cu used by the fused frame pipeline (scratch = [splat grid][blurred grid][luminance plane])
float *bilateral_lum_plane(void *scratch, bil::GridDims g);
const float *bilateral_blurred(const void *scratch, bil::GridDims g);
int bilateral_zero_grid(void *scratch, bil::GridDims g, cudaStream_t s);
int bilateral_blur(void *scratch, bil::GridDims g, cudaStream_t s);
int bilateral_build_grid(void *scratch, const float *lum, int width, int height, bil::GridDims g, float sigma_s, float sigma_r,
                         cudaStream_t s);

}  // namespace tdb
