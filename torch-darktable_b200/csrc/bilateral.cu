// Bilateral-grid local contrast (darktable "bilateral" / local contrast), splat -> blur -> slice.
//
// Reference: csrc/local_contrast/bilateral.cu:358-385 = zero_ + splat (8 global float atomics per pixel) + two 5-tap blur
// passes whose thread-x walks the z axis (uncoalesced) + z-derivative pass + slice; Bilateral.process_rgb adds a
// compute_luminance pass before and a modify_luminance pass after (local_contrast.py:110-114), about 116 B/px in total.
// Here:
//   grid  : splat and blur in ONE kernel without atomics and without the splatted grid in HBM: every grid column gathers its
//           pixels into private bins in shared memory and the CTA filters its columns on the spot (grid_build_kernel below);
//           the scatter with red.global.add.f32 + the separate fused blur remain for saturating grids;
//   slice : trilinear gather; the RGB variant applies modify_luminance in the same pass and shares the Lab conversion with
//           the luminance it needs for the lookup.
// Algorithmic traffic: lum->lum 12 B/px, rgb->rgb 36 B/px, plus the grid once each way.
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

// ---- splat + blur without atomics ------------------------------------------------------------------------------------
// The splat is a scatter with data-dependent z only: along x and y a pixel feeds the two cells around p / sigma_s with the hat
// weights (1 - f, f), whatever its value.  So a grid COLUMN (i, j) can gather instead: it visits the pixels whose cell index is i - 1
// or i (j - 1 or j), at most ceil(2 sigma_s) per axis, and adds every pixel's two z contributions into its own bins in shared
// memory -- no atomics, no zeroed grid, and the splatted grid never exists in HBM, because the CTA gathers its 32 x 8 columns plus
// the blur halo (36 x 12) and runs the x / y / z-derivative filter of blur_kernel on them straight away.
// At sigma_s = 8 the atomic splat spends 1.2 ms on a 50 MP frame fighting over 64 pixels per cell; this kernel has no such term.
// Per-pixel weights are the products of the scatter (w_x * w_y * w_z * contrib, same order); only the order of the additions
// differs, and that order was arbitrary before (atomics) and is fixed now.
// Not for saturating grids (more than 3000 cells per axis wanted, bilateral.cu:282-284: all pixels beyond the last cell pile up in
// it) nor for very fine ones (sigma_s < 1); those keep the scatter.
struct AxisSample {
  int i;
  float f;
};
__device__ __forceinline__ AxisSample axis_sample(int p, float sigma_s, int n) {  // the x / y part of make_sample
  const float gp = fminf(fmaxf(p / sigma_s, 0.0f), (float)(n - 1));
  const int i = min((int)gp, n - 2);
  return AxisSample{i, gp - (float)i};
}
constexpr int GPX = BX + 4;  // gathered columns per CTA row (blur halo 2)
constexpr int kMaxAxisPx = 1024;  // pixels a CTA's columns can span along one axis: (GPX + 1) * sigma_s + 4, sigma_s <= 24

// kBY: grid rows per CTA (8, or 16 for shallow grids: less halo work, 1.41x instead of 1.69x gathered columns per output column)
// kTables: fine grids (sigma_s < 3) look the per-pixel cell index / fraction up in per-CTA tables; for coarse grids neighbouring
// columns are sigma_s pixels apart and the table reads would collide in the same banks, so they recompute them instead
// kS2: sigma_s == 2 exactly (every camera preset of the reference): cell c is fed by the pixels 2c - 1, 2c, 2c + 1 with weights
// 1/2, 1, 1/2 -- no tables, no ranges, nine unrolled visits whose x / y weights are immediates (powers of two: the products are the
// ones of the general path bit for bit)
template <int kBY, bool kTables, bool kS2 = false>
__global__ void __launch_bounds__(kThreads) grid_build_kernel(const float *__restrict__ lum, float *__restrict__ out, int width, int height,
                                                              GridDims g, float sigma_s, float sigma_r, int pitch, const float *__restrict__ pile_x,
                                                              const float *__restrict__ pile_y) {
  // width / height: the pixels the gather may visit (for a saturating grid only those in front of the piles, see pile_rows_kernel);
  // pitch: floats per image row; pile_x[j][z] / pile_y[i][z]: what the piled pixels add to the last cell column / row (or null)
  constexpr int PY = kBY + 4;
  extern __shared__ float sm[];
  float *cells = sm;                     // [z][PY][GPX] splatted columns of this CTA
  // plane stride of the bins: a multiple of 32 words, so that the bank of a bin depends on the column only -- the lanes of a warp own
  // consecutive columns but hit different z planes (data dependent), and with the natural stride (720 = 16 mod 32) lanes whose
  // z differ by an odd number collided
  constexpr int ZS = (PY * GPX + 31) / 32 * 32;
  float *xb = sm + g.z * ZS;             // [z][PY][BX]  after the x pass
  // per-axis tables, built once per CTA: cell index and fraction of every pixel the CTA can touch, pixel range of every column
  __shared__ AxisSample ax_x[kMaxAxisPx], ax_y[kMaxAxisPx];
  __shared__ short2 rng_x[GPX], rng_y[PY];
  const int ci0 = blockIdx.x * BX - 2, cj0 = blockIdx.y * kBY - 2;
  const int xbase = max(0, (int)floorf((float)(ci0 - 1) * sigma_s) - 1), ybase = max(0, (int)floorf((float)(cj0 - 1) * sigma_s) - 1);
  const int xn = min(width - xbase, min(kMaxAxisPx, (int)((GPX + 1) * sigma_s) + 8));
  const int yn = min(height - ybase, min(kMaxAxisPx, (int)((PY + 1) * sigma_s) + 8));
  if (kTables && !kS2) {
    for (int t = threadIdx.x; t < max(xn, yn); t += kThreads) {
      if (t < xn) ax_x[t] = axis_sample(xbase + t, sigma_s, g.x);
      if (t < yn) ax_y[t] = axis_sample(ybase + t, sigma_s, g.y);
    }
    __syncthreads();
  }
  if (!kS2 && threadIdx.x < GPX + PY) {  // inclusive pixel range (relative to the base) of the pixels whose lower cell is c - 1 or c
    const bool is_x = threadIdx.x < GPX;
    const int c = is_x ? ci0 + threadIdx.x : cj0 + (threadIdx.x - GPX);
    const int n = is_x ? xn : yn, base = is_x ? xbase : ybase, cells_n = is_x ? g.x : g.y;
    auto feeds = [&](int t) {
      const AxisSample a = axis_sample(base + t, sigma_s, cells_n);
      return (a.i == c && a.f != 1.0f) || (a.i + 1 == c && a.f != 0.0f);
    };
    int lo = max(0, (int)floorf((float)(c - 1) * sigma_s) - 1 - base), hi = min(n - 1, (int)ceilf((float)(c + 1) * sigma_s) + 1 - base);
    while (lo <= hi && !feeds(lo)) lo++;
    while (hi >= lo && !feeds(hi)) hi--;
    if (is_x) rng_x[threadIdx.x] = make_short2((short)lo, (short)hi);
    else rng_y[threadIdx.x - GPX] = make_short2((short)lo, (short)hi);
  }
  __syncthreads();
  const float contrib = 1.0f / (sigma_s * sigma_s);
  const float inv_sigma_r = rcp_approx(sigma_r);  // v / sigma_r == v * inv_sigma_r under --use_fast_math
  // (Staging the clamped z coordinate of the CTA's 73 x 41 pixels in shared memory first -- even / odd pixel columns as separate
  // planes, conflict-free, each pixel loaded and clamped once instead of by up to four columns -- was measured for sigma_s = 2:
  // 0.077 -> 0.082 ms; the extra pass and barrier cost more than the L1-served loads they replace.)
  // (dealing the pixel rows of a column to several threads with private bins was measured for sigma_s = 8: no gain, the coarse
  // case is bound by its sigma_s-strided luminance loads, not by the length of the per-column chains)
  for (int col = threadIdx.x; col < PY * GPX; col += kThreads) {
    const int lj = col / GPX, li = col - lj * GPX;
    const int i = ci0 + li, j = cj0 + lj;
    float *bins = cells + lj * GPX + li;
    const bool in_grid = i >= 0 && j >= 0 && i < g.x && j < g.y;
    if (pile_x != nullptr && in_grid && (i == g.x - 1 || j == g.y - 1)) {  // last cell column / row of a saturating grid: start from the piles
      for (int z = 0; z < g.z; z++)
        bins[z * ZS] = (i == g.x - 1 ? __ldg(pile_x + j * g.z + z) : 0.0f) + (j == g.y - 1 ? __ldg(pile_y + i * g.z + z) : 0.0f);
    } else {
      for (int z = 0; z < g.z; z++) bins[z * ZS] = 0.0f;
    }
    if (!in_grid) continue;
    if (kS2) {
      if (i >= 1 && 2 * i + 1 < width && j >= 1 && 2 * j + 1 < height) {
        // All nine pixels lie inside the gathered image (every column but those of the grid's rim).  ncu on the loop below: every
        // visit sat in its own branch region -- its load could not start before the previous visit's additions had retired (22 %
        // of the stall samples waited for those loads), and each recomputed 1 / sigma_r and the shared-memory window base (41
        // instructions per visit).  Here the nine loads are issued back to back, the reciprocal is taken once per thread, and the
        // visits run without tests, in the same order with the same products: bit-identical bins.
        const float *c = lum + (int64_t)(2 * j) * pitch + 2 * i;
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; k++) v[k] = __ldg(c + (int64_t)(k / 3 - 1) * pitch + (k % 3 - 1));
#pragma unroll
        for (int k = 0; k < 9; k++) {
          const float wy = k / 3 == 1 ? 1.0f : 0.5f, wx = k % 3 == 1 ? 1.0f : 0.5f;
          const float gz = fminf(fmaxf(v[k] * inv_sigma_r, 0.0f), (float)(g.z - 1));
          const int iz = min((int)gz, g.z - 2);
          const float fz = gz - (float)iz, az = 1.0f - fz;
          const float w0 = wx * wy * az * contrib, w1 = wx * wy * fz * contrib;
          float *b = bins + iz * ZS;
          if (w0 != 0.0f) b[0] += w0;
          if (w1 != 0.0f) b[ZS] += w1;
        }
        continue;
      }
#pragma unroll
      for (int dy = -1; dy <= 1; dy++) {
        const int y = 2 * j + dy;
        if (y < 0 || y >= height) continue;
        const float wy = dy == 0 ? 1.0f : 0.5f;
        const float *row = lum + (int64_t)y * pitch + 2 * i;
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          if (2 * i + dx < 0 || 2 * i + dx >= width) continue;
          const float wx = dx == 0 ? 1.0f : 0.5f;
          const float gz = fminf(fmaxf(__ldg(row + dx) / sigma_r, 0.0f), (float)(g.z - 1));
          const int iz = min((int)gz, g.z - 2);
          const float fz = gz - (float)iz, az = 1.0f - fz;
          const float w0 = wx * wy * az * contrib, w1 = wx * wy * fz * contrib;
          float *b = bins + iz * ZS;
          if (w0 != 0.0f) b[0] += w0;
          if (w1 != 0.0f) b[ZS] += w1;
        }
      }
      continue;
    }
    const short2 rx = rng_x[li], ry = rng_y[lj];
    for (int ty = ry.x; ty <= ry.y; ty++) {
      const AxisSample sy = kTables ? ax_y[ty] : axis_sample(ybase + ty, sigma_s, g.y);
      const float wy = sy.i == j ? 1.0f - sy.f : sy.f;
      const float *row = lum + (int64_t)(ybase + ty) * pitch + xbase;
#pragma unroll 4
      for (int tx = rx.x; tx <= rx.y; tx++) {
        const AxisSample sx = kTables ? ax_x[tx] : axis_sample(xbase + tx, sigma_s, g.x);
        const float wx = sx.i == i ? 1.0f - sx.f : sx.f;
        const float gz = fminf(fmaxf(__ldg(row + tx) / sigma_r, 0.0f), (float)(g.z - 1));
        const int iz = min((int)gz, g.z - 2);
        const float fz = gz - (float)iz, az = 1.0f - fz;
        const float w0 = wx * wy * az * contrib, w1 = wx * wy * fz * contrib;
        float *b = bins + iz * ZS;
        if (w0 != 0.0f) b[0] += w0;
        if (w1 != 0.0f) b[ZS] += w1;
      }
    }
  }
  __syncthreads();
  // x pass (cells outside the grid are zero, which is what blur_kernel's bounds tests produce)
  const float w0 = 6.0f / 16.0f, w1 = 4.0f / 16.0f, w2 = 1.0f / 16.0f;
  for (int t = threadIdx.x; t < g.z * PY * BX; t += kThreads) {
    const int tx = t % BX, r = t / BX;  // r = z * PY + ly
    const float *c = cells + (r / PY) * ZS + (r % PY) * GPX + tx + 2;
    xb[r * BX + tx] = c[0] * w0 + w1 * (c[1] + c[-1]) + w2 * (c[2] + c[-2]);
  }
  __syncthreads();
  const int tx = threadIdx.x & 31;
  const int x = blockIdx.x * BX + tx;
  if (x >= g.x) return;
  const int64_t plane = (int64_t)g.x * g.y;
  const float d1 = 4.0f / 16.0f, d2 = 2.0f / 16.0f;
  for (int ty = threadIdx.x >> 5; ty < kBY; ty += kThreads / 32) {
    const int y = blockIdx.y * kBY + ty;
    if (y >= g.y) break;
    const float *b = xb + (ty + 2) * BX + tx;
    float *o = out + (int64_t)y * g.x + x;
    float m2 = 0.0f, m1 = 0.0f, c0 = 0.0f, p1 = 0.0f;
    for (int z = 0; z < g.z + 2; z++, b += PY * BX) {
      float p2 = 0.0f;
      if (z < g.z) p2 = b[0] * w0 + w1 * (b[BX] + b[-BX]) + w2 * (b[2 * BX] + b[-2 * BX]);
      if (z >= 2) o[plane * (z - 2)] = d1 * (p1 - m1) + d2 * (p2 - m2);
      m2 = m1, m1 = c0, c0 = p1, p1 = p2;
    }
  }
}

// ---- saturating grids ------------------------------------------------------------------------------------------------------------
// The reference clamps the cell count to 3000 per axis but keeps sampling with the RAW sigma_s (bilateral.cu:273-299 + :71-86): at
// 8192 px and sigma_s = 2 every pixel with x >= 6000 has the clamped coordinate g.x - 1, i.e. lower cell g.x - 2 with fraction 1, and
// lands in the LAST cell column with x weight 1 (likewise rows y >= 4500; the corner cell collects 3.6 M pixels).  The scatter with
// global atomics serialises on those cells (2.9 ms of 3.3 ms at 50 MP).  Instead: the piled pixels are reduced separately into
//   pile_x[j][z]  everything the pixels x >= x_pile add to cell column g.x - 1   (one CTA per image row: its pixels share the two cell
//                 rows and y weights, so the CTA sums their z contributions privately and issues 2 g.z atomics)
//   pile_y[i][z]  what the pixels y >= y_pile, x < x_pile add to cell row g.y - 1 (one CTA per 32 image columns, summed down the rows)
// and the gather kernel, restricted to the pixels in front of the piles, seeds the bins of the last column / row with them.  Same
// weights as the scatter (w_x * w_y * w_z / sigma_s^2); only the order of the additions differs, which was arbitrary before.
__global__ void __launch_bounds__(kThreads) pile_rows_kernel(const float *__restrict__ lum, float *__restrict__ pile_x, int width, int height,
                                                             int x_pile, GridDims g, float sigma_s, float sigma_r) {
  extern __shared__ float psm[];  // [g.z][kThreads] private z bins, then the per-z totals
  const int y = blockIdx.x;
  for (int z = 0; z < g.z; z++) psm[z * kThreads + threadIdx.x] = 0.0f;
  const float contrib = 1.0f / (sigma_s * sigma_s);
  const float *row = lum + (int64_t)y * width;
  for (int x = x_pile + threadIdx.x; x < width; x += kThreads) {
    const float gz = fminf(fmaxf(__ldg(row + x) / sigma_r, 0.0f), (float)(g.z - 1));
    const int iz = min((int)gz, g.z - 2);
    const float fz = gz - (float)iz;
    psm[iz * kThreads + threadIdx.x] += (1.0f - fz) * contrib;
    psm[(iz + 1) * kThreads + threadIdx.x] += fz * contrib;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const AxisSample sy = axis_sample(y, sigma_s, g.y);
  for (int z = warp; z < g.z; z += kThreads / 32) {
    float v = 0.0f;
    for (int t = lane; t < kThreads; t += 32) v += psm[z * kThreads + t];
    v = warp_sum(v);
    if (lane == 0 && v != 0.0f) {
      if (sy.f != 1.0f) atomicAdd(pile_x + sy.i * g.z + z, (1.0f - sy.f) * v);
      if (sy.f != 0.0f) atomicAdd(pile_x + (sy.i + 1) * g.z + z, sy.f * v);
    }
  }
}
__global__ void __launch_bounds__(kThreads) pile_columns_kernel(const float *__restrict__ lum, float *__restrict__ pile_y, int width, int height,
                                                                int x_pile, int y_pile, GridDims g, float sigma_s, float sigma_r) {
  extern __shared__ float psm[];  // [g.z][8 row lanes][32 columns]
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + cx;
  for (int z = 0; z < g.z; z++) psm[z * kThreads + threadIdx.x] = 0.0f;
  const float contrib = 1.0f / (sigma_s * sigma_s);
  if (x < x_pile) {
    for (int y = y_pile + ry; y < height; y += kThreads / 32) {
      const float gz = fminf(fmaxf(__ldg(lum + (int64_t)y * width + x) / sigma_r, 0.0f), (float)(g.z - 1));
      const int iz = min((int)gz, g.z - 2);
      const float fz = gz - (float)iz;
      psm[iz * kThreads + threadIdx.x] += (1.0f - fz) * contrib;
      psm[(iz + 1) * kThreads + threadIdx.x] += fz * contrib;
    }
  }
  __syncthreads();
  if (ry == 0 && x < x_pile) {
    const AxisSample sx = axis_sample(x, sigma_s, g.x);
    for (int z = 0; z < g.z; z++) {
      float v = 0.0f;
#pragma unroll
      for (int r = 0; r < kThreads / 32; r++) v += psm[z * kThreads + r * 32 + cx];
      if (v != 0.0f) {
        if (sx.f != 1.0f) atomicAdd(pile_y + sx.i * g.z + z, (1.0f - sx.f) * v);
        if (sx.f != 0.0f) atomicAdd(pile_y + (sx.i + 1) * g.z + z, sx.f * v);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads) luminance_plane_kernel(const float *__restrict__ rgb, float *__restrict__ lum, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    lum[i] = pub::luminance(rgb_t{__ldg(rgb + 3 * i), __ldg(rgb + 3 * i + 1), __ldg(rgb + 3 * i + 2)});
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
  const float *lum = in;
  if (kRgb) {
    float *plane = bilateral_lum_plane(scratch, g);
    const int64_t n = (int64_t)width * height;
    luminance_plane_kernel<<<(int)((n + kThreads - 1) / kThreads < kNumSMs * 16 ? (n + kThreads - 1) / kThreads : kNumSMs * 16), kThreads, 0, s>>>(in, plane, n);
    if (int e = check_launch("bilateral_luminance")) return e;
    lum = plane;
  }
  if (int e = bilateral_build_grid(scratch, lum, width, height, g, sigma_s, sigma_r, s)) return e;
  dim3 pgrid(div_up(width, 32), div_up(height, 8));
  slice_kernel<kRgb><<<pgrid, kThreads, 0, s>>>(in, bilateral_blurred(scratch, g), out, width, height, g, sigma_s, sigma_r, detail);
  return check_launch("bilateral_slice");
}

}  // namespace

float *bilateral_lum_plane(void *scratch, bil::GridDims g) { return static_cast<float *>(scratch) + 2 * (size_t)g.x * g.y * g.z; }
const float *bilateral_blurred(const void *scratch, bil::GridDims g) { return static_cast<const float *>(scratch) + (size_t)g.x * g.y * g.z; }

int bilateral_zero_grid(void *scratch, bil::GridDims g, cudaStream_t s) {
  cudaMemsetAsync(scratch, 0, (size_t)g.x * g.y * g.z * sizeof(float), s);
  return check_launch("bilateral_zero_grid");
}

int bilateral_blur(void *scratch, bil::GridDims g, cudaStream_t s) {
  static DeviceOnce attr;
  const size_t cells = (size_t)g.x * g.y * g.z;
  float *grid = static_cast<float *>(scratch), *blurred = grid + cells;
  const size_t bytes = (size_t)g.z * PYB * BX * sizeof(float);
  attr.run([&] {
    cudaFuncSetAttribute(blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 51 * PYB * BX * 4);
  });
  dim3 bgrid(div_up(g.x, BX), div_up(g.y, BY));
  blur_kernel<<<bgrid, kThreads, bytes, s>>>(grid, blurred, g);
  return check_launch("bilateral_blur");
}

// splatted + blurred grid of a luminance plane, into the second half of the grid scratch
int bilateral_build_grid(void *scratch, const float *lum, int width, int height, bil::GridDims g, float sigma_s, float sigma_r,
                         cudaStream_t s) {
  // the gather needs every pixel to feed the two cells around p / sigma_s: no saturation at the last cell (grid_dims clamps
  // the cell count to 3000 per axis), and a bounded number of pixels per cell
  // pixels at or beyond these coordinates have the clamped cell coordinate g.x - 1 (g.y - 1): they pile up in the last cell column (row)
  const int x_pile = (int)fminf((float)width, ceilf((float)(g.x - 1) * sigma_s)), y_pile = (int)fminf((float)height, ceilf((float)(g.y - 1) * sigma_s));
  const bool piled = x_pile < width || y_pile < height;
  // the gather needs a bounded number of pixels per cell: true in front of the piles
  const bool gather = sigma_s >= 1.0f && (GPX + 1) * sigma_s + 8.0f <= (float)kMaxAxisPx && x_pile >= 1 && y_pile >= 1 &&
                      (!piled || (size_t)g.z * kThreads * sizeof(float) <= 48 * 1024);  // (the pile kernels keep g.z x 256 private bins)
  if (gather) {
    static DeviceOnce attr;
    attr.run([&] {
      cudaFuncSetAttribute(grid_build_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(grid_build_kernel<16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * (((20 * GPX + 31) / 32 * 32) + 20 * BX) * 4);
      cudaFuncSetAttribute(grid_build_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * (((20 * GPX + 31) / 32 * 32) + 20 * BX) * 4);
    });
    float *blurred = static_cast<float *>(scratch) + (size_t)g.x * g.y * g.z;
    float *pile_x = nullptr, *pile_y = nullptr;
    if (piled) {  // the unused splat-grid half of the scratch holds the piles
      pile_x = static_cast<float *>(scratch), pile_y = pile_x + (size_t)g.y * g.z;
      cudaMemsetAsync(pile_x, 0, (size_t)(g.x + g.y) * g.z * sizeof(float), s);
      if (int e = check_launch("bilateral_zero_piles")) return e;
      const size_t psm = (size_t)g.z * kThreads * sizeof(float);
      if (x_pile < width) {
        pile_rows_kernel<<<height, kThreads, psm, s>>>(lum, pile_x, width, height, x_pile, g, sigma_s, sigma_r);
        if (int e = check_launch("bilateral_pile_rows")) return e;
      }
      if (y_pile < height) {
        pile_columns_kernel<<<div_up(x_pile, 32), kThreads, psm, s>>>(lum, pile_y, width, height, x_pile, y_pile, g, sigma_s, sigma_r);
        if (int e = check_launch("bilateral_pile_columns")) return e;
      }
    }
    const int gw = x_pile, gh = y_pile;  // the gather visits the pixels in front of the piles only
    if (g.z <= 16 && sigma_s == 2.0f) {
      grid_build_kernel<16, true, true><<<dim3(div_up(g.x, BX), div_up(g.y, 16)), kThreads, (size_t)g.z * (((20 * GPX + 31) / 32 * 32) + 20 * BX) * sizeof(float), s>>>(
          lum, blurred, gw, gh, g, sigma_s, sigma_r, width, pile_x, pile_y);
    } else if (g.z <= 16 && sigma_s < 3.0f) {
      grid_build_kernel<16, true><<<dim3(div_up(g.x, BX), div_up(g.y, 16)), kThreads, (size_t)g.z * (((20 * GPX + 31) / 32 * 32) + 20 * BX) * sizeof(float), s>>>(
          lum, blurred, gw, gh, g, sigma_s, sigma_r, width, pile_x, pile_y);
    } else {
      grid_build_kernel<8, false><<<dim3(div_up(g.x, BX), div_up(g.y, 8)), kThreads, (size_t)g.z * (((12 * GPX + 31) / 32 * 32) + 12 * BX) * sizeof(float), s>>>(
          lum, blurred, gw, gh, g, sigma_s, sigma_r, width, pile_x, pile_y);
    }
    return check_launch("bilateral_grid_build");
  }
  if (int e = bilateral_zero_grid(scratch, g, s)) return e;
  dim3 sgrid(div_up(width, TP), div_up(height, TP));
  splat_kernel<false><<<sgrid, kThreads, 0, s>>>(lum, static_cast<float *>(scratch), width, height, g, sigma_s, sigma_r);
  if (int e = check_launch("bilateral_splat")) return e;
  return bilateral_blur(scratch, g, s);
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
  return (2 * (size_t)g.x * g.y * g.z + (size_t)width * height) * sizeof(float);  // two grids + a luminance plane
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
  float *plane = bilateral_lum_plane(scratch, g);
  const int64_t n = (int64_t)width * height;
  luminance_plane_kernel<<<(int)((n + kThreads - 1) / kThreads < kNumSMs * 16 ? (n + kThreads - 1) / kThreads : kNumSMs * 16), kThreads, 0, s>>>(rgb, plane, n);
  if (int e = check_launch("bilateral_luminance")) return e;
  return bilateral_build_grid(scratch, plane, width, height, g, sigma_s, sigma_r, s);
}

}  // extern "C"
