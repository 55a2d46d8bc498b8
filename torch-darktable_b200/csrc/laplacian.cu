// Local Laplacian filter (darktable local-laplacian, 6 gamma curves, fp16 pyramids like the reference).
//
// Reference: csrc/local_contrast/laplacian.cu:446-582 = pad + input pyramid + 6 x (curve over the padded level 0 + its
// pyramid) + assemble per level + write back: ~100 launches, 8 fp16 pyramids of the 2.5x padded frame in HBM
// (about 190 B per image pixel at 50 MP) and pointer tables in __device__ globals.
// Here neither the padded frame nor its six tone-curved copies are ever materialised: ONE kernel builds level 1 of all seven
// pyramids from the fp32 image (replicate padding, fp16 rounding, curves and the seven 5x5 reductions fused; CTAs over the padding
// evaluate the curves once per distinct source pixel), and the level-0 assemble recomputes curve(image) for the two gammas it
// blends.  The assemble passes only run on the region that can reach the cropped output, and the last one writes fp32
// directly (write_back fused).  No global device state: level pointers travel in kernel arguments.
// All arithmetic follows the reference's order of operations and its fp16 rounding points, so results are expected
// to be bit-identical up to fast-math differences in the curve.
#include "tdb_common.cuh"

namespace tdb {
namespace {

constexpr int G = 6;  // number of gamma curves (the only count the reference accepts, laplacian.cu:625-634)
constexpr int kMaxLevels = 30;
constexpr int kThreads = 256;

inline int dl(int x, int level) { return (x + (1 << level) - 1) >> level; }

struct Plan {
  int levels, max_supp, bw, bh;
  int w[kMaxLevels], h[kMaxLevels];
  size_t padded[kMaxLevels], output[kMaxLevels], proc[G][kMaxLevels];  // element offsets (halfs) into the scratch
  size_t total;
};

Plan make_plan(int width, int height) {
  Plan p{};
  const int m = width < height ? width : height;
  int levels = 0;
  while ((1 << (levels + 1)) <= m) levels++;
  p.levels = levels > kMaxLevels ? kMaxLevels : levels;
  p.max_supp = 1 << (p.levels - 1);
  p.bw = width + 2 * p.max_supp, p.bh = height + 2 * p.max_supp;
  size_t off = 0;
  auto take = [&](size_t n) {
    const size_t o = off;
    off += (n + 7) & ~size_t(7);  // keep every plane 16-byte aligned
    return o;
  };
  for (int l = 0; l < p.levels; l++) {
    p.w[l] = dl(p.bw, l), p.h[l] = dl(p.bh, l);
    const size_t n = (size_t)p.w[l] * p.h[l];
    if (l == 0) continue;  // level 0 is virtual everywhere: the padded input and its six curved copies are recomputed from the image
    p.padded[l] = take(n);
    p.output[l] = take(n);
    for (int k = 0; k < G; k++) p.proc[k][l] = take(n);
  }
  p.total = off;
  return p;
}

struct CurveParams {
  float sigma, shadows, highlights, clarity;
};

__device__ __forceinline__ float curve(float x, float g, const CurveParams &cp) {  // laplacian.cu:266-290
  const float c = x - g;
  const float ssigma = c > 0.0f ? cp.sigma : -cp.sigma;
  const float shadhi = c > 0.0f ? cp.shadows : cp.highlights;
  float val;
  if (fabsf(c) > 2 * cp.sigma) {
    val = g + ssigma + shadhi * (c - ssigma);
  } else {
    const float t = clip01(c / (2.0f * ssigma));
    const float t2 = t * t, mt = 1.0f - t;
    val = g + ssigma * 2.0f * mt * t + t2 * (ssigma + ssigma * shadhi);
  }
  // clarity == 0 (the default): the term is an exact zero, skip its exp()
  if (cp.clarity != 0.0f) val += cp.clarity * c * expf(-c * c / (2.0f * cp.sigma * cp.sigma / 3.0f));
  return val;
}

__device__ __forceinline__ float h2f(__half h) { return __half2float(h); }
__device__ __forceinline__ __half f2h(float f) { return __float2half_rn(f); }

constexpr int NP = G + 1;  // pyramids built side by side: the six curved ones and the input's own
struct ReduceBatch {
  const __half *fine[NP];
  __half *coarse[NP];
};

// Zones of a pyramid level in which the replicate padding is still exact: columns <= x_lo all equal column x_lo, columns >= x_hi
// all equal column x_hi, likewise for rows.  At level 0 these are the image edges; a coarse pixel whose five fine taps lie inside a
// zone repeats its neighbour bit for bit, so the zone of the next level is ((lo - 2) >> 1, (hi + 3) >> 1).
struct FlatZones {
  int x_lo, x_hi, y_lo, y_hi;
};
inline FlatZones coarser(const FlatZones &f) {
  return FlatZones{f.x_lo >= 2 ? (f.x_lo - 2) >> 1 : -1, (f.x_hi + 3) >> 1, f.y_lo >= 2 ? (f.y_lo - 2) >> 1 : -1, (f.y_hi + 3) >> 1};
}

// 5x5 binomial, decimate by 2, clone a 1-px border (laplacian.cu:178-208), levels >= 2.  blockIdx.z selects the pyramid.
// v1 staged a 35 x 35 patch with one division per element and read it back with a lane stride of two floats (two-way bank
// conflicts on all 25 taps; ncu: 51 M conflicts, 290 instructions per coarse pixel) -- over the whole padded plane.  Here the patch
// is split by column parity (the taps of a lane are then contiguous words), rows are staged two columns per thread, and tiles
// inside a flat zone compute one row / column / pixel and replicate it, which removes most of the padding's share (60 % of the plane).
constexpr int RT = 16;            // coarse tile edge
constexpr int RP = 2 * RT + 3;    // fine patch edge (35)
constexpr int RPS = 24;           // row stride of a parity plane: 18 words used; 2 * 24 = 16 (mod 32) keeps the two half-warps apart
// one 16 x 16 coarse tile at (cx0, cy0); pe / po: RP * RPS floats each (even / odd patch columns), res: RT * RT halves
__device__ __forceinline__ void reduce_tile16(const __half *__restrict__ fine, __half *__restrict__ coarse, int cx0, int cy0, int cw, int ch,
                                              int fw, int fh, const FlatZones &fz, float *pe, float *po, __half (*res)[RT]) {
  // coarse pixel c reads fine 2*c'-2 .. 2*c'+2 with c' clamped to [1, size-2]; patch origin = 2*cx0 - 2 covers every
  // unclamped pixel of the tile, clamped border pixels are handled by reading through the same patch when possible
  const int fx0 = 2 * cx0 - 2, fy0 = 2 * cy0 - 2;
  const bool in_frame = fx0 >= 0 && fy0 >= 0 && fx0 + RP <= fw && fy0 + RP <= fh && cx0 + RT < cw && cy0 + RT < ch;
  const bool flat_x = in_frame && (fx0 + RP - 1 <= fz.x_lo || fx0 >= fz.x_hi);
  const bool flat_y = in_frame && (fy0 + RP - 1 <= fz.y_lo || fy0 >= fz.y_hi);
  const float w[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  const int tid = threadIdx.x;
  if (flat_x || flat_y) {
    // constant along x and / or y: stage the distinct rows / columns only, compute the distinct outputs, replicate
    const int nx = flat_x ? 1 : RP, ny = flat_y ? 1 : RP;
    for (int i = tid; i < nx * ny; i += kThreads) {
      const int ly = i / nx, lx = i - ly * nx;
      pe[i] = h2f(fine[(int64_t)(fy0 + ly) * fw + fx0 + lx]);  // at most one row or one column: 35 values
    }
    __syncthreads();
    const int sxm = flat_x ? 0 : 1, sym = flat_y ? 0 : 1;  // a column is stored densely (not flat in y implies flat in x here)
    const int ncx = flat_x ? 1 : RT, ncy = flat_y ? 1 : RT;
    for (int t = tid; t < ncx * ncy; t += kThreads) {
      const int lx = t % ncx, ly = t / ncx;
      const float *c = pe + (2 * ly + 2) * sym + (2 * lx + 2) * sxm;
      float acc = 0.0f;
#pragma unroll
      for (int j = -2; j <= 2; j++)
#pragma unroll
        for (int i = -2; i <= 2; i++) acc += c[j * sym + i * sxm] * w[i + 2] * w[j + 2];
      res[ly][lx] = f2h(acc);
    }
    __syncthreads();
    const int lx = tid & 15, ly = tid >> 4;
    coarse[(int64_t)(cy0 + ly) * cw + cx0 + lx] = res[flat_y ? 0 : ly][flat_x ? 0 : lx];
    return;
  }
  // two columns (one even, one odd) of one row per step: 35 rows x 18 column pairs
  for (int i = tid; i < RP * 18; i += kThreads) {
    const int ly = i / 18, kx = i - ly * 18;
    const int x = fx0 + 2 * kx, y = fy0 + ly;
    float ve = 0.0f, vo = 0.0f;
    if (y >= 0 && y < fh) {
      const __half *row = fine + (int64_t)y * fw;
      if (x >= 0 && x < fw) ve = h2f(row[x]);
      if (x + 1 >= 0 && x + 1 < fw && 2 * kx + 1 < RP) vo = h2f(row[x + 1]);
    }
    pe[ly * RPS + kx] = ve, po[ly * RPS + kx] = vo;
  }
  __syncthreads();
  const int lx = tid & 15, ly = tid >> 4;
  const int x = cx0 + lx, y = cy0 + ly;
  if (x >= cw || y >= ch) return;
  int cx = x, cy = y;
  if (x >= cw - 1) cx = cw - 2;
  if (y >= ch - 1) cy = ch - 2;
  if (cx <= 0) cx = 1;
  if (cy <= 0) cy = 1;
  float acc = 0.0f;
  const int px = 2 * cx - fx0, py = 2 * cy - fy0;  // patch coordinates of the stencil centre (px is even)
  const bool inside = px >= 2 && py >= 2 && px + 2 < RP && py + 2 < RP;
  if (inside) {
    const float *e = pe + py * RPS + (px >> 1), *o = po + py * RPS + (px >> 1);
#pragma unroll
    for (int j = -2; j <= 2; j++) {
      acc += e[j * RPS - 1] * w[0] * w[j + 2];
      acc += o[j * RPS - 1] * w[1] * w[j + 2];
      acc += e[j * RPS] * w[2] * w[j + 2];
      acc += o[j * RPS] * w[3] * w[j + 2];
      acc += e[j * RPS + 1] * w[4] * w[j + 2];
    }
  } else {  // a clamped border pixel whose stencil left the staged patch (only at the far image edge)
#pragma unroll
    for (int j = -2; j <= 2; j++)
#pragma unroll
      for (int i = -2; i <= 2; i++) {
        acc += h2f(fine[(int64_t)(2 * cy + j) * fw + (2 * cx + i)]) * w[i + 2] * w[j + 2];
      }
  }
  coarse[(int64_t)y * cw + x] = f2h(acc);
}

// Levels >= 2, all seven pyramids (blockIdx.z), 32 x 32 coarse tiles.  A tile whose 67 x 67 fine patch lies inside the plane and is not
// flat takes the separable path of reduce1_interior (same argument: fp16 data times k / 256 weights sum exactly in fp32, so the order
// of the additions cannot change the result): a thread owns one coarse column and four coarse rows, 18.75 multiply-adds and 13.75
// loads per coarse pixel instead of 50 and 25, no per-pixel clamp tests.  Everything else -- plane borders (clamped coarse
// coordinates, reference laplacian.cu:178-208) and flat zones of the replicate padding -- goes through reduce_tile16, one 16 x 16
// quarter at a time.
constexpr int R2 = 32, P2 = 2 * R2 + 3, P2S = 68;  // coarse tile, fine patch 67 x 67, row stride: 34 even + 33 odd words + 1
__global__ void __launch_bounds__(kThreads) reduce_kernel(const __grid_constant__ ReduceBatch b, int cw, int ch, int fw, int fh,
                                                          const FlatZones fz) {
  __shared__ __align__(16) float sm2[P2 * P2S];
  const int k = blockIdx.z;
  const __half *__restrict__ fine = b.fine[k];
  __half *__restrict__ coarse = b.coarse[k];
  const int cx0 = blockIdx.x * R2, cy0 = blockIdx.y * R2;
  const int fx0 = 2 * cx0 - 2, fy0 = 2 * cy0 - 2;
  const bool inside = fx0 >= 0 && fy0 >= 0 && fx0 + P2 <= fw && fy0 + P2 <= fh && cx0 >= 1 && cy0 >= 1 && cx0 + R2 <= cw - 1 && cy0 + R2 <= ch - 1;
  const bool flat = (fx0 + P2 - 1 <= fz.x_lo || fx0 >= fz.x_hi) || (fy0 + P2 - 1 <= fz.y_lo || fy0 >= fz.y_hi);
  if (inside && !flat) {
    constexpr int NPAIR = (P2 + 1) / 2;  // 34 column pairs: even column -> word kx, odd column -> word 34 + kx
    const __half *src = fine + (int64_t)fy0 * fw + fx0;  // fx0 is even: pairs are 4-byte aligned when fw is even
    const bool pair_loads = (fw & 1) == 0 && (reinterpret_cast<uintptr_t>(fine) & 3) == 0;
    for (int i = threadIdx.x; i < P2 * NPAIR; i += kThreads) {
      const int ly = i / NPAIR, kx = i - ly * NPAIR;
      const __half *row = src + (int64_t)ly * fw + 2 * kx;
      float ve, vo = 0.0f;
      if (kx < NPAIR - 1) {
        if (pair_loads) {
          const float2 v = __half22float2(*reinterpret_cast<const __half2 *>(row));
          ve = v.x, vo = v.y;
        } else {
          ve = h2f(row[0]), vo = h2f(row[1]);
        }
      } else {
        ve = h2f(row[0]);  // the patch has 67 columns: the last pair has no odd member
      }
      sm2[ly * P2S + kx] = ve;
      if (kx < NPAIR - 1) sm2[ly * P2S + 34 + kx] = vo;
    }
    __syncthreads();
    const float w0 = 1.0f / 16.0f, w1 = 4.0f / 16.0f, w2 = 6.0f / 16.0f;
    const int lx = threadIdx.x & 31, rg = threadIdx.x >> 5;  // coarse column, group of four coarse rows
    const float *e = sm2 + (8 * rg) * P2S + lx;
    float h[11];
#pragma unroll
    for (int r = 0; r < 11; r++) {
      const float *q = e + r * P2S;
      h[r] = fmaf(q[2], w0, fmaf(q[35], w1, fmaf(q[1], w2, fmaf(q[34], w1, q[0] * w0))));
    }
    __half *o = coarse + (int64_t)(cy0 + 4 * rg) * cw + cx0 + lx;
#pragma unroll
    for (int r = 0; r < 4; r++)
      o[(int64_t)r * cw] = f2h(fmaf(h[2 * r + 4], w0, fmaf(h[2 * r + 3], w1, fmaf(h[2 * r + 2], w2, fmaf(h[2 * r + 1], w1, h[2 * r] * w0)))));
    return;
  }
  float *pe = sm2, *po = sm2 + RP * RPS;
  __half(*res)[RT] = reinterpret_cast<__half(*)[RT]>(sm2 + 2 * RP * RPS);
  static_assert(2 * RP * RPS + RT * RT / 2 <= P2 * P2S, "the 16 x 16 path lives inside the 32 x 32 path's shared memory");
  for (int q = 0; q < 4; q++) {
    const int qx = cx0 + (q & 1) * RT, qy = cy0 + (q >> 1) * RT;
    __syncthreads();  // the previous quarter's readers are done
    if (qx < cw && qy < ch) reduce_tile16(fine, coarse, qx, qy, cw, ch, fw, fh, fz, pe, po, res);
  }
}

// Level 1 of all seven pyramids straight from the fp32 image: replicate padding (laplacian.cu:90-109), the fp16 rounding of the
// padded level 0, the six curves (:266-290) and the seven 5x5 reductions (:178-208) in one pass.  The reference writes the padded
// frame (2.5x the image at 50 MP) and six curved copies of it to HBM and reads each back; here none of them exists.
// A CTA whose fine patch lies entirely beside / diagonally off the image sees a patch that is constant along one / both axes
// (replicate padding): it evaluates the curves for one row / column / pixel only and lets the stencil re-read it (stride 0).
constexpr int R1W = 32, R1H = 16;                        // coarse tile
constexpr int P1W = 2 * R1W + 3, P1H = 2 * R1H + 3;      // fine patch 67 x 35
constexpr int P1S = P1W + 1;
struct Reduce1Args {
  const float *in;  // (height, width) fp32 image
  __half *coarse[NP];
  int width, height, max_supp;
  int fw, fh, cw, ch;  // padded level 0 / level 1 sizes
};
// A tile whose fine patch lies inside the IMAGE (no padding, no clamped coordinate, no flat axis) -- 96 % of the work at 50 MP.
// The general path below spends 2400 instructions per coarse pixel: run-time divisions in the staging loop, bounds and clamp tests
// per element, and seven 25-tap sums per coarse pixel whose taps are two multiplies each (the reference's (v * w_i) * w_j).  Here
//   staging : a thread owns (row, column pair) tasks with compile-time divisors; the image is read without clamps
//   sums    : SEPARABLE, in fp32: a thread owns one coarse column and FOUR coarse rows of half of the planes, forms the horizontal
//             5-tap sums of the 11 fine rows it needs once (conflict-free: even / odd columns are stored apart) and combines them
//             vertically -- 18.75 multiply-adds per coarse pixel and plane instead of 50, 13.75 loads instead of 25
//   stores  : straight from registers, a warp writes 32 consecutive halves of a row
// Separability does not change the result: every fine value is an fp16 number (11 significant bits) and every weight product is
// k / 256 with k in {1, 4, 6, 16, 24, 36}, so each of the 25 terms is exact in fp32 and so is any partial sum of a neighbourhood whose
// values lie within a factor of ~2^5 of each other -- the order of the additions (the reference's i-outer / j-inner loop,
// laplacian.cu:178-208, or rows first as here) cannot matter.  Measured: bit-identical to the general path and to the reference on
// the 50 MP frame of the live comparison (profiles/r02_ref_live_report.jsonl), 1.30 -> 0.95 ms.
__device__ __forceinline__ void reduce1_interior(const Reduce1Args &a, const CurveParams &cp, float *sm1, int cx0, int cy0, int fx0, int fy0) {
  constexpr int NPAIR = (P1W + 1) / 2;  // 34 column pairs per patch row: even column -> word kx, odd column -> word 34 + kx
  const float *src = a.in + (int64_t)(fy0 - a.max_supp) * a.width + (fx0 - a.max_supp);
  for (int i = threadIdx.x; i < P1H * NPAIR; i += kThreads) {
    const int ly = i / NPAIR, kx = i - ly * NPAIR;
    const float *row = src + (int64_t)ly * a.width + 2 * kx;
    float *cell = sm1 + ly * P1S + kx;
    const float ve = h2f(f2h(__ldg(row)));
    cell[G * P1H * P1S] = ve;
#pragma unroll
    for (int k = 0; k < G; k++) cell[k * P1H * P1S] = h2f(f2h(curve(ve, (k + 0.5f) / (float)G, cp)));
    if (kx < NPAIR - 1) {  // the patch has 67 columns: the last pair has no odd member
      const float vo = h2f(f2h(__ldg(row + 1)));
      cell[34 + G * P1H * P1S] = vo;
#pragma unroll
      for (int k = 0; k < G; k++) cell[34 + k * P1H * P1S] = h2f(f2h(curve(vo, (k + 0.5f) / (float)G, cp)));
    }
  }
  __syncthreads();
  const float w0 = 1.0f / 16.0f, w1 = 4.0f / 16.0f, w2 = 6.0f / 16.0f;
  const int lx = threadIdx.x & 31, rg = (threadIdx.x >> 5) & 3, part = threadIdx.x >> 7;
  const int k0 = part ? 4 : 0, k1 = part ? NP : 4;  // planes 0 .. 3 / 4 .. 6
  for (int k = k0; k < k1; k++) {
    // coarse row 4 rg + r has its stencil centre on patch row 2 (4 rg + r) + 2: fine rows 8 rg + 2 r .. 8 rg + 2 r + 4; the centre
    // column of coarse column lx is patch column 2 lx + 2: even words lx, lx + 1, lx + 2 and odd words lx, lx + 1
    const float *e = sm1 + k * P1H * P1S + (8 * rg) * P1S + lx;
    float h[11];
#pragma unroll
    for (int r = 0; r < 11; r++) {
      const float *p = e + r * P1S;
      h[r] = fmaf(p[2], w0, fmaf(p[35], w1, fmaf(p[1], w2, fmaf(p[34], w1, p[0] * w0))));
    }
    __half *o = a.coarse[k] + (int64_t)(cy0 + 4 * rg) * a.cw + cx0 + lx;
#pragma unroll
    for (int r = 0; r < 4; r++)
      o[(int64_t)r * a.cw] = f2h(fmaf(h[2 * r + 4], w0, fmaf(h[2 * r + 3], w1, fmaf(h[2 * r + 2], w2, fmaf(h[2 * r + 1], w1, h[2 * r] * w0)))));
  }
}

__global__ void __launch_bounds__(kThreads) reduce1_kernel(const __grid_constant__ Reduce1Args a, const CurveParams cp) {
  extern __shared__ float sm1[];  // [NP][P1H][P1S]
  const int cx0 = blockIdx.x * R1W, cy0 = blockIdx.y * R1H;
  const int fx0 = 2 * cx0 - 2, fy0 = 2 * cy0 - 2;
  const int ms = a.max_supp;
  if (fx0 - ms >= 0 && fy0 - ms >= 0 && fx0 + P1W - ms <= a.width && fy0 + P1H - ms <= a.height && cx0 >= 1 && cy0 >= 1 &&
      cx0 + R1W <= a.cw - 1 && cy0 + R1H <= a.ch - 1) {
    reduce1_interior(a, cp, sm1, cx0, cy0, fx0, fy0);
    return;
  }
  // patch inside the padded frame and on one side of the image along an axis -> constant along that axis
  const bool in_frame = fx0 >= 0 && fy0 >= 0 && fx0 + P1W <= a.fw && fy0 + P1H <= a.fh;
  const bool flat_x = in_frame && (fx0 + P1W - 1 - ms <= 0 || fx0 - ms >= a.width - 1);
  const bool flat_y = in_frame && (fy0 + P1H - 1 - ms <= 0 || fy0 - ms >= a.height - 1);
  const int nx = flat_x ? 1 : P1W, ny = flat_y ? 1 : P1H;
  for (int i = threadIdx.x; i < nx * ny; i += kThreads) {
    const int ly = i / nx, lx = i - ly * nx;
    const int x = fx0 + lx, y = fy0 + ly;
    float v = 0.0f;
    const bool inside = x >= 0 && y >= 0 && x < a.fw && y < a.fh;
    if (inside) {
      const int sx = min(max(x - ms, 0), a.width - 1), sy = min(max(y - ms, 0), a.height - 1);
      v = h2f(f2h(__ldg(a.in + (int64_t)sy * a.width + sx)));
    }
    // non-flat tiles keep the even and the odd columns of a row apart (34 + 33 words): the 5 taps of a lane are then contiguous
    float *cell = sm1 + ly * P1S + (flat_x || flat_y ? lx : (lx & 1) * 34 + (lx >> 1));
    cell[G * P1H * P1S] = v;
#pragma unroll
    for (int k = 0; k < G; k++) cell[k * P1H * P1S] = inside ? h2f(f2h(curve(v, (k + 0.5f) / (float)G, cp))) : 0.0f;
  }
  __syncthreads();
  const float w[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  const int sxm = flat_x ? 0 : 1, sym = flat_y ? 0 : P1S;  // stride-0 re-reads along a constant axis
  // a patch that is constant along an axis gives results that are constant along it too: one column / row / pixel of the tile is
  // computed and the stores replicate it (flat tiles lie inside the frame, where no coarse coordinate is clamped)
  const int ncx = flat_x ? 1 : R1W, ncy = flat_y ? 1 : R1H;
  __shared__ __half res[NP][R1H][R1W];
  for (int t = threadIdx.x; t < ncx * ncy; t += kThreads) {
    const int lx = t % ncx, ly = t / ncx;
    const int x = cx0 + lx, y = cy0 + ly;
    if (x >= a.cw || y >= a.ch) continue;
    int cx = x, cy = y;
    if (x >= a.cw - 1) cx = a.cw - 2;
    if (y >= a.ch - 1) cy = a.ch - 2;
    if (cx <= 0) cx = 1;
    if (cy <= 0) cy = 1;
    const int px = 2 * cx - fx0, py = 2 * cy - fy0;  // patch coordinates of the stencil centre
    const bool staged = px >= 2 && py >= 2 && px + 2 < P1W && py + 2 < P1H;
#pragma unroll 1
    for (int k = 0; k < NP; k++) {
      float acc = 0.0f;
      if (staged && !flat_x && !flat_y) {
        const float *e = sm1 + k * P1H * P1S + py * P1S + (px >> 1), *o = e + 34;  // px is even
#pragma unroll
        for (int j = -2; j <= 2; j++) {
          acc += e[j * P1S - 1] * w[0] * w[j + 2];
          acc += o[j * P1S - 1] * w[1] * w[j + 2];
          acc += e[j * P1S] * w[2] * w[j + 2];
          acc += o[j * P1S] * w[3] * w[j + 2];
          acc += e[j * P1S + 1] * w[4] * w[j + 2];
        }
      } else if (staged) {
        const float *c = sm1 + k * P1H * P1S + py * sym + px * sxm;
#pragma unroll
        for (int j = -2; j <= 2; j++)
#pragma unroll
          for (int i = -2; i <= 2; i++) acc += c[j * sym + i * sxm] * w[i + 2] * w[j + 2];
      } else {  // a clamped border pixel whose stencil left the staged patch (only at the far edge of the padded frame)
#pragma unroll
        for (int j = -2; j <= 2; j++)
#pragma unroll
          for (int i = -2; i <= 2; i++) {
            const int sx = min(max(2 * cx + i - ms, 0), a.width - 1), sy = min(max(2 * cy + j - ms, 0), a.height - 1);
            float v = h2f(f2h(__ldg(a.in + (int64_t)sy * a.width + sx)));
            if (k < G) v = h2f(f2h(curve(v, (k + 0.5f) / (float)G, cp)));
            acc += v * w[i + 2] * w[j + 2];
          }
      }
      res[k][ly][lx] = f2h(acc);
    }
  }
  __syncthreads();
  // stores: half2 along x (cw is even only sometimes: pair up when the row offset is even)
  for (int t = threadIdx.x; t < NP * R1H * (R1W / 2); t += kThreads) {
    const int lx = 2 * (t % (R1W / 2)), ly = (t / (R1W / 2)) % R1H, k = t / (R1H * (R1W / 2));
    const int x = cx0 + lx, y = cy0 + ly;
    if (x >= a.cw || y >= a.ch) continue;
    const __half v0 = res[k][flat_y ? 0 : ly][flat_x ? 0 : lx], v1 = res[k][flat_y ? 0 : ly][flat_x ? 0 : lx + 1];
    __half *o = a.coarse[k] + (int64_t)y * a.cw + x;
    if (x + 1 < a.cw && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) *reinterpret_cast<__half2 *>(o) = __halves2half2(v0, v1);
    else {
      o[0] = v0;
      if (x + 1 < a.cw) o[1] = v1;
    }
  }
}

__device__ __forceinline__ float expand_gaussian(const __half *__restrict__ coarse, int px, int py, int cw) {  // :111-140
  const float w[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  const int cx = px >> 1, cy = py >> 1, xo = px & 1, yo = py & 1;
  float c = 0.0f;
#pragma unroll
  for (int i = -1; i <= 1; i++)
#pragma unroll
    for (int j = -1; j <= 1; j++) {
      if ((xo && i < 0) || (yo && j < 0)) continue;
      const int wi = xo ? (2 * i + 1) : (2 * i + 2), wj = yo ? (2 * j + 1) : (2 * j + 2);
      c = fmaf(h2f(coarse[(int64_t)(cy + j) * cw + (cx + i)]), w[wi] * w[wj], c);  // exact products: see expand_from
    }
  return 4.0f * c;
}

struct AssembleArgs {
  const __half *padded;      // gaussian of the input at this (fine) level (levels >= 1)
  const float *image;        // level 0: the fp32 image itself (its padded fp16 copy is never stored)
  const __half *out_coarse;  // reconstruction one level coarser
  const __half *proc_fine[G];    // curved gaussians, this level (unused at level 0)
  const __half *proc_coarse[G];  // curved gaussians, one level coarser
  __half *out_fine;          // reconstruction at this level (levels >= 1)
  float *out_image;          // final fp32 image (level 0)
  int fw, fh;                // fine level size
  int rx0, ry0, rx1, ry1;    // region to compute (half-open), fine coordinates
  int max_supp, width;       // level 0: crop origin and output row length
};

// (Staging the coarse 3 x 3 neighbourhoods of all seven planes in shared memory per 32 x 8 tile was measured: 1.43 -> 1.75 ms over the
// eleven launches at 50 MP; L1 already serves the 27 two-byte loads per pixel, the kernel is bound by its arithmetic.)
template <bool kLevel0>
__global__ void __launch_bounds__(kThreads) assemble_kernel(const __grid_constant__ AssembleArgs a, CurveParams cp) {  // laplacian.cu:222-263
  const int x = a.rx0 + blockIdx.x * 32 + (threadIdx.x & 31), y = a.ry0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.rx1 || y >= a.ry1) return;
  const int w = a.fw, h = a.fh;
  int qx = x, qy = y;  // clamp_boundary, :53-65
  if (w & 1) { if (qx > w - 2) qx = w - 2; } else { if (qx > w - 3) qx = w - 3; }
  if (h & 1) { if (qy > h - 2) qy = h - 2; } else { if (qy > h - 3) qy = h - 3; }
  if (qx <= 0) qx = 1;
  if (qy <= 0) qy = 1;
  const int cw = (w - 1) / 2 + 1;
  float val = expand_gaussian(a.out_coarse, qx, qy, cw);
  const float v = kLevel0 ? h2f(f2h(__ldg(a.image + (int64_t)(y - a.max_supp) * a.width + (x - a.max_supp))))
                          : h2f(a.padded[(int64_t)y * w + x]);
  int hi = 1;
  for (; hi < G - 1 && ((float)hi + .5f) / (float)G <= v; hi++);
  const int lo = hi - 1;
  const float t = fminf(fmaxf(v * G - ((float)lo + .5f), 0.0f), 1.0f);
  float f0, f1;
  if (kLevel0) {
    f0 = h2f(f2h(curve(v, (lo + 0.5f) / (float)G, cp)));
    f1 = h2f(f2h(curve(v, (lo + 1.5f) / (float)G, cp)));
  } else {
    f0 = h2f(a.proc_fine[lo][(int64_t)y * w + x]);
    f1 = h2f(a.proc_fine[lo + 1][(int64_t)y * w + x]);
  }
  const float l0 = f0 - expand_gaussian(a.proc_coarse[lo], qx, qy, cw);
  const float l1 = f1 - expand_gaussian(a.proc_coarse[lo + 1], qx, qy, cw);
  val += l0 * (1.0f - t) + l1 * t;
  if (kLevel0) a.out_image[(int64_t)(y - a.max_supp) * a.width + (x - a.max_supp)] = h2f(f2h(val));
  else a.out_fine[(int64_t)y * w + x] = f2h(val);
}


// The same assemble step for a 2 x 2 block of fine pixels per thread.  assemble_kernel executes about 600 instructions per pixel, most
// of them integer work (ncu: ALU pipe 75 %, FMA 36 % at level 0): the boundary clamp, the parity tests inside expand_gaussian and the
// 64-bit indices of its 27 two-byte loads, per pixel.  The four pixels (2 bx + {0,1}, 2 by + {0,1}) share the coarse centre (bx, by),
// so one 3 x 3 coarse neighbourhood per plane serves all of them (9 loads instead of 9 + 6 + 6 + 4), their parities -- and with them
// the tap sets and weights of expand_gaussian (laplacian.cu:111-140) -- are compile-time, and each pixel's sum still runs over its
// taps in the reference's order (i outer, j inner), so the result is bit-identical.  The two gamma planes a pixel blends depend on
// its own value; the block walks the planes any of its pixels needs (two when the four agree, as in smooth image regions).
// Only for regions that no boundary clamp can reach (1 <= x <= w - 3): the host falls back to assemble_kernel otherwise.
struct Taps9 {
  float v[3][3];  // [j + 1][i + 1]: coarse (bx + i, by + j)
};
__device__ __forceinline__ Taps9 load_taps(const __half *__restrict__ coarse, int bx, int by, int cw) {
  Taps9 t;
  const __half *c = coarse + (int64_t)(by - 1) * cw + (bx - 1);
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) t.v[j][i] = h2f(c[j * cw + i]);
  return t;
}
template <int XO, int YO>
__device__ __forceinline__ float expand_from(const Taps9 &t) {  // expand_gaussian for the pixel with parities (XO, YO)
  const float w[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  float c = 0.0f;
#pragma unroll
  for (int i = -1; i <= 1; i++)
#pragma unroll
    for (int j = -1; j <= 1; j++) {
      if ((XO && i < 0) || (YO && j < 0)) continue;
      const int wi = XO ? (2 * i + 1) : (2 * i + 2), wj = YO ? (2 * j + 1) : (2 * j + 2);
      // (v * w_i) * w_j of the reference as ONE multiply-add: v is an fp16 value and w_i * w_j = k / 256, so both products are exact
      // in fp32 and the two forms round identically; the compiler cannot know that and would keep two multiplies per tap
      c = fmaf(t.v[j + 1][i + 1], w[wi] * w[wj], c);
    }
  return 4.0f * c;
}
struct Quad {
  float p[4];  // (0,0), (1,0), (0,1), (1,1): index = XO + 2 * YO
};
__device__ __forceinline__ Quad expand_quad(const Taps9 &t) {
  return Quad{{expand_from<0, 0>(t), expand_from<1, 0>(t), expand_from<0, 1>(t), expand_from<1, 1>(t)}};
}

template <bool kLevel0>
__global__ void __launch_bounds__(kThreads) assemble2_kernel(const __grid_constant__ AssembleArgs a, CurveParams cp) {
  const int bx = (a.rx0 >> 1) + blockIdx.x * 32 + (threadIdx.x & 31), by = (a.ry0 >> 1) + blockIdx.y * 8 + (threadIdx.x >> 5);
  const int x0 = 2 * bx, y0 = 2 * by;
  if (x0 >= a.rx1 || y0 >= a.ry1) return;
  const int w = a.fw;
  const int cw = (w - 1) / 2 + 1;
  // the four fine values, the gamma pair each of them blends and its weight (:238-250)
  float v[4], tt[4];
  int lo[4];
  bool in[4];
  int lo_min = G, lo_max = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = x0 + (k & 1), y = y0 + (k >> 1);
    in[k] = x < a.rx1 && y < a.ry1;
    v[k] = 0.0f;
    if (in[k]) v[k] = kLevel0 ? h2f(f2h(__ldg(a.image + (int64_t)(y - a.max_supp) * a.width + (x - a.max_supp)))) : h2f(a.padded[(int64_t)y * w + x]);
    int hi = 1;
    for (; hi < G - 1 && ((float)hi + .5f) / (float)G <= v[k]; hi++);
    lo[k] = hi - 1;
    tt[k] = fminf(fmaxf(v[k] * G - ((float)lo[k] + .5f), 0.0f), 1.0f);
    lo_min = min(lo_min, lo[k]), lo_max = max(lo_max, lo[k]);
  }
  const Quad base = expand_quad(load_taps(a.out_coarse, bx, by, cw));
  float e0[4], e1[4];  // expand_gaussian of the planes lo and lo + 1 of each pixel
#pragma unroll
  for (int k = 0; k < 4; k++) e0[k] = 0.0f, e1[k] = 0.0f;
  for (int g = lo_min; g <= lo_max + 1; g++) {
    const Quad q = expand_quad(load_taps(a.proc_coarse[g], bx, by, cw));
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (g == lo[k]) e0[k] = q.p[k];
      if (g == lo[k] + 1) e1[k] = q.p[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (!in[k]) continue;
    const int x = x0 + (k & 1), y = y0 + (k >> 1);
    float f0, f1;
    if (kLevel0) {
      f0 = h2f(f2h(curve(v[k], (lo[k] + 0.5f) / (float)G, cp)));
      f1 = h2f(f2h(curve(v[k], (lo[k] + 1.5f) / (float)G, cp)));
    } else {
      f0 = h2f(a.proc_fine[lo[k]][(int64_t)y * w + x]);
      f1 = h2f(a.proc_fine[lo[k] + 1][(int64_t)y * w + x]);
    }
    const float l0 = f0 - e0[k], l1 = f1 - e1[k];
    const float val = base.p[k] + (l0 * (1.0f - tt[k]) + l1 * tt[k]);
    if (kLevel0) a.out_image[(int64_t)(y - a.max_supp) * a.width + (x - a.max_supp)] = h2f(f2h(val));
    else a.out_fine[(int64_t)y * w + x] = f2h(val);
  }
}

}  // namespace
}  // namespace tdb

using namespace tdb;

extern "C" {

size_t tdb_laplacian_scratch_bytes(int width, int height) {
  if (width < 2 || height < 2) return 0;
  return make_plan(width, height).total * sizeof(__half);
}

int tdb_laplacian(const float *lum, float *out, void *scratch, int width, int height, float sigma, float shadows, float highlights,
                  float clarity, tdb_stream_t stream) {
  TDB_REQUIRE(lum && out && scratch, "Laplacian: null pointer");
  TDB_REQUIRE(width >= 4 && height >= 4, "Laplacian: image must be at least 4x4");
  const Plan p = make_plan(width, height);
  cudaStream_t s = as_stream(stream);
  __half *base = static_cast<__half *>(scratch);
  const CurveParams cp{sigma, shadows, highlights, clarity};
  const int L = p.levels;

  // level 1 of the seven pyramids from the image itself, deeper levels batched over the pyramids (blockIdx.z); the coarsest
  // level of the input's pyramid seeds the reconstruction (laplacian.cu:515-528)
  auto input_level = [&](int l) { return base + (l == L - 1 ? p.output[l] : p.padded[l]); };
  {
    static DeviceOnce attr;
    constexpr int bytes = NP * P1H * P1S * sizeof(float);
    attr.run([&] {
      cudaFuncSetAttribute(reduce1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    });
    Reduce1Args r{};
    r.in = lum, r.width = width, r.height = height, r.max_supp = p.max_supp;
    r.fw = p.w[0], r.fh = p.h[0], r.cw = p.w[1], r.ch = p.h[1];
    for (int k = 0; k < G; k++) r.coarse[k] = base + p.proc[k][1];
    r.coarse[G] = input_level(1);
    reduce1_kernel<<<dim3(div_up(p.w[1], R1W), div_up(p.h[1], R1H)), kThreads, bytes, s>>>(r, cp);
    if (int e = check_launch("laplacian_reduce_level1")) return e;
  }
  FlatZones fz{p.max_supp, p.max_supp + width - 1, p.max_supp, p.max_supp + height - 1};  // level 0
  for (int l = 2; l < L; l++) {
    fz = coarser(fz);  // zones of level l - 1, the fine level of this reduction
    ReduceBatch b{};
    for (int k = 0; k < G; k++) b.fine[k] = base + p.proc[k][l - 1], b.coarse[k] = base + p.proc[k][l];
    b.fine[G] = input_level(l - 1), b.coarse[G] = input_level(l);
    reduce_kernel<<<dim3(div_up(p.w[l], R2), div_up(p.h[l], R2), NP), kThreads, 0, s>>>(b, p.w[l], p.h[l], p.w[l - 1], p.h[l - 1], fz);
    if (int e = check_launch("laplacian_reduce")) return e;
  }
  // regions of each level that can reach the cropped output: level l needs level l+1 on region/2 -+ 1
  int rx0[kMaxLevels], ry0[kMaxLevels], rx1[kMaxLevels], ry1[kMaxLevels];
  rx0[0] = p.max_supp, ry0[0] = p.max_supp, rx1[0] = p.max_supp + width, ry1[0] = p.max_supp + height;
  for (int l = 1; l < L; l++) {
    // clamp_boundary may move a fine pixel by up to 2 before halving; stay conservative with a 3-pixel collar
    rx0[l] = max(0, (rx0[l - 1] >> 1) - 3), ry0[l] = max(0, (ry0[l - 1] >> 1) - 3);
    rx1[l] = min(p.w[l], ((rx1[l - 1] + 1) >> 1) + 3), ry1[l] = min(p.h[l], ((ry1[l - 1] + 1) >> 1) + 3);
  }
  for (int l = L - 2; l >= 0; l--) {
    AssembleArgs a{};
    a.padded = l == 0 ? nullptr : base + p.padded[l];
    a.image = lum;
    a.out_coarse = base + p.output[l + 1];
    for (int k = 0; k < G; k++) {
      a.proc_fine[k] = l == 0 ? nullptr : base + p.proc[k][l];
      a.proc_coarse[k] = base + p.proc[k][l + 1];
    }
    a.out_fine = l == 0 ? nullptr : base + p.output[l];
    a.out_image = out;
    a.fw = p.w[l], a.fh = p.h[l];
    a.rx0 = rx0[l], a.ry0 = ry0[l], a.rx1 = rx1[l], a.ry1 = ry1[l];
    a.max_supp = p.max_supp, a.width = width;
    // 2 x 2 pixels per thread where no boundary clamp can act (clamp_boundary leaves 1 <= x <= w - 3 alone): the region is widened
    // to even bounds (the extra pixels are valid results nobody reads)
    const int ex0 = a.rx0 & ~1, ey0 = a.ry0 & ~1, ex1 = min(a.fw, (a.rx1 + 1) & ~1), ey1 = min(a.fh, (a.ry1 + 1) & ~1);
    if (ex0 >= 2 && ey0 >= 2 && ex1 <= a.fw - 2 && ey1 <= a.fh - 2) {
      if (l > 0) a.rx0 = ex0, a.ry0 = ey0, a.rx1 = ex1, a.ry1 = ey1;  // level 0 keeps the exact image rectangle (max_supp is even)
      const dim3 grid2(div_up(a.rx1 - a.rx0, 64), div_up(a.ry1 - a.ry0, 16));
      if (l == 0) assemble2_kernel<true><<<grid2, kThreads, 0, s>>>(a, cp);
      else assemble2_kernel<false><<<grid2, kThreads, 0, s>>>(a, cp);
    } else {
      const dim3 grid(div_up(a.rx1 - a.rx0, 32), div_up(a.ry1 - a.ry0, 8));
      if (l == 0) assemble_kernel<true><<<grid, kThreads, 0, s>>>(a, cp);
      else assemble_kernel<false><<<grid, kThreads, 0, s>>>(a, cp);
    }
    if (int e = check_launch("laplacian_assemble")) return e;
  }
  return TDB_OK;
}

}  // extern "C"
