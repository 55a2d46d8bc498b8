// Local Laplacian filter (darktable local-laplacian, 6 gamma curves, fp16 pyramids like the reference).
//
// Reference: csrc/local_contrast/laplacian.cu:446-582 = pad + input pyramid + 6 x (curve over the padded level 0 + its
// pyramid) + assemble per level + write back: ~100 launches, 8 fp16 pyramids of the 2.5x padded frame in HBM
// (about 190 B per image pixel at 50 MP) and pointer tables in __device__ globals.
// Here the six tone-curved copies of level 0 are never materialised: the first reduction applies the curve while it
// stages its fine patch in shared memory, and the level-0 assemble recomputes curve(padded) for the two gammas it
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
    p.padded[l] = take(n);
    p.output[l] = take(n);
    for (int k = 0; k < G; k++) p.proc[k][l] = l == 0 ? 0 : take(n);  // level 0 of the curved pyramids is virtual
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
  val += cp.clarity * c * expf(-c * c / (2.0f * cp.sigma * cp.sigma / 3.0f));
  return val;
}

__device__ __forceinline__ float h2f(__half h) { return __half2float(h); }
__device__ __forceinline__ __half f2h(float f) { return __float2half_rn(f); }

// replicate padding, fp32 -> fp16 (laplacian.cu:90-109)
__global__ void __launch_bounds__(kThreads) pad_kernel(const float *__restrict__ in, __half *__restrict__ padded, int width, int height,
                                                       int max_supp, int bw, int bh) {
  const int x = blockIdx.x * 64 + 2 * (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= bw || y >= bh) return;
  const int cy = min(max(y - max_supp, 0), height - 1);
  const int cx0 = min(max(x - max_supp, 0), width - 1), cx1 = min(max(x + 1 - max_supp, 0), width - 1);
  const float v0 = __ldg(in + (int64_t)cy * width + cx0), v1 = __ldg(in + (int64_t)cy * width + cx1);
  __half *o = padded + (int64_t)y * bw + x;
  if (x + 1 < bw && (((int64_t)y * bw + x) & 1) == 0) *reinterpret_cast<__half2 *>(o) = __halves2half2(f2h(v0), f2h(v1));
  else {
    o[0] = f2h(v0);
    if (x + 1 < bw) o[1] = f2h(v1);
  }
}

struct ReduceBatch {
  const __half *fine[G];
  __half *coarse[G];
};

// 5x5 binomial, decimate by 2, clone a 1-px border (laplacian.cu:178-208).  blockIdx.z selects the pyramid.
// kCurve: the fine level is curve_k(padded level 0), evaluated while the patch is staged (never stored in HBM).
constexpr int RT = 16;            // coarse tile edge
constexpr int RP = 2 * RT + 3;    // fine patch edge
template <bool kCurve>
__global__ void __launch_bounds__(kThreads) reduce_kernel(ReduceBatch b, int cw, int ch, int fw, int fh, CurveParams cp) {
  __shared__ float patch[RP][RP + 1];
  const int k = blockIdx.z;
  const __half *__restrict__ fine = b.fine[kCurve ? 0 : k];
  __half *__restrict__ coarse = b.coarse[k];
  const int cx0 = blockIdx.x * RT, cy0 = blockIdx.y * RT;
  // coarse pixel c reads fine 2*c'-2 .. 2*c'+2 with c' clamped to [1, size-2]; patch origin = 2*cx0 - 2 covers every
  // unclamped pixel of the tile, clamped border pixels are handled by reading through the same patch when possible
  const int fx0 = 2 * cx0 - 2, fy0 = 2 * cy0 - 2;
  const float g = (k + 0.5f) / (float)G;
  for (int i = threadIdx.x; i < RP * RP; i += kThreads) {
    const int ly = i / RP, lx = i - ly * RP;
    const int x = fx0 + lx, y = fy0 + ly;
    float v = 0.0f;
    if (x >= 0 && y >= 0 && x < fw && y < fh) {
      v = h2f(fine[(int64_t)y * fw + x]);
      if (kCurve) v = h2f(f2h(curve(v, g, cp)));  // the reference stores the curved level 0 as fp16
    }
    patch[ly][lx] = v;
  }
  __syncthreads();
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  const int x = cx0 + lx, y = cy0 + ly;
  if (x >= cw || y >= ch) return;
  int cx = x, cy = y;
  if (x >= cw - 1) cx = cw - 2;
  if (y >= ch - 1) cy = ch - 2;
  if (cx <= 0) cx = 1;
  if (cy <= 0) cy = 1;
  const float w[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  float acc = 0.0f;
  const int px = 2 * cx - fx0, py = 2 * cy - fy0;  // patch coordinates of the stencil centre
  const bool inside = px >= 2 && py >= 2 && px + 2 < RP && py + 2 < RP;
  if (inside) {
#pragma unroll
    for (int j = -2; j <= 2; j++)
#pragma unroll
      for (int i = -2; i <= 2; i++) acc += patch[py + j][px + i] * w[i + 2] * w[j + 2];
  } else {  // a clamped border pixel whose stencil left the staged patch (only at the far image edge)
#pragma unroll
    for (int j = -2; j <= 2; j++)
#pragma unroll
      for (int i = -2; i <= 2; i++) {
        float v = h2f(fine[(int64_t)(2 * cy + j) * fw + (2 * cx + i)]);
        if (kCurve) v = h2f(f2h(curve(v, g, cp)));
        acc += v * w[i + 2] * w[j + 2];
      }
  }
  coarse[(int64_t)y * cw + x] = f2h(acc);
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
      c += h2f(coarse[(int64_t)(cy + j) * cw + (cx + i)]) * w[wi] * w[wj];
    }
  return 4.0f * c;
}

struct AssembleArgs {
  const __half *padded;      // gaussian of the input at this (fine) level
  const __half *out_coarse;  // reconstruction one level coarser
  const __half *proc_fine[G];    // curved gaussians, this level (unused at level 0)
  const __half *proc_coarse[G];  // curved gaussians, one level coarser
  __half *out_fine;          // reconstruction at this level (levels >= 1)
  float *out_image;          // final fp32 image (level 0)
  int fw, fh;                // fine level size
  int rx0, ry0, rx1, ry1;    // region to compute (half-open), fine coordinates
  int max_supp, width;       // level 0: crop origin and output row length
};

template <bool kLevel0>
__global__ void __launch_bounds__(kThreads) assemble_kernel(AssembleArgs a, CurveParams cp) {  // laplacian.cu:222-263
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
  const float v = h2f(a.padded[(int64_t)y * w + x]);
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

  pad_kernel<<<dim3(div_up(p.bw, 64), div_up(p.bh, 8)), kThreads, 0, s>>>(lum, base + p.padded[0], width, height, p.max_supp, p.bw, p.bh);
  if (int e = check_launch("laplacian_pad")) return e;

  // gaussian pyramid of the input; the coarsest level seeds the reconstruction (laplacian.cu:515-528)
  for (int l = 1; l < L; l++) {
    ReduceBatch b{};
    b.fine[0] = base + p.padded[l - 1];
    b.coarse[0] = base + (l == L - 1 ? p.output[l] : p.padded[l]);
    reduce_kernel<false><<<dim3(div_up(p.w[l], RT), div_up(p.h[l], RT), 1), kThreads, 0, s>>>(b, p.w[l], p.h[l], p.w[l - 1], p.h[l - 1], cp);
    if (int e = check_launch("laplacian_reduce")) return e;
  }
  // the six curved pyramids: level 1 straight from padded level 0 (curve fused), deeper levels batched over gammas
  for (int l = 1; l < L; l++) {
    ReduceBatch b{};
    for (int k = 0; k < G; k++) {
      b.fine[k] = l == 1 ? base + p.padded[0] : base + p.proc[k][l - 1];
      b.coarse[k] = base + p.proc[k][l];
    }
    const dim3 grid(div_up(p.w[l], RT), div_up(p.h[l], RT), G);
    if (l == 1) reduce_kernel<true><<<grid, kThreads, 0, s>>>(b, p.w[l], p.h[l], p.w[l - 1], p.h[l - 1], cp);
    else reduce_kernel<false><<<grid, kThreads, 0, s>>>(b, p.w[l], p.h[l], p.w[l - 1], p.h[l - 1], cp);
    if (int e = check_launch("laplacian_reduce_curves")) return e;
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
    a.padded = base + p.padded[l];
    a.out_coarse = base + p.output[l + 1];
    for (int k = 0; k < G; k++) {
      a.proc_fine[k] = l == 0 ? nullptr : base + p.proc[k][l];
      a.proc_coarse[k] = base + p.proc[k][l + 1];
    }
    a.out_fine = base + p.output[l];
    a.out_image = out;
    a.fw = p.w[l], a.fh = p.h[l];
    a.rx0 = rx0[l], a.ry0 = ry0[l], a.rx1 = rx1[l], a.ry1 = ry1[l];
    a.max_supp = p.max_supp, a.width = width;
    const dim3 grid(div_up(a.rx1 - a.rx0, 32), div_up(a.ry1 - a.ry0, 8));
    if (l == 0) assemble_kernel<true><<<grid, kThreads, 0, s>>>(a, cp);
    else assemble_kernel<false><<<grid, kThreads, 0, s>>>(a, cp);
    if (int e = check_launch("laplacian_assemble")) return e;
  }
  return TDB_OK;
}

}  // extern "C"
