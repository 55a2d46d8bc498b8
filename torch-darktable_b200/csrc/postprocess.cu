// Demosaic post-process: N x 3x3-median colour smoothing, global and local green equilibration.
//
// The reference (csrc/debayer/postprocess.cu:311-390) does copy_ + one kernel per smoothing pass + reduce kernel +
// torch sum + two blocking .item() reads + apply kernel + clone: about 156 B/px of HBM traffic and a host sync in the
// middle of the stream.  Here all smoothing passes run inside one shared-memory tile (halo = number of passes), the
// green sums are taken in the same kernel (smoothing never changes G apart from the >= 0 clamp), the ratio stays on the
// device, and one second kernel applies the global ratio and the local equilibration: 24 B/px per kernel, no sync.
#include "cfa_tile.cuh"
#include "frame_state.cuh"

namespace tdb {
namespace {

constexpr int T = 32;
constexpr int kThreads = 256;

struct Header {          // first bytes of the scratch buffer
  float sum1, sum2;      // G1 / G2 sums
  float ratio;           // sum2 / sum1 (or 1)
  float pad;
};

// statistics the smoothing kernel can leave behind: 1 = per-CTA green sums (the ratio kernel follows), 2 = the deferred form of
// the fused frame pipeline: green sums AND the min / max of every stride-th pixel, kept apart for the G1 sites (which the global
// equilibration multiplies by the ratio) and for everything else; frame_stats_kernel finishes them
struct SmoothStats {
  int mode;
  int stride;
  float *partials;
  FrameState *state;
  int first_in_set, last_in_set;
  const float *prev_bounds;  // EMA state (device float[2]) or null
  float moving_average;
  float *bounds_out;         // device float[2], written on the last frame of a set
  float *ratio_out;          // device float[1]
};

// ---- colour smoothing ------------------------------------------------------------------------------------------------
// v1 of this kernel ran the reference's 19-exchange network twice per pixel per pass on interleaved RGB in shared memory
// and was bound by the ALU pipe (FMNMX) and by index arithmetic (ncu: 830 thread instructions per output pixel).  Here:
//   * the tile lives as planes: G and the two colour differences R-G, B-G (ping-pong), so a pass reads 2 x 9 floats
//     instead of 27 and writes the differences the next pass needs
//   * a thread owns one column and walks down its row segment; each row contributes one SORTED horizontal triple per
//     channel (3 exchanges) that three consecutive outputs share, and the median of nine is
//     med3(max of the three lows, med3 of the three mids, min of the three highs) with FMNMX3 for the 3-input min/max:
//     16 min/max per channel per output instead of 38.  The median VALUE is the same as the reference network's.
//   * tile 56 x 32 outputs inside a 64-column patch: lanes map to consecutive columns (conflict-free planes), the patch
//     is staged and the result written with 128-bit accesses.
constexpr int SW = 56, SH = 32;      // output tile
constexpr int SHX = 4;               // x halo (keeps every patch row 16-byte aligned in global memory)
constexpr int SPW = SW + 2 * SHX;    // 64 patch columns
constexpr int kMaxFusedPasses = 4;   // <= SHX

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float med3(float a, float b, float c) { return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c)); }
struct Triple {
  float lo, mid, hi;
};
__device__ __forceinline__ Triple sort3(float a, float b, float c) {
  const float t1 = fminf(a, b), t2 = fmaxf(a, b);
  return Triple{fminf(t1, c), fmaxf(t1, fminf(t2, c)), fmaxf(t2, c)};
}
__device__ __forceinline__ float median_of_rows(const Triple &a, const Triple &b, const Triple &c) {
  return med3(max3(a.lo, b.lo, c.lo), med3(a.mid, b.mid, c.mid), min3(a.hi, b.hi, c.hi));
}

// One column segment of one pass: the thread walks down rows [rb, re) of its column.  Everything that does not change inside the
// walk is a template argument or an immediate: the plane stride (kPlane), whether the pass is the last one (writes the RGB tile
// and the green sums instead of the difference planes) and whether the patch lies inside the image (no per-pixel tests).  The row
// loop is unrolled by three with the sorted row triples renamed instead of moved; v2 of this loop spent more instructions on
// register rotation, five pointer walks and per-pixel predicates (IMAD 94, ISETP 31, VIADD 25, FSEL 16 per output pixel) than on
// the min/max network itself (143).
template <int kPasses, bool kLast, bool kInside>
__device__ __forceinline__ void smooth_rows(const float *__restrict__ src, float *__restrict__ dst, float *__restrict__ gpl,
                                            float *__restrict__ outt, int rb, int re, int col, int gx, int gy0, int width, int height,
                                            int we, int he, bool green_row0, int want_sums, float &s1, float &s2) {
  constexpr int kPlane = (SH + 2 * kPasses) * SPW;
  const bool col_in = gx >= 0 && gx < width;
  const float *p = src + (rb - 1) * SPW + col;
  Triple ra = sort3(p[-1], p[0], p[1]), ba = sort3(p[kPlane - 1], p[kPlane], p[kPlane + 1]);
  Triple rm = sort3(p[SPW - 1], p[SPW], p[SPW + 1]), bm = sort3(p[kPlane + SPW - 1], p[kPlane + SPW], p[kPlane + SPW + 1]);
  Triple rc, bc;
  p += 2 * SPW;                    // row rb + 1: the row below the first output
  float *pg = gpl + rb * SPW + col, *pd = dst + rb * SPW + col;
  float *po = outt + 3 * ((rb - kPasses) * SW + (col - SHX));
  // one output row: T2 = sorted triples of the row below (loaded here), T0 / T1 = the two rows above
#define TDB_SMOOTH_ROW(K, R0, B0, R1, B1, R2, B2)                                                                   \
  {                                                                                                                 \
    R2 = sort3(p[(K) * SPW - 1], p[(K) * SPW], p[(K) * SPW + 1]);                                                   \
    B2 = sort3(p[kPlane + (K) * SPW - 1], p[kPlane + (K) * SPW], p[kPlane + (K) * SPW + 1]);                        \
    const float mr = median_of_rows(R0, R1, R2), mb = median_of_rows(B0, B1, B2);                                   \
    const float g = pg[(K) * SPW];                                                                                  \
    float R = fmaxf(mr + g, 0.0f), B = fmaxf(mb + g, 0.0f), G = fmaxf(g, 0.0f);                                     \
    const int gy = gy0 + r + (K);                                                                                   \
    const bool pix_in = kInside || (col_in && gy >= 0 && gy < height);                                              \
    if (!kInside && !pix_in) R = 0.0f, G = 0.0f, B = 0.0f; /* outside stays zero for the next pass */               \
    if (!kLast) {                                                                                                   \
      pd[(K) * SPW] = R - G, pd[kPlane + (K) * SPW] = B - G;                                                        \
      pg[(K) * SPW] = G;                                                                                            \
    } else {                                                                                                        \
      po[(K) * 3 * SW] = R, po[(K) * 3 * SW + 1] = G, po[(K) * 3 * SW + 2] = B;                                     \
      if (want_sums) { /* green sums over the even-cropped image (postprocess.cu:195-203) */                        \
        const bool green = green_row0 != (bool)((r + (K)) & 1);                                                     \
        const bool counted = green && (kInside || (pix_in && gx < we && gy < he));                                  \
        const float gs = counted ? G : 0.0f;                                                                        \
        if (gy & 1) s2 += gs; else s1 += gs;                                                                        \
      }                                                                                                             \
    }                                                                                                               \
  }
  for (int r = rb; r < re; r += 3, p += 3 * SPW, pg += 3 * SPW, pd += 3 * SPW, po += 9 * SW) {
    TDB_SMOOTH_ROW(0, ra, ba, rm, bm, rc, bc)
    if (r + 1 < re) TDB_SMOOTH_ROW(1, rm, bm, rc, bc, ra, ba)
    if (r + 2 < re) TDB_SMOOTH_ROW(2, rc, bc, ra, ba, rm, bm)
  }
#undef TDB_SMOOTH_ROW
}

// kPasses = 1..kMaxFusedPasses smoothing passes (+ per-CTA green sums of the result)
template <int kPasses>
__global__ void __launch_bounds__(kThreads) smooth_kernel(const float *__restrict__ in, float *__restrict__ out, int width, int height,
                                                          uint32_t filters, const SmoothStats st) {
  extern __shared__ __align__(16) float sm[];
  constexpr int passes = kPasses;
  const int want_sums = st.mode;
  float *__restrict__ partials = st.partials;
  constexpr int PH = SH + 2 * passes;     // patch rows
  constexpr int plane = PH * SPW;
  // plane order: D0r D0b X D1r D1b G.  {X, D1r, D1b} stages the raw RGB patch; the output tile takes whichever
  // three contiguous planes the last pass does not read
  float *d0 = sm, *xpl = sm + 2 * plane, *d1 = sm + 3 * plane, *gpl = sm + 5 * plane;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * SW, y0 = blockIdx.y * SH;
  const int gx0 = x0 - SHX, gy0 = y0 - passes;  // image coordinates of patch cell (0, 0)

  // ---- stage the raw patch (zero outside the image, postprocess.cu:58-59)
  float *raw = xpl;  // [PH][SPW * 3]
  const bool vec = ((width & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  __shared__ __align__(8) uint64_t bar;
  if (vec && gx0 >= 0 && gy0 >= 0 && gx0 + SPW <= width && gy0 + PH <= height) {
    // patch inside the image: every patch row is 768 contiguous, 16-byte aligned bytes of the frame -> one bulk asynchronous copy
    // per row (cp.async.bulk, the 1-D form of the TMA engine) issued by the lanes of warp 0, completion through an mbarrier; no
    // thread spends issue slots or registers on moving the 29 KB
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid < 32) {
      if (tid == 0) mbar_arrive_expect_tx(&bar, (uint32_t)(PH * SPW * 3 * sizeof(float)));
      __syncwarp();
      for (int ly = tid; ly < PH; ly += 32)
        bulk_copy_g2s(raw + ly * SPW * 3, in + 3 * ((int64_t)(gy0 + ly) * width + gx0), (uint32_t)(SPW * 3 * sizeof(float)), &bar);
    }
    mbar_wait(&bar, 0);
  } else if (vec) {
    constexpr int QR = SPW * 3 / 4;  // float4 per patch row
    for (int i = tid; i < PH * QR; i += kThreads) {
      const int ly = i / QR, q = i - ly * QR;
      const int y = gy0 + ly, x = gx0 + (4 * q) / 3;  // first pixel touched; a float4 never straddles the image edge
      float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (y >= 0 && y < height && x >= 0 && x < width) v = ld_stream(reinterpret_cast<const float4 *>(in + 3 * ((int64_t)y * width + gx0)) + q);
      *reinterpret_cast<float4 *>(raw + ly * SPW * 3 + 4 * q) = v;
    }
  } else {
    for (int i = tid; i < PH * SPW * 3; i += kThreads) {
      const int ly = i / (SPW * 3), f = i - ly * (SPW * 3);
      const int y = gy0 + ly, x = gx0 + f / 3;
      raw[i] = (y >= 0 && y < height && x >= 0 && x < width) ? __ldg(in + 3 * ((int64_t)y * width + gx0) + f) : 0.0f;
    }
  }
  __syncthreads();
  for (int i = tid; i < PH * SPW; i += kThreads) {
    const float r = raw[3 * i], g = raw[3 * i + 1], b = raw[3 * i + 2];
    gpl[i] = g, d0[i] = r - g, d0[plane + i] = b - g;
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int cg = warp & 1, sg = warp >> 1;  // column group (2 x 32 lanes), row segment (4)
  float *outt = (passes & 1) ? xpl : d0;    // the last pass reads D[(passes-1)&1]; the tile takes the other three planes
  float s1 = 0.0f, s2 = 0.0f;
  float lo_g = FLT_MAX, hi_g = -FLT_MAX, lo_o = FLT_MAX, hi_o = -FLT_MAX;
  const int we = width & ~1, he = height & ~1;
  // patch fully inside the image (and inside its even crop): no per-pixel tests in the passes
  const bool inside = gx0 >= 0 && gy0 >= 0 && gx0 + SPW <= we && gy0 + PH <= he;
  const int green_par = fc(0, 0, filters) == 1 ? 0 : 1;  // green sites: (x + y) & 1 == green_par
#pragma unroll
  for (int m = 1; m <= passes; m++) {
    const float *src = ((m - 1) & 1) ? d1 : d0;
    float *dst = (m & 1) ? d1 : d0;
    const int grow = passes - m;  // how far beyond the output tile this pass must still be valid
    const int r_lo = passes - grow, r_hi = passes + SH + grow;  // rows [r_lo, r_hi)
    const int c_lo = SHX - grow, c_hi = SHX + SW + grow;
    const int seg = (r_hi - r_lo + 3) >> 2;
    const int rb = r_lo + sg * seg, re = min(rb + seg, r_hi);
    const int col = c_lo + cg * 32 + lane;
    if (col < c_hi && rb < re) {
      const int gx = gx0 + col;
      const bool green_row0 = ((gx ^ gy0) & 1) == green_par;  // is (gx, patch row 0) a green site
      if (m < passes) {
        if (inside) smooth_rows<kPasses, false, true>(src, dst, gpl, outt, rb, re, col, gx, gy0, width, height, we, he, green_row0, 0, s1, s2);
        else smooth_rows<kPasses, false, false>(src, dst, gpl, outt, rb, re, col, gx, gy0, width, height, we, he, green_row0, 0, s1, s2);
      } else {
        if (inside) smooth_rows<kPasses, true, true>(src, dst, gpl, outt, rb, re, col, gx, gy0, width, height, we, he, green_row0, want_sums, s1, s2);
        else smooth_rows<kPasses, true, false>(src, dst, gpl, outt, rb, re, col, gx, gy0, width, height, we, he, green_row0, want_sums, s1, s2);
      }
    }
    __syncthreads();
  }
  if (want_sums == 2) {
    // the pixels compute_image_bounds(stride) would read, straight from the finished tile (at most 5 x 8 of them)
    const int ly0 = (st.stride - y0 % st.stride) % st.stride, lx0 = (st.stride - x0 % st.stride) % st.stride;
    const int ny = ly0 < SH ? (SH - ly0 + st.stride - 1) / st.stride : 0, nx = lx0 < SW ? (SW - lx0 + st.stride - 1) / st.stride : 0;
    for (int i = tid; i < ny * nx; i += kThreads) {
      const int ly = ly0 + (i / nx) * st.stride, lx = lx0 + (i % nx) * st.stride;
      const int gy = y0 + ly, gx = x0 + lx;
      if (gy < height && gx < width) {
        const float *q = outt + 3 * (ly * SW + lx);
        const float R = q[0], G = q[1], B = q[2];
        if ((((gx ^ gy) & 1) == green_par) && !(gy & 1)) {  // G1 site: its green takes the ratio later
          lo_g = fminf(lo_g, G), hi_g = fmaxf(hi_g, G);
          lo_o = fminf(lo_o, fminf(R, B)), hi_o = fmaxf(hi_o, fmaxf(R, B));
        } else {
          lo_o = fminf(lo_o, fminf(fminf(R, G), B)), hi_o = fmaxf(hi_o, fmaxf(fmaxf(R, G), B));
        }
      }
    }
  }
  store_rgb_tile(outt, SW * 3, out, x0, y0, SW, SH, width, height);
  if (want_sums) {
    __shared__ float red[6][kThreads / 32];
    s1 = warp_sum(s1), s2 = warp_sum(s2);
    if (lane == 0) red[0][warp] = s1, red[1][warp] = s2;
    if (want_sums == 2) {
      lo_g = warp_min(lo_g), hi_g = warp_max(hi_g), lo_o = warp_min(lo_o), hi_o = warp_max(hi_o);
      if (lane == 0) red[2][warp] = lo_g, red[3][warp] = hi_g, red[4][warp] = lo_o, red[5][warp] = hi_o;
    }
    __syncthreads();
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
      float a = 0.0f, b = 0.0f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; w++) a += red[0][w], b += red[1][w];
      if (want_sums == 1) {
        partials[2 * blk] = a, partials[2 * blk + 1] = b;
      } else {
        float v[4] = {FLT_MAX, -FLT_MAX, FLT_MAX, -FLT_MAX};
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++)
          v[0] = fminf(v[0], red[2][w]), v[1] = fmaxf(v[1], red[3][w]), v[2] = fminf(v[2], red[4][w]), v[3] = fmaxf(v[3], red[5][w]);
        float *p = partials + 6 * blk;
        p[0] = a, p[1] = b, p[2] = v[0], p[3] = v[1], p[4] = v[2], p[5] = v[3];
      }
    }
  }
}

// Finishes the statistics of a smoothed frame from the per-CTA partials (one CTA; a ticket + __threadfence at the end of the
// smoothing kernel itself was measured: the fence holds every CTA until its tile stores have landed, +12 % on that kernel):
// ratio (postprocess.cu:362-366), bounds of the equilibrated image (x -> max(0, x * ratio) is monotone, so the extrema of the G1
// class commute with it), image-set merge and moving average (image_processor.py:288-290)
constexpr int kStatThreads = 1024;
__global__ void __launch_bounds__(kStatThreads) frame_stats_kernel(const SmoothStats st, unsigned int nblk) {
  __shared__ double dsum[2][kStatThreads / 32];
  __shared__ float red[4][kStatThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double a = 0.0, b = 0.0;
  float v[4] = {FLT_MAX, -FLT_MAX, FLT_MAX, -FLT_MAX};
  for (unsigned int i = tid; i < nblk; i += kStatThreads) {
    const float2 *p = reinterpret_cast<const float2 *>(st.partials + 6 * i);  // 24-byte records in a 256-byte aligned array
    const float2 s = __ldg(p), g = __ldg(p + 1), o = __ldg(p + 2);
    a += s.x, b += s.y;
    v[0] = fminf(v[0], g.x), v[1] = fmaxf(v[1], g.y), v[2] = fminf(v[2], o.x), v[3] = fmaxf(v[3], o.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o), b += __shfl_xor_sync(0xffffffffu, b, o);
  v[0] = warp_min(v[0]), v[1] = warp_max(v[1]), v[2] = warp_min(v[2]), v[3] = warp_max(v[3]);
  if (lane == 0) dsum[0][warp] = a, dsum[1][warp] = b, red[0][warp] = v[0], red[1][warp] = v[1], red[2][warp] = v[2], red[3][warp] = v[3];
  __syncthreads();
  if (tid == 0) {
    a = b = 0.0;
    for (int w = 0; w < kStatThreads / 32; w++) {
      a += dsum[0][w], b += dsum[1][w];
      v[0] = fminf(v[0], red[0][w]), v[1] = fmaxf(v[1], red[1][w]), v[2] = fminf(v[2], red[2][w]), v[3] = fmaxf(v[3], red[3][w]);
    }
    const float sum1 = (float)a, sum2 = (float)b;
    const float ratio = (sum1 > 0.0f && sum2 > 0.0f) ? sum2 / sum1 : 1.0f;
    *st.ratio_out = ratio;
    float lo = v[2], hi = v[3];
    if (v[0] <= v[1]) lo = fminf(lo, fmaxf(v[0] * ratio, 0.0f)), hi = fmaxf(hi, fmaxf(v[1] * ratio, 0.0f));
    FrameState *fs = st.state;
    if (!st.first_in_set) lo = fminf(lo, fs->set_lo), hi = fmaxf(hi, fs->set_hi);
    fs->set_lo = lo, fs->set_hi = hi;
    if (st.last_in_set) st.bounds_out[0] = ema(st.prev_bounds, 0, lo, st.moving_average), st.bounds_out[1] = ema(st.prev_bounds, 1, hi, st.moving_average);
  }
}

// The same partials reduced over a band of CTA rows only, left RAW (G1 sum, G2 sum, min / max of the sampled G1 greens, min / max of
// every other sampled value): what a rank contributes when one frame is split into row bands across GPUs (pipeline/tiled.py).
__global__ void __launch_bounds__(kStatThreads) band_stats_kernel(const float *__restrict__ partials, int ctas_x, int row_lo, int row_hi,
                                                                  float *__restrict__ raw_out) {
  __shared__ double dsum[2][kStatThreads / 32];
  __shared__ float red[4][kStatThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double a = 0.0, b = 0.0;
  float v[4] = {FLT_MAX, -FLT_MAX, FLT_MAX, -FLT_MAX};
  for (int i = row_lo * ctas_x + tid; i < row_hi * ctas_x; i += kStatThreads) {
    const float2 *p = reinterpret_cast<const float2 *>(partials + 6 * i);
    const float2 s = __ldg(p), g = __ldg(p + 1), o = __ldg(p + 2);
    a += s.x, b += s.y;
    v[0] = fminf(v[0], g.x), v[1] = fmaxf(v[1], g.y), v[2] = fminf(v[2], o.x), v[3] = fmaxf(v[3], o.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o), b += __shfl_xor_sync(0xffffffffu, b, o);
  v[0] = warp_min(v[0]), v[1] = warp_max(v[1]), v[2] = warp_min(v[2]), v[3] = warp_max(v[3]);
  if (lane == 0) dsum[0][warp] = a, dsum[1][warp] = b, red[0][warp] = v[0], red[1][warp] = v[1], red[2][warp] = v[2], red[3][warp] = v[3];
  __syncthreads();
  if (tid == 0) {
    a = b = 0.0;
    for (int w = 0; w < kStatThreads / 32; w++) {
      a += dsum[0][w], b += dsum[1][w];
      v[0] = fminf(v[0], red[0][w]), v[1] = fmaxf(v[1], red[1][w]), v[2] = fminf(v[2], red[2][w]), v[3] = fmaxf(v[3], red[3][w]);
    }
    raw_out[0] = (float)a, raw_out[1] = (float)b, raw_out[2] = v[0], raw_out[3] = v[1], raw_out[4] = v[2], raw_out[5] = v[3];
  }
}

// green sums of an image that is not smoothed (passes == 0): one CTA per 32 x 32 tile
__global__ void __launch_bounds__(kThreads) green_sums_kernel(const float *__restrict__ in, int width, int height, uint32_t filters,
                                                              float *__restrict__ partials) {
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * T, y0 = blockIdx.y * T;
  float s1 = 0.0f, s2 = 0.0f;
  const int we = width & ~1, he = height & ~1;
  for (int i = tid; i < T * T; i += kThreads) {
    const int x = x0 + (i & (T - 1)), y = y0 + i / T;
    if (x < we && y < he && fc(y & 1, x & 1, filters) == 1) {
      const float g = __ldg(in + 3 * ((int64_t)y * width + x) + 1);
      if (y & 1) s2 += g; else s1 += g;
    }
  }
  __shared__ float red[2][kThreads / 32];
  s1 = warp_sum(s1), s2 = warp_sum(s2);
  if ((tid & 31) == 0) red[0][tid >> 5] = s1, red[1][tid >> 5] = s2;
  __syncthreads();
  if (tid == 0) {
    float a = 0.0f, b = 0.0f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) a += red[0][w], b += red[1][w];
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    partials[2 * blk] = a, partials[2 * blk + 1] = b;
  }
}

// fixed-order reduction of the per-CTA partial sums -> ratio (postprocess.cu:362-366, without the host round trip)
__global__ void __launch_bounds__(1024) ratio_kernel(const float *__restrict__ partials, int n, Header *hdr) {
  __shared__ double s[2][32];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += partials[2 * i], b += partials[2 * i + 1];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o), b += __shfl_xor_sync(0xffffffffu, b, o);
  if ((threadIdx.x & 31) == 0) s[0][threadIdx.x >> 5] = a, s[1][threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    a = b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) a += s[0][w], b += s[1][w];
    const float sum1 = (float)a, sum2 = (float)b;
    hdr->sum1 = sum1, hdr->sum2 = sum2;
    hdr->ratio = (sum1 > 0.0f && sum2 > 0.0f) ? sum2 / sum1 : 1.0f;
  }
}

// global ratio on G1 sites (+ clamp) and local equilibration of G2 sites (postprocess.cu:84-169, :234-255)
__global__ void __launch_bounds__(kThreads) green_eq_kernel(const float *__restrict__ in, float *__restrict__ out, int width, int height,
                                                            uint32_t filters, int do_global, int do_local, float threshold,
                                                            const float *__restrict__ ratio_ptr) {
  constexpr int PW = T + 4;
  __shared__ float gsm[PW * PW];
  const float ratio = do_global ? __ldg(ratio_ptr) : 1.0f;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * T, y0 = blockIdx.y * T;
  if (do_local) {
    for (int i = tid; i < PW * PW; i += kThreads) {
      const int ly = i / PW, lx = i - ly * PW;
      const int x = x0 - 2 + lx, y = y0 - 2 + ly;
      float g = 0.0f;
      if (x >= 0 && y >= 0 && x < width && y < height) {
        g = __ldg(in + 3 * ((int64_t)y * width + x) + 1);
        if (do_global) g = fmaxf(g * ((fc(y & 1, x & 1, filters) == 1 && !(y & 1)) ? ratio : 1.0f), 0.0f);
      }
      gsm[i] = g;
    }
    __syncthreads();
  }
  for (int i = tid; i < T * T; i += kThreads) {
    const int ly = i / T, lx = i - ly * T;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= width || y >= height) continue;
    const float *p = in + 3 * ((int64_t)y * width + x);
    float r = __ldg(p), g = __ldg(p + 1), b = __ldg(p + 2);
    const int c = fc(y & 1, x & 1, filters);
    if (do_global) {
      g *= (c == 1 && !(y & 1)) ? ratio : 1.0f;
      r = fmaxf(r, 0.0f), g = fmaxf(g, 0.0f), b = fmaxf(b, 0.0f);
    }
    if (do_local) {
      float o = g;
      if (c == 1 && (y & 1)) {
        const float *q = gsm + (ly + 2) * PW + lx + 2;
        const float o1_1 = q[-PW - 1], o1_2 = q[-PW + 1], o1_3 = q[PW - 1], o1_4 = q[PW + 1];
        const float o2_1 = q[-2 * PW], o2_2 = q[2 * PW], o2_3 = q[-2], o2_4 = q[2];
        const float m1 = (o1_1 + o1_2 + o1_3 + o1_4) / 4.0f, m2 = (o2_1 + o2_2 + o2_3 + o2_4) / 4.0f;
        if (m2 > 0.0f && m1 > 0.0f && m1 / m2 < 2.0f) {
          const float c1 = (fabsf(o1_1 - o1_2) + fabsf(o1_1 - o1_3) + fabsf(o1_1 - o1_4) + fabsf(o1_2 - o1_3) + fabsf(o1_3 - o1_4) + fabsf(o1_2 - o1_4)) / 6.0f;
          const float c2 = (fabsf(o2_1 - o2_2) + fabsf(o2_1 - o2_3) + fabsf(o2_1 - o2_4) + fabsf(o2_2 - o2_3) + fabsf(o2_3 - o2_4) + fabsf(o2_2 - o2_4)) / 6.0f;
          if (o < 0.95f && c1 < threshold && c2 < threshold) o *= m1 / m2;
        }
      }
      g = fmaxf(o, 0.0f);
    }
    float *d = out + 3 * ((int64_t)y * width + x);
    d[0] = r, d[1] = g, d[2] = b;
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace tdb

using namespace tdb;

extern "C" {

size_t tdb_postprocess_scratch_bytes(int width, int height) {
  const size_t nblk = (size_t)div_up(width, T) * div_up(height, T);
  return align_up(sizeof(Header), 256) + align_up(nblk * 8 * sizeof(float), 256) + 2 * (size_t)width * height * 3 * sizeof(float);
}

// the smoothing chain: all passes fused in one launch when they fit (the pipeline default is 3), longer chains in chunks.
// `final_dst` receives the last pass; `stats` applies to the last launch.  Returns the number of CTAs of that launch in *nctas.
static int run_smoothing(const float *in, float *final_dst, float *img_a, float *img_b, int width, int height, uint32_t filters, int passes,
                         const SmoothStats &stats, size_t *nctas, cudaStream_t s) {
  static DeviceOnce attr;
  attr.run([&] {
    cudaFuncSetAttribute(smooth_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(6 * (SH + 2 * 1) * SPW * sizeof(float)));
    cudaFuncSetAttribute(smooth_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(6 * (SH + 2 * 2) * SPW * sizeof(float)));
    cudaFuncSetAttribute(smooth_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(6 * (SH + 2 * 3) * SPW * sizeof(float)));
    cudaFuncSetAttribute(smooth_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(6 * (SH + 2 * 4) * SPW * sizeof(float)));
  });
  const float *cur = in;
  int remaining = passes;
  while (remaining > 0) {
    const int chunk = remaining < kMaxFusedPasses ? remaining : kMaxFusedPasses;
    const bool last = (remaining - chunk) == 0;
    float *dst = last ? final_dst : (cur == img_a ? img_b : img_a);
    SmoothStats st = stats;
    if (!last) st.mode = 0;
    dim3 sgrid(div_up(width, SW), div_up(height, SH));
    const size_t bytes = 6 * (SH + 2 * chunk) * SPW * sizeof(float);
    switch (chunk) {
      case 1: smooth_kernel<1><<<sgrid, kThreads, bytes, s>>>(cur, dst, width, height, filters, st); break;
      case 2: smooth_kernel<2><<<sgrid, kThreads, bytes, s>>>(cur, dst, width, height, filters, st); break;
      case 3: smooth_kernel<3><<<sgrid, kThreads, bytes, s>>>(cur, dst, width, height, filters, st); break;
      default: smooth_kernel<4><<<sgrid, kThreads, bytes, s>>>(cur, dst, width, height, filters, st); break;
    }
    if (int e = check_launch("color_smoothing")) return e;
    if (nctas) *nctas = (size_t)sgrid.x * sgrid.y;
    cur = dst;
    remaining -= chunk;
  }
  return TDB_OK;
}

int tdb_postprocess(const float *in, float *out, void *scratch, int width, int height, uint32_t filters, int passes,
                    int green_eq_local, int green_eq_global, float green_eq_threshold, tdb_stream_t stream) {
  TDB_REQUIRE(in && out && scratch, "PostProcess: null pointer");
  TDB_REQUIRE(width > 0 && height > 0 && passes >= 0, "PostProcess: bad arguments");
  cudaStream_t s = as_stream(stream);
  const size_t nblk = (size_t)div_up(width, T) * div_up(height, T);
  char *base = static_cast<char *>(scratch);
  Header *hdr = reinterpret_cast<Header *>(base);
  float *partials = reinterpret_cast<float *>(base + align_up(sizeof(Header), 256));
  float *img_a = reinterpret_cast<float *>(base + align_up(sizeof(Header), 256) + align_up(nblk * 8 * sizeof(float), 256));
  float *img_b = img_a + (size_t)width * height * 3;
  dim3 grid(div_up(width, T), div_up(height, T));
  const bool eq = green_eq_local || green_eq_global;

  const float *cur = in;
  size_t npartials = nblk;
  if (passes > 0) {
    SmoothStats st{};
    st.mode = green_eq_global ? 1 : 0, st.partials = partials;
    // the last pass lands in `out` unless an equilibration pass follows (then in whichever scratch image is free)
    float *final_dst = !eq ? out : ((((passes + kMaxFusedPasses - 1) / kMaxFusedPasses) & 1) ? img_a : img_b);
    if (int e = run_smoothing(in, final_dst, img_a, img_b, width, height, filters, passes, st, &npartials, s)) return e;
    cur = final_dst;
  }
  if (green_eq_global && passes == 0) {
    green_sums_kernel<<<grid, kThreads, 0, s>>>(cur, width, height, filters, partials);
    if (int e = check_launch("green_sums")) return e;
  }
  if (green_eq_global) {
    ratio_kernel<<<1, 1024, 0, s>>>(partials, (int)npartials, hdr);
    if (int e = check_launch("green_eq_ratio")) return e;
  }
  if (eq) {
    green_eq_kernel<<<grid, kThreads, 0, s>>>(cur, out, width, height, filters, green_eq_global, green_eq_local,
                                              (float)(green_eq_threshold / 100.), &hdr->ratio);
    return check_launch("green_equilibration");
  }
  if (passes == 0) {  // nothing to do: the reference still returns a copy
    cudaMemcpyAsync(out, in, (size_t)width * height * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s);
    check_launch("postprocess_copy");
  }
  return TDB_OK;
}

// ---- fused frame pipeline, step A: smoothing with the statistics of the (not yet applied) global green equilibration
size_t tdb_frame_state_bytes(void) { return sizeof(FrameState); }

int tdb_postprocess_deferred(const float *in, float *out, void *scratch, int width, int height, uint32_t filters, int passes,
                             int bounds_stride, void *frame_state, int first_in_set, int last_in_set, const float *prev_bounds,
                             float moving_average, float *bounds_out, float *ratio_out, tdb_stream_t stream) {
  TDB_REQUIRE(in && out && scratch && frame_state && bounds_out && ratio_out, "postprocess_deferred: null pointer");
  TDB_REQUIRE(width > 0 && height > 0 && passes >= 1 && bounds_stride >= 1, "postprocess_deferred: bad arguments");
  cudaStream_t s = as_stream(stream);
  const size_t nblk = (size_t)div_up(width, T) * div_up(height, T);
  char *base = static_cast<char *>(scratch);
  float *partials = reinterpret_cast<float *>(base + align_up(sizeof(Header), 256));
  float *img_a = reinterpret_cast<float *>(base + align_up(sizeof(Header), 256) + align_up(nblk * 8 * sizeof(float), 256));
  float *img_b = img_a + (size_t)width * height * 3;
  SmoothStats st{};
  st.mode = 2, st.stride = bounds_stride, st.partials = partials, st.state = static_cast<FrameState *>(frame_state);
  st.first_in_set = first_in_set, st.last_in_set = last_in_set, st.prev_bounds = prev_bounds, st.moving_average = moving_average;
  st.bounds_out = bounds_out, st.ratio_out = ratio_out;
  size_t nctas = 0;
  if (int e = run_smoothing(in, out, img_a, img_b, width, height, filters, passes, st, &nctas, s)) return e;
  frame_stats_kernel<<<1, kStatThreads, 0, s>>>(st, (unsigned int)nctas);
  return check_launch("frame_stats");
}

int tdb_postprocess_deferred_band(const float *in, float *out, void *scratch, int width, int height, uint32_t filters, int passes,
                                  int bounds_stride, int row_lo, int row_hi, float *raw_out, tdb_stream_t stream) {
  TDB_REQUIRE(in && out && scratch && raw_out, "postprocess_deferred_band: null pointer");
  TDB_REQUIRE(width > 0 && height > 0 && passes >= 1 && bounds_stride >= 1, "postprocess_deferred_band: bad arguments");
  TDB_REQUIRE(row_lo >= 0 && row_lo < row_hi && row_hi <= height && row_lo % SH == 0 && (row_hi % SH == 0 || row_hi == height),
              "postprocess_deferred_band: the band [%d, %d) must consist of whole %d-row tiles", row_lo, row_hi, SH);
  cudaStream_t s = as_stream(stream);
  const size_t nblk = (size_t)div_up(width, T) * div_up(height, T);
  char *base = static_cast<char *>(scratch);
  float *partials = reinterpret_cast<float *>(base + align_up(sizeof(Header), 256));
  float *img_a = reinterpret_cast<float *>(base + align_up(sizeof(Header), 256) + align_up(nblk * 8 * sizeof(float), 256));
  float *img_b = img_a + (size_t)width * height * 3;
  SmoothStats st{};
  st.mode = 2, st.stride = bounds_stride, st.partials = partials;
  if (int e = run_smoothing(in, out, img_a, img_b, width, height, filters, passes, st, nullptr, s)) return e;
  band_stats_kernel<<<1, kStatThreads, 0, s>>>(partials, div_up(width, SW), row_lo / SH, div_up(row_hi, SH), raw_out);
  return check_launch("band_stats");
}

// ---- pieces of the post-process for a frame that is split into row tiles across GPUs: the green sums of the rows a rank owns
// (sums[0] = G1, sums[1] = G2; the caller all-reduces them and forms the ratio), and the green equilibration with a given ratio.
int tdb_green_sums(const float *rgb, int width, int height, uint32_t filters, void *scratch, float *sums, tdb_stream_t stream) {
  TDB_REQUIRE(rgb && scratch && sums && width > 0 && height > 0, "green_sums: bad arguments");
  cudaStream_t s = as_stream(stream);
  const size_t nblk = (size_t)div_up(width, T) * div_up(height, T);
  char *base = static_cast<char *>(scratch);
  Header *hdr = reinterpret_cast<Header *>(base);
  float *partials = reinterpret_cast<float *>(base + align_up(sizeof(Header), 256));
  dim3 grid(div_up(width, T), div_up(height, T));
  green_sums_kernel<<<grid, kThreads, 0, s>>>(rgb, width, height, filters, partials);
  if (int e = check_launch("green_sums")) return e;
  ratio_kernel<<<1, 1024, 0, s>>>(partials, (int)nblk, hdr);
  if (int e = check_launch("green_eq_ratio")) return e;
  cudaMemcpyAsync(sums, hdr, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s);
  return check_launch("green_sums_copy");
}

int tdb_green_eq_apply(const float *in, float *out, int width, int height, uint32_t filters, int green_eq_local, float green_eq_threshold,
                       const float *ratio, tdb_stream_t stream) {
  TDB_REQUIRE(in && out && ratio && width > 0 && height > 0, "green_eq_apply: bad arguments");
  dim3 grid(div_up(width, T), div_up(height, T));
  green_eq_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(in, out, width, height, filters, 1, green_eq_local,
                                                          (float)(green_eq_threshold / 100.), ratio);
  return check_launch("green_equilibration");
}

}  // extern "C"
