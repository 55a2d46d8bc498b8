// Bilinear 5x5 and PPG demosaic as single fused shared-memory stencils.
//
// bilinear: replaces csrc/debayer/bilinear.cu (13 clamped global loads per pixel, 12-byte scattered stores).
// PPG:      replaces the 3-4 kernel chain of csrc/debayer/ppg.cu:413-463 (border -> [pre-median] -> green -> red/blue,
//           each a full HBM round trip plus a torch::zeros) by ONE kernel: the CFA patch is staged once (float plane or
//           12-bit packed bytes), the optional median, the green plane and the red/blue fill all live in shared memory,
//           and the RGB tile leaves through 128-bit stores.  Algorithmic traffic: 4 (or 1.5) B in + 12 B out per pixel.
#include <cstdlib>

#include "cfa_tile.cuh"

namespace tdb {
namespace {

constexpr int kTile = 32;      // output tile edge
constexpr int kThreads = 256;  // 16 x 16 launch, 4 pixels per thread

// ------------------------------------------------------------------------------------------------------------------
// bilinear 5x5: weights are indexed by the pixel's position in the RGGB-ordered quad (0 = R site, 1 = G on an R row,
// 2 = G on a B row, 3 = B site); taps outside the image use clamped coordinates (reference bilinear.cu:90).
// v1 of this kernel looked the 39 weights of a pixel up in __constant__ memory by its site type, which differs between
// neighbouring lanes: the constant cache serialises divergent addresses, and the kernel ran three times slower than the
// reference (0.457 vs 0.146 ms at 24 MP).  Here a thread owns one 2x2 quad, so all four site types sit in one thread and
// every weight is a compile-time immediate (zero weights vanish; a zero weight contributes an exact +0, so the sums are
// unchanged), the 6x6 neighbourhood is loaded once as 64-bit words and shared by the four pixels, and the CFA pattern is a
// template argument.  The accumulation order per channel is the tap order of the reference (bilinear.cu:17-23).
__host__ __device__ constexpr int bil_dx(int k) { constexpr int d[13] = {-2, -1, -1, -1, 0, 0, 0, 0, 0, 1, 1, 1, 2}; return d[k]; }
__host__ __device__ constexpr int bil_dy(int k) { constexpr int d[13] = {0, -1, 0, 1, -2, -1, 0, 1, 2, -1, 0, 1, 0}; return d[k]; }
__host__ __device__ constexpr float bil_w(int type, int k, int c) {
  constexpr float w[4][13][3] = {
      {{0, -2, -3}, {0, 0, 4}, {0, 4, 0}, {0, 0, 4}, {0, -2, -3}, {0, 4, 0}, {16, 8, 12}, {0, 4, 0}, {0, -2, -3}, {0, 0, 4}, {0, 4, 0}, {0, 0, 4}, {0, -2, -3}},
      {{-2, 0, 1}, {-2, 0, -2}, {8, 0, 0}, {-2, 0, -2}, {1, 0, -2}, {0, 0, 8}, {10, 16, 10}, {0, 0, 8}, {1, 0, -2}, {-2, 0, -2}, {8, 0, 0}, {-2, 0, -2}, {-2, 0, 1}},
      {{1, 0, -2}, {-2, 0, -2}, {0, 0, 8}, {-2, 0, -2}, {-2, 0, 1}, {8, 0, 0}, {10, 16, 10}, {8, 0, 0}, {-2, 0, 1}, {-2, 0, -2}, {0, 0, 8}, {-2, 0, -2}, {1, 0, -2}},
      {{-3, -2, 0}, {4, 0, 0}, {0, 4, 0}, {4, 0, 0}, {-3, -2, 0}, {0, 4, 0}, {12, 8, 16}, {0, 4, 0}, {-3, -2, 0}, {4, 0, 0}, {0, 4, 0}, {4, 0, 0}, {-3, -2, 0}}};
  return w[type][k][c];
}

// pixel type of quad position c = (x&1) + 2*(y&1); two bits per entry:
// RGGB {0,1,2,3} = 0xE4, BGGR {3,1,2,0} = 0x27, GRBG {1,0,3,2} = 0xB1, GBRG {1,3,0,2} = 0x8D
__host__ __device__ constexpr uint32_t quad_code(uint32_t filters) {
  return filters == TDB_FILTERS_RGGB ? 0xE4u : filters == TDB_FILTERS_BGGR ? 0x27u : filters == TDB_FILTERS_GRBG ? 0xB1u : 0x8Du;
}

template <int kType, int kC, int kTap>
struct BilTaps {  // acc = fma(w, v, acc) for taps kTap..12 in order, non-zero weights only
  static __device__ __forceinline__ float run(const float (&nb)[6][6], int dy, int dx, float acc) {
    constexpr float w = bil_w(kType, kTap, kC);
    if (w != 0.0f) acc = fmaf(w, nb[2 + dy + bil_dy(kTap)][2 + dx + bil_dx(kTap)], acc);
    if constexpr (kTap < 12) return BilTaps<kType, kC, kTap + 1>::run(nb, dy, dx, acc);
    else return acc;
  }
};

template <int kType>
__device__ __forceinline__ void bil_pixel(const float (&nb)[6][6], int dy, int dx, float *out) {
  // every weight column sums to 16
  out[0] = BilTaps<kType, 0, 0>::run(nb, dy, dx, 0.0f) * 0.0625f;
  out[1] = BilTaps<kType, 1, 0>::run(nb, dy, dx, 0.0f) * 0.0625f;
  out[2] = BilTaps<kType, 2, 0>::run(nb, dy, dx, 0.0f) * 0.0625f;
}

// One CTA-wide tensor-map load of a (box_w x box_h) float patch whose first element is image pixel (px0, py0): thread 0 arms the
// barrier and issues the copy, everybody waits for its bytes.  Pixels outside the image arrive as zeros.
__device__ __forceinline__ void stage_patch_tma(float *patch, uint64_t *bar, const CUtensorMap *tmap, int px0, int py0, uint32_t bytes) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(bar, bytes);
    tma_load_2d(patch, tmap, px0, py0, bar);
  }
  mbar_wait(bar, 0);
}

// use_tma: `tmap` describes src.cfa (float plane, width % 4 == 0).  Tiles whose 36 x 36 patch lies inside the image take the tensor-map
// load (SASS UTMALDG); the others need the reference's clamp-to-edge coordinates (bilinear.cu:90), which the TMA unit cannot produce
template <uint32_t kCode>
__global__ void __launch_bounds__(kThreads) bilinear_kernel(CfaSource src, float *__restrict__ rgb, int width, int height,
                                                            uint32_t filters, const __grid_constant__ CUtensorMap tmap, int use_tma) {
  // Patch rows hold image columns x0 - 4 .. x0 + 35 (stride 40), of which the stencil reads x0 - 2 .. x0 + 33: a tensor-map box must start
  // on a 16-byte boundary of its row (a box at x0 - 2 faulted with 'illegal instruction' on B200), and x0 - 4 is one when width % 4 == 0
  constexpr int P = kTile + 4, S = kTile + 8, X0 = 2;  // X0: patch column of image column x0 - 2
  extern __shared__ __align__(128) float smem[];
  float *patch = smem, *outt = smem + P * S;
  __shared__ uint64_t bar;
  resolve_gains(src, filters);
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  if (use_tma && x0 >= 2 && y0 >= 2 && x0 + kTile + 2 <= width && y0 + kTile + 2 <= height) {
    stage_patch_tma(patch, &bar, &tmap, x0 - 4, y0 - 2, S * P * sizeof(float));  // the box may reach past the image: zeros nobody reads
  } else {
    stage_patch<Oob::kClamp, false>(patch + X0, S, x0 - 2, y0 - 2, P, P, src, width, height);
    __syncthreads();
  }
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int qx = tid & 15, qy = tid >> 4;  // one quad per thread, tile origin even in both axes
  float nb[6][6];                          // patch rows 2qy .. 2qy+5, columns 2qx .. 2qx+5
#pragma unroll
  for (int r = 0; r < 6; r++) {
    const float2 *row = reinterpret_cast<const float2 *>(patch + X0 + (2 * qy + r) * S + 2 * qx);
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const float2 v = row[j];
      nb[r][2 * j] = v.x, nb[r][2 * j + 1] = v.y;
    }
  }
  float *o = outt + 3 * (2 * qy * kTile + 2 * qx);
  bil_pixel<(kCode >> 0) & 3>(nb, 0, 0, o);
  bil_pixel<(kCode >> 2) & 3>(nb, 0, 1, o + 3);
  bil_pixel<(kCode >> 4) & 3>(nb, 1, 0, o + 3 * kTile);
  bil_pixel<(kCode >> 6) & 3>(nb, 1, 1, o + 3 * kTile + 3);
  __syncthreads();
  store_rgb_tile(outt, kTile * 3, rgb, x0, y0, kTile, kTile, width, height);
}

// ------------------------------------------------------------------------------------------------------------------
// PPG.  Stages (all on shared memory, image coordinates in comments):
//   cfa  : raw patch, halo 6 (4 without the median), zero outside the image
//   med  : optional thresholded 9-tap same-colour median, halo 4                     (ppg.cu:21-113)
//   tmp  : per-pixel RGB after border_interpolate(3) / green fill, halo 1            (ppg.cu:342-389, :120-223)
//   out  : red/blue fill from tmp                                                    (ppg.cu:230-337)
// One pixel of the `tmp` stage the slow way: outside the image, or in the 3-px ring that border_interpolate owns (ppg.cu:342-389)
__device__ __forceinline__ void ppg_tmp_border(const float *cfa_px, int sc, int x, int y, int width, int height, uint32_t filters, float *t) {
  float r = 0.0f, g = 0.0f, b = 0.0f;  // zero outside the image (ppg.cu:270)
  if (x >= 0 && y >= 0 && x < width && y < height) {
    const int c = fc(y, x, filters);
    // 3x3 same-colour averages of the RAW cfa clamped at 0
    float sum[3] = {0, 0, 0};
    int cnt[3] = {0, 0, 0};
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        const int xx = x + dx, yy = y + dy;
        if (xx >= 0 && yy >= 0 && xx < width && yy < height) {
          const int f = fc(yy, xx, filters);
          const float v = fmaxf(0.0f, cfa_px[dy * sc + dx]);
          sum[0] += f == 0 ? v : 0.0f, sum[1] += f == 1 ? v : 0.0f, sum[2] += f == 2 ? v : 0.0f;
          cnt[0] += f == 0, cnt[1] += f == 1, cnt[2] += f == 2;
        }
      }
    const float v = fmaxf(0.0f, cfa_px[0]);
    r = cnt[0] > 0 ? sum[0] / cnt[0] : v;
    g = cnt[1] > 0 ? sum[1] / cnt[1] : v;
    b = cnt[2] > 0 ? sum[2] / cnt[2] : v;
    if (c == 0) r = v; else if (c == 2) b = v; else g = v;
  }
  t[0] = r, t[1] = g, t[2] = b;
}

// One interior pixel of the `tmp` stage: the native sample, and PPG green at R / B sites (ppg.cu:120-223).
// kType: 0 = R site, 1 = G on an R row, 2 = G on a B row, 3 = B site (compile time: a thread owns a whole 2x2 quad)
template <int kType>
__device__ __forceinline__ void ppg_tmp_pixel(const float *p, int gs, float *t) {
  const float pc = p[0];
  float r = 0.0f, g = 0.0f, b = 0.0f;
  if (kType == 0) r = pc; else if (kType == 3) b = pc; else g = pc;
  if (kType == 0 || kType == 3) {
    const float pym = p[-gs], pym2 = p[-2 * gs], pym3 = p[-3 * gs];
    const float pyM = p[gs], pyM2 = p[2 * gs], pyM3 = p[3 * gs];
    const float pxm = p[-1], pxm2 = p[-2], pxm3 = p[-3], pxM = p[1], pxM2 = p[2], pxM3 = p[3];
    const float guessx = (pxm + pc + pxM) * 2.0f - pxM2 - pxm2;
    const float diffx = (fabsf(pxm2 - pc) + fabsf(pxM2 - pc) + fabsf(pxm - pxM)) * 3.0f + (fabsf(pxM3 - pxM) + fabsf(pxm3 - pxm)) * 2.0f;
    const float guessy = (pym + pc + pyM) * 2.0f - pyM2 - pym2;
    const float diffy = (fabsf(pym2 - pc) + fabsf(pyM2 - pc) + fabsf(pym - pyM)) * 3.0f + (fabsf(pyM3 - pyM) + fabsf(pym3 - pym)) * 2.0f;
    if (diffx > diffy) g = fmaxf(fminf(guessy * 0.25f, fmaxf(pym, pyM)), fminf(pym, pyM));
    else g = fmaxf(fminf(guessx * 0.25f, fmaxf(pxm, pxM)), fminf(pxm, pxM));
  }
  t[0] = fmaxf(r, 0.0f), t[1] = fmaxf(g, 0.0f), t[2] = fmaxf(b, 0.0f);
}

// One pixel of the red / blue fill (ppg.cu:230-337); p = this pixel in tmp, R = floats per tmp row
template <int kType, int R>
__device__ __forceinline__ void ppg_fill_pixel(const float *p, bool border, float *o) {
  float r = p[0], g = p[1], b = p[2];
  if (!border) {
    if (kType == 1 || kType == 2) {
      const float *nt = p - R, *nb = p + R, *nl = p - 3, *nr = p + 3;
      if (kType == 1) {  // the horizontal neighbours are red
        b = (nt[2] + nb[2] + 2.0f * g - nt[1] - nb[1]) * 0.5f;
        r = (nl[0] + nr[0] + 2.0f * g - nl[1] - nr[1]) * 0.5f;
      } else {
        r = (nt[0] + nb[0] + 2.0f * g - nt[1] - nb[1]) * 0.5f;
        b = (nl[2] + nr[2] + 2.0f * g - nl[1] - nr[1]) * 0.5f;
      }
    } else {
      const float *ntl = p - R - 3, *ntr = p - R + 3, *nbl = p + R - 3, *nbr = p + R + 3;
      constexpr int k = (kType == 0) ? 2 : 0;
      const float diff1 = fabsf(ntl[k] - nbr[k]) + fabsf(ntl[1] - g) + fabsf(nbr[1] - g);
      const float guess1 = ntl[k] + nbr[k] + 2.0f * g - ntl[1] - nbr[1];
      const float diff2 = fabsf(ntr[k] - nbl[k]) + fabsf(ntr[1] - g) + fabsf(nbl[1] - g);
      const float guess2 = ntr[k] + nbl[k] + 2.0f * g - ntr[1] - nbl[1];
      const float v = diff1 > diff2 ? guess2 * 0.5f : (diff1 < diff2 ? guess1 * 0.5f : (guess1 + guess2) * 0.25f);
      if (kType == 0) b = v; else r = v;
    }
  }
  o[0] = fmaxf(r, 0.0f), o[1] = fmaxf(g, 0.0f), o[2] = fmaxf(b, 0.0f);
}

// v1 of the two stages below ran one pixel per thread with the site colour evaluated at run time: half of the lanes idled in the
// green stage and the two branches of the fill ran one after the other.  Here a thread owns a 2x2 quad, so all four site types
// sit in one thread and the CFA pattern is a template argument.
// site type of quad position (row & 1, col & 1) from the real CFA colours (the table of the bilinear kernel mirrors the
// reference's get_pixel_type, which is not the same thing for BGGR / GBRG)
__host__ __device__ constexpr int cfa_color(int row, int col, uint32_t filters) { return (filters >> ((((row << 1) & 14) + (col & 1)) << 1)) & 3u; }
__host__ __device__ constexpr int ppg_type(int row, int col, uint32_t filters) {
  const int c = cfa_color(row, col, filters);
  return c == 0 ? 0 : (c == 2 ? 3 : (cfa_color(row, col + 1, filters) == 0 ? 1 : 2));
}

template <bool kMedian, uint32_t kFilters>
__global__ void __launch_bounds__(kThreads) ppg_kernel(CfaSource src, float *__restrict__ rgb, int width, int height,
                                                       uint32_t filters, float threshold, const __grid_constant__ CUtensorMap tmap, int use_tma) {
  constexpr int HC = kMedian ? 6 : 4;             // cfa halo
  // cfa patch: PC x PC pixels from image column x0 - HC.  A tensor-map box must start on a 16-byte boundary of its row, so with the
  // median (HC = 6) the box starts two columns earlier, at x0 - 8, and is 48 wide; XA = patch column of image column x0 - HC
  constexpr int PC = kTile + 2 * HC, XA = (HC & 3), SC = PC + 2 * XA;
  constexpr int PM = kTile + 8, SM = PM + 1;      // median patch (halo 4)
  constexpr int PT = kTile + 2;                   // tmp patch (halo 1)
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  float *cfa = smem + XA;                               // PC rows of SC floats; cfa[ly * SC + lx] = image pixel (x0 - HC + lx, y0 - HC + ly)
  float *med = smem + PC * SC;                          // PM*SM (only with the median)
  float *tmp = med + (kMedian ? PM * SM : 0);           // PT*PT*3
  float *outt = tmp + PT * PT * 3;                      // kTile*kTile*3 (16-byte aligned by construction below)
  outt = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(outt) + 15) & ~uintptr_t(15));

  resolve_gains(src, filters);
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  // float plane: every tile, frame tiles included, is ONE tensor-map copy -- the zeros the TMA unit delivers for coordinates outside
  // the image are the zero-filled halo of the reference (ppg.cu:61,159,270); packed frames are unpacked by the threads
  if (use_tma) {
    stage_patch_tma(smem, &bar, &tmap, x0 - HC - XA, y0 - HC, SC * PC * sizeof(float));
  } else {
    stage_patch<Oob::kZero, false>(cfa, SC, x0 - HC, y0 - HC, PC, PC, src, width, height);
    __syncthreads();
  }

  const float *g_in;  // plane the green stage reads, with halo 4 around the tile
  int g_stride;
  if (kMedian) {
    constexpr int lim[5] = {0, 1, 2, 1, 0};
    for (int i = tid; i < PM * PM; i += kThreads) {
      const int my = i / PM, mx = i - my * PM;
      const int x = x0 - 4 + mx, y = y0 - 4 + my;
      float result = 0.0f;  // outside the image the next stage must see zeros (its own halo fill, ppg.cu:159)
      if (x >= 0 && y >= 0 && x < width && y < height) {
        const float *c = cfa + (my + 2) * SC + mx + 2;
        const float center = c[0];
        float v[9];
        int cnt = 0, k = 0;
#pragma unroll
        for (int r = 0; r < 5; r++)
#pragma unroll
          for (int j = -lim[r]; j <= lim[r]; j += 2) {
            const float t = c[(r - 2) * SC + j];
            if (fabsf(t - center) < threshold) v[k++] = t, cnt++;
            else v[k++] = 64.0f + t;
          }
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
          for (int b = a + 1; b < 9; b++)
            if (v[a] > v[b]) { const float t = v[a]; v[a] = v[b]; v[b] = t; }
        float color = center;
        if (fc(y, x, filters) & 1) {
          // med[(cnt-1)/2] with a runtime index: select without local memory
          const int want = (cnt == 1) ? 4 : (cnt - 1) / 2;
          float target = v[0];
#pragma unroll
          for (int a = 1; a < 9; a++) target = (a == want) ? v[a] : target;
          if (cnt == 1) target -= 64.0f;
          color = center + fminf(fmaxf(target - center, -threshold), threshold);
        }
        result = fmaxf(color, 0.0f);
      }
      med[my * SM + mx] = result;
    }
    __syncthreads();
    g_in = med, g_stride = SM;
  } else {
    g_in = cfa, g_stride = SC;  // HC == 4: same halo as the median patch
  }

  // tmp (34 x 34, origin (x0-1, y0-1)): border_interpolate for the outer 3 px ring, PPG green elsewhere.  Quads are aligned to even
  // image coordinates, so the 18 x 18 quads over [x0-2, x0+34) cover it; their outermost pixels fall outside tmp and are skipped.
  constexpr int T0 = ppg_type(0, 0, kFilters), T1 = ppg_type(0, 1, kFilters), T2 = ppg_type(1, 0, kFilters), T3 = ppg_type(1, 1, kFilters);
  for (int i = tid; i < 18 * 18; i += kThreads) {
    const int qy = i / 18, qx = i - qy * 18;
    const int tx = 2 * qx - 1, ty = 2 * qy - 1;          // tmp coordinates of the quad's first pixel
    const int x = x0 - 1 + tx, y = y0 - 1 + ty;          // image coordinates (even)
    const bool interior = x >= 3 && y >= 3 && x + 1 < width - 3 && y + 1 < height - 3;
    const bool c0 = tx >= 0, c1 = tx + 1 < PT, r0 = ty >= 0, r1 = ty + 1 < PT;
    float *t = tmp + 3 * (ty * PT + tx);
    if (interior) {
      const float *p = g_in + (ty - 1 + 4) * g_stride + (tx - 1 + 4);
      if (r0 && c0) ppg_tmp_pixel<T0>(p, g_stride, t);
      if (r0 && c1) ppg_tmp_pixel<T1>(p + 1, g_stride, t + 3);
      if (r1 && c0) ppg_tmp_pixel<T2>(p + g_stride, g_stride, t + 3 * PT);
      if (r1 && c1) ppg_tmp_pixel<T3>(p + g_stride + 1, g_stride, t + 3 * PT + 3);
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int dx = k & 1, dy = k >> 1;
        if (!((dy ? r1 : r0) && (dx ? c1 : c0))) continue;
        const int xx = x + dx, yy = y + dy;
        float *tp = t + 3 * (dy * PT + dx);
        if (xx >= 3 && yy >= 3 && xx < width - 3 && yy < height - 3) {
          const float *p = g_in + (ty + dy - 1 + 4) * g_stride + (tx + dx - 1 + 4);
          if (k == 0) ppg_tmp_pixel<T0>(p, g_stride, tp);
          else if (k == 1) ppg_tmp_pixel<T1>(p, g_stride, tp);
          else if (k == 2) ppg_tmp_pixel<T2>(p, g_stride, tp);
          else ppg_tmp_pixel<T3>(p, g_stride, tp);
        } else {
          ppg_tmp_border(cfa + (ty + dy - 1 + HC) * SC + (tx + dx - 1 + HC), SC, xx, yy, width, height, filters, tp);
        }
      }
    }
  }
  __syncthreads();

  // red / blue fill: one quad per thread
  {
    const int qx = tid & 15, qy = tid >> 4;
    const int lx = 2 * qx, ly = 2 * qy;
    const int x = x0 + lx, y = y0 + ly;
    constexpr int R = 3 * PT;  // one tmp row
    const float *p = tmp + 3 * ((ly + 1) * PT + lx + 1);
    float *o = outt + 3 * (ly * kTile + lx);
    // pixels outside the image are never stored; the outermost image ring keeps its tmp value
    const bool bx0 = x >= width || x == 0 || x == width - 1, bx1 = x + 1 >= width || x + 1 == width - 1;
    const bool by0 = y >= height || y == 0 || y == height - 1, by1 = y + 1 >= height || y + 1 == height - 1;
    ppg_fill_pixel<T0, R>(p, bx0 || by0, o);
    ppg_fill_pixel<T1, R>(p + 3, bx1 || by0, o + 3);
    ppg_fill_pixel<T2, R>(p + R, bx0 || by1, o + 3 * kTile);
    ppg_fill_pixel<T3, R>(p + R + 3, bx1 || by1, o + 3 * kTile + 3);
  }
  __syncthreads();
  store_rgb_tile(outt, kTile * 3, rgb, x0, y0, kTile, kTile, width, height);
}

template <bool kMedian>
constexpr size_t ppg_smem_bytes() {
  constexpr int HC = kMedian ? 6 : 4, PC = kTile + 2 * HC, SC = PC + 2 * (HC & 3), PM = kTile + 8, SM = PM + 1, PT = kTile + 2;
  return sizeof(float) * (PC * SC + (kMedian ? PM * SM : 0) + PT * PT * 3 + kTile * kTile * 3) + 16;
}

int check_frame(const char *name, int width, int height) {
  if (width < 16 || height < 16 || (width & 1) || (height & 1)) {
    set_error("%s: width and height must be even and >= 16 (got %dx%d)", name, width, height);
    return TDB_EINVAL;
  }
  return TDB_OK;
}

}  // namespace

int launch_bilinear(const CfaSource &src, float *rgb, int width, int height, uint32_t filters, cudaStream_t s) {
  dim3 block(16, 16), grid(div_up(width, kTile), div_up(height, kTile));
  CUtensorMap tmap;
  static const bool tma_on = getenv("TDB_BILINEAR_TMA") == nullptr || atoi(getenv("TDB_BILINEAR_TMA")) != 0;
  const int tma = tma_on && make_tensor_map_f32(&tmap, src.cfa, width, height, kTile + 8, kTile + 4) ? 1 : 0;
  constexpr size_t bytes = ((kTile + 4) * (kTile + 8) + kTile * kTile * 3) * sizeof(float);
  switch (quad_code(filters)) {
    case 0xE4u: bilinear_kernel<0xE4u><<<grid, block, bytes, s>>>(src, rgb, width, height, filters, tmap, tma); break;
    case 0x27u: bilinear_kernel<0x27u><<<grid, block, bytes, s>>>(src, rgb, width, height, filters, tmap, tma); break;
    case 0xB1u: bilinear_kernel<0xB1u><<<grid, block, bytes, s>>>(src, rgb, width, height, filters, tmap, tma); break;
    default: bilinear_kernel<0x8Du><<<grid, block, bytes, s>>>(src, rgb, width, height, filters, tmap, tma); break;
  }
  return check_launch("bilinear5x5_demosaic");
}

int launch_ppg(const CfaSource &src, float *rgb, int width, int height, uint32_t filters, float median_threshold, cudaStream_t s) {
  dim3 block(16, 16), grid(div_up(width, kTile), div_up(height, kTile));
  CUtensorMap tmap;
  const int patch = kTile + 2 * (median_threshold > 0.0f ? 6 : 4), box_w = median_threshold > 0.0f ? patch + 4 : patch;
  const int tma = make_tensor_map_f32(&tmap, src.cfa, width, height, box_w, patch) ? 1 : 0;
#define TDB_PPG(CODE)                                                                                                             \
  if (median_threshold > 0.0f) {                                                                                                  \
    cudaFuncSetAttribute(ppg_kernel<true, CODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ppg_smem_bytes<true>());       \
    ppg_kernel<true, CODE><<<grid, block, ppg_smem_bytes<true>(), s>>>(src, rgb, width, height, filters, median_threshold / 100.0f, tmap, tma); \
  } else {                                                                                                                        \
    ppg_kernel<false, CODE><<<grid, block, ppg_smem_bytes<false>(), s>>>(src, rgb, width, height, filters, 0.0f, tmap, tma);      \
  }
  switch (filters) {
    case TDB_FILTERS_RGGB: TDB_PPG(TDB_FILTERS_RGGB) break;
    case TDB_FILTERS_BGGR: TDB_PPG(TDB_FILTERS_BGGR) break;
    case TDB_FILTERS_GRBG: TDB_PPG(TDB_FILTERS_GRBG) break;
    default: TDB_PPG(TDB_FILTERS_GBRG) break;
  }
#undef TDB_PPG
  return check_launch("ppg_demosaic");
}

}  // namespace tdb

using namespace tdb;

extern "C" {

int tdb_bilinear5x5(const float *cfa, float *rgb, int width, int height, uint32_t filters, tdb_stream_t stream) {
  TDB_REQUIRE(cfa && rgb, "bilinear5x5_demosaic: null pointer");
  if (int e = check_frame("bilinear5x5_demosaic", width, height)) return e;
  CfaSource src{};
  src.cfa = cfa;
  return launch_bilinear(src, rgb, width, height, filters, as_stream(stream));
}

int tdb_ppg(const float *cfa, float *rgb, int width, int height, uint32_t filters, float median_threshold, tdb_stream_t stream) {
  TDB_REQUIRE(cfa && rgb, "PPG: null pointer");
  if (int e = check_frame("PPG", width, height)) return e;
  CfaSource src{};
  src.cfa = cfa;
  return launch_ppg(src, rgb, width, height, filters, median_threshold, as_stream(stream));
}

}  // extern "C"
