/* CPU oracle (test infrastructure, see oracle.h): bilinear 5x5, PPG, RCD demosaic and the demosaic post-process.
 * Follows torch_darktable/csrc/debayer/{bilinear,ppg,rcd,postprocess}.cu; each function cites the lines. */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FC(r, c) orc_fc((r), (c), filters)

/* ------------------------------------------------------------------------------------------------
 * bilinear 5x5 (13-tap diamond), bilinear.cu:17-99.  Pixel type = position inside the RGGB-ordered
 * 2x2 quad (bayer_device.h:14-33): 0 = R site, 1 = G on the R row, 2 = G on the B row, 3 = B site.
 * Coordinates are clamped to the edge (bilinear.cu:90).                                          */
static const int bil_dx[13] = {-2, -1, -1, -1, 0, 0, 0, 0, 0, 1, 1, 1, 2};
static const int bil_dy[13] = {0, -1, 0, 1, -2, -1, 0, 1, 2, -1, 0, 1, 0};
/* weights[type][tap][channel]; same numbers as bilinear.cu:28-61 (a Malvar-style kernel / 16) */
static const float bil_w[4][13][3] = {
    {{0, -2, -3}, {0, 0, 4}, {0, 4, 0}, {0, 0, 4}, {0, -2, -3}, {0, 4, 0}, {16, 8, 12}, {0, 4, 0}, {0, -2, -3}, {0, 0, 4}, {0, 4, 0}, {0, 0, 4}, {0, -2, -3}},
    {{-2, 0, 1}, {-2, 0, -2}, {8, 0, 0}, {-2, 0, -2}, {1, 0, -2}, {0, 0, 8}, {10, 16, 10}, {0, 0, 8}, {1, 0, -2}, {-2, 0, -2}, {8, 0, 0}, {-2, 0, -2}, {-2, 0, 1}},
    {{1, 0, -2}, {-2, 0, -2}, {0, 0, 8}, {-2, 0, -2}, {-2, 0, 1}, {8, 0, 0}, {10, 16, 10}, {8, 0, 0}, {-2, 0, 1}, {-2, 0, -2}, {0, 0, 8}, {-2, 0, -2}, {1, 0, -2}},
    {{-3, -2, 0}, {4, 0, 0}, {0, 4, 0}, {4, 0, 0}, {-3, -2, 0}, {0, 4, 0}, {12, 8, 16}, {0, 4, 0}, {-3, -2, 0}, {4, 0, 0}, {0, 4, 0}, {4, 0, 0}, {-3, -2, 0}}};

static int quad_type(uint32_t filters, int c) { /* bayer_device.h:14-33 */
  static const int rggb[4] = {0, 1, 2, 3}, bggr[4] = {3, 1, 2, 0}, grbg[4] = {1, 0, 3, 2}, gbrg[4] = {1, 3, 0, 2};
  switch (filters) {
    case 0x94949494u: return rggb[c];
    case 0x16161616u: return bggr[c];
    case 0x61616161u: return grbg[c];
    default: return gbrg[c];
  }
}

void orc_bilinear5x5(const float *cfa, float *rgb, int width, int height, uint32_t filters) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      const int type = quad_type(filters, (x & 1) + 2 * (y & 1)); /* offset2x2: C%2 = dx, C/2 = dy */
      float acc[3] = {0, 0, 0}, sum[3] = {0, 0, 0};
      for (int k = 0; k < 13; k++) {
        /* the reference's int2 offsets are {x, y} initialisers: offsets[k] = {dx, dy} (bilinear.cu:17-23) */
        int cx = x + bil_dx[k], cy = y + bil_dy[k];
        cx = cx < 0 ? 0 : (cx > width - 1 ? width - 1 : cx);
        cy = cy < 0 ? 0 : (cy > height - 1 ? height - 1 : cy);
        const float v = cfa[(long)cy * width + cx];
        for (int c = 0; c < 3; c++) {
          acc[c] = fmaf(bil_w[type][k][c], v, acc[c]);
          sum[c] += bil_w[type][k][c];
        }
      }
      float *o = rgb + 3 * ((long)y * width + x);
      for (int c = 0; c < 3; c++) o[c] = acc[c] / sum[c];
    }
}

/* ------------------------------------------------------------------------------------------------
 * shared PPG-style pieces                                                                        */

/* ppg.cu:342-389: 3x3 same-colour average for the outer `border` pixels */
static void border_interpolate(const float *in, float *out, int width, int height, uint32_t filters, int border) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      if (x >= border && x < width - border && y >= border && y < height - border) continue;
      float sum[4] = {0, 0, 0, 0};
      int count[4] = {0, 0, 0, 0};
      for (int j = y - 1; j <= y + 1; j++)
        for (int i = x - 1; i <= x + 1; i++)
          if (j >= 0 && i >= 0 && j < height && i < width) {
            const int f = FC(j, i);
            sum[f] += fmaxf(0.0f, in[(long)j * width + i]);
            count[f]++;
          }
      const float v = fmaxf(0.0f, in[(long)y * width + x]);
      float o[3];
      o[0] = count[0] > 0 ? sum[0] / count[0] : v;
      o[1] = count[1] + count[3] > 0 ? (sum[1] + sum[3]) / (count[1] + count[3]) : v;
      o[2] = count[2] > 0 ? sum[2] / count[2] : v;
      const int f = FC(y, x);
      if (f == 0) o[0] = v;
      else if (f == 2) o[2] = v;
      else o[1] = v;
      memcpy(out + 3 * ((long)y * width + x), o, sizeof o);
    }
}

/* zero outside the image, optionally clamped at 0 inside (rcd.cu:320 vs ppg.cu:159) */
static inline float tap(const float *in, int width, int height, int x, int y, int clamp0) {
  if (x < 0 || y < 0 || x >= width || y >= height) return 0.0f;
  const float v = in[(long)y * width + x];
  return clamp0 ? fmaxf(0.0f, v) : v;
}

/* ppg.cu:120-223 / rcd.cu:285-384: green at red/blue sites from the H/V gradient choice.
 * Writes only pixels >= 3 from every edge and (for RCD) inside the `border` frame. */
static void ppg_green(const float *in, float *out, int width, int height, uint32_t filters, int clamp0, int border) {
#pragma omp parallel for schedule(static)
  for (int y = 3; y < height - 3; y++)
    for (int x = 3; x < width - 3; x++) {
      if (border > 0 && x >= border && x < width - border && y >= border && y < height - border) continue;
      const int c = FC(y, x);
      float col[3] = {0, 0, 0};
#define T(dx, dy) tap(in, width, height, x + (dx), y + (dy), clamp0)
      const float pc = T(0, 0);
      if (c == 0) col[0] = pc;
      else if (c == 2) col[2] = pc;
      else col[1] = pc;
      if (c == 0 || c == 2) {
        const float pym = T(0, -1), pym2 = T(0, -2), pym3 = T(0, -3), pyM = T(0, 1), pyM2 = T(0, 2), pyM3 = T(0, 3);
        const float pxm = T(-1, 0), pxm2 = T(-2, 0), pxm3 = T(-3, 0), pxM = T(1, 0), pxM2 = T(2, 0), pxM3 = T(3, 0);
        const float guessx = (pxm + pc + pxM) * 2.0f - pxM2 - pxm2;
        const float diffx = (fabsf(pxm2 - pc) + fabsf(pxM2 - pc) + fabsf(pxm - pxM)) * 3.0f + (fabsf(pxM3 - pxM) + fabsf(pxm3 - pxm)) * 2.0f;
        const float guessy = (pym + pc + pyM) * 2.0f - pyM2 - pym2;
        const float diffy = (fabsf(pym2 - pc) + fabsf(pyM2 - pc) + fabsf(pym - pyM)) * 3.0f + (fabsf(pyM3 - pyM) + fabsf(pym3 - pym)) * 2.0f;
        if (diffx > diffy) {
          const float m = fminf(pym, pyM), M = fmaxf(pym, pyM);
          col[1] = fmaxf(fminf(guessy * 0.25f, M), m);
        } else {
          const float m = fminf(pxm, pxM), M = fmaxf(pxm, pxM);
          col[1] = fmaxf(fminf(guessx * 0.25f, M), m);
        }
      }
#undef T
      float *o = out + 3 * ((long)y * width + x);
      for (int k = 0; k < 3; k++) o[k] = fmaxf(col[k], 0.0f);
    }
}

static inline void px3(const float *in, int width, int height, int x, int y, int clamp0, float o[3]) {
  if (x < 0 || y < 0 || x >= width || y >= height) {
    o[0] = o[1] = o[2] = 0.0f;
    return;
  }
  const float *p = in + 3 * ((long)y * width + x);
  for (int k = 0; k < 3; k++) o[k] = clamp0 ? fmaxf(0.0f, p[k]) : p[k];
}

/* ppg.cu:230-337 / rcd.cu:387-493: red/blue from colour differences.  `in` may alias `out` for the RCD
 * variant (the reference runs it in place; the fields it reads are never the ones it writes, so the
 * result does not depend on the order). */
static void ppg_redblue(const float *in, float *out, int width, int height, uint32_t filters, int clamp0, int border) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      if (border > 0 && x >= border && x < width - border && y >= border && y < height - border) continue;
      const int c = FC(y, x);
      float col[3];
      px3(in, width, height, x, y, clamp0, col);
      if (!(x == 0 || y == 0 || x == width - 1 || y == height - 1)) {
#define P(name, dx, dy) \
  float name[3];        \
  px3(in, width, height, x + (dx), y + (dy), clamp0, name)
        if (c == 1 || c == 3) {
          P(nt, 0, -1);
          P(nb, 0, 1);
          P(nl, -1, 0);
          P(nr, 1, 0);
          if (FC(y, x + 1) == 0) {
            col[2] = (nt[2] + nb[2] + 2.0f * col[1] - nt[1] - nb[1]) * 0.5f;
            col[0] = (nl[0] + nr[0] + 2.0f * col[1] - nl[1] - nr[1]) * 0.5f;
          } else {
            col[0] = (nt[0] + nb[0] + 2.0f * col[1] - nt[1] - nb[1]) * 0.5f;
            col[2] = (nl[2] + nr[2] + 2.0f * col[1] - nl[1] - nr[1]) * 0.5f;
          }
        } else {
          P(ntl, -1, -1);
          P(ntr, 1, -1);
          P(nbl, -1, 1);
          P(nbr, 1, 1);
          const int k = (c == 0) ? 2 : 0; /* red site fills blue and vice versa */
          const float diff1 = fabsf(ntl[k] - nbr[k]) + fabsf(ntl[1] - col[1]) + fabsf(nbr[1] - col[1]);
          const float guess1 = ntl[k] + nbr[k] + 2.0f * col[1] - ntl[1] - nbr[1];
          const float diff2 = fabsf(ntr[k] - nbl[k]) + fabsf(ntr[1] - col[1]) + fabsf(nbl[1] - col[1]);
          const float guess2 = ntr[k] + nbl[k] + 2.0f * col[1] - ntr[1] - nbl[1];
          if (diff1 > diff2) col[k] = guess2 * 0.5f;
          else if (diff1 < diff2) col[k] = guess1 * 0.5f;
          else col[k] = (guess1 + guess2) * 0.25f;
        }
#undef P
      }
      float *o = out + 3 * ((long)y * width + x);
      for (int k = 0; k < 3; k++) o[k] = fmaxf(col[k], 0.0f);
    }
}

/* ppg.cu:21-113: thresholded same-colour 9-tap median on green sites */
static void pre_median(const float *in, float *out, int width, int height, uint32_t filters, float threshold) {
  static const int lim[5] = {0, 1, 2, 1, 0};
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      const float center = tap(in, width, height, x, y, 0);
      float med[9];
      int cnt = 0, k = 0;
      for (int i = 0; i < 5; i++)
        for (int j = -lim[i]; j <= lim[i]; j += 2) {
          const float v = tap(in, width, height, x + j, y + i - 2, 0);
          if (fabsf(v - center) < threshold) {
            med[k++] = v;
            cnt++;
          } else
            med[k++] = 64.0f + v;
        }
      for (int i = 0; i < 8; i++)
        for (int ii = i + 1; ii < 9; ii++)
          if (med[i] > med[ii]) {
            const float t = med[i];
            med[i] = med[ii];
            med[ii] = t;
          }
      float color = center;
      if (FC(y, x) & 1) {
        const float target = (cnt == 1) ? (med[4] - 64.0f) : med[(cnt - 1) / 2];
        color = center + fminf(fmaxf(target - center, -threshold), threshold);
      }
      out[(long)y * width + x] = fmaxf(color, 0.0f);
    }
}

void orc_ppg(const float *cfa, float *rgb, int width, int height, uint32_t filters, float median_threshold) {
  const long n = (long)width * height;
  float *tmp = calloc(3 * n, sizeof(float));
  float *med = NULL;
  const float *src = cfa;
  border_interpolate(cfa, tmp, width, height, filters, 3); /* ppg.cu:441 */
  if (median_threshold > 0.0f) {                           /* ppg.cu:445-452 */
    med = malloc(n * sizeof(float));
    pre_median(cfa, med, width, height, filters, median_threshold / 100.0f);
    src = med;
  }
  ppg_green(src, tmp, width, height, filters, 0, 0);   /* ppg.cu:454 */
  ppg_redblue(tmp, rgb, width, height, filters, 0, 0); /* ppg.cu:458 */
  free(tmp);
  free(med);
}

/* ------------------------------------------------------------------------------------------------
 * RCD, rcd.cu.  Literal flat-buffer restatement: same eight scratch planes, same launch order and the
 * same aliasing (VP_diff/HQ_diff hold v/h_diff at full-resolution indices in step 1 and p/q_diff at
 * idx/2 in step 4; lpf_PQ holds lpf then PQ_dir), because a thin band inside the 7-px margin depends
 * on it.                                                              */
static inline float sq(float x) { return x * x; }
static inline float mixf(float a, float b, float t) { return (1.0f - t) * a + t * b; }

void orc_rcd(const float *in, float *out, int width, int height, uint32_t filters, float *scratch) {
  const long n = (long)width * height;
  const int w = width, w2 = 2 * width, w3 = 3 * width, w4 = 4 * width;
  float *cfa = scratch, *rgb0 = scratch + n, *rgb1 = scratch + 2 * n, *rgb2 = scratch + 3 * n;
  float *VH_dir = scratch + 4 * n, *VP_diff = scratch + 5 * n, *HQ_diff = scratch + 6 * n, *lpf_PQ = scratch + 7 * n;
  float *rgbp[3] = {rgb0, rgb1, rgb2};

  border_interpolate(in, out, width, height, filters, 3); /* rcd.cu:616 */
  ppg_green(in, out, width, height, filters, 1, 32);      /* rcd.cu:622 */
  ppg_redblue(out, out, width, height, filters, 1, 16);   /* rcd.cu:628 (in place) */

  /* populate, rcd.cu:30-46 (scale = 1) */
#pragma omp parallel for schedule(static)
  for (int row = 0; row < height; row++)
    for (int col = 0; col < width; col++) {
      const long idx = (long)row * w + col;
      const float val = fmaxf(0.0f, in[idx]);
      cfa[idx] = val;
      rgbp[FC(row, col) == 1 ? 1 : (FC(row, col) == 2 ? 2 : 0)][idx] = val;
    }
  /* step 1.1, rcd.cu:63-75 */
  float *v_diff = VP_diff, *h_diff = HQ_diff;
#pragma omp parallel for schedule(static)
  for (int row = 3; row <= height - 4; row++)
    for (int col = 3; col <= width - 4; col++) {
      const long i = (long)row * w + col;
      v_diff[i] = sq(cfa[i - w3] - 3.0f * cfa[i - w2] - cfa[i - w] + 6.0f * cfa[i] - cfa[i + w] - 3.0f * cfa[i + w2] + cfa[i + w3]);
      h_diff[i] = sq(cfa[i - 3] - 3.0f * cfa[i - 2] - cfa[i - 1] + 6.0f * cfa[i] - cfa[i + 1] - 3.0f * cfa[i + 2] + cfa[i + 3]);
    }
  /* step 1.2, rcd.cu:78-90 */
#pragma omp parallel for schedule(static)
  for (int row = 2; row <= height - 3; row++)
    for (int col = 2; col <= width - 3; col++) {
      const long i = (long)row * w + col;
      const float V = fmaxf(1e-10f, v_diff[i - w] + v_diff[i] + v_diff[i + w]);
      const float Hs = fmaxf(1e-10f, h_diff[i - 1] + h_diff[i] + h_diff[i + 1]);
      VH_dir[i] = V / (V + Hs);
    }
  /* step 2.1, rcd.cu:93-104: low-pass at red/blue sites, stored at idx/2 */
  float *lpf = lpf_PQ;
#pragma omp parallel for schedule(static)
  for (int row = 2; row <= height - 2; row++)
    for (int col = 2 + (FC(row, 0) & 1); col <= width - 2; col += 2) {
      const long i = (long)row * w + col;
      lpf[i / 2] = cfa[i] + 0.5f * (cfa[i - w] + cfa[i + w] + cfa[i - 1] + cfa[i + 1]) +
                   0.25f * (cfa[i - w - 1] + cfa[i - w + 1] + cfa[i + w - 1] + cfa[i + w + 1]);
    }
  /* step 3.1, rcd.cu:107-146: green at red/blue sites */
#pragma omp parallel for schedule(static)
  for (int row = 4; row <= height - 5; row++)
    for (int col = 4 + (FC(row, 0) & 1); col <= width - 5; col += 2) {
      const long i = (long)row * w + col, l = i / 2;
      const float eps = 1e-5f;
      const float c0 = VH_dir[i];
      const float nb = 0.25f * (VH_dir[i - w - 1] + VH_dir[i - w + 1] + VH_dir[i + w - 1] + VH_dir[i + w + 1]);
      const float disc = (fabsf(0.5f - c0) < fabsf(0.5f - nb)) ? nb : c0;
      const float ci = cfa[i];
      const float Ng = eps + fabsf(cfa[i - w] - cfa[i + w]) + fabsf(ci - cfa[i - w2]) + fabsf(cfa[i - w] - cfa[i - w3]) + fabsf(cfa[i - w2] - cfa[i - w4]);
      const float Sg = eps + fabsf(cfa[i + w] - cfa[i - w]) + fabsf(ci - cfa[i + w2]) + fabsf(cfa[i + w] - cfa[i + w3]) + fabsf(cfa[i + w2] - cfa[i + w4]);
      const float Wg = eps + fabsf(cfa[i - 1] - cfa[i + 1]) + fabsf(ci - cfa[i - 2]) + fabsf(cfa[i - 1] - cfa[i - 3]) + fabsf(cfa[i - 2] - cfa[i - 4]);
      const float Eg = eps + fabsf(cfa[i + 1] - cfa[i - 1]) + fabsf(ci - cfa[i + 2]) + fabsf(cfa[i + 1] - cfa[i + 3]) + fabsf(cfa[i + 2] - cfa[i + 4]);
      const float li = lpf[l];
      const float Ne = cfa[i - w] * (li + li) / (eps + li + lpf[l - w]);
      const float Se = cfa[i + w] * (li + li) / (eps + li + lpf[l + w]);
      const float We = cfa[i - 1] * (li + li) / (eps + li + lpf[l - 1]);
      const float Ee = cfa[i + 1] * (li + li) / (eps + li + lpf[l + 1]);
      const float Ve = (Sg * Ne + Ng * Se) / (Ng + Sg);
      const float He = (Wg * Ee + Eg * We) / (Eg + Wg);
      rgb1[i] = mixf(Ve, He, disc);
    }
  /* step 4.1, rcd.cu:149-163: P/Q diagonal high-pass on odd columns of every row, stored at idx/2 */
  float *p_diff = VP_diff, *q_diff = HQ_diff;
#pragma omp parallel for schedule(static)
  for (int row = 3; row <= height - 4; row++)
    for (int col = 3; col <= width - 4; col += 2) {
      const long i = (long)row * w + col;
      p_diff[i / 2] = sq((cfa[i - w3 - 3] - cfa[i - w - 1] - cfa[i + w + 1] + cfa[i + w3 + 3]) - 3.0f * (cfa[i - w2 - 2] + cfa[i + w2 + 2]) + 6.0f * cfa[i]);
      q_diff[i / 2] = sq((cfa[i - w3 + 3] - cfa[i - w + 1] - cfa[i + w - 1] + cfa[i + w3 - 3]) - 3.0f * (cfa[i - w2 + 2] + cfa[i + w2 - 2]) + 6.0f * cfa[i]);
    }
  /* step 4.2, rcd.cu:166-182.  Reads precede writes in the reference only by launch order; PQ_dir
   * overwrites lpf, which nothing reads any more. */
  float *PQ_dir = lpf_PQ;
#pragma omp parallel for schedule(static)
  for (int row = 2; row <= height - 3; row++)
    for (int col = 2 + (FC(row, 0) & 1); col <= width - 3; col += 2) {
      const long i = (long)row * w + col;
      const long i2 = i / 2, i3 = (i - w - 1) / 2, i4 = (i + w - 1) / 2;
      const float P = fmaxf(1e-10f, p_diff[i3] + p_diff[i2] + p_diff[i4 + 1]);
      const float Q = fmaxf(1e-10f, q_diff[i3 + 1] + q_diff[i2] + q_diff[i4]);
      PQ_dir[i2] = P / (P + Q);
    }
  /* step 5.1, rcd.cu:185-224: the opposite colour at red/blue sites along the diagonals */
#pragma omp parallel for schedule(static)
  for (int row = 4; row <= height - 4; row++)
    for (int col = 4 + (FC(row, 0) & 1); col <= width - 4; col += 2) {
      const int color = 2 - FC(row, col);
      float *rc = rgbp[color == 1 ? 1 : (color == 2 ? 2 : 0)];
      const long i = (long)row * w + col;
      const long q1 = i / 2, q2 = (i - w - 1) / 2, q3 = (i + w - 1) / 2;
      const float eps = 1e-5f;
      const float c0 = PQ_dir[q1];
      const float nb = 0.25f * (PQ_dir[q2] + PQ_dir[q2 + 1] + PQ_dir[q3] + PQ_dir[q3 + 1]);
      const float disc = (fabsf(0.5f - c0) < fabsf(0.5f - nb)) ? nb : c0;
      const float NWg = eps + fabsf(rc[i - w - 1] - rc[i + w + 1]) + fabsf(rc[i - w - 1] - rc[i - w3 - 3]) + fabsf(rgb1[i] - rgb1[i - w2 - 2]);
      const float NEg = eps + fabsf(rc[i - w + 1] - rc[i + w - 1]) + fabsf(rc[i - w + 1] - rc[i - w3 + 3]) + fabsf(rgb1[i] - rgb1[i - w2 + 2]);
      const float SWg = eps + fabsf(rc[i - w + 1] - rc[i + w - 1]) + fabsf(rc[i + w - 1] - rc[i + w3 - 3]) + fabsf(rgb1[i] - rgb1[i + w2 - 2]);
      const float SEg = eps + fabsf(rc[i - w - 1] - rc[i + w + 1]) + fabsf(rc[i + w + 1] - rc[i + w3 + 3]) + fabsf(rgb1[i] - rgb1[i + w2 + 2]);
      const float NWe = rc[i - w - 1] - rgb1[i - w - 1], NEe = rc[i - w + 1] - rgb1[i - w + 1];
      const float SWe = rc[i + w - 1] - rgb1[i + w - 1], SEe = rc[i + w + 1] - rgb1[i + w + 1];
      const float Pe = (NWg * SEe + SEg * NWe) / (NWg + SEg);
      const float Qe = (NEg * SWe + SWg * NEe) / (NEg + SWg);
      rc[i] = rgb1[i] + mixf(Pe, Qe, disc);
    }
  /* step 5.2, rcd.cu:227-282: red and blue at green sites */
#pragma omp parallel for schedule(static)
  for (int row = 4; row <= height - 4; row++)
    for (int col = 4 + (FC(row, 1) & 1); col <= width - 4; col += 2) {
      const long i = (long)row * w + col;
      const float eps = 1e-5f;
      const float c0 = VH_dir[i];
      const float nb = 0.25f * (VH_dir[i - w - 1] + VH_dir[i - w + 1] + VH_dir[i + w - 1] + VH_dir[i + w + 1]);
      const float disc = (fabsf(0.5f - c0) < fabsf(0.5f - nb)) ? nb : c0;
      const float g = rgb1[i];
      const float N1 = eps + fabsf(g - rgb1[i - w2]), S1 = eps + fabsf(g - rgb1[i + w2]);
      const float W1 = eps + fabsf(g - rgb1[i - 2]), E1 = eps + fabsf(g - rgb1[i + 2]);
      const float gN = rgb1[i - w], gS = rgb1[i + w], gW = rgb1[i - 1], gE = rgb1[i + 1];
      for (int c = 0; c <= 2; c += 2) {
        float *rc = rgbp[c];
        const float SN = fabsf(rc[i - w] - rc[i + w]), EW = fabsf(rc[i - 1] - rc[i + 1]);
        const float Ng = N1 + SN + fabsf(rc[i - w] - rc[i - w3]);
        const float Sg = S1 + SN + fabsf(rc[i + w] - rc[i + w3]);
        const float Wg = W1 + EW + fabsf(rc[i - 1] - rc[i - 3]);
        const float Eg = E1 + EW + fabsf(rc[i + 1] - rc[i + 3]);
        const float Ne = rc[i - w] - gN, Se = rc[i + w] - gS, We = rc[i - 1] - gW, Ee = rc[i + 1] - gE;
        const float Ve = (Ng * Se + Sg * Ne) / (Ng + Sg);
        const float He = (Eg * We + Wg * Ee) / (Eg + Wg);
        rc[i] = g + mixf(Ve, He, disc);
      }
    }
  /* write_output, rcd.cu:49-60: inside the 7-px margin only */
#pragma omp parallel for schedule(static)
  for (int row = 7; row < height - 7; row++)
    for (int col = 7; col < width - 7; col++) {
      const long i = (long)row * w + col;
      out[3 * i] = fmaxf(rgb0[i], 0.0f);
      out[3 * i + 1] = fmaxf(rgb1[i], 0.0f);
      out[3 * i + 2] = fmaxf(rgb2[i], 0.0f);
    }
}

/* ------------------------------------------------------------------------------------------------
 * post-process, postprocess.cu                                                                    */
static inline void cas(float *a, float *b) { /* reduction.h:85-91 */
  const float x = *a;
  const int c = *a > *b;
  *a = c ? *b : *a;
  *b = c ? x : *b;
}

static float median9(float s[9]) { /* reduction.h:93-116, same exchange sequence */
  cas(&s[1], &s[2]); cas(&s[4], &s[5]); cas(&s[7], &s[8]);
  cas(&s[0], &s[1]); cas(&s[3], &s[4]); cas(&s[6], &s[7]);
  cas(&s[1], &s[2]); cas(&s[4], &s[5]); cas(&s[7], &s[8]);
  cas(&s[0], &s[3]); cas(&s[5], &s[8]); cas(&s[4], &s[7]);
  cas(&s[3], &s[6]); cas(&s[1], &s[4]); cas(&s[2], &s[5]);
  cas(&s[4], &s[7]); cas(&s[4], &s[2]); cas(&s[6], &s[4]);
  cas(&s[4], &s[2]);
  return s[4];
}

static void color_smoothing(const float *in, float *out, int width, int height) { /* postprocess.cu:24-78 */
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      float dr[9], db[9];
      int k = 0;
      for (int j = -1; j <= 1; j++)
        for (int i = -1; i <= 1; i++, k++) {
          float p[3];
          px3(in, width, height, x + i, y + j, 0, p);
          dr[k] = p[0] - p[1];
          db[k] = p[2] - p[1];
        }
      const float *o = in + 3 * ((long)y * width + x);
      float *d = out + 3 * ((long)y * width + x);
      d[0] = fmaxf(median9(dr) + o[1], 0.0f);
      d[1] = fmaxf(o[1], 0.0f);
      d[2] = fmaxf(median9(db) + o[1], 0.0f);
    }
}

void orc_postprocess(const float *in, float *out, int width, int height, uint32_t filters, int smoothing_passes,
                     int green_eq_local, int green_eq_global, float green_eq_threshold) {
  const long n = (long)width * height;
  float *a = malloc(3 * n * sizeof(float)), *b = malloc(3 * n * sizeof(float));
  memcpy(a, in, 3 * n * sizeof(float));
  for (int p = 0; p < smoothing_passes; p++) {
    color_smoothing(a, b, width, height);
    float *t = a; a = b; b = t;
  }
  if (green_eq_global) { /* postprocess.cu:175-255, :352-377 */
    double sum1 = 0, sum2 = 0;
    const int we = 2 * (width / 2), he = 2 * (height / 2);
    for (int y = 0; y < he; y++)
      for (int x = 0; x < we; x++)
        if (FC(y, x) == 1) {
          if (y & 1) sum2 += a[3 * ((long)y * width + x) + 1];
          else sum1 += a[3 * ((long)y * width + x) + 1];
        }
    const float s1 = (float)sum1, s2 = (float)sum2;
    const float ratio = (s1 > 0.0f && s2 > 0.0f) ? s2 / s1 : 1.0f;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; y++)
      for (int x = 0; x < width; x++) {
        const float *s = a + 3 * ((long)y * width + x);
        float *d = b + 3 * ((long)y * width + x);
        const int g1 = FC(y, x) == 1 && !(y & 1);
        d[0] = fmaxf(s[0], 0.0f);
        d[1] = fmaxf(s[1] * (g1 ? ratio : 1.0f), 0.0f);
        d[2] = fmaxf(s[2], 0.0f);
      }
    float *t = a; a = b; b = t;
  }
  if (green_eq_local) { /* postprocess.cu:84-169, threshold/100 at :383 */
    const float thr = (float)(green_eq_threshold / 100.);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; y++)
      for (int x = 0; x < width; x++) {
        const float *s = a + 3 * ((long)y * width + x);
        float *d = b + 3 * ((long)y * width + x);
        float o = s[1];
        if (FC(y, x) == 1 && (y & 1)) {
#define G(dx, dy) ((x + (dx) < 0 || y + (dy) < 0 || x + (dx) >= width || y + (dy) >= height) ? 0.0f : a[3 * ((long)(y + (dy)) * width + x + (dx)) + 1])
          const float o1_1 = G(-1, -1), o1_2 = G(1, -1), o1_3 = G(-1, 1), o1_4 = G(1, 1);
          const float o2_1 = G(0, -2), o2_2 = G(0, 2), o2_3 = G(-2, 0), o2_4 = G(2, 0);
#undef G
          const float m1 = (o1_1 + o1_2 + o1_3 + o1_4) / 4.0f, m2 = (o2_1 + o2_2 + o2_3 + o2_4) / 4.0f;
          if (m2 > 0.0f && m1 > 0.0f && m1 / m2 < 2.0f) {
            const float c1 = (fabsf(o1_1 - o1_2) + fabsf(o1_1 - o1_3) + fabsf(o1_1 - o1_4) + fabsf(o1_2 - o1_3) + fabsf(o1_3 - o1_4) + fabsf(o1_2 - o1_4)) / 6.0f;
            const float c2 = (fabsf(o2_1 - o2_2) + fabsf(o2_1 - o2_3) + fabsf(o2_1 - o2_4) + fabsf(o2_2 - o2_3) + fabsf(o2_3 - o2_4) + fabsf(o2_2 - o2_4)) / 6.0f;
            if (o < 0.95f && c1 < thr && c2 < thr) o *= m1 / m2;
          }
        }
        d[0] = s[0];
        d[1] = fmaxf(o, 0.0f);
        d[2] = s[2];
      }
    float *t = a; a = b; b = t;
  }
  memcpy(out, a, 3 * n * sizeof(float));
  free(a);
  free(b);
}
