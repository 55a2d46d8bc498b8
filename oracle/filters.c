/* CPU oracle (test infrastructure, see oracle.h): Wiener tile denoiser, bilateral grid, local Laplacian.
 * Follows torch_darktable/csrc/denoise/{denoise.cu,fft.h,window.h} and
 * torch_darktable/csrc/local_contrast/{bilateral,laplacian}.cu. */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* =====================================  Wiener  ===================================== */

/* in-place radix-2 DIT FFT of length n (16 or 32); sign = -1 forward, +1 inverse; no scaling (fft.h:133-166) */
static void fft1d(float *re, float *im, int n, int stride, int sign) {
  int bits = 0;
  while ((1 << bits) < n) bits++;
  for (int i = 0; i < n; i++) {
    int r = 0;
    for (int b = 0; b < bits; b++) r |= ((i >> b) & 1) << (bits - 1 - b);
    if (r > i) {
      float t = re[i * stride]; re[i * stride] = re[r * stride]; re[r * stride] = t;
      t = im[i * stride]; im[i * stride] = im[r * stride]; im[r * stride] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1;
    for (int k = 0; k < half; k++) {
      const double ang = sign * 2.0 * M_PI * k / len;
      const float wr = (float)cos(ang), wi = (float)sin(ang);
      for (int s = 0; s < n; s += len) {
        const int a = (s + k) * stride, b = (s + k + half) * stride;
        const float tr = re[b] * wr - im[b] * wi, ti = re[b] * wi + im[b] * wr;
        re[b] = re[a] - tr, im[b] = im[a] - ti;
        re[a] += tr, im[a] += ti;
      }
    }
  }
}

static inline int reflect_index(int x, int limit) { /* denoise.cu:118-122 (single reflection) */
  if (x < 0) x = -x;
  if (x >= limit) x = 2 * limit - x - 1;
  /* the reference reads out of bounds past this point (sides shorter than 2K-1); keep the oracle defined */
  if (x < 0) x = 0;
  if (x >= limit) x = limit - 1;
  return x;
}

void orc_wiener(const float *in, float *out, int width, int height, int channels, int K, int overlap,
                const float *sigmas) {
  const int stride = K / overlap;
  const int hp = height + 2 * K, wp = width + 2 * K;
  const int grid_start = -(K / stride);
  const int grid_h = (height + K + stride - 1) / stride - grid_start; /* denoise.cu:283-285 */
  const int grid_w = (width + K + stride - 1) / stride - grid_start;
  const float eps = 1e-15f;

  /* window.h:18-43: both windows exp(-r^2 / (0.3 (K/2)^2)), L2-normalised */
  float win[32];
  {
    const float half = K / 2.0f, scale = 0.3f * half * half;
    double norm = 0;
    for (int i = 0; i < K; i++) {
      const float r = -half + 0.5f + (float)i * ((half - 0.5f) - (-half + 0.5f)) / (float)(K - 1);
      win[i] = expf(-(r * r) / scale);
      norm += (double)win[i] * win[i];
    }
    const float inv = (float)(1.0 / sqrt(norm));
    for (int i = 0; i < K; i++) win[i] *= inv;
  }

  float *acc = calloc((size_t)hp * wp * channels, sizeof(float));
  float *mask = calloc((size_t)hp * wp, sizeof(float));

  /* tiles are independent up to the accumulation; rows of tiles that are K apart never touch the same
   * accumulator cells, so process tile rows in `overlap`-spaced phases for a race-free parallel loop */
  for (int phase = 0; phase < overlap; phase++) {
#pragma omp parallel for schedule(dynamic)
    for (int gy = phase; gy < grid_h; gy += overlap) {
      float *re = malloc(sizeof(float) * K * K), *im = malloc(sizeof(float) * K * K);
      float *val = malloc(sizeof(float) * K * K * channels);
      for (int gx = 0; gx < grid_w; gx++) {
        const int ox = (gx + grid_start) * stride, oy = (gy + grid_start) * stride; /* denoise.cu:146 */
        float mean[3] = {0, 0, 0};
        for (int ty = 0; ty < K; ty++)
          for (int tx = 0; tx < K; tx++) {
            const int sx = reflect_index(ox + tx, width), sy = reflect_index(oy + ty, height);
            for (int c = 0; c < channels; c++) {
              const float v = in[((long)sy * width + sx) * channels + c];
              val[(ty * K + tx) * channels + c] = v;
              mean[c] += v;
            }
          }
        for (int c = 0; c < channels; c++) mean[c] /= (float)(K * K); /* block_mean, denoise.cu:85-101 */
        for (int c = 0; c < channels; c++) {
          for (int ty = 0; ty < K; ty++)
            for (int tx = 0; tx < K; tx++) {
              re[ty * K + tx] = (val[(ty * K + tx) * channels + c] - mean[c]) * (win[tx] * win[ty]);
              im[ty * K + tx] = 0.0f;
            }
          for (int r = 0; r < K; r++) fft1d(re + r * K, im + r * K, K, 1, -1);
          for (int q = 0; q < K; q++) fft1d(re + q, im + q, K, K, -1);
          const float s2 = sigmas[c] * sigmas[c];
          for (int i = 0; i < K * K; i++) { /* apply_gain, denoise.cu:181-185 */
            const float power = re[i] * re[i] + im[i] * im[i] + eps;
            const float gain = fmaxf(power - s2, 0.0f) / power;
            re[i] *= gain, im[i] *= gain;
          }
          for (int r = 0; r < K; r++) fft1d(re + r * K, im + r * K, K, 1, +1);
          for (int q = 0; q < K; q++) fft1d(re + q, im + q, K, K, +1);
          const float inv = (1.0f / K) * (1.0f / K); /* fft.h:209-230 */
          for (int i = 0; i < K * K; i++) val[i * channels + c] = re[i] * inv;
        }
        for (int ty = 0; ty < K; ty++)
          for (int tx = 0; tx < K; tx++) { /* store_pixel, denoise.cu:150-178 */
            const int px = ox + tx + K, py = oy + ty + K;
            if (py < 0 || px < 0 || py >= hp || px >= wp) continue;
            const float wf = win[tx] * win[ty], wi = win[tx] * win[ty];
            const long o = (long)py * wp + px;
            for (int c = 0; c < channels; c++) acc[o * channels + c] += (val[(ty * K + tx) * channels + c] + mean[c] * wf) * wi;
            mask[o] += wf * wi;
          }
      }
      free(re), free(im), free(val);
    }
  }
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) { /* normalize_and_crop, denoise.cu:224-242 */
      const long p = (long)(y + K) * wp + (x + K);
      for (int c = 0; c < channels; c++) out[((long)y * width + x) * channels + c] = acc[p * channels + c] / (mask[p] + eps);
    }
  free(acc);
  free(mask);
}

/* =====================================  bilateral grid  ===================================== */

static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

void orc_bilateral_grid_size(int width, int height, float sigma_s, float sigma_r, int size[3]) { /* bilateral.cu:273-299 */
  float ss = sigma_s;
  if (ss < 0.5f) ss = 0.5f;
  const float gx = clampf(roundf(width / ss), 4.0f, 3000.0f);
  const float gy = clampf(roundf(height / ss), 4.0f, 3000.0f);
  const float gz = clampf(roundf(1.0f / sigma_r), 4.0f, 50.0f);
  const float eff_s = fmaxf(height / gy, width / gx), eff_r = 1.0f / gz;
  size[0] = (int)ceilf(width / eff_s) + 1;
  size[1] = (int)ceilf(height / eff_s) + 1;
  size[2] = (int)ceilf(1.0f / eff_r) + 1;
}

typedef struct {
  long gi;
  float fx, fy, fz;
} grid_sample;

static inline grid_sample make_sample(int x, int y, float L, const int size[3], float sigma_s, float sigma_r) { /* :71-86 */
  const float gx = clampf(x / sigma_s, 0.0f, (float)(size[0] - 1));
  const float gy = clampf(y / sigma_s, 0.0f, (float)(size[1] - 1));
  const float gz = clampf(L / sigma_r, 0.0f, (float)(size[2] - 1));
  int ix = (int)gx, iy = (int)gy, iz = (int)gz;
  if (ix > size[0] - 2) ix = size[0] - 2;
  if (iy > size[1] - 2) iy = size[1] - 2;
  if (iz > size[2] - 2) iz = size[2] - 2;
  grid_sample s = {ix + (long)size[0] * (iy + (long)size[1] * iz), gx - (float)ix, gy - (float)iy, gz - (float)iz};
  return s;
}

/* zero-padded 5-tap filter along one axis; taps[] = coefficients for offsets -2..+2 (:132-203) */
static void blur_axis(const float *in, float *out, const int size[3], int axis, const float taps[5]) {
  const long step[3] = {1, size[0], (long)size[0] * size[1]};
  const long total = (long)size[0] * size[1] * size[2];
  const int n = size[axis];
#pragma omp parallel for schedule(static)
  for (long i = 0; i < total; i++) {
    const int pos = (int)((i / step[axis]) % n);
    float acc = 0.0f;
    for (int d = -2; d <= 2; d++)
      if (pos + d >= 0 && pos + d < n) acc += taps[d + 2] * in[i + d * step[axis]];
    out[i] = acc;
  }
}

void orc_bilateral(const float *lum, float *out, int width, int height, float sigma_s, float sigma_r, float detail) {
  int size[3];
  orc_bilateral_grid_size(width, height, sigma_s, sigma_r, size);
  const long total = (long)size[0] * size[1] * size[2];
  const long ox = 1, oy = size[0], oz = (long)size[0] * size[1];
  float *grid = calloc(total, sizeof(float)), *tmp = calloc(total, sizeof(float));
  const float contrib = 1.0f / (sigma_s * sigma_s);
  for (int y = 0; y < height; y++) /* splat, :99-112 (atomics in the reference; order-dependent in the last bits) */
    for (int x = 0; x < width; x++) {
      const grid_sample s = make_sample(x, y, lum[(long)y * width + x], size, sigma_s, sigma_r);
      const float ax = 1.0f - s.fx, ay = 1.0f - s.fy, az = 1.0f - s.fz;
      float *g = grid + s.gi;
      g[0] += ax * ay * az * contrib;
      g[ox] += s.fx * ay * az * contrib;
      g[oy] += ax * s.fy * az * contrib;
      g[oy + ox] += s.fx * s.fy * az * contrib;
      g[oz] += ax * ay * s.fz * contrib;
      g[oz + ox] += s.fx * ay * s.fz * contrib;
      g[oz + oy] += ax * s.fy * s.fz * contrib;
      g[oz + oy + ox] += s.fx * s.fy * s.fz * contrib;
    }
  const float gauss[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  const float deriv[5] = {-2.0f / 16.0f, -4.0f / 16.0f, 0.0f, 4.0f / 16.0f, 2.0f / 16.0f}; /* :171-203 */
  blur_axis(grid, tmp, size, 0, gauss); /* :319-326 */
  blur_axis(tmp, grid, size, 1, gauss);
  blur_axis(grid, tmp, size, 2, deriv); /* :335-340 */
  const float norm = -detail * sigma_r * 4.0f;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; y++) /* slice, :208-228 */
    for (int x = 0; x < width; x++) {
      const float L = lum[(long)y * width + x];
      const grid_sample s = make_sample(x, y, L, size, sigma_s, sigma_r);
      const float ax = 1.0f - s.fx, ay = 1.0f - s.fy, az = 1.0f - s.fz;
      const float *g = tmp + s.gi;
      const float d = g[0] * ax * ay * az + g[ox] * s.fx * ay * az + g[oy] * ax * s.fy * az + g[oy + ox] * s.fx * s.fy * az +
                      g[oz] * ax * ay * s.fz + g[oz + ox] * s.fx * ay * s.fz + g[oz + oy] * ax * s.fy * s.fz +
                      g[oz + oy + ox] * s.fx * s.fy * s.fz;
      out[(long)y * width + x] = fmaxf(0.0f, L + norm * d);
    }
  free(grid);
  free(tmp);
}

/* =====================================  local Laplacian  ===================================== */

#define LAP_GAMMAS 6
#define H16(x) orc_half_round(x) /* every buffer of the reference is at::Half */

static inline int dl(int x, int level) { return (x + (1 << level) - 1) >> level; } /* laplacian.cu:50 */

static void gauss_reduce(const float *fine, float *coarse, int cw, int ch, int fw) { /* :178-208 */
  static const float w[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
#pragma omp parallel for schedule(static)
  for (int y = 0; y < ch; y++)
    for (int x = 0; x < cw; x++) {
      int cx = x, cy = y;
      if (x >= cw - 1) cx = cw - 2;
      if (y >= ch - 1) cy = ch - 2;
      if (cx <= 0) cx = 1;
      if (cy <= 0) cy = 1;
      float v = 0.0f;
      for (int j = -2; j <= 2; j++)
        for (int i = -2; i <= 2; i++) v += fine[(long)(2 * cy + j) * fw + (2 * cx + i)] * w[i + 2] * w[j + 2];
      coarse[(long)y * cw + x] = H16(v);
    }
}

static inline float expand_gaussian(const float *coarse, int px, int py, int cw) { /* :111-140 */
  static const float w[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  const int cx = px / 2, cy = py / 2, xo = px & 1, yo = py & 1;
  float c = 0.0f;
  for (int i = (xo ? 0 : -1); i <= 1; i++)
    for (int j = (yo ? 0 : -1); j <= 1; j++) {
      const float p = coarse[(long)(cy + j) * cw + (cx + i)];
      const int wi = xo ? (2 * i + 1) : (2 * i + 2), wj = yo ? (2 * j + 1) : (2 * j + 2);
      c += p * w[wi] * w[wj];
    }
  return 4.0f * c;
}

static inline float curve(float x, float g, float sigma, float shadows, float highlights, float clarity) { /* :266-290 */
  const float c = x - g;
  float val;
  const float ssigma = c > 0.0f ? sigma : -sigma;
  const float shadhi = c > 0.0f ? shadows : highlights;
  if (fabsf(c) > 2 * sigma)
    val = g + ssigma + shadhi * (c - ssigma);
  else {
    const float t = fminf(fmaxf(c / (2.0f * ssigma), 0.0f), 1.0f);
    const float t2 = t * t, mt = 1.0f - t;
    val = g + ssigma * 2.0f * mt * t + t2 * (ssigma + ssigma * shadhi);
  }
  val += clarity * c * expf(-c * c / (2.0f * sigma * sigma / 3.0f));
  return val;
}

void orc_laplacian(const float *lum, float *out, int width, int height, float sigma, float shadows, float highlights,
                   float clarity) {
  int num_levels = 0;
  {
    const int m = width < height ? width : height; /* :415: min(30, floor(log2(min(w,h)))) */
    while ((1 << (num_levels + 1)) <= m) num_levels++;
    if (num_levels > 30) num_levels = 30;
  }
  const int max_supp = 1 << (num_levels - 1);
  const int bw = width + 2 * max_supp, bh = height + 2 * max_supp;
  float **padded = calloc(num_levels, sizeof(float *)), **output = calloc(num_levels, sizeof(float *));
  float **proc[LAP_GAMMAS];
  for (int k = 0; k < LAP_GAMMAS; k++) proc[k] = calloc(num_levels, sizeof(float *));
  for (int l = 0; l < num_levels; l++) {
    const size_t n = (size_t)dl(bw, l) * dl(bh, l);
    padded[l] = malloc(n * sizeof(float));
    output[l] = malloc(n * sizeof(float));
    for (int k = 0; k < LAP_GAMMAS; k++) proc[k][l] = malloc(n * sizeof(float));
  }
  /* pad_input_half, :90-109 */
#pragma omp parallel for schedule(static)
  for (int y = 0; y < bh; y++)
    for (int x = 0; x < bw; x++) {
      int cx = x - max_supp, cy = y - max_supp;
      if (cx >= width) cx = width - 1;
      if (cy >= height) cy = height - 1;
      if (cx < 0) cx = 0;
      if (cy < 0) cy = 0;
      padded[0][(long)y * bw + x] = H16(lum[(long)cy * width + cx]);
    }
  /* gaussian pyramid of the input; the coarsest level lands in the output pyramid, :515-528 */
  for (int l = 1; l < num_levels; l++)
    gauss_reduce(padded[l - 1], (l == num_levels - 1) ? output[l] : padded[l], dl(bw, l), dl(bh, l), dl(bw, l - 1));
  /* six tone-curved copies and their pyramids, :530-553 */
  for (int k = 0; k < LAP_GAMMAS; k++) {
    const float g = (k + 0.5f) / (float)LAP_GAMMAS;
    const long n0 = (long)bw * bh;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n0; i++) proc[k][0][i] = H16(curve(padded[0][i], g, sigma, shadows, highlights, clarity));
    for (int l = 1; l < num_levels; l++) gauss_reduce(proc[k][l - 1], proc[k][l], dl(bw, l), dl(bh, l), dl(bw, l - 1));
  }
  /* assemble coarse -> fine, :222-263, :555-582 */
  for (int l = num_levels - 2; l >= 0; l--) {
    const int pw = dl(bw, l), ph = dl(bh, l), cw = (pw - 1) / 2 + 1;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ph; y++)
      for (int x = 0; x < pw; x++) {
        int qx = x, qy = y; /* clamp_boundary, :53-65 */
        if (pw & 1) { if (qx > pw - 2) qx = pw - 2; } else { if (qx > pw - 3) qx = pw - 3; }
        if (ph & 1) { if (qy > ph - 2) qy = ph - 2; } else { if (qy > ph - 3) qy = ph - 3; }
        if (qx <= 0) qx = 1;
        if (qy <= 0) qy = 1;
        float v_out = expand_gaussian(output[l + 1], qx, qy, cw);
        const float v = padded[l][(long)y * pw + x];
        int hi = 1;
        for (; hi < LAP_GAMMAS - 1 && ((float)hi + .5f) / (float)LAP_GAMMAS <= v; hi++);
        const int lo = hi - 1;
        const float a = fminf(fmaxf(v * LAP_GAMMAS - ((float)lo + .5f), 0.0f), 1.0f);
        const float l0 = proc[lo][l][(long)y * pw + x] - expand_gaussian(proc[lo][l + 1], qx, qy, cw);
        const float l1 = proc[lo + 1][l][(long)y * pw + x] - expand_gaussian(proc[lo + 1][l + 1], qx, qy, cw);
        v_out += l0 * (1.0f - a) + l1 * a;
        output[l][(long)y * pw + x] = H16(v_out);
      }
  }
  /* write_back_half, :372-384 */
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) out[(long)y * width + x] = output[0][(long)(y + max_supp) * bw + (x + max_supp)];
  for (int l = 0; l < num_levels; l++) {
    free(padded[l]), free(output[l]);
    for (int k = 0; k < LAP_GAMMAS; k++) free(proc[k][l]);
  }
  free(padded), free(output);
  for (int k = 0; k < LAP_GAMMAS; k++) free(proc[k]);
}
