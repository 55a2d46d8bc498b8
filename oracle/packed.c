/* CPU oracle (test infrastructure, see oracle.h): 12-bit packed Bayer codec and white balance.
 * Follows torch_darktable/csrc/packed.cu:8-31 (byte layouts), :34-155 (kernels) and
 * torch_darktable/csrc/white_balance.cu:10-42. */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* IEEE binary32 -> binary16 bits, round-to-nearest-even (what at::Half / __float2half do). */
uint16_t orc_float_to_half_bits(float x) {
  uint32_t f;
  memcpy(&f, &x, 4);
  uint32_t sign = (f >> 16) & 0x8000u;
  uint32_t mag = f & 0x7fffffffu;
  if (mag >= 0x7f800000u) return (uint16_t)(sign | (mag > 0x7f800000u ? 0x7e00u : 0x7c00u));
  if (mag >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u); /* rounds to inf */
  if (mag < 0x33000001u) return (uint16_t)sign;             /* below half of the smallest subnormal */
  int32_t exp = (int32_t)(mag >> 23) - 127;
  uint32_t man = (mag & 0x7fffffu) | 0x800000u;
  uint32_t shift, half;
  if (exp < -14) {
    shift = (uint32_t)(13 + (-14 - exp)); /* subnormal target */
    half = 0;
  } else {
    shift = 13;
    half = (uint32_t)(exp + 15) << 10;
  }
  uint32_t q = man >> shift;
  uint32_t rem = man & ((1u << shift) - 1u);
  uint32_t mid = 1u << (shift - 1);
  if (rem > mid || (rem == mid && (q & 1u))) q++;
  if (exp < -14) return (uint16_t)(sign | q); /* may carry into the normal range, which is correct */
  return (uint16_t)(sign | (half + (q - 0x400u)));
}

static float half_bits_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t f;
  if (exp == 0) {
    if (man == 0) {
      f = sign;
    } else {
      int e = -1;
      do {
        man <<= 1;
        e++;
      } while (!(man & 0x400u));
      f = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) {
    f = sign | 0x7f800000u | (man << 13);
  } else {
    f = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float out;
  memcpy(&out, &f, 4);
  return out;
}

float orc_half_round(float x) { return half_bits_to_float(orc_float_to_half_bits(x)); }

static inline void unpack_pair(const uint8_t *b, int ids, uint16_t *p0, uint16_t *p1) {
  if (ids) { /* packed.cu:28-31 */
    *p0 = (uint16_t)(((uint16_t)b[0] << 4) | (b[2] & 0xf));
    *p1 = (uint16_t)(((uint16_t)b[1] << 4) | (b[2] >> 4));
  } else { /* packed.cu:15-18 */
    *p0 = (uint16_t)((((uint16_t)b[1] & 0xf) << 8) | b[0]);
    *p1 = (uint16_t)(((uint16_t)b[2] << 4) | (b[1] >> 4));
  }
}

static inline void pack_pair(uint16_t p0, uint16_t p1, int ids, uint8_t *b) {
  if (ids) { /* packed.cu:21-25 -- note p0's low nibble lands in the HIGH nibble of b[2] (reference quirk) */
    b[0] = (uint8_t)(p0 >> 4);
    b[1] = (uint8_t)(p1 >> 4);
    b[2] = (uint8_t)(((p0 & 0xf) << 4) | (p1 & 0xf));
  } else { /* packed.cu:8-12 */
    b[0] = (uint8_t)(p0 & 0xff);
    b[1] = (uint8_t)(((p1 & 0xf) << 4) | (p0 >> 8));
    b[2] = (uint8_t)(p1 >> 4);
  }
}

void orc_decode12_f32(const uint8_t *in, float *out, long npairs, int ids, int scaled) {
  const float scale = scaled ? (1.0f / 4095.0f) : 1.0f; /* packed.cu:216 */
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npairs; i++) {
    uint16_t p0, p1;
    unpack_pair(in + 3 * i, ids, &p0, &p1);
    out[2 * i] = (float)p0 * scale;
    out[2 * i + 1] = (float)p1 * scale;
  }
}

void orc_decode12_f16(const uint8_t *in, uint16_t *out_bits, long npairs, int ids, int scaled) {
  const float scale = scaled ? (1.0f / 4095.0f) : 1.0f;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npairs; i++) {
    uint16_t p0, p1;
    unpack_pair(in + 3 * i, ids, &p0, &p1);
    out_bits[2 * i] = orc_float_to_half_bits((float)p0 * scale);
    out_bits[2 * i + 1] = orc_float_to_half_bits((float)p1 * scale);
  }
}

void orc_decode12_u16(const uint8_t *in, uint16_t *out, long npairs, int ids) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npairs; i++) unpack_pair(in + 3 * i, ids, &out[2 * i], &out[2 * i + 1]);
}

void orc_encode12_u16(const uint16_t *in, uint8_t *out, long npairs, int ids) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npairs; i++) {
    uint16_t p0 = in[2 * i] > 4095 ? 4095 : in[2 * i]; /* packed.cu:46-47 */
    uint16_t p1 = in[2 * i + 1] > 4095 ? 4095 : in[2 * i + 1];
    pack_pair(p0, p1, ids, out + 3 * i);
  }
}

/* uint16_t(roundf(f)) on the GPU saturates: negatives and NaN -> 0, large -> 65535; then min(.,4095). */
static inline uint16_t quantise12(float f) {
  float r = roundf(f);
  if (!(r > 0.0f)) return 0;
  if (r >= 4095.0f) return 4095;
  return (uint16_t)r;
}

void orc_encode12_f32(const float *in, uint8_t *out, long npairs, int ids, int scaled) {
  const float scale = scaled ? 4095.0f : 1.0f; /* packed.cu:190 */
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npairs; i++) pack_pair(quantise12(in[2 * i] * scale), quantise12(in[2 * i + 1] * scale), ids, out + 3 * i);
}

int orc_fc(int row, int col, uint32_t filters) { return (int)((filters >> ((((row << 1) & 14) + (col & 1)) << 1)) & 3u); }

void orc_white_balance(const float *in, float *out, int width, int height, uint32_t filters, const float gains[3]) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      const int c = orc_fc(y, x, filters);
      const float g = c == 0 ? gains[0] : (c == 2 ? gains[2] : gains[1]);
      out[(long)y * width + x] = fminf(fmaxf(in[(long)y * width + x] * g, 0.0f), 1.0f);
    }
}

static int cmp_float(const void *a, const void *b) {
  float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}

void orc_estimate_white_balance(const float *const *images, int n_images, int width, int height, uint32_t filters,
                                float quantile, int stride, float gains[3]) {
  const int sw = width / stride, sh = height / stride;
  long cap = (long)n_images * sw * sh, n = 0;
  float *cr = malloc(sizeof(float) * cap), *cg = malloc(sizeof(float) * cap), *inten = malloc(sizeof(float) * cap);
  for (int k = 0; k < n_images; k++) {
    const float *im = images[k];
    /* white_balance.cu:69-71: threads with pos+1 >= size/stride return; patch origin is pos*2 (not pos*stride) */
    for (int py = 0; py + 1 < sh; py++)
      for (int px = 0; px + 1 < sw; px++) {
        const int x = px * 2, y = py * 2;
        const float p00 = im[(long)y * width + x], p01 = im[(long)y * width + x + 1];
        const float p10 = im[(long)(y + 1) * width + x], p11 = im[(long)(y + 1) * width + x + 1];
        float r, g, b;
        switch (filters) { /* bayer_device.h:36-44 */
          case 0x94949494u: r = p00, g = (p01 + p10) * 0.5f, b = p11; break;
          case 0x16161616u: r = p11, g = (p01 + p10) * 0.5f, b = p00; break;
          case 0x61616161u: r = p01, g = (p00 + p11) * 0.5f, b = p10; break;
          default: r = p10, g = (p00 + p11) * 0.5f, b = p01; break;
        }
        const float mx = fmaxf(fmaxf(p00, p01), fmaxf(p10, p11));
        if (!(mx < 1.0f)) continue;
        const float s = r + g + b;
        cr[n] = r / s, cg[n] = g / s, inten[n] = s, n++;
      }
  }
  gains[0] = gains[1] = gains[2] = 1.0f;
  if (n > 0) {
    float *sorted = malloc(sizeof(float) * n);
    memcpy(sorted, inten, sizeof(float) * n);
    qsort(sorted, n, sizeof(float), cmp_float);
    /* torch.quantile, linear interpolation */
    const double pos = (double)quantile * (double)(n - 1);
    const long lo = (long)floor(pos), hi = lo + 1 < n ? lo + 1 : lo;
    const float thr = (float)(sorted[lo] + (sorted[hi] - sorted[lo]) * (pos - (double)lo));
    double sr = 0, sg = 0;
    long m = 0;
    for (long i = 0; i < n; i++)
      if (inten[i] >= thr) sr += cr[i], sg += cg[i], m++;
    if (m > 0) {
      const float mr = (float)(sr / m), mg = (float)(sg / m);
      gains[0] = mr / mg, gains[1] = 1.0f, gains[2] = (1.0f - mr - mg) / mg;
    }
    free(sorted);
  }
  free(cr), free(cg), free(inten);
}
