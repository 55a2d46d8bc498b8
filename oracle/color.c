/* CPU oracle (test infrastructure, see oracle.h): colour conversions, image statistics and the tone mappers.
 * Two different Lab definitions coexist in the reference and both are restated here:
 *   - public colour ops:  torch_darktable/csrc/device_conversions.h  (powf(t,1/3), 0.008856/7.787, normalised L)
 *   - tonemap vibrance:   torch_darktable/csrc/device_color_conversions.h (cbrtf, (6/29)^3, L/100)            */
#include "oracle.h"

#include <math.h>

static inline float clip01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

/* ---------------- device_conversions.h ---------------- */
static inline float srgb_to_linear(float x) { /* :14-25 */
  return (x > 0.04045f) ? powf((x + 0.055f) / (1.0f + 0.055f), 2.4f) : x * (1.0f / 12.92f);
}
static inline float linear_to_srgb(float x) { /* :27-38 */
  return (x > 0.0031308f) ? (1.0f + 0.055f) * powf(x, 1.0f / 2.4f) - 0.055f : x * 12.92f;
}
static inline float lab_f(float t) { return (t > 0.008856f) ? powf(t, 1.0f / 3.0f) : (t * 7.787f + 16.0f / 116.0f); }
static inline float lab_f_inv(float t) {
  const float c = t * t * t;
  return (c > 0.008856f) ? c : (t - 16.0f / 116.0f) / 7.787f;
}

static void rgb_to_xyz(const float rgb[3], float xyz[3]) { /* :73-83 */
  const float r = srgb_to_linear(rgb[0]), g = srgb_to_linear(rgb[1]), b = srgb_to_linear(rgb[2]);
  xyz[0] = 0.4124564f * r + 0.3575761f * g + 0.1804375f * b;
  xyz[1] = 0.2126729f * r + 0.7151522f * g + 0.0721750f * b;
  xyz[2] = 0.0193339f * r + 0.1191920f * g + 0.9503041f * b;
}
static void xyz_to_lab(const float xyz[3], float lab[3]) { /* :85-97 */
  const float fx = lab_f(xyz[0] / 0.95047f), fy = lab_f(xyz[1] / 1.0f), fz = lab_f(xyz[2] / 1.08883f);
  lab[0] = (116.0f / 100.0f) * fy - (16.0f / 100.0f);
  lab[1] = (500.0f / 128.0f) * (fx - fy);
  lab[2] = (200.0f / 128.0f) * (fy - fz);
}
static void lab_to_xyz(const float lab[3], float xyz[3]) { /* :99-112 */
  const float fy = lab[0] * (100.0f / 116.0f) + (16.0f / 116.0f);
  const float fx = lab[1] * (128.0f / 500.0f) + fy, fz = fy - lab[2] * (128.0f / 200.0f);
  xyz[0] = lab_f_inv(fx) * 0.95047f;
  xyz[1] = lab_f_inv(fy) * 1.0f;
  xyz[2] = lab_f_inv(fz) * 1.08883f;
}
static void xyz_to_rgb(const float xyz[3], float rgb[3]) { /* :114-123 */
  rgb[0] = linear_to_srgb(3.2404542f * xyz[0] + -1.5371385f * xyz[1] + -0.4985314f * xyz[2]);
  rgb[1] = linear_to_srgb(-0.9692660f * xyz[0] + 1.8760108f * xyz[1] + 0.0415560f * xyz[2]);
  rgb[2] = linear_to_srgb(0.0556434f * xyz[0] + -0.2040259f * xyz[1] + 1.0572252f * xyz[2]);
}
static void rgb_to_lab(const float rgb[3], float lab[3]) {
  float xyz[3];
  rgb_to_xyz(rgb, xyz);
  xyz_to_lab(xyz, lab);
}
static void lab_to_rgb(const float lab[3], float rgb[3]) {
  float xyz[3];
  lab_to_xyz(lab, xyz);
  xyz_to_rgb(xyz, rgb);
}

static void rgb_to_hsl(const float rgb[3], float hsl[3]) { /* :144-168 */
  const float mx = fmaxf(fmaxf(rgb[0], rgb[1]), rgb[2]), mn = fminf(fminf(rgb[0], rgb[1]), rgb[2]);
  const float delta = mx - mn;
  float h = 0.0f, s = 0.0f;
  const float l = (mx + mn) * 0.5f;
  if (delta > 1e-6f) {
    s = (l < 0.5f) ? delta / (mx + mn) : delta / (2.0f - mx - mn);
    if (mx == rgb[0]) h = (rgb[1] - rgb[2]) / delta + (rgb[1] < rgb[2] ? 6.0f : 0.0f);
    else if (mx == rgb[1]) h = (rgb[2] - rgb[0]) / delta + 2.0f;
    else h = (rgb[0] - rgb[1]) / delta + 4.0f;
    h /= 6.0f;
  }
  hsl[0] = h, hsl[1] = s, hsl[2] = l;
}
static float hue_to_rgb(float p, float q, float t) { /* :170-177 */
  if (t < 0.0f) t += 1.0f;
  if (t > 1.0f) t -= 1.0f;
  if (t < 1.0f / 6.0f) return p + (q - p) * 6.0f * t;
  if (t < 1.0f / 2.0f) return q;
  if (t < 2.0f / 3.0f) return p + (q - p) * (2.0f / 3.0f - t) * 6.0f;
  return p;
}
static void hsl_to_rgb(const float hsl[3], float rgb[3]) { /* :179-195 */
  const float h = hsl[0], s = hsl[1], l = hsl[2];
  if (s < 1e-6f) {
    rgb[0] = rgb[1] = rgb[2] = l;
    return;
  }
  const float q = (l < 0.5f) ? l * (1.0f + s) : l + s - l * s, p = 2.0f * l - q;
  rgb[0] = hue_to_rgb(p, q, h + 1.0f / 3.0f);
  rgb[1] = hue_to_rgb(p, q, h);
  rgb[2] = hue_to_rgb(p, q, h - 1.0f / 3.0f);
}

static float lab_l_of(const float rgb_in[3]) { /* rgb_to_lab_l :197-207 on clip(rgb), color_conversions.cu:168-172 */
  const float r = srgb_to_linear(clip01(rgb_in[0])), g = srgb_to_linear(clip01(rgb_in[1])), b = srgb_to_linear(clip01(rgb_in[2]));
  const float y = 0.2126729f * r + 0.7151522f * g + 0.0721750f * b;
  return fmaxf(0.0f, (116.0f / 100.0f) * lab_f(y) - (16.0f / 100.0f));
}

void orc_color_convert(const float *in, float *out, long npixels, int op, const float *p) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npixels; i++) {
    const float *a = in + 3 * i;
    float *o = out + 3 * i;
    float t[3], u[3];
    switch (op) {
      case ORC_RGB_TO_XYZ: rgb_to_xyz(a, o); break;
      case ORC_XYZ_TO_LAB: xyz_to_lab(a, o); break;
      case ORC_LAB_TO_XYZ: lab_to_xyz(a, o); break;
      case ORC_XYZ_TO_RGB: xyz_to_rgb(a, o); break;
      case ORC_RGB_TO_LAB: rgb_to_lab(a, o); break;
      case ORC_LAB_TO_RGB: lab_to_rgb(a, o); break;
      case ORC_MODIFY_HSL: { /* device_conversions.h:227-239 */
        rgb_to_hsl(a, t);
        float h = t[0] + p[0];
        if (h < 0.0f) h += 1.0f;
        if (h > 1.0f) h -= 1.0f;
        u[0] = h;
        u[1] = powf(t[1], 1.0f / (1.0f + p[1]));
        u[2] = powf(t[2], 1.0f / (1.0f + p[2]));
        hsl_to_rgb(u, t);
        for (int k = 0; k < 3; k++) o[k] = clip01(t[k]);
        break;
      }
      case ORC_MODIFY_VIBRANCE: { /* device_conversions.h:242-261 */
        rgb_to_lab(a, t);
        const float chroma = sqrtf(t[1] * t[1] + t[2] * t[2]);
        const float ls = 1.0f - p[0] * chroma * 0.25f, ss = 1.0f + p[0] * chroma;
        u[0] = t[0] * ls, u[1] = t[1] * ss, u[2] = t[2] * ss;
        lab_to_rgb(u, t);
        for (int k = 0; k < 3; k++) o[k] = clip01(t[k]);
        break;
      }
      case ORC_MATRIX_3X3: /* device_conversions.h:209-211, device_math.h:107-113 */
        for (int k = 0; k < 3; k++) o[k] = clip01(p[3 * k] * a[0] + p[3 * k + 1] * a[1] + p[3 * k + 2] * a[2]);
        break;
    }
  }
}

void orc_compute_luminance(const float *rgb, float *lum, long npixels) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npixels; i++) lum[i] = lab_l_of(rgb + 3 * i);
}

void orc_compute_log_luminance(const float *rgb, float *lum, long npixels, float eps) { /* color_conversions.cu:174-183 */
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npixels; i++) lum[i] = logf(fmaxf(eps, lab_l_of(rgb + 3 * i)));
}

static void replace_l(const float *rgb, float l, float *out) { /* device_conversions.h:213-218 */
  float lab[3], t[3];
  rgb_to_lab(rgb, lab);
  lab[0] = fmaxf(0.0f, fminf(1.0f, l));
  lab_to_rgb(lab, t);
  for (int k = 0; k < 3; k++) out[k] = clip01(t[k]);
}

void orc_modify_luminance(const float *rgb, const float *lum, float *out, long npixels) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npixels; i++) replace_l(rgb + 3 * i, lum[i], out + 3 * i);
}

void orc_modify_log_luminance(const float *rgb, const float *loglum, float *out, long npixels, float eps) {
  (void)eps; /* the device code ignores eps on the way back, device_conversions.h:220-225 */
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npixels; i++) replace_l(rgb + 3 * i, expf(loglum[i]), out + 3 * i);
}

/* ---------------- statistics, tonemap/color_adaption.cu ---------------- */
void orc_bounds_accumulate(const float *rgb, int width, int height, int stride, float bounds[2]) { /* :12-36 */
  float lo = bounds[0], hi = bounds[1];
  for (int y = 0; y < height; y += stride)
    for (int x = 0; x < width; x += stride) {
      const float *p = rgb + 3 * ((long)y * width + x);
      lo = fminf(lo, fminf(fminf(p[0], p[1]), p[2]));
      hi = fmaxf(hi, fmaxf(fmaxf(p[0], p[1]), p[2]));
    }
  bounds[0] = lo, bounds[1] = hi;
}

void orc_metrics_accumulate(const float *rgb, int width, int height, int stride, float min_gray, const float bounds[2],
                            double sums[6]) { /* :39-84 */
  const float range = bounds[1] - bounds[0] + 1e-6f;
  for (int y = 0; y < height; y += stride)
    for (int x = 0; x < width; x += stride) {
      const float *p = rgb + 3 * ((long)y * width + x);
      const float r = (p[0] - bounds[0]) / range, g = (p[1] - bounds[0]) / range, b = (p[2] - bounds[0]) / range;
      const float mask = (r >= 0.99f || g >= 0.99f || b >= 0.99f) ? 0.0f : 1.0f;
      const float gray = r * 0.299f + g * 0.587f + b * 0.114f; /* device_math.h:460-462 */
      const float log_gray = logf(fmaxf(gray, min_gray));
      sums[0] += log_gray * mask, sums[1] += gray * mask, sums[2] += r * mask, sums[3] += g * mask, sums[4] += b * mask;
      sums[5] += mask;
    }
}

/* ---------------- device_color_conversions.h (tonemap flavour of Lab) ---------------- */
static inline float tm_srgb_to_linear(float x) { return x <= 0.04045f ? x / 12.92f : powf((x + 0.055f) / 1.055f, 2.4f); }
static inline float tm_linear_to_srgb(float x) { return x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f; }
static inline float tm_lab_f(float t) {
  const float d = 6.0f / 29.0f;
  return (t > d * d * d) ? cbrtf(t) : (1.0f / (3.0f * d * d)) * t + 4.0f / 29.0f;
}
static inline float tm_lab_f_inv(float t) {
  const float d = 6.0f / 29.0f;
  return (t > d) ? t * t * t : (3.0f * d * d) * (t - 4.0f / 29.0f);
}

static void tm_vibrance(const float rgb[3], float amount, float out[3]) { /* :199-213 with :21-113 */
  const float r = tm_srgb_to_linear(rgb[0]), g = tm_srgb_to_linear(rgb[1]), b = tm_srgb_to_linear(rgb[2]);
  const float X = 0.4124564f * r + 0.3575761f * g + 0.1804375f * b;
  const float Y = 0.2126729f * r + 0.7151522f * g + 0.0721750f * b;
  const float Z = 0.0193339f * r + 0.1191920f * g + 0.9503041f * b;
  const float fx = tm_lab_f(X / 0.95047f), fy = tm_lab_f(Y / 1.0f), fz = tm_lab_f(Z / 1.08883f);
  const float L0 = (116.0f * fy - 16.0f) / 100.0f, a0 = (500.0f * (fx - fy)) / 128.0f, b0 = (200.0f * (fy - fz)) / 128.0f;
  const float chroma = sqrtf(a0 * a0 + b0 * b0);
  const float ls = 1.0f - amount * chroma * 0.25f, ss = 1.0f + amount * chroma;
  const float L = L0 * ls * 100.0f, A = a0 * ss * 128.0f, B = b0 * ss * 128.0f;
  const float gy = (L + 16.0f) / 116.0f, gx = A / 500.0f + gy, gz = gy - B / 200.0f;
  const float x2 = tm_lab_f_inv(gx) * 0.95047f, y2 = tm_lab_f_inv(gy) * 1.0f, z2 = tm_lab_f_inv(gz) * 1.08883f;
  out[0] = clip01(tm_linear_to_srgb(3.2404542f * x2 + -1.5371385f * y2 + -0.4985314f * z2));
  out[1] = clip01(tm_linear_to_srgb(-0.9692660f * x2 + 1.8760108f * y2 + 0.0415560f * z2));
  out[2] = clip01(tm_linear_to_srgb(0.0556434f * x2 + -0.2040259f * y2 + 1.0572252f * z2));
}

static void aces_fit(const float in[3], float out[3]) { /* aces.cu:13-34 */
  static const float mi[9] = {0.59719f, 0.35458f, 0.04823f, 0.07600f, 0.90834f, 0.01566f, 0.02840f, 0.13383f, 0.83777f};
  static const float mo[9] = {1.60475f, -0.53108f, -0.07367f, -0.10208f, 1.10813f, -0.00605f, -0.00327f, -0.07276f, 1.07602f};
  float v[3], c[3];
  for (int k = 0; k < 3; k++) v[k] = mi[3 * k] * in[0] + mi[3 * k + 1] * in[1] + mi[3 * k + 2] * in[2];
  for (int k = 0; k < 3; k++) {
    const float a = v[k] * (v[k] + 0.0245786f) - 0.000090537f;
    const float b = v[k] * (0.983729f * v[k] + 0.4329510f) + 0.238081f;
    c[k] = a / b;
  }
  for (int k = 0; k < 3; k++) out[k] = mo[3 * k] * c[0] + mo[3 * k + 1] * c[1] + mo[3 * k + 2] * c[2];
}

static inline uint8_t to_u8(float x) { /* device_math.h:347-349; the cast saturates on the GPU */
  const float r = fminf(roundf(x * 255.0f), 255.0f);
  return (uint8_t)(r > 0.0f ? r : 0.0f);
}

void orc_tonemap(const float *rgb, uint8_t *out, long npixels, int op, const float metrics[5], float gamma,
                 float intensity, float light_adapt, float vibrance) {
  float map_key = 0.0f, exposure = 1.0f;
  if (op != ORC_TM_ACES) { /* color_adaption.h:17-44 */
    const float normalized = fmaxf(0.0f, fminf(1.0f, (-metrics[0]) / 9.21034f));
    map_key = 0.3f + 0.7f * powf(normalized, 1.4f);
    exposure = expf(intensity);
  }
  const float inv_gamma = 1.0f / gamma;
  const float aces_gain = powf(2.0f, intensity);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < npixels; i++) {
    const float *p = rgb + 3 * i;
    float t[3], adapt[3], g[3], v[3];
    if (op != ORC_TM_ACES)
      for (int k = 0; k < 3; k++) { /* color_adaption.h:46-76; lerp(t,a,b) = a + t*(b-a), device_math.h:353-355 */
        const float mean = metrics[2 + k] + light_adapt * (p[k] - metrics[2 + k]);
        adapt[k] = powf(mean / exposure, map_key);
      }
    switch (op) {
      case ORC_TM_REINHARD: /* reinhard.cu:34-36 */
        for (int k = 0; k < 3; k++) t[k] = p[k] / (adapt[k] + p[k]);
        break;
      case ORC_TM_LINEAR: /* linear.cu:31-32 */
        for (int k = 0; k < 3; k++) t[k] = p[k] / adapt[k];
        break;
      case ORC_TM_ADAPTIVE_ACES: /* aces.cu:56-58 */
        for (int k = 0; k < 3; k++) v[k] = p[k] / adapt[k];
        aces_fit(v, t);
        break;
      default: /* aces.cu:83 */
        for (int k = 0; k < 3; k++) v[k] = p[k] * aces_gain;
        aces_fit(v, t);
        break;
    }
    for (int k = 0; k < 3; k++) g[k] = powf(fmaxf(t[k], 0.0f), inv_gamma);
    tm_vibrance(g, vibrance, v);
    for (int k = 0; k < 3; k++) out[3 * i + k] = to_u8(v[k]);
  }
}
