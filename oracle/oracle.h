/*
 * CPU oracle for the torch-darktable RAW->sRGB hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's CUDA kernels
 * (uc-vision/torch-darktable v0.2.3, paths below are relative to torch_darktable/csrc/).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
 * the product path (torch-darktable_b200/) never does.
 *
 * Parity pin: the reference ships no golden vectors (its only test is a JSON round trip), so the
 * oracle is pinned against outputs of the reference extension itself, generated on a B200 by
 * tests/golden/make_golden.py and committed under tests/golden/.  Tolerances are stated in
 * tests/test_oracle_golden.py (the reference is compiled with --use_fast_math, this file is not).
 *
 * Layout: row-major, channels-last, float32 unless stated.  All functions are reentrant.
 */
#ifndef TDB_ORACLE_H
#define TDB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- packed 12-bit codec: packed.cu:8-31 (pair layouts), :34-155 (kernels) ---- */
void orc_decode12_f32(const uint8_t *in, float *out, long npairs, int ids, int scaled);
void orc_decode12_f16(const uint8_t *in, uint16_t *out_bits, long npairs, int ids, int scaled);
void orc_decode12_u16(const uint8_t *in, uint16_t *out, long npairs, int ids);
void orc_encode12_u16(const uint16_t *in, uint8_t *out, long npairs, int ids);
void orc_encode12_f32(const float *in, uint8_t *out, long npairs, int ids, int scaled);

/* ---- CFA helpers: debayer/bayer_device.h:9-11 ---- */
int orc_fc(int row, int col, uint32_t filters);

/* ---- white balance: white_balance.cu:10-42 ---- */
void orc_white_balance(const float *in, float *out, int width, int height, uint32_t filters, const float gains[3]);
/* white_balance.cu:57-161 (estimate); returns gains[3]. Uninitialised last row/col samples of the
 * reference are excluded here (they are garbage there). */
void orc_estimate_white_balance(const float *const *images, int n_images, int width, int height, uint32_t filters,
                                float quantile, int stride, float gains[3]);

/* ---- demosaic ---- */
/* debayer/bilinear.cu:17-99 */
void orc_bilinear5x5(const float *cfa, float *rgb, int width, int height, uint32_t filters);
/* debayer/ppg.cu:21-389, host order :441-461.  median_threshold as passed to the Python ctor (percent). */
void orc_ppg(const float *cfa, float *rgb, int width, int height, uint32_t filters, float median_threshold);
/* debayer/rcd.cu:30-282 (steps), :285-493 (borders), :601-671 (launch order and buffer reuse).
 * scratch = 8*width*height floats that persist between calls of one workspace; zero it for a fresh one. */
void orc_rcd(const float *cfa, float *rgb, int width, int height, uint32_t filters, float *scratch);
/* debayer/postprocess.cu:24-255, host order :311-390 */
void orc_postprocess(const float *in, float *out, int width, int height, uint32_t filters, int smoothing_passes,
                     int green_eq_local, int green_eq_global, float green_eq_threshold);

/* ---- colour ops: device_conversions.h, color_conversions.cu ---- */
enum {
  ORC_RGB_TO_XYZ = 0,
  ORC_XYZ_TO_LAB = 1,
  ORC_LAB_TO_XYZ = 2,
  ORC_XYZ_TO_RGB = 3,
  ORC_RGB_TO_LAB = 4,
  ORC_LAB_TO_RGB = 5,
  ORC_MODIFY_HSL = 6,      /* p[0..2] = hue, sat, lum adjust */
  ORC_MODIFY_VIBRANCE = 7, /* p[0] = amount */
  ORC_MATRIX_3X3 = 8       /* p[0..8] = row-major matrix, result clipped to [0,1] */
};
void orc_color_convert(const float *in, float *out, long npixels, int op, const float *p);
void orc_compute_luminance(const float *rgb, float *lum, long npixels);
void orc_compute_log_luminance(const float *rgb, float *lum, long npixels, float eps);
void orc_modify_luminance(const float *rgb, const float *lum, float *out, long npixels);
void orc_modify_log_luminance(const float *rgb, const float *loglum, float *out, long npixels, float eps);

/* ---- image statistics + tonemap: tonemap/color_adaption.cu, color_adaption.h, reinhard.cu, aces.cu, linear.cu ---- */
void orc_bounds_accumulate(const float *rgb, int width, int height, int stride, float bounds[2]);
/* sums[6] = log_gray, gray, r, g, b, valid_count (double accumulators); bounds = {b0, b1} */
void orc_metrics_accumulate(const float *rgb, int width, int height, int stride, float min_gray, const float bounds[2],
                            double sums[6]);
enum { ORC_TM_REINHARD = 0, ORC_TM_ACES = 1, ORC_TM_ADAPTIVE_ACES = 2, ORC_TM_LINEAR = 3 };
void orc_tonemap(const float *rgb, uint8_t *out, long npixels, int op, const float metrics[5], float gamma,
                 float intensity, float light_adapt, float vibrance);

/* ---- Wiener: denoise/denoise.cu:85-242,267-331, window.h:18-43, fft.h ---- */
void orc_wiener(const float *in, float *out, int width, int height, int channels, int tile, int overlap,
                const float *sigmas);

/* ---- bilateral grid: local_contrast/bilateral.cu ---- */
void orc_bilateral_grid_size(int width, int height, float sigma_s, float sigma_r, int size[3]);
void orc_bilateral(const float *lum, float *out, int width, int height, float sigma_s, float sigma_r, float detail);

/* ---- local Laplacian: local_contrast/laplacian.cu ---- */
void orc_laplacian(const float *lum, float *out, int width, int height, float sigma, float shadows, float highlights,
                   float clarity);

/* fp16 round trip used by the Laplacian (at::Half storage) */
float orc_half_round(float x);
uint16_t orc_float_to_half_bits(float x);

#ifdef __cplusplus
}
#endif
#endif
