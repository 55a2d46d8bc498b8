/*
 * libtdb200 -- C ABI of the B200-native (sm_100a) torch-darktable hot path.
 *
 * This header is the drop-in boundary.  It replaces the reference's pybind11/torch module
 * `torch_darktable.torch_darktable_extension` (torch_darktable/csrc/extension.cpp:50-248) with plain
 * `extern "C"` entry points: raw DEVICE pointers, sizes, scalars and a CUDA stream.  No torch types, no hidden
 * allocation, no device synchronisation, no process-global device state.  The Python shim
 * torch-darktable_b200/torch_darktable/extension.py presents the reference's attribute surface on top.
 *
 * Conventions
 *   - every function returns 0 on success or a TDB_E* code; tdb_last_error() gives a thread-local message.
 *   - images are row-major, channels-last float32: CFA (H,W), RGB (H,W,3); byte buffers uint8.
 *   - `filters` is the darktable CFA word (reference csrc/debayer/demosaic.h:7-12).
 *   - `stream` is a cudaStream_t passed as void*; NULL = legacy default stream.
 *   - pointers documented "device" must be device-accessible; scalars are passed by value (the reference
 *     reads gains / sigmas / reduction results back with blocking .item() calls; here they stay on the device).
 *   - workspaces ("scratch") are caller-provided; tdb_*_scratch_bytes() reports their size.
 */
#ifndef TDB200_H
#define TDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDB_OK 0
#define TDB_EINVAL 1   /* bad argument (maps to RuntimeError/ValueError in the shim) */
#define TDB_ECUDA 2    /* CUDA runtime error (launch failure etc.) */
#define TDB_EUNSUPPORTED 3
#define TDB_EJPEG 4    /* nvJPEG status != success (maps to JpegException in the shim) */

#define TDB_FILTERS_RGGB 0x94949494u
#define TDB_FILTERS_BGGR 0x16161616u
#define TDB_FILTERS_GRBG 0x61616161u
#define TDB_FILTERS_GBRG 0x49494949u

typedef void *tdb_stream_t;

int tdb_version(void);
const char *tdb_last_error(void);
/* number of kernel launches issued through this library by the calling process (for bench.py's gpu_launches) */
uint64_t tdb_launch_count(void);
/* Hint for the CALLING THREAD: the launches that follow belong to `lanes` frames that are in flight on different streams at the same
 * time (0 or 1: one stream at a time, the default).  Kernels whose register allocation would otherwise fill an SM on their own then
 * launch a variant that leaves room for another kernel's CTA (the Wiener tile kernel: 104 instead of 128 registers -- 5 % slower alone,
 * 1.3 % faster frames when two are in flight).  Results do not depend on it. */
void tdb_set_concurrency_hint(int lanes);
/* Optional per-kernel timing (the counterpart of the reference's CudaTimer, csrc/cuda_utils.h:40-85).  Between
 * tdb_timing_begin and tdb_timing_end every launch of the calling thread is followed by a cudaEventRecord on its
 * stream; tdb_timing_end synchronises and writes "kernel_name,launches,total_ms" lines into buf (returns the number of
 * bytes needed).  A kernel's time is the gap to the previous event on the same stream, so it is exact for launches that
 * queue back to back on one stream.  Off by default; costs nothing when off. */
void tdb_timing_begin(tdb_stream_t stream);
size_t tdb_timing_end(char *buf, size_t buf_bytes);

/* ---------------------------------------------------------------------------------------------------------
 * 12-bit packed codec.  Replaces decode12_float/half/u16, encode12_u16/float
 * (extension.cpp:159-169, csrc/packed.cu:158-280).  npairs = pixels / 2 = bytes / 3.                       */
int tdb_decode12_f32(const uint8_t *packed, float *out, int64_t npairs, int ids_format, int scaled, tdb_stream_t stream);
int tdb_decode12_f16(const uint8_t *packed, uint16_t *out_half_bits, int64_t npairs, int ids_format, int scaled, tdb_stream_t stream);
int tdb_decode12_u16(const uint8_t *packed, uint16_t *out, int64_t npairs, int ids_format, tdb_stream_t stream);
int tdb_encode12_u16(const uint16_t *values, uint8_t *packed, int64_t npairs, int ids_format, tdb_stream_t stream);
int tdb_encode12_f32(const float *values, uint8_t *packed, int64_t npairs, int ids_format, int scaled, tdb_stream_t stream);

/* Fused ingest (north_star: unpack + black level + white balance in one pass; the reference runs
 * decode12_float, then clone + apply_white_balance, pipeline/image_processor.py:190-240).
 *   v = decode(p) / 4095;  v = v - black;  if gains: v = clamp(v * gains[fc(y,x)], 0, 1)
 * black = 0 and gains = NULL reproduce decode12_float bit for bit.  gains: device float[3] or NULL.         */
int tdb_unpack12_wb(const uint8_t *packed, float *cfa, int width, int height, int ids_format, uint32_t filters,
                    float black, const float *gains, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * White balance.  Replaces apply_white_balance (extension.cpp:209, csrc/white_balance.cu:164-183).
 * gains: device float[3].  in == out is allowed.                                                           */
int tdb_white_balance(const float *in, float *out, int width, int height, uint32_t filters, const float *gains, tdb_stream_t stream);
/* Replaces estimate_white_balance (extension.cpp:211, csrc/white_balance.cu:94-162): phase 1 collects the
 * bright-patch samples of one image; phase 2 (csrc/white_balance.cu:135-161: masked gathers, torch::quantile, mean) is one
 * single-CTA kernel over the sample arrays of all images: exact selection of the two order statistics torch.quantile
 * interpolates between, then the mean chromaticity of the samples at or above the threshold.
 * chroma: (n_samples,2), intensity: (n_samples), valid: (n_samples) uint8; n_samples = (W/stride)*(H/stride) per image.
 * gains: device float[3] = (R/G, 1, B/G), or (1, 1, 1) when no sample is valid.                                        */
int tdb_wb_collect_samples(const float *cfa, int width, int height, uint32_t filters, int stride, float *chroma,
                           float *intensity, uint8_t *valid, tdb_stream_t stream);
int tdb_wb_estimate_gains(const float *chroma, const float *intensity, const uint8_t *valid, int64_t n_samples, float quantile,
                          float *gains, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Demosaic.  cfa (H,W) -> rgb (H,W,3).  width and height must be even and >= 16.
 * tdb_bilinear5x5 replaces bilinear5x5_demosaic (extension.cpp:205, csrc/debayer/bilinear.cu:104-148).
 * tdb_ppg replaces PPG.process (extension.cpp:57-65, csrc/debayer/ppg.cu:413-463); median_threshold in percent.
 * tdb_rcd replaces RCD.process (extension.cpp:67-74, csrc/debayer/rcd.cu:601-671) with the semantics of a
 *   FRESH reference workspace (the reference's dependence on the previous frame is not reproduced).
 * The *_packed variants read the 12-bit packed frame directly and apply tdb_unpack12_wb's arithmetic on the fly. */
int tdb_bilinear5x5(const float *cfa, float *rgb, int width, int height, uint32_t filters, tdb_stream_t stream);
int tdb_ppg(const float *cfa, float *rgb, int width, int height, uint32_t filters, float median_threshold, tdb_stream_t stream);
int tdb_rcd(const float *cfa, float *rgb, int width, int height, uint32_t filters, tdb_stream_t stream);

#define TDB_DEMOSAIC_BILINEAR 0
#define TDB_DEMOSAIC_PPG 1
#define TDB_DEMOSAIC_RCD 2
int tdb_demosaic_packed(const uint8_t *packed, float *rgb, int width, int height, int ids_format, uint32_t filters,
                        int method, float black, const float *gains, float ppg_median_threshold, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Demosaic post-process.  Replaces PostProcess.process (extension.cpp:77-90, csrc/debayer/postprocess.cu:311-390):
 * `passes` 3x3 median colour smoothing, global green equilibration (ratio computed on the device, no read-back),
 * local green equilibration (threshold in percent).
 * scratch: tdb_postprocess_scratch_bytes(width,height) bytes of device memory.                              */
size_t tdb_postprocess_scratch_bytes(int width, int height);
int tdb_postprocess(const float *in, float *out, void *scratch, int width, int height, uint32_t filters, int passes,
                    int green_eq_local, int green_eq_global, float green_eq_threshold, tdb_stream_t stream);
/* The two halves of the global green equilibration (postprocess.cu:355-384) for a frame whose rows are split across GPUs
 * (SURVEY.md 8e): sums[0..1] = green sums of the G1 / G2 sites of the given rows (the caller all-reduces them over the ranks
 * and forms ratio = sums[1] / sums[0]), then the equilibration with that ratio (device float).  scratch as above.          */
int tdb_green_sums(const float *rgb, int width, int height, uint32_t filters, void *scratch, float *sums, tdb_stream_t stream);
int tdb_green_eq_apply(const float *in, float *out, int width, int height, uint32_t filters, int green_eq_local,
                       float green_eq_threshold, const float *ratio, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Colour ops (extension.cpp:127-156, csrc/color_conversions.cu, csrc/device_conversions.h).  npixels = H*W.  */
#define TDB_RGB_TO_XYZ 0
#define TDB_XYZ_TO_LAB 1
#define TDB_LAB_TO_XYZ 2
#define TDB_XYZ_TO_RGB 3
#define TDB_RGB_TO_LAB 4
#define TDB_LAB_TO_RGB 5
#define TDB_MODIFY_HSL 6       /* p0..p2 = hue, sat, lum adjust */
#define TDB_MODIFY_VIBRANCE 7  /* p0 = amount */
int tdb_color_convert(const float *in, float *out, int64_t npixels, int op, float p0, float p1, float p2, tdb_stream_t stream);
/* color_transform_3x3: out = clip(M * rgb); matrix: DEVICE float[9], row-major (the reference dereferences this
 * device pointer on the host, csrc/color_conversions.cu:158-159, which faults; here it is read on the device). */
int tdb_color_transform_3x3(const float *in, float *out, int64_t npixels, const float *matrix, tdb_stream_t stream);
int tdb_compute_luminance(const float *rgb, float *lum, int64_t npixels, tdb_stream_t stream);
int tdb_compute_log_luminance(const float *rgb, float *loglum, int64_t npixels, float eps, tdb_stream_t stream);
int tdb_modify_luminance(const float *rgb, const float *lum, float *out, int64_t npixels, tdb_stream_t stream);
int tdb_modify_log_luminance(const float *rgb, const float *loglum, float *out, int64_t npixels, float eps, tdb_stream_t stream);
/* pipeline/util.py:8-10 normalize_image (a torch.compile'd helper in the reference): (rgb - b0) / (b1 - b0),
 * bounds: device float[2].  nvalues = H*W*3. */
int tdb_normalize(const float *in, float *out, int64_t nvalues, const float *bounds, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Image statistics (extension.cpp:182-185, csrc/tonemap/color_adaption.cu:90-166).  Accumulating calls so a
 * list of images maps to one call per image on the same device accumulators, then one finalize.
 *   bounds: device float[2], initialise with tdb_bounds_init (FLT_MAX, -FLT_MAX).
 *   sums: device float[6] (log_gray, gray, r, g, b, valid count), zero-initialised by tdb_metrics_init.
 *   tdb_metrics_finalize: metrics[5] = sums[0..4] / max(sums[5], 1).                                         */
int tdb_bounds_init(float *bounds, tdb_stream_t stream);
int tdb_bounds_accumulate(const float *rgb, int width, int height, int stride, float *bounds, tdb_stream_t stream);
int tdb_metrics_init(float *sums, tdb_stream_t stream);
int tdb_metrics_accumulate(const float *rgb, int width, int height, int stride, float min_gray, const float *bounds /* device float[2] or NULL = (0,1) */,
                           float *sums, tdb_stream_t stream);
int tdb_metrics_finalize(const float *sums, float *metrics, tdb_stream_t stream);
/* a + (b - a) * t on n device floats (pipeline/util.py:4 lerp used for the EMA of bounds / metrics) */
int tdb_lerp(const float *a, const float *b, float t, float *out, int n, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Tone mapping -> uint8 (extension.cpp:188-195; csrc/tonemap/{reinhard,aces,linear}.cu, color_adaption.h).
 * metrics: device float[5] (ignored for TDB_TM_ACES).  matrix: optional device float[9] colour matrix applied to
 * the linear input first (north_star's fused 3x3; NULL = identity, which is what the reference pipeline does).
 * transform: one of TDB_TF_* (pipeline/transform.py:39-56) fused into the store; out has the transformed shape. */
#define TDB_TM_REINHARD 0
#define TDB_TM_ACES 1
#define TDB_TM_ADAPTIVE_ACES 2
#define TDB_TM_LINEAR 3
#define TDB_TF_NONE 0
#define TDB_TF_ROTATE_90 1
#define TDB_TF_ROTATE_180 2
#define TDB_TF_ROTATE_270 3
#define TDB_TF_TRANSPOSE 4
#define TDB_TF_FLIP_HORIZ 5
#define TDB_TF_FLIP_VERT 6
#define TDB_TF_TRANSVERSE 7
int tdb_tonemap(const float *rgb, uint8_t *out, int width, int height, int op, const float *metrics, float gamma,
                float intensity, float light_adapt, float vibrance, const float *matrix, int transform, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Wiener tile denoiser.  Replaces Wiener.process (extension.cpp:215-223, csrc/denoise/denoise.cu:267-331).
 * in/out (H,W,C), C in {1,3}; tile in {16,32}; overlap in {2,4,8}; sigmas: device float[C].
 * scratch: tdb_wiener_scratch_bytes() bytes.
 * tdb_wiener_log_luminance = Wiener.process_log_luminance (denoise.py:54-58) fused: log-luminance is computed
 * while tiles are loaded and the Lab write-back happens in the normalisation pass; `noise` by value.        */
size_t tdb_wiener_scratch_bytes(int width, int height, int channels, int tile);
int tdb_wiener(const float *in, float *out, void *scratch, int width, int height, int channels, int tile, int overlap,
               const float *sigmas, tdb_stream_t stream);
int tdb_wiener_log_luminance(const float *rgb, float *out, void *scratch, int width, int height, int tile, int overlap,
                             float noise, float eps, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Bilateral-grid local contrast.  Replaces Bilateral.process (extension.cpp:111-121,
 * csrc/local_contrast/bilateral.cu:358-385).  grid_size as the reference computes it (:273-299).
 * scratch: 2 * gx*gy*gz floats (tdb_bilateral_scratch_bytes).
 * tdb_bilateral_rgb = Bilateral.process_rgb (local_contrast.py:110-114) with compute_luminance fused into the
 * splat and modify_luminance fused into the slice.                                                          */
int tdb_bilateral_grid_size(int width, int height, float sigma_s, float sigma_r, int size[3]);
size_t tdb_bilateral_scratch_bytes(int width, int height, float sigma_s, float sigma_r);
int tdb_bilateral(const float *lum, float *out, void *scratch, int width, int height, float sigma_s, float sigma_r,
                  float detail, tdb_stream_t stream);
int tdb_bilateral_rgb(const float *rgb, float *out, void *scratch, int width, int height, float sigma_s, float sigma_r,
                      float detail, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Local Laplacian.  Replaces Laplacian.process (extension.cpp:94-108, csrc/local_contrast/laplacian.cu:446-480),
 * num_gamma = 6, fp16 storage like the reference.  scratch: tdb_laplacian_scratch_bytes().                  */
size_t tdb_laplacian_scratch_bytes(int width, int height);
int tdb_laplacian(const float *lum, float *out, void *scratch, int width, int height, float sigma, float shadows,
                  float highlights, float clarity, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused frame pipeline: what ImageProcessor.process_image_set (pipeline/image_processor.py:284-300) runs when the
 * post-process, the Wiener denoiser and the bilateral local contrast are enabled.  Same arithmetic per pixel as the stage
 * calls above, rearranged so that every intermediate image crosses HBM once and no statistic needs a launch of its own:
 *
 *   A  tdb_postprocess_deferred     smoothing; leaves the green ratio of the frame and the bounds of the image set behind
 *   B  tdb_frame_prepare            green equilibration + normalisation + log-luminance + accumulator clear, one pass
 *   C  tdb_wiener_log_luminance_fused  Wiener tiles; the normalisation pass also splats the bilateral grid; grid blur
 *   D  tdb_metrics_sliced           tone-mapping metrics of the bilateral output, evaluated at the sampled pixels only
 *   E  tdb_bilateral_slice_tonemap  slice + modify_luminance + tone map + gamma + vibrance + uint8 + transform, one pass
 *
 * frame_state: tdb_frame_state_bytes() of device memory per ImageProcessor, zero-filled ONCE by the caller (the last CTA
 * of the kernels that gather statistics reduces their per-CTA partials and resets its own ticket).  An image set is a run of
 * frames bracketed by first_in_set / last_in_set; bounds_out / metrics_out are written on the last frame as
 * prev + (value - prev) * moving_average (prev == NULL: the value itself), i.e. pipeline/util.py:4 lerp.                 */
size_t tdb_frame_state_bytes(void);
/* scratch: tdb_postprocess_scratch_bytes().  out: smoothed image, green equilibration NOT applied.  ratio_out: device
 * float[1], sum(G2)/sum(G1) of this frame (postprocess.cu:355-366).  bounds: over every bounds_stride-th pixel of the
 * EQUILIBRATED image (compute_image_bounds semantics; x -> max(0, x * ratio) is monotone, so extrema commute with it).  */
int tdb_postprocess_deferred(const float *in, float *out, void *scratch, int width, int height, uint32_t filters, int passes,
                             int bounds_stride, void *frame_state, int first_in_set, int last_in_set, const float *prev_bounds,
                             float moving_average, float *bounds_out, float *ratio_out, tdb_stream_t stream);
/* c = normalize(green_eq_global(rgb, ratio), bounds); ratio == NULL skips the equilibration.
 * wiener_scratch == NULL: out (H,W,3) = c.
 * wiener_scratch != NULL: out (H,W,2) = the Lab (a, b) pair of c (rgb_to_lab), log(max(eps, L(c))) goes into the scratch's
 * luminance plane and its accumulator and job counters are cleared: what tdb_wiener_log_luminance_fused(prepared = 2) needs. */
int tdb_frame_prepare(const float *rgb, float *out, void *wiener_scratch, int width, int height, uint32_t filters,
                      const float *ratio, const float *bounds, float eps, tdb_stream_t stream);
/* tdb_wiener_log_luminance with two options.  prepared: 0 = plain; 1 = the scratch holds log-luminance and a cleared
 * accumulator, rgb is the colour image; 2 = as 1 and `rgb` is the (H,W,2) Lab (a, b) plane written by tdb_frame_prepare.
 * bilateral_scratch != NULL = also build the blurred bilateral grid of the OUTPUT image there (luminance plane + grid);
 * `out` then receives rgb_to_lab(result) instead of the result: pass it on with lab_input = 1 below.                     */
int tdb_wiener_log_luminance_fused(const float *rgb, float *out, void *scratch, int width, int height, int tile, int overlap,
                                   float noise, float eps, int prepared, void *bilateral_scratch, float sigma_s, float sigma_r,
                                   tdb_stream_t stream);
/* compute_image_metrics(stride, min_gray) of Bilateral.process_rgb(rgb, detail) given the blurred grid in
 * bilateral_scratch (NULL: of rgb itself), accumulated over the image set.  metrics_out: device float[5].
 * lab_input != 0: `rgb` holds rgb_to_lab of the image (the output of tdb_wiener_log_luminance_fused with a grid).        */
int tdb_metrics_sliced(const float *rgb, int lab_input, const void *bilateral_scratch, int width, int height, float sigma_s, float sigma_r,
                       float detail, int stride, float min_gray, void *frame_state, int first_in_set, int last_in_set,
                       const float *prev_metrics, float moving_average, float *metrics_out, tdb_stream_t stream);
/* tdb_tonemap of Bilateral.process_rgb(rgb, detail) given the blurred grid in bilateral_scratch.                         */
int tdb_bilateral_slice_tonemap(const float *rgb, int lab_input, const void *bilateral_scratch, uint8_t *out, int width, int height,
                                float sigma_s, float sigma_r, float detail, int op, const float *metrics, float gamma,
                                float intensity, float light_adapt, float vibrance, const float *matrix, int transform,
                                tdb_stream_t stream);
/* The two statistics steps for ONE FRAME SPLIT INTO ROW BANDS ACROSS GPUS (pipeline/tiled.py): the rank's padded band goes
 * through the kernels above, but only the rows it owns, [row_lo, row_hi) of the band, may count, and the results stay raw so
 * that the ranks can all-reduce them: raw_out = {G1 sum, G2 sum, min, max of the sampled G1 greens, min, max of every other
 * sampled value} (row_lo a multiple of 32, row_hi a multiple of 32 or the band height); raw_sums = the six metric sums.      */
int tdb_postprocess_deferred_band(const float *in, float *out, void *scratch, int width, int height, uint32_t filters, int passes,
                                  int bounds_stride, int row_lo, int row_hi, float *raw_out, tdb_stream_t stream);
int tdb_metrics_sliced_band(const float *rgb, int lab_input, const void *bilateral_scratch, int width, int height, float sigma_s,
                            float sigma_r, float detail, int stride, float min_gray, void *frame_state, int row_lo, int row_hi,
                            float *raw_sums, tdb_stream_t stream);
/* Where the ranks' partial statistics of a split frame meet (after the all-gather of the six raw_out floats of every rank, and after
 * the all-reduce of the six metric sums): ratio, bounds of the equilibrated image and metrics with their moving averages, one
 * single-thread kernel each.  gathered: device float[world][6] in rank order.  prev_* == NULL: no previous value (first frame);
 * bounds_out / metrics_out may alias prev_* (updated in place).                                                                */
int tdb_band_stats_finish(const float *gathered, int world, const float *prev_bounds, float moving_average, float *bounds_out, float *ratio_out,
                          tdb_stream_t stream);
int tdb_band_metrics_finish(const float *sums, const float *prev_metrics, float moving_average, float *metrics_out, tdb_stream_t stream);
/* the first half of tdb_bilateral_rgb: zero + splat + blur, leaving the blurred grid in scratch                          */
int tdb_bilateral_grid_rgb(const float *rgb, void *scratch, int width, int height, float sigma_s, float sigma_r,
                           tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * JPEG output of the uint8 sRGB result.  Replaces the `Jpeg` class (extension.cpp:228-233,
 * csrc/jpeg_encoder.cu:104-180): nvJPEG (vendor library, bound lazily with dlopen) driven stream-ordered on the
 * caller's stream.  The enum values are the reference's (csrc/jpeg_encoder.h:6-17).
 *   tdb_jpeg_encode    device image -> encoded stream held in the coder; *length = its size in bytes.
 *                      Interleaved formats: (H, W, 3) with `row_pitch` bytes per row (plane_stride ignored); planar
 *                      formats: 3 planes of (H, row_pitch) bytes, `plane_stride` bytes apart.  quality 1..100,
 *                      optimised Huffman tables always on (jpeg_encoder.cu:119).
 *   tdb_jpeg_retrieve  copies the stream into HOST memory (synchronises `stream`, as the reference does).          */
#define TDB_JPEG_BGR 0
#define TDB_JPEG_RGB 1
#define TDB_JPEG_BGRI 2
#define TDB_JPEG_RGBI 3
#define TDB_JPEG_CSS_444 0
#define TDB_JPEG_CSS_422 1
#define TDB_JPEG_CSS_GRAY 2
int tdb_jpeg_available(void);
int tdb_jpeg_create(void **coder);
int tdb_jpeg_destroy(void *coder);
int tdb_jpeg_encode(void *coder, const uint8_t *image, int width, int height, int64_t row_pitch, int64_t plane_stride,
                    int input_format, int quality, int subsampling, int progressive, size_t *length, tdb_stream_t stream);
int tdb_jpeg_retrieve(void *coder, uint8_t *host_out, size_t capacity, size_t *length, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Per-channel noise estimate: replaces estimate_channel_noise (reference denoise.py:131-158, a conv2d over the whole image +
 * strided slice + two torch.median) by a sampled kernel and an exact on-device selection.  sigma: device float[3] =
 * median(|r - median(r)|) / 0.6745 per channel, r = 4-neighbour Laplacian response at every `stride`-th pixel (zero padding).
 * scratch: tdb_channel_noise_scratch_bytes() bytes.  The result can be passed to tdb_wiener as its `sigmas`.               */
size_t tdb_channel_noise_scratch_bytes(int width, int height, int stride);
int tdb_channel_noise(const float *rgb, int width, int height, int stride, void *scratch, float *sigma, tdb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Pipe-throughput probes (csrc/probe.cu): kernels that keep ONLY the FP32 FMA pipe / ONLY the MUFU unit busy, for the
 * denominators of the non-HBM rooflines in bench.py (SURVEY.md 8d asks for FLOP/s beside B/px for the Wiener tiles;
 * the reference has no counterpart).  The caller times the launch with CUDA events; *flops / *ops = the work it does. */
int tdb_probe_fp32(float *sink, int iters, double *flops, tdb_stream_t stream);
int tdb_probe_mufu(float *sink, int iters, double *ops, tdb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TDB200_H */
