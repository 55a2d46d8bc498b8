// Overlapped-tile Wiener denoiser (FFT domain shrinkage) with register-resident FFTs.
//
// Reference (csrc/denoise/denoise.cu:191-242, fft.h): one CTA of KxK threads per tile, one pixel per thread, radix-2
// FFT by warp shuffles (2 shuffles per butterfly), shared-memory transposes, two global float atomics per tile pixel
// (value + weight mask) into a padded accumulator, sigmas / windows in process-global __constant__ symbols.
// Every pixel is covered by overlap^2 tiles, so at K=32/overlap 4 the stage is arithmetic-bound, not HBM-bound.
//
// Here ONE WARP owns TWO real tiles packed as a single complex tile z = a + i*b:
//   lane = column, registers = rows; column FFT in registers (radix-2 DIF, compile-time twiddles, no shuffles),
//   transpose through shared memory, row FFT in registers; the two spectra are separated with the Hermitian identity
//   A = (Z + conj(Z~))/2, B = (Z - conj(Z~))/(2i) (one shuffle per value), shrunk with their own Wiener gains, merged
//   again, and taken back through the inverse (DIT) transforms.  Half the FFT work of the reference, no shuffles inside
//   the butterflies, coalesced loads and stores (a warp touches 32 consecutive pixels of a row at a time).
//   The weight mask is separable and input-independent, so it is evaluated in closed form in the normalisation pass
//   instead of being accumulated with atomics; the accumulator has no padding.
// No process-global device state: windows and sigmas travel as kernel arguments.
//
// Three kernels share this file: wiener_tile_kernel (K = 16, the organisation above as first written), wiener32_kernel (K = 32, the
// same organisation rebuilt around instruction count; used for overlap 2 and 8) and shr::wiener32_shared_kernel (K = 32, overlap 4 --
// the default and the frame pipeline's configuration), which shares the column transforms between overlapping tiles and is
// described where it is defined.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "bilateral.cuh"
#include "fft32.cuh"
#include "fft_quad.cuh"
#include "wiener_layout.cuh"

namespace tdb {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr float kEps = 1e-15f;

// cos/sin(2*pi*k/32), k = 0..15, as literals so that unrolled butterflies fold them into immediates
__device__ __forceinline__ constexpr float tw_cos32(int k) {
  switch (k) {
    case 0: return 1.0f; case 1: return 0.98078528040323043f; case 2: return 0.92387953251128674f; case 3: return 0.83146961230254524f;
    case 4: return 0.70710678118654757f; case 5: return 0.55557023301960218f; case 6: return 0.38268343236508978f;
    case 7: return 0.19509032201612825f; case 8: return 0.0f; case 9: return -0.19509032201612825f; case 10: return -0.38268343236508978f;
    case 11: return -0.55557023301960218f; case 12: return -0.70710678118654757f; case 13: return -0.83146961230254524f;
    case 14: return -0.92387953251128674f; default: return -0.98078528040323043f;
  }
}
__device__ __forceinline__ constexpr float tw_sin32(int k) { return k < 8 ? tw_cos32(8 - k) : tw_cos32(k - 8); }  // sin(x) = cos(pi/2 - x)

// forward DIF radix-2: natural order in, bit-reversed order out.  W = exp(-2*pi*i/N)
template <int N>
__device__ __forceinline__ void fft_dif(float (&re)[N], float (&im)[N]) {
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1) {
#pragma unroll
    for (int s = 0; s < N; s += 2 * half) {
#pragma unroll
      for (int k = 0; k < half; k++) {
        const int a = s + k, b = s + k + half;
        const float ar = re[a], ai = im[a], br = re[b], bi = im[b];
        re[a] = ar + br, im[a] = ai + bi;
        const float dr = ar - br, di = ai - bi;
        const int t = k * (16 / half);  // twiddle index on the 32-point circle
        if (t == 0) {
          re[b] = dr, im[b] = di;
        } else if (t == 8) {  // multiply by -i
          re[b] = di, im[b] = -dr;
        } else {
          const float c = tw_cos32(t), sn = tw_sin32(t);  // (dr + i di) * (c - i sn)
          re[b] = dr * c + di * sn, im[b] = di * c - dr * sn;
        }
      }
    }
  }
}

// inverse DIT radix-2: bit-reversed order in, natural order out, unscaled.  W = exp(+2*pi*i/N)
template <int N>
__device__ __forceinline__ void fft_dit_inv(float (&re)[N], float (&im)[N]) {
#pragma unroll
  for (int half = 1; half < N; half <<= 1) {
#pragma unroll
    for (int s = 0; s < N; s += 2 * half) {
#pragma unroll
      for (int k = 0; k < half; k++) {
        const int a = s + k, b = s + k + half;
        const int t = k * (16 / half);
        float tr, ti;
        if (t == 0) {
          tr = re[b], ti = im[b];
        } else if (t == 8) {  // multiply by +i
          tr = -im[b], ti = re[b];
        } else {
          const float c = tw_cos32(t), sn = tw_sin32(t);  // (br + i bi) * (c + i sn)
          tr = re[b] * c - im[b] * sn, ti = im[b] * c + re[b] * sn;
        }
        const float ar = re[a], ai = im[a];
        re[a] = ar + tr, im[a] = ai + ti;
        re[b] = ar - tr, im[b] = ai - ti;
      }
    }
  }
}

template <int K>
__device__ __forceinline__ constexpr int brev(int v) {
  int r = 0;
  for (int b = 1; b < K; b <<= 1) r = (r << 1) | ((v & b) ? 1 : 0);
  return r;
}

// transpose a KxK complex block held as (lane = j, register = i) into (lane = i, register = j) through shared memory
template <int K>
__device__ __forceinline__ void transpose(float (&re)[K], float (&im)[K], float *sre, float *sim, int lane_in) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < K; i++) sre[lane_in * (K + 1) + i] = re[i], sim[lane_in * (K + 1) + i] = im[i];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < K; i++) re[i] = sre[i * (K + 1) + lane_in], im[i] = sim[i * (K + 1) + lane_in];
}

__device__ __forceinline__ int reflect_index(int x, int limit) {
  if (x < 0) x = -x;
  if (x >= limit) x = 2 * limit - x - 1;
  // the reference reads out of bounds here when a side is shorter than 2K-1; stay defined instead
  return min(max(x, 0), limit - 1);
}

struct WienerArgs {
  const float *in;     // (H, W, C)
  float *acc;          // (H, W, C) accumulator, zeroed by the caller
  const float *sigmas; // device float[C], or null when sigma_value is used
  float sigma_value;
  int width, height, channels;
  int stride;          // K / overlap
  int grid_w, grid_h;  // tiles per row / column
  int pairs_w;         // tile pairs per row
  int64_t njobs;       // grid_h * pairs_w * channels
  int px_lo, px_hi, gy_lo, gy_hi;  // K = 32: rectangle of tile pairs that lie fully inside the image (empty if px_hi < px_lo)
  int64_t njobs_interior;
  unsigned int *counters;  // two job counters (interior, border) in the scratch buffer, zeroed with the accumulator
  float win[32];       // 1-D window (both the FFT and the interpolation window of the reference)
};

template <int K>
__global__ void __launch_bounds__(kThreads) wiener_tile_kernel(const WienerArgs a) {
  constexpr int SUB = 32 / K;  // tile pairs handled by one warp at a time
  extern __shared__ float s_t[];  // [kWarps * SUB][2][K * (K + 1)] transpose staging (dynamic: 66 KB at K = 32)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / K, c = lane % K;  // c = this lane's column inside the tile
  float *sre = s_t + (size_t)(warp * SUB + sub) * 2 * K * (K + 1), *sim = sre + K * (K + 1);
  const float wc = a.win[c];
  const unsigned sub_mask = SUB == 1 ? 0xffffffffu : (0xffffu << (16 * sub));
  // lane holding frequency -kx after the row/column exchange: registers/lane indices are in bit-reversed order
  const int partner_lane = sub * K + brev<K>((K - brev<K>(c)) & (K - 1));

  const int64_t job0 = ((int64_t)blockIdx.x * kWarps + warp) * SUB + sub;
  const int64_t job_step = (int64_t)gridDim.x * kWarps * SUB;
  // all lanes of a warp must run the same number of iterations (shuffles / __syncwarp inside)
  const int64_t iters = (a.njobs + job_step - 1) / job_step;
  for (int64_t it = 0; it < iters; it++) {
    const int64_t job = job0 + it * job_step;
    const bool live = job < a.njobs;
    int ch = 0, gy = 0, px = 0;
    if (live) {
      ch = (int)(job % a.channels);
      const int64_t t = job / a.channels;
      px = (int)(t % a.pairs_w), gy = (int)(t / a.pairs_w);
    }
    const int shift = K / a.stride;  // the tile grid starts one tile early (denoise.cu:146)
    const int oy = (gy - shift) * a.stride;
    const int ox0 = (2 * px - shift) * a.stride, ox1 = (2 * px + 1 - shift) * a.stride;
    const bool has_b = live && (2 * px + 1) < a.grid_w;

    float re[K], im[K];
    float sum_a = 0.0f, sum_b = 0.0f;
    {
      const int xa = reflect_index(ox0 + c, a.width), xb = reflect_index(ox1 + c, a.width);
#pragma unroll
      for (int r = 0; r < K; r++) {
        const int y = reflect_index(oy + r, a.height);
        const float va = live ? __ldg(a.in + ((int64_t)y * a.width + xa) * a.channels + ch) : 0.0f;
        const float vb = has_b ? __ldg(a.in + ((int64_t)y * a.width + xb) * a.channels + ch) : 0.0f;
        re[r] = va, im[r] = vb;
        sum_a += va, sum_b += vb;
      }
    }
#pragma unroll
    for (int o = K / 2; o > 0; o >>= 1) {
      sum_a += __shfl_xor_sync(sub_mask, sum_a, o);
      sum_b += __shfl_xor_sync(sub_mask, sum_b, o);
    }
    const float mean_a = sum_a / (float)(K * K), mean_b = sum_b / (float)(K * K);
#pragma unroll
    for (int r = 0; r < K; r++) {
      const float w = a.win[r] * wc;  // reference: window_fft[pos.x] * window_fft[pos.y]
      re[r] = (re[r] - mean_a) * w, im[r] = (im[r] - mean_b) * w;
    }

    fft_dif<K>(re, im);                          // along y (registers)
    transpose<K>(re, im, sre, sim, c);           // lane = ky (bit-reversed), registers = x
    fft_dif<K>(re, im);                          // along x

    // separate the two real-input spectra, apply the Wiener gain to each, merge  (apply_gain: denoise.cu:181-185)
    {
      const float sg = a.sigmas ? __ldg(a.sigmas + ch) : a.sigma_value;
      const float s2 = sg * sg;
      // register p holds kx = brev(p); its mirror -kx sits in register brev((K - brev(p)) % K)
#pragma unroll
      for (int p = 0; p < K; p++) {
        const int q = brev<K>((K - brev<K>(p)) & (K - 1));
        if (q < p) continue;  // handled together with its mirror
        const float zr_p = re[p], zi_p = im[p], zr_q = re[q], zi_q = im[q];
        // mirrored bin of (lane, p) is (partner_lane, q) and vice versa
        const float mr_p = __shfl_sync(sub_mask, zr_q, partner_lane), mi_p = __shfl_sync(sub_mask, zi_q, partner_lane);
        const float mr_q = __shfl_sync(sub_mask, zr_p, partner_lane), mi_q = __shfl_sync(sub_mask, zi_p, partner_lane);
        {
          const float ar = 0.5f * (zr_p + mr_p), ai = 0.5f * (zi_p - mi_p);
          const float br = 0.5f * (zi_p + mi_p), bi = -0.5f * (zr_p - mr_p);
          const float pa = ar * ar + ai * ai + kEps, pb = br * br + bi * bi + kEps;
          const float ga = fmaxf(pa - s2, 0.0f) / pa, gb = fmaxf(pb - s2, 0.0f) / pb;
          re[p] = ga * ar - gb * bi, im[p] = ga * ai + gb * br;
        }
        if (q != p) {
          const float ar = 0.5f * (zr_q + mr_q), ai = 0.5f * (zi_q - mi_q);
          const float br = 0.5f * (zi_q + mi_q), bi = -0.5f * (zr_q - mr_q);
          const float pa = ar * ar + ai * ai + kEps, pb = br * br + bi * bi + kEps;
          const float ga = fmaxf(pa - s2, 0.0f) / pa, gb = fmaxf(pb - s2, 0.0f) / pb;
          re[q] = ga * ar - gb * bi, im[q] = ga * ai + gb * br;
        }
      }
    }

    fft_dit_inv<K>(re, im);                      // along x
    transpose<K>(re, im, sre, sim, c);           // lane = x, registers = ky (bit-reversed)
    fft_dit_inv<K>(re, im);                      // along y

    // overlap-add: (value + mean * w_fft) * w_interp   (store_pixel, denoise.cu:150-178)
    const float inv = 1.0f / (float)(K * K);
    const int xa = ox0 + c, xb = ox1 + c;
#pragma unroll
    for (int r = 0; r < K; r++) {
      const int y = oy + r;
      if (y < 0 || y >= a.height) continue;
      const float w = a.win[r] * wc;
      if (live && xa >= 0 && xa < a.width) atomicAdd(a.acc + ((int64_t)y * a.width + xa) * a.channels + ch, (re[r] * inv + mean_a * w) * w);
      if (has_b && xb >= 0 && xb < a.width) atomicAdd(a.acc + ((int64_t)y * a.width + xb) * a.channels + ch, (im[r] * inv + mean_b * w) * w);
    }
  }
}

// Wiener shrinkage of the two real tiles packed as z = a + i b  (apply_gain: denoise.cu:181-185), in place on the spectrum held
// as (lane = ky in some order, register p = kx = brev(p)); `partner` = the lane that holds -ky.
// With M = conj(Z(-k)):  A = (Z + M)/2, iB = (Z - M)/2, |A|^2 = (sr^2 + di^2)/4, |B|^2 = (si^2 + dr^2)/4 where
// s = Z + Z(-k), d = Z - Z(-k) componentwise; shrunk Z' = gA A + i gB B.  The Hermitian pair {(ky, kx), (-ky, -kx)} shares one gain
// evaluation: lane ky computes both shrunk bins and hands the mirrored one to lane -ky.  The scale 1/K^2 of the two inverse
// transforms is folded into the gain numerators (kq = 1/(2 K^2) against the factor 2 of s and d).
__device__ __forceinline__ void wiener_gains(float (&re)[32], float (&im)[32], int partner, float sg) {
  constexpr int K = 32;
  constexpr float kq = 1.0f / (2.0f * K * K);
  const float n0 = kq * (kEps - sg * sg);
#pragma unroll
  for (int p = 0; p < K; p++) {
    const int q = fft::brev<K>((K - fft::brev<K>(p)) & (K - 1));  // register holding -kx
    if (q < p) continue;                                           // written by the exchange below
    const float zr = re[p], zi = im[p];
    const float mr = __shfl_sync(0xffffffffu, re[q], partner), mi = __shfl_sync(0xffffffffu, im[q], partner);
    const float sr = zr + mr, dr = zr - mr, si = zi + mi, di = zi - mi;
    const float qa = fmaf(sr, sr, di * di), qb = fmaf(si, si, dr * dr);
    const float ga = __fdividef(fmaxf(fmaf(qa, 0.25f * kq, n0), 0.0f), fmaf(qa, 0.25f, kEps));
    const float gb = __fdividef(fmaxf(fmaf(qb, 0.25f * kq, n0), 0.0f), fmaf(qb, 0.25f, kEps));
    const float t0 = ga * sr, t1 = gb * dr, t2 = ga * di, t3 = gb * si;
    re[p] = t0 + t1, im[p] = t2 + t3;
    if (q != p) {  // the mirrored bin (-ky, -kx) lives in the partner lane's register q
      re[q] = __shfl_sync(0xffffffffu, t0 - t1, partner);
      im[q] = __shfl_sync(0xffffffffu, t3 - t2, partner);
    }
  }
}

// ---- K = 32 fast path ------------------------------------------------------------------------------------------------
// Same warp = tile-pair organisation, rebuilt around instruction count (the v1 kernel above executed about 6400 warp
// instructions per tile pair, half of them integer address arithmetic; ncu: profiles/r01_wiener_tile_v1_sass_hist.txt):
//   * FMA-form DIT butterflies in both directions (fft32.cuh): natural -> bit-reversed -> natural, no reorder pass
//   * interior tile pairs (the overwhelming majority) load and accumulate through one running pointer with
//     immediate offsets; only border pairs take the reflecting path
//   * transposes move (re, im) as 64-bit words: 64 + 64 shared-memory instructions instead of 128 + 128
//   * the Hermitian pair {(ky, kx), (-ky, -kx)} shares one gain evaluation: lane ky computes both shrunk bins and
//     hands the mirrored one to lane -ky; the 1/(2*K*K) scale is folded into the gain numerators
// kC1: single-channel image (the log-luminance path of the pipeline): the second tile of a pair sits at an immediate offset
template <int STRIDE, bool kC1>
__global__ void __launch_bounds__(kThreads, 2) wiener32_kernel(const WienerArgs a) {
  constexpr int K = 32, LD = K + 1;
  extern __shared__ float2 s_z[];  // [kWarps][K * LD] transpose staging
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2 *st = s_z + (size_t)warp * K * LD;
  const float wc = a.win[lane];
  const int partner = (K - lane) & (K - 1);  // lane holding -ky
  constexpr int shift = K / STRIDE;          // the tile grid starts one tile early (denoise.cu:146)
  const int cs = kC1 ? 1 : a.channels;
  const int64_t row_step = (int64_t)a.width * cs;

  // Interior tile pairs (both tiles inside the image) form the rectangle [px_lo, px_hi] x [gy_lo, gy_hi]; they take a path
  // without any reflecting code (ncu showed 29 % of the stall samples waiting for instruction fetch when every pair carried it).
  // The job queue lists the interior pairs first and the border pairs last, so the reflecting code is only fetched while the
  // grid drains.  kBorder is uniform per warp.  All warps of a CTA meet at a barrier per job, which keeps the four warps of a
  // scheduler in the same stretch of this long, loop-free instruction stream.
  const int64_t njobs = a.njobs;
  unsigned int *counter = a.counters;
  __shared__ unsigned int s_base;
  for (;;) {
    // dynamic scheduling: a CTA claims kWarps consecutive jobs (neighbours in x: shared lines in L1) at a time.
    // (Claiming one step ahead and prefetching the next pair's rows into L1 -- prefetch.global.L1, SASS CCTL.E.PF1 -- was
    // measured: 0.285 -> 0.300 ms; without the prefetch the single-barrier variant is a wash.)
    __syncthreads();
    if (threadIdx.x == 0) s_base = atomicAdd(counter, (unsigned int)kWarps);
    __syncthreads();
    const int64_t base = s_base;
    if (base >= njobs) break;
    int64_t job = base + warp;
    if (job >= njobs) continue;  // whole warps: no divergence around the shuffles
    const bool kBorder = job >= a.njobs_interior;
    if (kBorder) job -= a.njobs_interior;
    const int ch = (int)(job % cs);
    const int64_t t = job / cs;
    int px, gy;
    if (kBorder) {
      // tile pairs outside the interior rectangle, enumerated without gaps: the rows above it, the rows below it, then
      // the left and right flanks of the rows it spans
      const int64_t above = (int64_t)a.gy_lo * a.pairs_w, below = (int64_t)(a.grid_h - 1 - a.gy_hi) * a.pairs_w;
      if (a.njobs_interior == 0) {
        px = (int)(t % a.pairs_w), gy = (int)(t / a.pairs_w);
      } else if (t < above) {
        px = (int)(t % a.pairs_w), gy = (int)(t / a.pairs_w);
      } else if (t < above + below) {
        const int64_t u = t - above;
        px = (int)(u % a.pairs_w), gy = a.gy_hi + 1 + (int)(u / a.pairs_w);
      } else {
        const int64_t u = t - above - below;
        const int flank = a.px_lo + (a.pairs_w - 1 - a.px_hi);
        const int f = (int)(u % flank);
        gy = a.gy_lo + (int)(u / flank);
        px = f < a.px_lo ? f : a.px_hi + 1 + (f - a.px_lo);
      }
    } else {
      const int npx = a.px_hi - a.px_lo + 1;
      px = a.px_lo + (int)(t % npx), gy = a.gy_lo + (int)(t / npx);
    }
    const int oy = (gy - shift) * STRIDE;
    const int ox0 = (2 * px - shift) * STRIDE, ox1 = ox0 + STRIDE;
    const bool has_b = (2 * px + 1) < a.grid_w;

    float re[K], im[K];
    float sum_a = 0.0f, sum_b = 0.0f;
    if (!kBorder) {
      // 32-bit element offsets from the plane base (check_args: they fit): one IMAD.WIDE per row instead of the 64-bit pointer
      // walks the compiler made of `p += row_step` (8 integer instructions per row, ncu)
      int off = (int)(((int64_t)oy * a.width + ox0 + lane) * cs + ch);
      const int boff = kC1 ? STRIDE : STRIDE * cs, rs = (int)row_step;
#pragma unroll
      for (int r = 0; r < K; r++) {
        const float *p = a.in + off;
        re[r] = __ldg(p), im[r] = __ldg(p + boff);
        off += rs;
      }
    } else {
      const int xa = reflect_index(ox0 + lane, a.width), xb = reflect_index(ox1 + lane, a.width);
#pragma unroll
      for (int r = 0; r < K; r++) {
        const int64_t row = (int64_t)reflect_index(oy + r, a.height) * a.width;
        re[r] = __ldg(a.in + (row + xa) * cs + ch);
        im[r] = has_b ? __ldg(a.in + (row + xb) * cs + ch) : 0.0f;
      }
    }
#pragma unroll
    for (int r = 0; r < K; r++) sum_a += re[r], sum_b += im[r];
    const float mean_a = warp_sum(sum_a) * (1.0f / (K * K)), mean_b = warp_sum(sum_b) * (1.0f / (K * K));
#pragma unroll
    for (int r = 0; r < K; r++) {
      const float w = a.win[r] * wc;  // reference: window_fft[pos.x] * window_fft[pos.y]
      re[r] = (re[r] - mean_a) * w, im[r] = (im[r] - mean_b) * w;
    }

    fft::fft_fwd<K>(re, im);  // along y: register p holds ky = brev(p)
    __syncwarp();
#pragma unroll
    for (int p = 0; p < K; p++) st[fft::brev<K>(p) * LD + lane] = make_float2(re[p], im[p]);
    __syncwarp();
#pragma unroll
    for (int x = 0; x < K; x++) {
      const float2 v = st[lane * LD + x];  // lane = ky, register = x
      re[x] = v.x, im[x] = v.y;
    }
    fft::fft_fwd<K>(re, im);  // along x: register p holds kx = brev(p)

    wiener_gains(re, im, partner, a.sigmas ? __ldg(a.sigmas + ch) : a.sigma_value);

    fft::fft_inv<K>(re, im);  // along x: register = x
    __syncwarp();
#pragma unroll
    for (int x = 0; x < K; x++) st[lane * LD + x] = make_float2(re[x], im[x]);
    __syncwarp();
#pragma unroll
    for (int p = 0; p < K; p++) {
      const float2 v = st[fft::brev<K>(p) * LD + lane];  // lane = x, register p holds ky = brev(p)
      re[p] = v.x, im[p] = v.y;
    }
    fft::fft_inv<K>(re, im);  // along y: register = row

    // overlap-add: (value + mean * w_fft) * w_interp   (store_pixel, denoise.cu:150-178)
    if (!kBorder) {
      int off = (int)(((int64_t)oy * a.width + ox0 + lane) * cs + ch);
      const int boff = kC1 ? STRIDE : STRIDE * cs, rs = (int)row_step;
#pragma unroll
      for (int r = 0; r < K; r++) {
        const float w = a.win[r] * wc;
        float *o = a.acc + off;
        atomicAdd(o, fmaf(mean_a, w, re[r]) * w);
        atomicAdd(o + boff, fmaf(mean_b, w, im[r]) * w);
        off += rs;
      }
    } else {
      const int xa = ox0 + lane, xb = ox1 + lane;
      const bool in_a = xa >= 0 && xa < a.width, in_b = has_b && xb >= 0 && xb < a.width;
#pragma unroll
      for (int r = 0; r < K; r++) {
        const int y = oy + r;
        if (y < 0 || y >= a.height) continue;
        const float w = a.win[r] * wc;
        if (in_a) atomicAdd(a.acc + ((int64_t)y * a.width + xa) * cs + ch, fmaf(mean_a, w, re[r]) * w);
        if (in_b) atomicAdd(a.acc + ((int64_t)y * a.width + xb) * cs + ch, fmaf(mean_b, w, im[r]) * w);
      }
    }
  }
}

// ---- K = 32, stride 8 (the default overlap 4): column transforms shared between overlapping tiles -----------------------------------
// wiener32_kernel runs four 32-point transform passes per tile pair although the four tiles that overlap horizontally apply the
// SAME windowed column transform to the columns they share.  Here tiles are paired VERTICALLY (tile rows oy and oy + 8 as the real
// and the imaginary part), so the column spectrum U_x[ky] = FFT_r(w_r (v[oy + r][x] + i v[oy + 8 + r][x])) depends on the image
// column x alone and is computed once per column and tile-row pair.  The window is separable (window.h:18-43) and the tile mean
// comes out in the frequency domain: with mu = mean_a + i mean_b and What = FFT(w),
//     Z[ky][kx] = FFT_c( w_c (U_{ox+c}[ky] - mu What[ky]) ),
// the shrinkage is the one of wiener32_kernel, and the inverse column transform is linear, so the four tiles covering a column
// add up in the frequency domain first:
//     C_x[ky] += w_c (IFFT_kx(Z')[c] + mu What[ky] w_c / 32),      out[oy + r][x] = w_r IFFT_ky(C_x)[r]
// (real part: tile row oy, imaginary part: tile row oy + 8).  Per tile pair: 8 + 32 + 32 + 8 one-dimensional transforms instead of
// 128, 2.5 instead of 16 loads and 4 instead of 16 atomics per pixel.
// A CTA walks along a tile-row pair in steps of eight tile pairs (one per warp): spectra and accumulators of the 88 columns those
// tiles touch live in shared memory ([column][36] float2: conflict-free for the column phase, 4 lanes x 8 registers per column, and
// for the row phase, lane = frequency); the 24 columns shared with the next step are carried over.  Column transforms use four
// lanes per column (fft_quad.cuh) so that the 64 new columns of a step occupy all 256 threads.  The work is split statically: every
// CTA gets the same number of steps of the linearised (tile-row pair, step) sequence -- all steps cost the same, reflecting loads
// and bounds-checked atomics included, so no interior / border split is needed.
// (Measured and dropped: the row phase on two-wide FP32 instructions -- PTX fma/add/mul.rn.f32x2, SASS FFMA2 / FADD2 with the free
// .LO_HI half swap, (re, im) in one register pair -- halves the FP instruction count (194 instead of 388 per 32-point transform, 1740
// instead of 2330 warp instructions per step, same results) but runs at 0.2215 ms against 0.2169 ms: the kernel is bound by the FP32
// lanes, not by issue slots, and a two-wide instruction occupies them twice as long.)
namespace shr {

constexpr int K = 32, ST = 8;
constexpr int TPS = kWarps;          // tile pairs per step
constexpr int NEWC = TPS * ST;       // 64 new columns per step
constexpr int CARRY = K - ST;        // 24 columns shared with the next step
constexpr int BUFC = NEWC + CARRY;   // 88
constexpr int LD = 36;               // float2 per column; = 4 (mod 16): the 16 lanes of a half-warp (4 columns x 4 quarters) hit 16 banks
constexpr int kSmemBytes = (2 * BUFC * LD + BUFC + 32) * (int)sizeof(float2);

struct Args {
  const float *in;      // (H, W, C)
  float *acc;           // (H, W, C), zeroed by the caller
  const float *sigmas;  // device float[C] or null
  float sigma_value;
  int width, height, channels;
  int n_pairs;          // tile-row pairs per channel
  int n_tx;             // tiles per row that touch the image: ox = -24 + 8 t
  int steps_per_row, total_steps;
  float win[32];        // 1-D window
  float w2[32];         // win^2 / 32
  fft::cpx what[32];    // FFT of the window
  fft::cpx tw[32];      // fft_quad twiddles
};

// kCtas: resident CTAs per SM the register allocation is sized for (2: 127 registers; 3: 80 registers, 16 bytes of spills -- 24 warps per SM)
__device__ __forceinline__ void wiener32_shared_body(const Args &a) {
  extern __shared__ float2 s_shr[];
  float2 *spec = s_shr, *accu = s_shr + BUFC * LD, *csum = accu + BUFC * LD, *twt = csum + BUFC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = a.width, H = a.height;
  // column phase: lane = (column of the chunk, quarter); row phase: lane = position inside a column = frequency quad_freq(p, j)
  const int cj = lane & 3, cx = lane >> 2;
  const int ky = fft::quad_freq(lane >> 2, lane & 3);
  const int nky = (K - ky) & (K - 1);
  const int partner = fft::quad_reg_of(nky) * 4 + fft::quad_lane_of(nky);
  const float whx = a.what[ky].x, why = a.what[ky].y;
  if (tid < 32) twt[tid] = make_float2(a.tw[tid].x, a.tw[tid].y);
  __syncthreads();  // the table is read by every warp's first column transform
  const float2 *twl = twt + 8 * cj;
  const int cs = a.channels, Wc = W * cs;  // channels are interleaved: element (y, x, ch) at (y * W + x) * cs + ch

  const int L0 = (int)((int64_t)blockIdx.x * a.total_steps / gridDim.x), L1 = (int)((int64_t)(blockIdx.x + 1) * a.total_steps / gridDim.x);
  const int w4 = 4 * Wc;
  int P = L0 / a.steps_per_row, k = L0 - P * a.steps_per_row;  // the sequence runs over (channel, tile-row pair, step)
  int ch = P / a.n_pairs;
  P -= ch * a.n_pairs;
  for (int L = L0; L < L1; L++, k++) {
    if (k == a.steps_per_row) {
      k = 0;
      if (++P == a.n_pairs) P = 0, ch++;
    }
    const float sg = a.sigmas ? __ldg(a.sigmas + ch) : a.sigma_value;
    const bool first = L == L0 || k == 0, last = L == L1 - 1 || k == a.steps_per_row - 1;
    const int oy = -CARRY + 2 * ST * P;   // tile row of the real part; the imaginary part is the tile row oy + 8
    const int cb = -CARRY + NEWC * k;     // image column of buffer column 0; tile pair w of this step starts at cb + 8 w
    const bool rows_inside = oy >= 0 && oy + K + ST <= H;

    // ---- column phase: spectra of the new columns -------------------------------------------------------------------------
    // (every warp runs the same instruction stream: chunks beyond the buffer are computed on a clamped column and not stored, so
    // the shuffles sit in uniform control flow)
    if (first)
      for (int i = tid; i < BUFC * LD; i += kThreads) accu[i] = make_float2(0.0f, 0.0f);
    auto column_forward = [&](int c0) {
      const bool valid = c0 < BUFC;
      const int bc = (valid ? c0 : 0) + cx;
      const float *col = a.in + (reflect_index(cb + bc, W) * cs + ch);
      float v[10];
      if (rows_inside) {  // rows oy .. oy + 39 inside the image (uniform per step): no reflection, one multiply-add per row
        const int off = (oy + cj) * Wc;  // element offsets fit 32 bits (check_args)
#pragma unroll
        for (int m = 0; m < 10; m++) v[m] = __ldg(col + (off + m * w4));
      } else {
#pragma unroll
        for (int m = 0; m < 10; m++) v[m] = __ldg(col + (int64_t)reflect_index(oy + 4 * m + cj, H) * Wc);
      }
      float re[8], im[8], sa = 0.0f, sb = 0.0f;
#pragma unroll
      for (int m = 0; m < 8; m++) {
        const float wm = a.win[4 * m + cj];
        sa += v[m], sb += v[m + 2];
        re[m] = v[m] * wm, im[m] = v[m + 2] * wm;
      }
      sa += __shfl_xor_sync(0xffffffffu, sa, 1), sb += __shfl_xor_sync(0xffffffffu, sb, 1);
      sa += __shfl_xor_sync(0xffffffffu, sa, 2), sb += __shfl_xor_sync(0xffffffffu, sb, 2);
      fft::quad_fwd_local(re, im, twl);
#pragma unroll
      for (int p = 0; p < 8; p++) {
        const float pr = __shfl_xor_sync(0xffffffffu, re[p], 2), pi = __shfl_xor_sync(0xffffffffu, im[p], 2);
        fft::quad_fwd_a(re[p], im[p], pr, pi, cj);
      }
#pragma unroll
      for (int p = 0; p < 8; p++) {
        const float pr = __shfl_xor_sync(0xffffffffu, re[p], 1), pi = __shfl_xor_sync(0xffffffffu, im[p], 1);
        fft::quad_fwd_b(re[p], im[p], pr, pi, cj);
      }
      if (valid) {
        if (cj == 0) csum[bc] = make_float2(sa, sb);
        float2 *dst = spec + bc * LD + cj;
#pragma unroll
        for (int p = 0; p < 8; p++) dst[4 * p] = make_float2(re[p], im[p]);
      }
    };
    column_forward((first ? 0 : CARRY) + ST * warp);
    if (first) column_forward(NEWC + ST * warp);  // the 24 columns a step normally inherits
    __syncthreads();

    // ---- row phase: one tile pair per warp -------------------------------------------------------------------------------------
    const bool active = TPS * k + warp < a.n_tx;  // uniform per warp
    float re[K], im[K];
    float q32r, q32i;
    {  // warps without a tile (end of a tile row) run on whatever the buffer holds and skip the accumulation
      const float2 *src = spec + (ST * warp) * LD + lane;
#pragma unroll
      for (int c = 0; c < K; c++) {
        const float2 z = src[c * LD];
        re[c] = z.x, im[c] = z.y;
      }
      const float2 cs = csum[ST * warp + lane];
      const float ma = warp_sum(cs.x) * (1.0f / (K * K)), mb = warp_sum(cs.y) * (1.0f / (K * K));
      const float qr = ma * whx - mb * why, qi = ma * why + mb * whx;  // (mean_a + i mean_b) What[ky]
#pragma unroll
      for (int c = 0; c < K; c++) re[c] = (re[c] - qr) * a.win[c], im[c] = (im[c] - qi) * a.win[c];
      fft::fft_fwd<K>(re, im);  // along x: register p holds kx = brev(p)
      wiener_gains(re, im, partner, sg);
      fft::fft_inv<K>(re, im);  // register = column of the tile, scaled by 1 / K^2
      q32r = qr, q32i = qi;     // a.w2 carries the 1/32
    }
    // the four column blocks of a tile are shared with the neighbouring warps' tiles: block j of every warp in turn
#pragma unroll
    for (int j = 0; j < K / ST; j++) {
      if (active) {
        float2 *dst = accu + (ST * warp + ST * j) * LD + lane;
#pragma unroll
        for (int c = 0; c < ST; c++) {
          float2 z = dst[c * LD];
          z.x = fmaf(a.win[ST * j + c], re[ST * j + c], fmaf(a.w2[ST * j + c], q32r, z.x));
          z.y = fmaf(a.win[ST * j + c], im[ST * j + c], fmaf(a.w2[ST * j + c], q32i, z.y));
          dst[c * LD] = z;
        }
      }
      __syncthreads();
    }

    // ---- column phase back: the columns no later tile touches ---------------------------------------------------------------------
    auto column_back = [&](int c0) {
      const bool valid = c0 < BUFC;
      const int bc = (valid ? c0 : 0) + cx;
      const float2 *src = accu + bc * LD + cj;
      float cr[8], ci[8];
#pragma unroll
      for (int p = 0; p < 8; p++) {
        const float2 z = src[4 * p];
        cr[p] = z.x, ci[p] = z.y;
      }
#pragma unroll
      for (int p = 0; p < 8; p++) {
        const float pr = __shfl_xor_sync(0xffffffffu, cr[p], 1), pi = __shfl_xor_sync(0xffffffffu, ci[p], 1);
        fft::quad_inv_b(cr[p], ci[p], pr, pi, cj);
      }
#pragma unroll
      for (int p = 0; p < 8; p++) {
        const float pr = __shfl_xor_sync(0xffffffffu, cr[p], 2), pi = __shfl_xor_sync(0xffffffffu, ci[p], 2);
        fft::quad_inv_a(cr[p], ci[p], pr, pi, cj);
      }
      fft::quad_inv_local(cr, ci, twl);  // register m = row 4 m + cj of the tile
      const int x = cb + bc;
      if (valid && x >= 0 && x < W) {
        float wr[8];
#pragma unroll
        for (int m = 0; m < 8; m++) wr[m] = a.win[4 * m + cj];
        float val[10];  // rows oy + 4 m + cj: tile row oy contributes m < 8, tile row oy + 8 contributes m >= 2
#pragma unroll
        for (int m = 0; m < 10; m++) {
          val[m] = m < 8 ? cr[m] * wr[m] : 0.0f;
          if (m >= 2) val[m] = fmaf(ci[m - 2], wr[m - 2], val[m]);
        }
        float *dst = a.acc + (x * cs + ch);
        if (rows_inside) {
          const int off = (oy + cj) * Wc;
#pragma unroll
          for (int m = 0; m < 10; m++) atomicAdd(dst + (off + m * w4), val[m]);
        } else {
#pragma unroll
          for (int m = 0; m < 10; m++) {
            const int y = oy + 4 * m + cj;
            if (y >= 0 && y < H) atomicAdd(dst + (int64_t)y * Wc, val[m]);
          }
        }
      }
    };
    column_back(ST * warp);
    if (last) column_back(NEWC + ST * warp);
    __syncthreads();
    if (!last) {
      // carry the 24 columns the next step shares (spectra, partial accumulators, column sums), clear the rest of the accumulators
      constexpr int NC = CARRY * LD;  // 864 float2
      float2 ks[(NC + kThreads - 1) / kThreads], ka[(NC + kThreads - 1) / kThreads];
#pragma unroll
      for (int i = 0; i < (NC + kThreads - 1) / kThreads; i++) {
        const int e = tid + i * kThreads;
        if (e < NC) ks[i] = spec[NEWC * LD + e], ka[i] = accu[NEWC * LD + e];
      }
      float2 kc = make_float2(0.0f, 0.0f);
      if (tid < CARRY) kc = csum[NEWC + tid];
      __syncthreads();
#pragma unroll
      for (int i = 0; i < (NC + kThreads - 1) / kThreads; i++) {
        const int e = tid + i * kThreads;
        if (e < NC) spec[e] = ks[i], accu[e] = ka[i];
      }
      if (tid < CARRY) csum[tid] = kc;
      for (int i = NC + tid; i < BUFC * LD; i += kThreads) accu[i] = make_float2(0.0f, 0.0f);
      __syncthreads();
    }
  }
}

template <int kCtas>
__global__ void __launch_bounds__(kThreads, kCtas) wiener32_shared_kernel(const Args a) { wiener32_shared_body(a); }
// The same body capped at kRegs registers per thread without a residency promise: two CTAs of it leave a quarter of the SM's
// register file to a CTA of ANOTHER kernel, which matters when two frames are in flight (ImageProcessor.submit)
template <int kRegs>
__global__ void __maxnreg__(kRegs) wiener32_shared_kernel_capped(const Args a) { wiener32_shared_body(a); }

}  // namespace shr

// closed-form weight mask: sum over the overlap^2 covering tiles of (w_fft * w_interp)(x) * (w_fft * w_interp)(y)
struct NormArgs {
  const float *acc;
  const float *rgb;  // only for the log-luminance composite
  float *out;
  int width, height, channels, K, stride;
  float win[32];
  // kSplat: the denoised colour leaves as Lab, and its L is stored once more as a plane (4 B/px): it is what the bilateral grid
  // is built from (Bilateral.process_rgb's first step), so that stage does not read the image again
  float *lum_out;
};

// kAb: `rgb` holds the Lab (a, b) pairs of the colour (frame_prepare) instead of the colour itself
template <bool kLogLum, bool kSplat, bool kAb = false>
__global__ void __launch_bounds__(256) wiener_normalize_kernel(const NormArgs a) {
  __shared__ float m1[32];  // 1-D mask factor per phase: sum_j win[r + j*stride]^2
  if (threadIdx.x < a.stride) {
    float s = 0.0f;
    for (int j = threadIdx.x; j < a.K; j += a.stride) s += a.win[j] * a.win[j];
    m1[threadIdx.x] = s;
  }
  __syncthreads();
  const int64_t n = (int64_t)a.width * a.height;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / a.width), x = (int)(i - (int64_t)y * a.width);
    const float mask = m1[x % a.stride] * m1[y % a.stride];
    if (kLogLum) {
      const float l = __ldg(a.acc + i) / (mask + kEps);
      rgb_t r;
      if (kAb) {  // pub::with_luminance with rgb_to_lab(c).a/b already at hand
        const float2 ab = __ldg(reinterpret_cast<const float2 *>(a.rgb) + i);
        r = clip01(pub::lab_to_rgb(rgb_t{fmaxf(0.0f, fminf(1.0f, expf(l))), ab.x, ab.y}));
      } else {
        const rgb_t c{__ldg(a.rgb + 3 * i), __ldg(a.rgb + 3 * i + 1), __ldg(a.rgb + 3 * i + 2)};
        r = pub::with_luminance(c, expf(l));
      }
      if (kSplat) {
        // the bilateral stage follows: hand it rgb_to_lab(r) instead of r (r is inside [0,1]^3, so Lab L is compute_luminance(r),
        // what the grid is built from, and the slice needs exactly this Lab for its modify_luminance): one conversion, not three
        const rgb_t lab = pub::rgb_to_lab(r);
        a.out[3 * i] = lab.x, a.out[3 * i + 1] = lab.y, a.out[3 * i + 2] = lab.z;
        a.lum_out[i] = fmaxf(0.0f, lab.x);
      } else {
        a.out[3 * i] = r.x, a.out[3 * i + 1] = r.y, a.out[3 * i + 2] = r.z;
      }
    } else {
      for (int ch = 0; ch < a.channels; ch++) a.out[i * a.channels + ch] = __ldg(a.acc + i * a.channels + ch) / (mask + kEps);
    }
  }
}

// The log-luminance composite of the frame pipeline (kLogLum above) with four consecutive pixels of a row per thread: 128-bit loads
// and stores, the pixel's row / column kept incrementally instead of a 64-bit division per pixel, the stride (K / overlap, a power
// of two) applied as a bit mask, and 1 / (mask + eps) read from a stride x stride shared table filled with the very MUFU.RCP the
// per-pixel division compiled to -- ncu counted 215 instructions and 26 XU operations per pixel in the scalar kernel, a third of
// them index arithmetic.  Same arithmetic per pixel, bit-identical output.  Requires width % 4 == 0 and 16-byte aligned planes.
template <bool kSplat, bool kAb>
__global__ void __launch_bounds__(256) wiener_normalize_lum4_kernel(const NormArgs a, const bool lum_aligned) {
  __shared__ float m1[16];       // 1-D mask factor per phase: sum_j win[r + j*stride]^2
  __shared__ float inv[16][16];  // [y phase][x phase]: 1 / (mask + eps)
  const int st = a.stride, tid = threadIdx.x;
  if (tid < st) {
    float s = 0.0f;
    for (int j = tid; j < a.K; j += st) s += a.win[j] * a.win[j];
    m1[tid] = s;
  }
  __syncthreads();
  if (tid < st * st) {
    const int py = tid / st, px = tid - py * st;
    inv[py][px] = rcp_approx(fmaf(m1[px], m1[py], kEps));
  }
  __syncthreads();
  const unsigned wq = (unsigned)a.width >> 2, ngroups = wq * (unsigned)a.height;  // < 2^29 (check_args)
  const unsigned step = gridDim.x * 256u, dy = step / wq, dq = step - dy * wq;
  unsigned g = blockIdx.x * 256u + tid;
  unsigned y = g / wq, q = g - y * wq;
  const float4 *acc4 = reinterpret_cast<const float4 *>(a.acc), *in4 = reinterpret_cast<const float4 *>(a.rgb);
  float4 *out4 = reinterpret_cast<float4 *>(a.out);
  for (; g < ngroups; g += step) {
    const float4 av = __ldg(acc4 + g);
    const float *ir = inv[y & (st - 1)];
    const unsigned xp = (4u * q) & (st - 1);
    const float l[4] = {av.x * ir[xp], av.y * ir[(xp + 1) & (st - 1)], av.z * ir[(xp + 2) & (st - 1)], av.w * ir[(xp + 3) & (st - 1)]};
    rgb_t r[4];
    if (kAb) {  // pub::with_luminance with rgb_to_lab(c).a/b already at hand
      const float4 ab0 = __ldg(in4 + 2 * (size_t)g), ab1 = __ldg(in4 + 2 * (size_t)g + 1);
      const float pa[4] = {ab0.x, ab0.z, ab1.x, ab1.z}, pb[4] = {ab0.y, ab0.w, ab1.y, ab1.w};
#pragma unroll
      for (int i = 0; i < 4; i++) r[i] = clip01(pub::lab_to_rgb(rgb_t{fmaxf(0.0f, fminf(1.0f, expf(l[i]))), pa[i], pb[i]}));
    } else {
      rgb_t c[4];
      unpack4(__ldg(in4 + 3 * (size_t)g), __ldg(in4 + 3 * (size_t)g + 1), __ldg(in4 + 3 * (size_t)g + 2), c);
#pragma unroll
      for (int i = 0; i < 4; i++) r[i] = pub::with_luminance(c[i], expf(l[i]));
    }
    if (kSplat) {  // see wiener_normalize_kernel: hand the bilateral stage rgb_to_lab(r) and its L as a plane
      float lum[4];
#pragma unroll
      for (int i = 0; i < 4; i++) r[i] = pub::rgb_to_lab(r[i]), lum[i] = fmaxf(0.0f, r[i].x);
      if (lum_aligned) {
        reinterpret_cast<float4 *>(a.lum_out)[g] = make_float4(lum[0], lum[1], lum[2], lum[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) a.lum_out[4 * (size_t)g + i] = lum[i];
      }
    }
    float4 o0, o1, o2;
    pack4(r, o0, o1, o2);
    out4[3 * (size_t)g] = o0, out4[3 * (size_t)g + 1] = o1, out4[3 * (size_t)g + 2] = o2;
    q += dq, y += dy;
    if (q >= wq) q -= wq, y++;
  }
}

struct LogLum {
  float eps;
  __device__ float operator()(rgb_t c) const { return logf(fmaxf(eps, pub::luminance(c))); }
};
__global__ void __launch_bounds__(256) loglum_kernel(const float *__restrict__ rgb, float *__restrict__ out, int64_t n, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = LogLum{eps}(rgb_t{__ldg(rgb + 3 * i), __ldg(rgb + 3 * i + 1), __ldg(rgb + 3 * i + 2)});
}

// the reference builds the windows with torch (float32): linspace, exp(-r^2/scale), divide by the L2 norm (window.h:18-43)
void make_window(int K, float *win) {
  const float half = K / 2.0f, scale = 0.3f * half * half;
  float norm2 = 0.0f;
  for (int i = 0; i < K; i++) {
    const float r = (-half + 0.5f) + (float)i;
    win[i] = expf(-(r * r) / scale);
    norm2 += win[i] * win[i];
  }
  const float norm = sqrtf(norm2);
  for (int i = 0; i < K; i++) win[i] /= norm;
  for (int i = K; i < 32; i++) win[i] = 0.0f;
}

// the shared-column kernel for K = 32, stride 8 (the configuration of the frame pipeline and the default of Wiener.process).  TDB_WIENER_SHARED=0 keeps
// wiener32_kernel (A/B runs).
bool use_shared_columns() {
  static const bool on = [] {
    const char *e = getenv("TDB_WIENER_SHARED");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}

int run_tiles_shared(const float *in, float *acc, int width, int height, int channels, const float *sigmas, float sigma_value,
                     cudaStream_t s) {
  shr::Args a{};
  a.in = in, a.acc = acc, a.sigmas = sigmas, a.sigma_value = sigma_value, a.width = width, a.height = height, a.channels = channels;
  a.n_tx = (width - 1 + shr::CARRY) / shr::ST + 1;
  const int n_ty = (height - 1 + shr::CARRY) / shr::ST + 1;
  a.steps_per_row = (a.n_tx + shr::TPS - 1) / shr::TPS;
  a.n_pairs = (n_ty + 1) / 2;
  a.total_steps = a.steps_per_row * a.n_pairs * channels;
  // window, its squares / 32, its transform and the four-lane twiddles depend on nothing but K: built once per process
  static const shr::Args tables = [] {
    shr::Args t{};
    make_window(32, t.win);
    for (int i = 0; i < 32; i++) t.w2[i] = t.win[i] * t.win[i] * (1.0f / 32.0f);
    for (int k = 0; k < 32; k++) {
      double re = 0.0, im = 0.0;
      for (int r = 0; r < 32; r++) {
        const double ang = -2.0 * 3.14159265358979323846 * k * r / 32.0;
        re += t.win[r] * cos(ang), im += t.win[r] * sin(ang);
      }
      t.what[k] = fft::cpx{(float)re, (float)im};
    }
    fft::make_quad_twiddles(t.tw);
    return t;
  }();
  memcpy(a.win, tables.win, sizeof a.win), memcpy(a.w2, tables.w2, sizeof a.w2);
  memcpy(a.what, tables.what, sizeof a.what), memcpy(a.tw, tables.tw, sizeof a.tw);
  static DeviceOnce attr;
  attr.run([&] {
    cudaFuncSetAttribute(shr::wiener32_shared_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, shr::kSmemBytes);
    cudaFuncSetAttribute(shr::wiener32_shared_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, shr::kSmemBytes);
    cudaFuncSetAttribute(shr::wiener32_shared_kernel_capped<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, shr::kSmemBytes);
    cudaFuncSetAttribute(shr::wiener32_shared_kernel_capped<104>, cudaFuncAttributeMaxDynamicSharedMemorySize, shr::kSmemBytes);
  });
  static const int ctas = [] {
    const char *e = getenv("TDB_WIENER_CTAS");
    return e && atoi(e) == 3 ? 3 : 2;  // measured at 4K: 0.2165 ms with two CTAs per SM, 0.2221 ms with three (the kernel is not latency-bound)
  }();
  // with another frame in flight on a second stream (tdb_set_concurrency_hint) the 104-register variant lets other kernels' CTAs
  // share the SM: measured at 4K with two lanes 10.50 -> 10.63 GP/s, although the kernel alone runs 0.219 -> 0.231 ms
  static const int forced_regs = [] {
    const char *e = getenv("TDB_WIENER_REGS");
    return e ? atoi(e) : -1;
  }();
  const int regs = forced_regs >= 0 ? forced_regs : (concurrent_lanes() > 1 ? 104 : 0);
  // CTAs per SM of the (persistent, statically split) grid while another frame is in flight: TDB_WIENER_LANE_CTAS=1 leaves half of
  // every SM to the other lane's kernels for the whole launch instead of only at its tail (A/B runs)
  static const int lane_ctas = [] {
    const char *e = getenv("TDB_WIENER_LANE_CTAS");
    return e && atoi(e) == 1 ? 1 : 0;
  }();
  const int per_sm = lane_ctas && concurrent_lanes() > 1 ? lane_ctas : ctas;
  const int grid = a.total_steps < per_sm * kNumSMs ? a.total_steps : per_sm * kNumSMs;
  if (regs == 96) shr::wiener32_shared_kernel_capped<96><<<grid, kThreads, shr::kSmemBytes, s>>>(a);
  else if (regs == 104) shr::wiener32_shared_kernel_capped<104><<<grid, kThreads, shr::kSmemBytes, s>>>(a);
  else if (ctas == 3) shr::wiener32_shared_kernel<3><<<grid, kThreads, shr::kSmemBytes, s>>>(a);
  else shr::wiener32_shared_kernel<2><<<grid, kThreads, shr::kSmemBytes, s>>>(a);
  return check_launch("wiener_tiles");
}

int run_tiles(const float *in, float *acc, int width, int height, int channels, int tile, int overlap, const float *sigmas,
              float sigma_value, cudaStream_t s, bool cleared = false) {
  WienerArgs a{};
  a.in = in, a.acc = acc, a.sigmas = sigmas, a.sigma_value = sigma_value;
  a.width = width, a.height = height, a.channels = channels;
  a.stride = tile / overlap;
  const int start = tile / a.stride;
  a.grid_h = (height + tile + a.stride - 1) / a.stride + start;
  a.grid_w = (width + tile + a.stride - 1) / a.stride + start;
  a.pairs_w = (a.grid_w + 1) / 2;
  a.njobs = (int64_t)a.grid_h * a.pairs_w * channels;
  make_window(tile, a.win);
  // scratch layout: [64 floats: job counters][accumulator][extra plane]; one memset clears counters and accumulator
  a.counters = reinterpret_cast<unsigned int *>(acc) - 64;
  if (!cleared) {
    const cudaError_t cleared_err = cudaMemsetAsync(a.counters, 0, ((size_t)width * height * channels + 64) * sizeof(float), s);
    if (cleared_err != cudaSuccess) {
      set_error("wiener_zero_accumulator: %s", cudaGetErrorString(cleared_err));
      return TDB_ECUDA;
    }
    if (int e = check_launch("wiener_zero_accumulator")) return e;
  }
  const int sub = 32 / tile;
  const int64_t warps_needed = (a.njobs + sub - 1) / sub;
  int64_t ctas = (warps_needed + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)kNumSMs * 8;  // persistent-style grid: a few CTAs per SM, each warp loops over tile pairs
  if (ctas > cap) ctas = cap;
  const size_t smem = (size_t)kWarps * sub * 2 * tile * (tile + 1) * sizeof(float);
  static DeviceOnce attr;
  attr.run([&] {
    cudaFuncSetAttribute(wiener_tile_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * 2 * 2 * 16 * 17 * 4);
  });
  if (tile == 32 && a.stride == shr::ST && use_shared_columns()) return run_tiles_shared(in, acc, width, height, channels, sigmas, sigma_value, s);
  if (tile == 32) {
    const size_t smem32 = (size_t)kWarps * 32 * 33 * sizeof(float2);
    static DeviceOnce attr32;
    attr32.run([&] {
      cudaFuncSetAttribute(wiener32_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
      cudaFuncSetAttribute(wiener32_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
      cudaFuncSetAttribute(wiener32_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
      cudaFuncSetAttribute(wiener32_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
      cudaFuncSetAttribute(wiener32_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
      cudaFuncSetAttribute(wiener32_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
    });
    // interior pairs: oy = (gy - shift) * stride in [0, height - 32], ox0 = (2 px - shift) * stride >= 0, ox0 + stride + 32 <= width
    const int st = a.stride, shift = 32 / st;
    a.gy_lo = shift, a.gy_hi = (height - 32) / st + shift;
    a.px_lo = (shift + 1) / 2, a.px_hi = ((width - 32 - st) / st + shift) / 2;
    if (width < 32 + st || height < 32) a.px_hi = a.px_lo - 1;
    while (a.px_hi >= a.px_lo && 2 * a.px_hi + 1 >= a.grid_w) a.px_hi--;
    const bool any = a.px_hi >= a.px_lo && a.gy_hi >= a.gy_lo;
    a.njobs_interior = any ? (int64_t)(a.px_hi - a.px_lo + 1) * (a.gy_hi - a.gy_lo + 1) * channels : 0;
    auto grid_for = [](int64_t jobs) {
      int64_t c = (jobs + kWarps - 1) / kWarps;  // persistent grid: two CTAs per SM
      return (int)(c > 2 * kNumSMs ? 2 * kNumSMs : (c < 1 ? 1 : c));
    };
    const int g = grid_for(a.njobs);
    if (channels == 1) {
      if (st == 8) wiener32_kernel<8, true><<<g, kThreads, smem32, s>>>(a);
      else if (st == 4) wiener32_kernel<4, true><<<g, kThreads, smem32, s>>>(a);
      else wiener32_kernel<16, true><<<g, kThreads, smem32, s>>>(a);
    } else {
      if (st == 8) wiener32_kernel<8, false><<<g, kThreads, smem32, s>>>(a);
      else if (st == 4) wiener32_kernel<4, false><<<g, kThreads, smem32, s>>>(a);
      else wiener32_kernel<16, false><<<g, kThreads, smem32, s>>>(a);
    }
    return check_launch("wiener_tiles");
  } else {
    wiener_tile_kernel<16><<<(int)ctas, kThreads, smem, s>>>(a);
  }
  return check_launch("wiener_tiles");
}

int check_args(int width, int height, int channels, int tile, int overlap) {
  if (width <= 0 || height <= 0) { set_error("Wiener: image dimensions must be positive"); return TDB_EINVAL; }
  if ((int64_t)width * height * channels >= (int64_t)1 << 31) { set_error("Wiener: image too large (2^31 samples)"); return TDB_EINVAL; }
  if (channels != 1 && channels != 3) { set_error("input channels must be 1 or 3, got %d", channels); return TDB_EINVAL; }
  if (tile != 16 && tile != 32) { set_error("tile_size must be 16 or 32, got %d", tile); return TDB_EINVAL; }
  if (overlap != 2 && overlap != 4 && overlap != 8) { set_error("overlap_factor must be 2, 4, or 8"); return TDB_EINVAL; }
  return TDB_OK;
}

}  // namespace
}  // namespace tdb

using namespace tdb;

extern "C" {

size_t tdb_wiener_scratch_bytes(int width, int height, int channels, int tile) {
  (void)tile;
  if (width <= 0 || height <= 0) return 0;
  // accumulator (C planes interleaved) + one extra plane for the log-luminance composite
  return ((size_t)width * height * (channels + 1)) * sizeof(float) + 256;  // + job counters of the K = 32 kernels
}

int tdb_wiener(const float *in, float *out, void *scratch, int width, int height, int channels, int tile, int overlap,
               const float *sigmas, tdb_stream_t stream) {
  TDB_REQUIRE(in && out && scratch && sigmas, "Wiener: null pointer");
  if (int e = check_args(width, height, channels, tile, overlap)) return e;
  cudaStream_t s = as_stream(stream);
  float *acc = wiener_scratch(scratch, width, height, channels).acc;
  if (int e = run_tiles(in, acc, width, height, channels, tile, overlap, sigmas, 0.0f, s)) return e;
  NormArgs n{};
  n.acc = acc, n.rgb = nullptr, n.out = out, n.width = width, n.height = height, n.channels = channels, n.K = tile, n.stride = tile / overlap;
  make_window(tile, n.win);
  const int64_t px = (int64_t)width * height;
  const int grid = (int)((px + 255) / 256 < kNumSMs * 16 ? (px + 255) / 256 : kNumSMs * 16);
  wiener_normalize_kernel<false, false><<<grid, 256, 0, s>>>(n);
  return check_launch("wiener_normalize");
}

static int run_log_luminance(const float *rgb, float *out, void *scratch, int width, int height, int tile, int overlap, float noise,
                             float eps, int prepared, void *bilateral_scratch, float sigma_s, float sigma_r, cudaStream_t s) {
  const WienerScratch ws = wiener_scratch(scratch, width, height, 1);
  const int64_t px = (int64_t)width * height;
  const int grid = (int)((px + 255) / 256 < kNumSMs * 16 ? (px + 255) / 256 : kNumSMs * 16);
  if (!prepared) {
    loglum_kernel<<<grid, 256, 0, s>>>(rgb, ws.lum, px, eps);
    if (int e = check_launch("wiener_log_luminance")) return e;
  }
  NormArgs n{};
  bil::GridDims g{};
  if (bilateral_scratch) {
    g = bil::grid_dims(width, height, sigma_s, sigma_r);
    n.lum_out = bilateral_lum_plane(bilateral_scratch, g);
  }
  if (int e = run_tiles(ws.lum, ws.acc, width, height, 1, tile, overlap, nullptr, noise, s, prepared != 0)) return e;
  n.acc = ws.acc, n.rgb = rgb, n.out = out, n.width = width, n.height = height, n.channels = 1, n.K = tile, n.stride = tile / overlap;
  make_window(tile, n.win);
  const bool ab = prepared == 2;
  auto aligned = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  // four pixels per thread (TDB_WIENER_NORM4=0 keeps the scalar kernel for A/B runs)
  static const bool norm4 = [] {
    const char *e = getenv("TDB_WIENER_NORM4");
    return e ? atoi(e) != 0 : true;
  }();
  const bool vec = norm4 && width % 4 == 0 && aligned(n.acc) && aligned(rgb) && aligned(out);
  const int grid4 = (int)((px / 4 + 255) / 256 < kNumSMs * 16 ? (px / 4 + 255) / 256 : kNumSMs * 16);
  if (bilateral_scratch) {
    if (vec) {
      if (ab) wiener_normalize_lum4_kernel<true, true><<<grid4, 256, 0, s>>>(n, aligned(n.lum_out));
      else wiener_normalize_lum4_kernel<true, false><<<grid4, 256, 0, s>>>(n, aligned(n.lum_out));
    } else {
      if (ab) wiener_normalize_kernel<true, true, true><<<grid, 256, 0, s>>>(n);
      else wiener_normalize_kernel<true, true><<<grid, 256, 0, s>>>(n);
    }
    if (int e = check_launch("wiener_normalize_lum")) return e;
    return bilateral_build_grid(bilateral_scratch, n.lum_out, width, height, g, sigma_s, sigma_r, s);
  }
  if (vec) {
    if (ab) wiener_normalize_lum4_kernel<false, true><<<grid4, 256, 0, s>>>(n, false);
    else wiener_normalize_lum4_kernel<false, false><<<grid4, 256, 0, s>>>(n, false);
  } else {
    if (ab) wiener_normalize_kernel<true, false, true><<<grid, 256, 0, s>>>(n);
    else wiener_normalize_kernel<true, false><<<grid, 256, 0, s>>>(n);
  }
  return check_launch("wiener_normalize");
}

int tdb_wiener_log_luminance(const float *rgb, float *out, void *scratch, int width, int height, int tile, int overlap, float noise,
                             float eps, tdb_stream_t stream) {
  TDB_REQUIRE(rgb && out && scratch, "Wiener: null pointer");
  TDB_REQUIRE(eps > 0.0f, "Epsilon must be positive");
  if (int e = check_args(width, height, 1, tile, overlap)) return e;
  return run_log_luminance(rgb, out, scratch, width, height, tile, overlap, noise, eps, 0, nullptr, 0.0f, 0.0f, as_stream(stream));
}

int tdb_wiener_log_luminance_fused(const float *rgb, float *out, void *scratch, int width, int height, int tile, int overlap, float noise,
                                   float eps, int prepared, void *bilateral_scratch, float sigma_s, float sigma_r, tdb_stream_t stream) {
  TDB_REQUIRE(rgb && out && scratch, "Wiener: null pointer");
  TDB_REQUIRE(eps > 0.0f, "Epsilon must be positive");
  TDB_REQUIRE(!bilateral_scratch || (sigma_r > 0.0f && sigma_s > 0.0f), "Bilateral: invalid sigmas");
  if (int e = check_args(width, height, 1, tile, overlap)) return e;
  TDB_REQUIRE(prepared >= 0 && prepared <= 2, "Wiener: prepared must be 0, 1 or 2");
  return run_log_luminance(rgb, out, scratch, width, height, tile, overlap, noise, eps, prepared, bilateral_scratch, sigma_s, sigma_r,
                           as_stream(stream));
}

}  // extern "C"
