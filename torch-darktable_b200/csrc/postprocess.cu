// Demosaic post-process: N x 3x3-median colour smoothing, global and local green equilibration.
//
// The reference (csrc/debayer/postprocess.cu:311-390) does copy_ + one kernel per smoothing pass + reduce kernel +
// torch sum + two blocking .item() reads + apply kernel + clone: about 156 B/px of HBM traffic and a host sync in the
// middle of the stream.  Here all smoothing passes run inside one shared-memory tile (halo = number of passes), the
// green sums are taken in the same kernel (smoothing never changes G apart from the >= 0 clamp), the ratio stays on the
// device, and one second kernel applies the global ratio and the local equilibration: 24 B/px per kernel, no sync.
#include "tdb_common.cuh"

namespace tdb {
namespace {

constexpr int T = 32;
constexpr int kThreads = 256;
constexpr int kMaxFusedPasses = 6;

__device__ __forceinline__ void cas(float &a, float &b) {
  const float x = a;
  const bool c = a > b;
  a = c ? b : a;
  b = c ? x : b;
}
// same exchange sequence as the reference's 19-exchange network (csrc/reduction.h:93-116)
__device__ __forceinline__ float median9(float s0, float s1, float s2, float s3, float s4, float s5, float s6, float s7, float s8) {
  cas(s1, s2); cas(s4, s5); cas(s7, s8);
  cas(s0, s1); cas(s3, s4); cas(s6, s7);
  cas(s1, s2); cas(s4, s5); cas(s7, s8);
  cas(s0, s3); cas(s5, s8); cas(s4, s7);
  cas(s3, s6); cas(s1, s4); cas(s2, s5);
  cas(s4, s7); cas(s4, s2); cas(s6, s4);
  cas(s4, s2);
  return s4;
}

struct Header {          // first bytes of the scratch buffer
  float sum1, sum2;      // G1 / G2 sums
  float ratio;           // sum2 / sum1 (or 1)
  float pad;
};

// smoothing passes + (optionally) per-CTA green sums of the result
__global__ void __launch_bounds__(kThreads) smooth_kernel(const float *__restrict__ in, float *__restrict__ out, int width, int height,
                                                          uint32_t filters, int passes, int want_sums, int clamp_sums,
                                                          float *__restrict__ partials) {
  extern __shared__ __align__(16) float sm[];
  const int halo = passes;
  const int PW = T + 2 * halo;
  float *buf0 = sm, *buf1 = sm + PW * PW * 3;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * T, y0 = blockIdx.y * T;
  // stage: zero outside the image (postprocess.cu:58-59)
  for (int i = tid; i < PW * PW; i += kThreads) {
    const int ly = i / PW, lx = i - ly * PW;
    const int x = x0 - halo + lx, y = y0 - halo + ly;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    if (x >= 0 && y >= 0 && x < width && y < height) {
      const float *p = in + 3 * ((int64_t)y * width + x);
      r = __ldg(p), g = __ldg(p + 1), b = __ldg(p + 2);
    }
    buf0[3 * i] = r, buf0[3 * i + 1] = g, buf0[3 * i + 2] = b;
  }
  __syncthreads();
  float *src = buf0, *dst = buf1;
  for (int pass = 0; pass < passes; pass++) {
    const int m = pass + 1;  // the valid region shrinks by one ring per pass
    const int N = PW - 2 * m;
    for (int i = tid; i < N * N; i += kThreads) {
      const int ly = m + i / N, lx = m + i % N;
      const int x = x0 - halo + lx, y = y0 - halo + ly;
      float r = 0.0f, g = 0.0f, b = 0.0f;
      if (x >= 0 && y >= 0 && x < width && y < height) {  // outside stays zero for the next pass
        const float *c = src + 3 * (ly * PW + lx);
        const int R = 3 * PW;
#define DR(o) (c[(o)] - c[(o) + 1])
#define DB(o) (c[(o) + 2] - c[(o) + 1])
        const float mr = median9(DR(-R - 3), DR(-R), DR(-R + 3), DR(-3), DR(0), DR(3), DR(R - 3), DR(R), DR(R + 3));
        const float mb = median9(DB(-R - 3), DB(-R), DB(-R + 3), DB(-3), DB(0), DB(3), DB(R - 3), DB(R), DB(R + 3));
#undef DR
#undef DB
        g = c[1];
        r = fmaxf(mr + g, 0.0f), b = fmaxf(mb + g, 0.0f), g = fmaxf(g, 0.0f);
      }
      float *d = dst + 3 * (ly * PW + lx);
      d[0] = r, d[1] = g, d[2] = b;
    }
    __syncthreads();
    float *t = src; src = dst; dst = t;
  }
  // write the tile + green sums over the even-cropped image (postprocess.cu:195-203)
  float s1 = 0.0f, s2 = 0.0f;
  const int we = width & ~1, he = height & ~1;
  for (int i = tid; i < T * T; i += kThreads) {
    const int ly = i / T, lx = i - ly * T;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= width || y >= height) continue;
    const float *c = src + 3 * ((ly + halo) * PW + lx + halo);
    if (want_sums && x < we && y < he && fc(y & 1, x & 1, filters) == 1) {
      const float g = clamp_sums ? fmaxf(c[1], 0.0f) : c[1];
      if (y & 1) s2 += g; else s1 += g;
    }
    if (out) {
      float *o = out + 3 * ((int64_t)y * width + x);
      o[0] = c[0], o[1] = c[1], o[2] = c[2];
    }
  }
  if (want_sums) {
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
                                                            const Header *__restrict__ hdr) {
  constexpr int PW = T + 4;
  __shared__ float gsm[PW * PW];
  const float ratio = do_global ? hdr->ratio : 1.0f;
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
  return align_up(sizeof(Header), 256) + align_up(nblk * 2 * sizeof(float), 256) + 2 * (size_t)width * height * 3 * sizeof(float);
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
  float *img_a = reinterpret_cast<float *>(base + align_up(sizeof(Header), 256) + align_up(nblk * 2 * sizeof(float), 256));
  float *img_b = img_a + (size_t)width * height * 3;
  dim3 grid(div_up(width, T), div_up(height, T));
  const bool eq = green_eq_local || green_eq_global;

  static bool attr = false;
  if (!attr) {
    const int pw = T + 2 * kMaxFusedPasses;
    cudaFuncSetAttribute(smooth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * pw * pw * 3 * sizeof(float)));
    attr = true;
  }
  const float *cur = in;
  int remaining = passes;
  bool sums_done = false;
  // all passes fused in one launch when they fit (the pipeline default is 3); longer chains go in chunks
  while (remaining > 0 || (green_eq_global && !sums_done)) {
    const int chunk = remaining < kMaxFusedPasses ? remaining : kMaxFusedPasses;
    const bool last = (remaining - chunk) == 0;
    float *dst = nullptr;
    if (chunk > 0) dst = (last && !eq) ? out : (cur == img_a ? img_b : img_a);
    const int want_sums = last && green_eq_global;
    const int pw = T + 2 * chunk;
    smooth_kernel<<<grid, kThreads, 2 * pw * pw * 3 * sizeof(float), s>>>(cur, dst, width, height, filters, chunk, want_sums,
                                                                          passes > 0, partials);
    if (int e = check_launch("color_smoothing")) return e;
    if (dst) cur = dst;
    remaining -= chunk;
    if (want_sums) sums_done = true;
  }
  if (green_eq_global) {
    ratio_kernel<<<1, 1024, 0, s>>>(partials, (int)nblk, hdr);
    if (int e = check_launch("green_eq_ratio")) return e;
  }
  if (eq) {
    green_eq_kernel<<<grid, kThreads, 0, s>>>(cur, out, width, height, filters, green_eq_global, green_eq_local,
                                              (float)(green_eq_threshold / 100.), hdr);
    return check_launch("green_equilibration");
  }
  if (passes == 0) {  // nothing to do: the reference still returns a copy
    cudaMemcpyAsync(out, in, (size_t)width * height * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s);
    check_launch("postprocess_copy");
  }
  return TDB_OK;
}

}  // extern "C"
