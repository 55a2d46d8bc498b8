// Shared-memory CFA patch staging for the demosaic stencils.
//
// A CfaSource is either a float32 CFA plane or a 12-bit packed frame that is unpacked (+ black level + white balance)
// while the patch is staged, so the demosaic kernels never see an intermediate float CFA in HBM (north_star's
// "unpack fused with black-level subtraction and white balance").
#pragma once

#include "tdb_common.cuh"

namespace tdb {

struct CfaSource {
  const float *cfa;        // float plane, or null
  const uint8_t *packed;   // packed 12-bit frame, or null
  int ids;                 // IDS byte layout
  int apply;               // 0: v/4095 ; 1: v/4095 - black ; 2: clamp((v/4095 - black) * gain, 0, 1)
  float black;
  float gain[2][2];        // per CFA phase [row & 1][col & 1] (filled on the device from gains_dev when apply == 2)
  const float *gains_dev;  // device float[3] (R, G, B) or null
};

__device__ __forceinline__ void resolve_gains(CfaSource &s, uint32_t filters) {
  if (s.packed && s.apply == 2) {
    const float gr = __ldg(s.gains_dev), gg = __ldg(s.gains_dev + 1), gb = __ldg(s.gains_dev + 2);
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int col = fc(r, c, filters);
        s.gain[r][c] = col == 0 ? gr : (col == 2 ? gb : gg);
      }
  }
}

__device__ __forceinline__ float finish_sample(const CfaSource &s, uint32_t raw, int y, int x) {
  float v = __fmul_rn((float)raw, 1.0f / 4095.0f);  // no FMA contraction with the black level: fused = decode12_float(x) - black
  if (s.apply) {
    v = __fsub_rn(v, s.black);
    // selects on compile-time indices: a dynamically indexed member would push the whole struct into local memory
    if (s.apply == 2) v = clip01(v * ((y & 1) ? ((x & 1) ? s.gain[1][1] : s.gain[1][0]) : ((x & 1) ? s.gain[0][1] : s.gain[0][0])));
  }
  return v;
}

// in-bounds sample
__device__ __forceinline__ float cfa_at(const CfaSource &s, int x, int y, int width) {
  if (s.cfa) return __ldg(s.cfa + (int64_t)y * width + x);
  const int64_t i = (int64_t)y * width + x;
  const uint32_t raw = s.ids ? packed_sample<true>(s.packed, i) : packed_sample<false>(s.packed, i);
  return finish_sample(s, raw, y, x);
}

enum class Oob { kZero, kClamp };

// Stage the (pw x ph) patch whose top-left image coordinate is (px0, py0) into smem[ph][stride].
// px0 must be even so that packed pairs never straddle the patch edge.  All threads of the CTA participate.
template <Oob kOob, bool kClampNegative>
__device__ __forceinline__ void stage_patch(float *smem, int stride, int px0, int py0, int pw, int ph, const CfaSource &s,
                                            int width, int height) {
  const int nthreads = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (s.cfa) {
    for (int i = tid; i < pw * ph; i += nthreads) {
      const int ly = i / pw, lx = i - ly * pw;
      int x = px0 + lx, y = py0 + ly;
      float v;
      if (kOob == Oob::kClamp) {
        x = max(0, min(x, width - 1)), y = max(0, min(y, height - 1));
        v = __ldg(s.cfa + (int64_t)y * width + x);
      } else {
        v = (x >= 0 && y >= 0 && x < width && y < height) ? __ldg(s.cfa + (int64_t)y * width + x) : 0.0f;
      }
      smem[ly * stride + lx] = kClampNegative ? fmaxf(v, 0.0f) : v;
    }
  } else {
    // one packed pair (3 bytes -> 2 samples) per step
    const int pairs_w = pw >> 1;
    for (int i = tid; i < pairs_w * ph; i += nthreads) {
      const int ly = i / pairs_w, lp = i - ly * pairs_w;
      const int lx = lp * 2;
      int x = px0 + lx, y = py0 + ly;
      float v0 = 0.0f, v1 = 0.0f;
      const bool row_in = (y >= 0 && y < height);
      if (kOob == Oob::kClamp) {
        y = max(0, min(y, height - 1));
        const int xa = max(0, min(x, width - 1)), xb = max(0, min(x + 1, width - 1));
        v0 = cfa_at(s, xa, y, width), v1 = cfa_at(s, xb, y, width);
      } else if (row_in && x >= 0 && x + 1 < width) {
        // the pair's three bytes out of the one or two aligned 32-bit words that hold them (one or two 4-byte loads that neighbouring
        // lanes share through L1, instead of three 1-byte loads); the last pairs of the frame keep the byte loads, so nothing
        // beyond the frame's last byte is read
        const int64_t off = (((int64_t)y * width + x) >> 1) * 3;
        uint32_t w;
        if ((reinterpret_cast<uintptr_t>(s.packed) & 3) == 0 && off + 8 <= (int64_t)width * height * 3 / 2) {
          const uint32_t *wp = reinterpret_cast<const uint32_t *>(s.packed + (off & ~(int64_t)3));
          const uint32_t sh = (uint32_t)(off & 3) * 8u;
          const uint32_t w0 = __ldg(wp), w1 = sh > 8u ? __ldg(wp + 1) : 0u;
          w = __funnelshift_r(w0, w1, sh);
        } else {
          const uint8_t *b = s.packed + off;
          w = (uint32_t)__ldg(b) | ((uint32_t)__ldg(b + 1) << 8) | ((uint32_t)__ldg(b + 2) << 16);
        }
        uint32_t p0, p1;
        if (s.ids) unpack_pair<true>(w, p0, p1); else unpack_pair<false>(w, p0, p1);
        v0 = finish_sample(s, p0, y, x), v1 = finish_sample(s, p1, y, x + 1);
      }
      smem[ly * stride + lx] = kClampNegative ? fmaxf(v0, 0.0f) : v0;
      smem[ly * stride + lx + 1] = kClampNegative ? fmaxf(v1, 0.0f) : v1;
    }
  }
}

// Write a (tw x th) RGB tile held in smem (row stride `sstride` floats, 3 floats per pixel) to the image.  All threads of the CTA
// call it, after the barrier that completes the tile.  When the destination rows are 16-byte aligned (width % 4 == 0 and
// x0 % 4 == 0) every row leaves as ONE bulk asynchronous copy (cp.async.bulk shared -> global, the 1-D form of the TMA engine),
// issued by the lanes of warp 0; otherwise with 128-bit or scalar stores.
__device__ __forceinline__ void store_rgb_tile(const float *smem, int sstride, float *__restrict__ rgb, int x0, int y0, int tw,
                                               int th, int width, int height) {
  const int nthreads = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int vw = min(tw, width - x0), vh = min(th, height - y0);
  if (vw <= 0 || vh <= 0) return;
  const bool vec = ((width & 3) == 0) && ((x0 & 3) == 0) && ((vw & 3) == 0) && ((sstride & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(rgb) & 15) == 0);
  if (vec && (smem_addr(smem) & 15) == 0) {
    bulk_store_fence();   // this thread's writes of the tile -> visible to the asynchronous proxy
    __syncthreads();
    if (tid < 32) {
      for (int ly = tid; ly < vh; ly += 32)
        bulk_copy_s2g(rgb + 3 * ((int64_t)(y0 + ly) * width + x0), smem + ly * sstride, (uint32_t)(vw * 3 * sizeof(float)));
      bulk_store_commit_and_drain();
    }
  } else if (vec) {
    const int qrow = vw * 3 / 4;  // float4 per row
    for (int i = tid; i < qrow * vh; i += nthreads) {
      const int ly = i / qrow, q = i - ly * qrow;
      const float4 v = *reinterpret_cast<const float4 *>(smem + ly * sstride + 4 * q);
      st_stream(reinterpret_cast<float4 *>(rgb + 3 * ((int64_t)(y0 + ly) * width + x0)) + q, v);
    }
  } else {
    const int frow = vw * 3;
    for (int i = tid; i < frow * vh; i += nthreads) {
      const int ly = i / frow, f = i - ly * frow;
      rgb[3 * ((int64_t)(y0 + ly) * width + x0) + f] = smem[ly * sstride + f];
    }
  }
}

}  // namespace tdb
