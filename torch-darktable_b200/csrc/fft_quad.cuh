// A 32-point complex FFT spread over FOUR lanes (8 points each): 32 = 4 x 8 decimation.
//
//   time side      : lane j (0..3) holds x[4m + j] in register m = 0..7
//   frequency side : lane j holds X[k] with k = quad_freq(p, j) = brev8(p) + 8 * brev2(j) in register p
//
// forward : 8-point register FFT over m (fft32.cuh) -> twiddle W32^(j*k2) -> 4-point DFT across the lanes (two exchange stages)
// inverse : the same steps backwards with conjugated twiddles, unscaled.
// The exchange stages are written as per-lane functions of (own value, partner's value), so the device code only adds the two
// __shfl_xor_sync per value and tests/test_fft32.py can run the very same functions on four emulated lanes on the CPU.
// Used by the shared-column Wiener kernel (wiener.cu), whose column transforms have only eight new columns per warp and step:
// four lanes per column keep all 32 lanes busy.
#pragma once

#include "fft32.cuh"

namespace tdb {
namespace fft {

__host__ __device__ __forceinline__ constexpr int quad_freq(int p, int j) { return brev<8>(p) + 8 * brev<4>(j); }
// position (register p, lane j) that holds frequency k
__host__ __device__ __forceinline__ constexpr int quad_reg_of(int k) { return brev<8>(k & 7); }
__host__ __device__ __forceinline__ constexpr int quad_lane_of(int k) { return brev<4>(k >> 3); }

struct cpx {
  float x, y;
};

// twiddle table entry (j, p): exp(-2*pi*i * j * brev8(p) / 32); filled on the host (make_quad_twiddles)
inline void make_quad_twiddles(cpx (&tw)[32]) {
  for (int j = 0; j < 4; j++)
    for (int p = 0; p < 8; p++) {
      const double a = -2.0 * 3.14159265358979323846 * j * brev<8>(p) / 32.0;
      tw[j * 8 + p] = cpx{(float)__builtin_cos(a), (float)__builtin_sin(a)};
    }
}

// ---- forward ---------------------------------------------------------------------------------------------------------
// local part: 8-point FFT over m, then the twiddle of this lane.  tw = this lane's eight table entries.
template <class TW>
__host__ __device__ __forceinline__ void quad_fwd_local(float (&re)[8], float (&im)[8], const TW &tw) {
  fft_fwd<8>(re, im);  // register p holds k2 = brev8(p)
#pragma unroll
  for (int p = 1; p < 8; p++) {  // p = 0: k2 = 0, twiddle 1
    const float c = tw[p].x, s = tw[p].y, r = re[p], i = im[p];
    re[p] = r * c - i * s, im[p] = r * s + i * c;
  }
}
// exchange A (partner = lane ^ 2): lanes 0,1 keep the sum, lanes 2,3 the difference (partner - own); lane 3 then turns its value
// by -i (the W4^1 of the odd branch)
__host__ __device__ __forceinline__ void quad_fwd_a(float &re, float &im, float pr, float pi, int j) {
  const float s = (j & 2) ? -1.0f : 1.0f;
  const float r = pr + s * re, i = pi + s * im;
  re = (j == 3) ? i : r;
  im = (j == 3) ? -r : i;
}
// exchange B (partner = lane ^ 1): even lanes keep the sum, odd lanes partner - own.  Lane j ends with k1 = brev2(j).
__host__ __device__ __forceinline__ void quad_fwd_b(float &re, float &im, float pr, float pi, int j) {
  const float s = (j & 1) ? -1.0f : 1.0f;
  re = pr + s * re, im = pi + s * im;
}

// ---- inverse ---------------------------------------------------------------------------------------------------------
// exchange B' (partner = lane ^ 1): even lanes sum, odd lanes partner - own; lane 3 then turns its value by +i
__host__ __device__ __forceinline__ void quad_inv_b(float &re, float &im, float pr, float pi, int j) {
  const float s = (j & 1) ? -1.0f : 1.0f;
  const float r = pr + s * re, i = pi + s * im;
  re = (j == 3) ? -i : r;
  im = (j == 3) ? r : i;
}
// exchange A' (partner = lane ^ 2): lanes 0,1 sum, lanes 2,3 partner - own.  Lane j ends with time residue j.
__host__ __device__ __forceinline__ void quad_inv_a(float &re, float &im, float pr, float pi, int j) {
  const float s = (j & 2) ? -1.0f : 1.0f;
  re = pr + s * re, im = pi + s * im;
}
// local part: conjugated twiddle, then the inverse 8-point FFT (register m = time index 4m + j)
template <class TW>
__host__ __device__ __forceinline__ void quad_inv_local(float (&re)[8], float (&im)[8], const TW &tw) {
#pragma unroll
  for (int p = 1; p < 8; p++) {
    const float c = tw[p].x, s = -tw[p].y, r = re[p], i = im[p];
    re[p] = r * c - i * s, im[p] = r * s + i * c;
  }
  fft_inv<8>(re, im);
}

}  // namespace fft
}  // namespace tdb
