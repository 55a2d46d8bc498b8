// Register-resident 32-point (and 16-point) complex FFTs with compile-time indices and FMA-form butterflies.
//
// Both directions are decimation-in-time, so the twiddle multiplies the `b` input BEFORE the add/sub and the whole
// butterfly fuses into six FFMAs:  with w = c -/+ i*s and tn = s/c,
//     u = br +/- tn*bi,  v = bi -/+ tn*br,  a' = a + c*(u, v),  b' = a - c*(u, v).
// Twiddles 1 and -/+i cost four FADDs.  Per 32-point transform: 46 trivial + 34 FMA butterflies = 388 instructions
// (the textbook radix-2 DIF with separate twiddle multiplies needs 456).
//   fft_fwd : natural order in, BIT-REVERSED order out (register p holds frequency brev(p)), W = exp(-2*pi*i/N)
//   fft_inv : bit-reversed order in, natural order out, unscaled,                            W = exp(+2*pi*i/N)
// so fwd -> (pointwise work on bit-reversed bins) -> inv needs no reordering pass at all.
// The functions are __host__ __device__ so that tests/test_fft32.py can check them on the CPU.
#pragma once

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace tdb {
namespace fft {

// cos/sin(2*pi*k/32), k = 0..15, as literals so that unrolled butterflies fold them into immediates
__host__ __device__ __forceinline__ constexpr float cos32(int k) {
  switch (k) {
    case 0: return 1.0f; case 1: return 0.98078528040323043f; case 2: return 0.92387953251128674f; case 3: return 0.83146961230254524f;
    case 4: return 0.70710678118654757f; case 5: return 0.55557023301960218f; case 6: return 0.38268343236508978f;
    case 7: return 0.19509032201612825f; case 8: return 0.0f; case 9: return -0.19509032201612825f; case 10: return -0.38268343236508978f;
    case 11: return -0.55557023301960218f; case 12: return -0.70710678118654757f; case 13: return -0.83146961230254524f;
    case 14: return -0.92387953251128674f; default: return -0.98078528040323043f;
  }
}
__host__ __device__ __forceinline__ constexpr float sin32(int k) { return k < 8 ? cos32(8 - k) : cos32(k - 8); }
// tan(2*pi*k/32), k != 8
__host__ __device__ __forceinline__ constexpr float tan32(int k) {
  switch (k) {
    case 0: return 0.0f; case 1: return 0.19891236737965800f; case 2: return 0.41421356237309503f; case 3: return 0.66817863791929888f;
    case 4: return 1.0f; case 5: return 1.4966057626654890f; case 6: return 2.4142135623730949f; case 7: return 5.0273394921258481f;
    case 9: return -5.0273394921258481f; case 10: return -2.4142135623730949f; case 11: return -1.4966057626654890f;
    case 12: return -1.0f; case 13: return -0.66817863791929888f; case 14: return -0.41421356237309503f;
    default: return -0.19891236737965800f;
  }
}

template <int N>
__host__ __device__ __forceinline__ constexpr int brev(int v) {
  int r = 0;
  for (int b = 1; b < N; b <<= 1) r = (r << 1) | ((v & b) ? 1 : 0);
  return r;
}

#ifdef __CUDA_ARCH__
#define TDB_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define TDB_FMA(a, b, c) ((a) * (b) + (c))
#endif

// a' = a + b*w, b' = a - b*w with w = exp(-/+ 2*pi*i*t/32); kInv selects the sign
template <bool kInv, int t>
__host__ __device__ __forceinline__ void butterfly(float &ar, float &ai, float &br, float &bi) {
  if (t == 0) {
    const float xr = br, xi = bi;
    br = ar - xr, bi = ai - xi;
    ar = ar + xr, ai = ai + xi;
  } else if (t == 8) {  // w = -i (forward) / +i (inverse)
    const float xr = kInv ? -bi : bi, xi = kInv ? br : -br;
    br = ar - xr, bi = ai - xi;
    ar = ar + xr, ai = ai + xi;
  } else {
    constexpr float c = cos32(t), tn = kInv ? tan32(t) : -tan32(t);
    // b*w = c * ((br - tn*bi) + i (bi + tn*br))   [w = c (1 + i tn)]
    const float u = TDB_FMA(-tn, bi, br), v = TDB_FMA(tn, br, bi);
    br = TDB_FMA(-c, u, ar), bi = TDB_FMA(-c, v, ai);
    ar = TDB_FMA(c, u, ar), ai = TDB_FMA(c, v, ai);
  }
}

namespace detail {
// compile-time loops (plain `#pragma unroll` loops would also fold, but the twiddle index must be a template argument)
template <int N, int H, int G, int J>
struct FwdInner {
  // group G of the stage with half-span H: twiddle index brev(G) * H on the N-circle, scaled to the 32-circle
  __host__ __device__ static __forceinline__ void run(float (&re)[N], float (&im)[N]) {
    constexpr int groups = N / (2 * H);
    constexpr int k = brev<groups>(G);
    constexpr int t = k * H * (32 / N);
    constexpr int a = G * 2 * H + J, b = a + H;
    butterfly<false, t>(re[a], im[a], re[b], im[b]);
    if constexpr (J + 1 < H) FwdInner<N, H, G, J + 1>::run(re, im);
    else if constexpr (G + 1 < groups) FwdInner<N, H, G + 1, 0>::run(re, im);
  }
};
template <int N, int H>
struct FwdStage {
  __host__ __device__ static __forceinline__ void run(float (&re)[N], float (&im)[N]) {
    FwdInner<N, H, 0, 0>::run(re, im);
    if constexpr (H > 1) FwdStage<N, H / 2>::run(re, im);
  }
};
template <int N, int M, int B, int K>
struct InvInner {
  __host__ __device__ static __forceinline__ void run(float (&re)[N], float (&im)[N]) {
    constexpr int t = K * (16 / M);
    constexpr int a = B + K, b = a + M;
    butterfly<true, t>(re[a], im[a], re[b], im[b]);
    if constexpr (K + 1 < M) InvInner<N, M, B, K + 1>::run(re, im);
    else if constexpr (B + 2 * M < N) InvInner<N, M, B + 2 * M, 0>::run(re, im);
  }
};
template <int N, int M>
struct InvStage {
  __host__ __device__ static __forceinline__ void run(float (&re)[N], float (&im)[N]) {
    InvInner<N, M, 0, 0>::run(re, im);
    if constexpr (2 * M < N) InvStage<N, 2 * M>::run(re, im);
  }
};
}  // namespace detail

template <int N>
__host__ __device__ __forceinline__ void fft_fwd(float (&re)[N], float (&im)[N]) {
  detail::FwdStage<N, N / 2>::run(re, im);
}
template <int N>
__host__ __device__ __forceinline__ void fft_inv(float (&re)[N], float (&im)[N]) {
  detail::InvStage<N, 1>::run(re, im);
}

}  // namespace fft
}  // namespace tdb
