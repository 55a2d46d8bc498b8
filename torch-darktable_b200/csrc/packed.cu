// 12-bit packed Bayer codec and the fused ingest (unpack + black level + white balance).
//
// Replaces the reference's byte-per-thread kernels (csrc/packed.cu:34-155: one thread per pixel pair, three
// 1-byte loads, launched on the legacy default stream and followed by cudaDeviceSynchronize).
//
// Layout of the work: one thread owns 16 pixels = 24 packed bytes.  24 B groups are 8-byte aligned from the start
// of the buffer, so the thread issues three 64-bit loads; lane pairs then swap half of their results with
// __shfl_xor so that every store instruction writes whole 32-byte sectors (a lone thread would otherwise write
// 16 B out of each sector per instruction).  HBM-bound: 1.5 B in + 4 B out per pixel (f32), 1.5 + 2 (f16/u16).
#include "tdb_common.cuh"

namespace tdb {
namespace {

constexpr int kThreads = 256;

struct WbParams {   // per 2x2 phase gain, [row&1][col&1]; apply == 0 -> plain decode
  float g[2][2];
  float black;
  int apply;        // 0: v*scale ; 1: (v*scale - black) ; 2: clamp((v*scale - black) * gain, 0, 1)
};

template <bool kIds>
__device__ __forceinline__ void decode16(const uint2 a, const uint2 b, const uint2 c, uint32_t (&px)[16]) {
  // 24 bytes = 8 triples; assemble each triple into the low 24 bits of a word
  const uint32_t w[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int bit = 24 * t;           // bit offset inside the 192-bit group
    const int wi = bit >> 5, sh = bit & 31;
    uint32_t v = w[wi] >> sh;
    if (sh > 8) v |= w[wi + 1] << (32 - sh);
    unpack_pair<kIds>(v & 0xffffffu, px[2 * t], px[2 * t + 1]);
  }
}

// ---- decode to f32 (optionally fused with black level + white balance) -------------------------------------
template <bool kIds, bool kFused>
__global__ void __launch_bounds__(kThreads) decode12_f32_kernel(const uint8_t *__restrict__ in, float *__restrict__ out,
                                                                int64_t ngroups, int64_t npairs, float scale, int width,
                                                                WbParams wb, const float *__restrict__ gains_dev,
                                                                uint32_t filters) {
  if (kFused && wb.apply == 2) {  // gains live on the device: no host read-back (the reference does 3 x .item())
    const float gr = __ldg(gains_dev), gg = __ldg(gains_dev + 1), gb = __ldg(gains_dev + 2);
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int col = fc(r, c, filters);
        wb.g[r][c] = col == 0 ? gr : (col == 2 ? gb : gg);
      }
  }
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < ngroups; g += stride) {
    const uint2 *src = reinterpret_cast<const uint2 *>(in + g * 24);
    const uint2 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
    uint32_t px[16];
    decode16<kIds>(a, b, c, px);
    float f[16];
    if (kFused && wb.apply) {
      const int64_t i0 = g * 16;
      int y = (int)(i0 / width), x = (int)(i0 - (int64_t)y * width);
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        // two roundings, not one FMA: the fused result must equal decode12_float(x) - black of the stage-by-stage path bit for bit
        float v0 = __fsub_rn(__fmul_rn((float)px[k], scale), wb.black), v1 = __fsub_rn(__fmul_rn((float)px[k + 1], scale), wb.black);
        if (wb.apply == 2) {
          v0 = clip01(v0 * wb.g[y & 1][0]);  // x is even for the first sample of a pair (width is even)
          v1 = clip01(v1 * wb.g[y & 1][1]);
        }
        f[k] = v0, f[k + 1] = v1;
        x += 2;
        if (x >= width) x -= width, y++;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; k++) f[k] = (float)px[k] * scale;
    }
    // lane pair exchange: even lane keeps quads 0,2 of the 128-byte pair block, odd lane quads 1,3 ...
    const bool odd = threadIdx.x & 1;
    float4 q[4] = {make_float4(f[0], f[1], f[2], f[3]), make_float4(f[4], f[5], f[6], f[7]),
                   make_float4(f[8], f[9], f[10], f[11]), make_float4(f[12], f[13], f[14], f[15])};
    const bool partner_active = (g ^ 1) < ngroups;  // the whole warp is converged except possibly at the very end
    const unsigned mask = __activemask();
    float4 s0 = odd ? q[0] : q[1], s1 = odd ? q[2] : q[3];
    float4 r0, r1;
    r0.x = __shfl_xor_sync(mask, s0.x, 1), r0.y = __shfl_xor_sync(mask, s0.y, 1);
    r0.z = __shfl_xor_sync(mask, s0.z, 1), r0.w = __shfl_xor_sync(mask, s0.w, 1);
    r1.x = __shfl_xor_sync(mask, s1.x, 1), r1.y = __shfl_xor_sync(mask, s1.y, 1);
    r1.z = __shfl_xor_sync(mask, s1.z, 1), r1.w = __shfl_xor_sync(mask, s1.w, 1);
    float4 *dst = reinterpret_cast<float4 *>(out + g * 16);
    if (partner_active) {
      // pair block = 8 quads: a0 a1 a2 a3 b0 b1 b2 b3 ; instruction k writes quads 2k (even lane) and 2k+1 (odd lane)
      float4 *pb = reinterpret_cast<float4 *>(out + (g & ~(int64_t)1) * 16);
      if (!odd) {
        st_stream(pb + 0, q[0]), st_stream(pb + 2, q[2]), st_stream(pb + 4, r0), st_stream(pb + 6, r1);
      } else {
        st_stream(pb + 1, r0), st_stream(pb + 3, r1), st_stream(pb + 5, q[1]), st_stream(pb + 7, q[3]);
      }
    } else {
      dst[0] = q[0], dst[1] = q[1], dst[2] = q[2], dst[3] = q[3];
    }
  }
  // tail: pixel pairs beyond the last full 16-pixel group, byte-wise
  const int64_t tail0 = ngroups * 8;
  for (int64_t p = tail0 + (int64_t)blockIdx.x * kThreads + threadIdx.x; p < npairs; p += stride) {
    const uint8_t *b = in + p * 3;
    uint32_t p0, p1;
    unpack_pair<kIds>((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16), p0, p1);
    float v0 = __fmul_rn((float)p0, scale), v1 = __fmul_rn((float)p1, scale);
    if (kFused && wb.apply) {
      const int y = (int)((2 * p) / width);
      v0 = __fsub_rn(v0, wb.black), v1 = __fsub_rn(v1, wb.black);
      if (wb.apply == 2) v0 = clip01(v0 * wb.g[y & 1][0]), v1 = clip01(v1 * wb.g[y & 1][1]);
    }
    out[2 * p] = v0, out[2 * p + 1] = v1;
  }
}

// ---- decode to 16-bit outputs (f16 bits or u16) -------------------------------------------------------------
template <bool kIds, bool kHalf>
__global__ void __launch_bounds__(kThreads) decode12_16_kernel(const uint8_t *__restrict__ in, uint16_t *__restrict__ out,
                                                               int64_t ngroups, int64_t npairs, float scale) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < ngroups; g += stride) {
    const uint2 *src = reinterpret_cast<const uint2 *>(in + g * 24);
    const uint2 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
    uint32_t px[16];
    decode16<kIds>(a, b, c, px);
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uint32_t lo = px[2 * k], hi = px[2 * k + 1];
      if (kHalf) {
        lo = __half_as_ushort(__float2half_rn((float)lo * scale));
        hi = __half_as_ushort(__float2half_rn((float)hi * scale));
      }
      w[k] = lo | (hi << 16);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(out + g * 16);  // 32 B per thread = exactly one sector
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
  const int64_t tail0 = ngroups * 8;
  for (int64_t p = tail0 + (int64_t)blockIdx.x * kThreads + threadIdx.x; p < npairs; p += stride) {
    const uint8_t *b = in + p * 3;
    uint32_t p0, p1;
    unpack_pair<kIds>((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16), p0, p1);
    if (kHalf) {
      p0 = __half_as_ushort(__float2half_rn((float)p0 * scale));
      p1 = __half_as_ushort(__float2half_rn((float)p1 * scale));
    }
    out[2 * p] = (uint16_t)p0, out[2 * p + 1] = (uint16_t)p1;
  }
}

// ---- encode ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t quantise12(float f) {
  // uint16_t(roundf(f)) saturates on the GPU (negative / NaN -> 0); then min(., 4095)   (packed.cu:74-75)
  const float r = roundf(f);
  return r > 0.0f ? (uint32_t)fminf(r, 4095.0f) : 0u;
}

template <bool kIds, bool kFloat>
__global__ void __launch_bounds__(kThreads) encode12_kernel(const void *__restrict__ in_, uint8_t *__restrict__ out,
                                                            int64_t ngroups, int64_t npairs, float scale) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < ngroups; g += stride) {
    uint32_t px[16];
    if (kFloat) {
      const float4 *src = reinterpret_cast<const float4 *>(static_cast<const float *>(in_) + g * 16);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float4 v = __ldg(src + k);
        px[4 * k] = quantise12(v.x * scale), px[4 * k + 1] = quantise12(v.y * scale);
        px[4 * k + 2] = quantise12(v.z * scale), px[4 * k + 3] = quantise12(v.w * scale);
      }
    } else {
      const uint4 *src = reinterpret_cast<const uint4 *>(static_cast<const uint16_t *>(in_) + g * 16);
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const uint4 v = __ldg(src + k);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          px[8 * k + 2 * j] = min(w[j] & 0xffffu, 4095u);
          px[8 * k + 2 * j + 1] = min(w[j] >> 16, 4095u);
        }
      }
    }
    uint32_t w[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const uint32_t v = pack_pair<kIds>(px[2 * t], px[2 * t + 1]);
      const int bit = 24 * t, wi = bit >> 5, sh = bit & 31;
      w[wi] |= v << sh;
      if (sh > 8) w[wi + 1] |= v >> (32 - sh);
    }
    uint2 *dst = reinterpret_cast<uint2 *>(out + g * 24);
    dst[0] = make_uint2(w[0], w[1]), dst[1] = make_uint2(w[2], w[3]), dst[2] = make_uint2(w[4], w[5]);
  }
  const int64_t tail0 = ngroups * 8;
  for (int64_t p = tail0 + (int64_t)blockIdx.x * kThreads + threadIdx.x; p < npairs; p += stride) {
    uint32_t p0, p1;
    if (kFloat) {
      const float *f = static_cast<const float *>(in_);
      p0 = quantise12(f[2 * p] * scale), p1 = quantise12(f[2 * p + 1] * scale);
    } else {
      const uint16_t *u = static_cast<const uint16_t *>(in_);
      p0 = min((uint32_t)u[2 * p], 4095u), p1 = min((uint32_t)u[2 * p + 1], 4095u);
    }
    const uint32_t v = pack_pair<kIds>(p0, p1);
    out[3 * p] = (uint8_t)v, out[3 * p + 1] = (uint8_t)(v >> 8), out[3 * p + 2] = (uint8_t)(v >> 16);
  }
}

inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

inline int stream_grid(int64_t ngroups, int64_t npairs) {
  // enough CTAs to fill the machine several times over, capped so that the grid-stride loop amortises setup
  const int64_t want = (ngroups > 0 ? ngroups : npairs) / kThreads + 1;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(want < cap ? want : cap);
}

}  // namespace
}  // namespace tdb

using namespace tdb;

extern "C" {

static int decode_f32_impl(const uint8_t *packed, float *out, int64_t npairs, int ids, float scale, int width,
                           const WbParams &wb, const float *gains, uint32_t filters, bool fused, tdb_stream_t stream) {
  TDB_REQUIRE(packed && out, "decode12: null pointer");
  TDB_REQUIRE(npairs >= 0, "decode12: negative size");
  if (npairs == 0) return TDB_OK;
  // vector path needs 8-byte aligned input and 16-byte aligned output; otherwise everything goes down the tail loop
  const int64_t ngroups = (aligned(packed, 8) && aligned(out, 16)) ? npairs / 8 : 0;
  const int grid = stream_grid(ngroups, npairs);
  cudaStream_t s = as_stream(stream);
#define LAUNCH(IDS, FUSED) \
  decode12_f32_kernel<IDS, FUSED><<<grid, kThreads, 0, s>>>(packed, out, ngroups, npairs, scale, width, wb, gains, filters)
  if (fused) {
    if (ids) LAUNCH(true, true); else LAUNCH(false, true);
  } else {
    if (ids) LAUNCH(true, false); else LAUNCH(false, false);
  }
#undef LAUNCH
  return check_launch("decode12_f32");
}

int tdb_decode12_f32(const uint8_t *packed, float *out, int64_t npairs, int ids_format, int scaled, tdb_stream_t stream) {
  WbParams wb{};
  return decode_f32_impl(packed, out, npairs, ids_format, scaled ? (1.0f / 4095.0f) : 1.0f, 2, wb, nullptr, 0, false, stream);
}

int tdb_unpack12_wb(const uint8_t *packed, float *cfa, int width, int height, int ids_format, uint32_t filters, float black,
                    const float *gains, tdb_stream_t stream) {
  TDB_REQUIRE(width > 0 && height > 0 && (width % 2) == 0, "unpack12_wb: width must be even and positive");
  WbParams wb{};
  wb.black = black;
  wb.apply = gains ? 2 : (black != 0.0f ? 1 : 0);
  return decode_f32_impl(packed, cfa, (int64_t)width * height / 2, ids_format, 1.0f / 4095.0f, width, wb, gains, filters, true, stream);
}

static int decode16_impl(const uint8_t *packed, uint16_t *out, int64_t npairs, int ids, bool half, float scale, tdb_stream_t stream) {
  TDB_REQUIRE(packed && out, "decode12: null pointer");
  if (npairs <= 0) return TDB_OK;
  const int64_t ngroups = (aligned(packed, 8) && aligned(out, 16)) ? npairs / 8 : 0;
  const int grid = stream_grid(ngroups, npairs);
  cudaStream_t s = as_stream(stream);
  if (half) {
    if (ids) decode12_16_kernel<true, true><<<grid, kThreads, 0, s>>>(packed, out, ngroups, npairs, scale);
    else decode12_16_kernel<false, true><<<grid, kThreads, 0, s>>>(packed, out, ngroups, npairs, scale);
  } else {
    if (ids) decode12_16_kernel<true, false><<<grid, kThreads, 0, s>>>(packed, out, ngroups, npairs, scale);
    else decode12_16_kernel<false, false><<<grid, kThreads, 0, s>>>(packed, out, ngroups, npairs, scale);
  }
  return check_launch("decode12_16");
}

int tdb_decode12_f16(const uint8_t *packed, uint16_t *out, int64_t npairs, int ids_format, int scaled, tdb_stream_t stream) {
  return decode16_impl(packed, out, npairs, ids_format, true, scaled ? (1.0f / 4095.0f) : 1.0f, stream);
}

int tdb_decode12_u16(const uint8_t *packed, uint16_t *out, int64_t npairs, int ids_format, tdb_stream_t stream) {
  return decode16_impl(packed, out, npairs, ids_format, false, 1.0f, stream);
}

static int encode_impl(const void *values, uint8_t *packed, int64_t npairs, int ids, bool is_float, float scale, tdb_stream_t stream) {
  TDB_REQUIRE(values && packed, "encode12: null pointer");
  if (npairs <= 0) return TDB_OK;
  const int64_t ngroups = (aligned(values, 16) && aligned(packed, 8)) ? npairs / 8 : 0;
  const int grid = stream_grid(ngroups, npairs);
  cudaStream_t s = as_stream(stream);
  if (is_float) {
    if (ids) encode12_kernel<true, true><<<grid, kThreads, 0, s>>>(values, packed, ngroups, npairs, scale);
    else encode12_kernel<false, true><<<grid, kThreads, 0, s>>>(values, packed, ngroups, npairs, scale);
  } else {
    if (ids) encode12_kernel<true, false><<<grid, kThreads, 0, s>>>(values, packed, ngroups, npairs, scale);
    else encode12_kernel<false, false><<<grid, kThreads, 0, s>>>(values, packed, ngroups, npairs, scale);
  }
  return check_launch("encode12");
}

int tdb_encode12_u16(const uint16_t *values, uint8_t *packed, int64_t npairs, int ids_format, tdb_stream_t stream) {
  return encode_impl(values, packed, npairs, ids_format, false, 1.0f, stream);
}

int tdb_encode12_f32(const float *values, uint8_t *packed, int64_t npairs, int ids_format, int scaled, tdb_stream_t stream) {
  return encode_impl(values, packed, npairs, ids_format, true, scaled ? 4095.0f : 1.0f, stream);
}

}  // extern "C"
