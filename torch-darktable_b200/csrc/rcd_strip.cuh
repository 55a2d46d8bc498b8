// RCD interior, strip-marching form: a CTA walks DOWN a 64-pixel-wide column strip, 16 output rows per iteration, with every
// intermediate plane of the algorithm kept as a ROLLING window of rows in shared memory.
//
// The tile kernel of rcd_planar.cuh pays the 10-row dependency radius of RCD above and below every 32-row tile: an 88 x 56 patch
// per 64 x 32 outputs, 2.4x the pixels staged and 1.7 - 1.9x the pixels computed in every step (ncu: 112 M warp instructions per 4K
// frame).  Here the vertical halo is paid once per strip SEGMENT (1.1 iterations of pipeline fill for ~19 iterations of output)
// and only the horizontal one remains (88 / 64 staged, 80 / 64 computed).
//
// Pipeline.  Iteration `it` produces output rows [Y, Y + 16), Y = y_start + 16 it.  Every step runs a fixed number of rows AHEAD of
// the output so that its inputs are complete (offsets relative to Y; all even, so a thread's 2 x 4 pixel block keeps both Bayer row
// parities and the compile-time parities of the tile kernel carry over).  Six barrier-separated phases per iteration:
//     A  1.1 v/h diff [10, 26), 2.1 lpf [8, 24), 4.1 p/q diff [8, 24)      (read the CFA rows [-3, 29) only)
//     B  1.2 VH_dir [8, 24), 4.2 PQ_dir [6, 22)
//     C  3.1 G at R/B [6, 22)
//     D  5.1 opposite colour [4, 20)
//     E  5.2 + output [0, 16); beside it the 16 CFA rows of the NEXT iteration are unpacked into a staging plane
//     F  the windows move up by 16 rows
// A plane keeps the rows between its oldest reader and its newest writer (cfa 32, VH_dir 26, G / opposite colour 24, the others 20)
// and is moved up by 16 rows once its last reader of the iteration is done -- a plain copy of 4 .. 16 rows, which keeps every
// address in the steps an immediate offset (a ring buffer would need a modulo per row access).  The arithmetic of every step is the
// tile kernel's, expression by expression: outputs are bit-identical to it.
//
// Staging (north_star: halos staged by TMA).  The 16 new CFA rows of an iteration arrive through the TMA engine while the previous
// iteration computes: for a float CFA plane ONE 2-D tensor-map copy (cp.async.bulk.tensor.2d, SASS UTMALDG: box 88 x 16 floats), for a
// 12-bit packed frame one 1-D bulk copy per row (cp.async.bulk, SASS UBLKCP; the row pitch of a packed frame is not a multiple of
// 16 bytes in general -- 9000 bytes at W = 6000 -- so no tensor map can describe it, and each row copy starts at the enclosing
// 16-byte boundary), all completing on one mbarrier.  The threads then unpack (+ black level + white balance) from shared memory into
// the phase-planar layout.  One raw buffer suffices: it is refilled right after the barrier that ends its unpacking, a whole
// iteration before it is needed.
//
// Shared memory: 63.5 KB of planes + 5.6 KB unpacked next rows + 5.5 KB raw rows = 74.6 KB per CTA of 160 threads (8 row pairs x
// 20 quads = one 2 x 4 block per thread and step), three CTAs per SM.
#pragma once

#include "rcd_planar.cuh"

namespace tdb {
namespace v4 {

using v3::fo;
using v3::ho;
using v3::hp7;
using v3::HX;
using v3::PQ;
using v3::PW;
using v3::RH;
using v3::RS;
using v3::TW;

constexpr int R = 16;     // output rows per iteration
constexpr int NT = 160;   // threads: 8 row pairs x 20 quads
constexpr int NQ = 20;
static_assert(NT == (R / 2) * NQ, "one 2 x 4 block per thread and step");

// plane heights and the row (relative to Y) of their local row 0
constexpr int CFA_H = 32, CFA_B = -3;  // rows [Y - 3, Y + 29): step 5.2 reaches 3 rows up, step 1.1 needs row Y + 28
constexpr int VD_H = 20, VD_B = 6;    // v diff and h diff
constexpr int VH_H = 26, VH_B = -2;
constexpr int LPF_H = 20, LPF_B = 4;
constexpr int PD_H = 20, PD_B = 4;    // p diff and q diff
constexpr int GRB_H = 24, GRB_B = -2;
constexpr int PQD_H = 20, PQD_B = 2;  // PQ_dir
constexpr int CRB_H = 24, CRB_B = -4;

constexpr int O_CFA = 0, O_VD = O_CFA + CFA_H * RS, O_HD = O_VD + VD_H * RS, O_VH = O_HD + VD_H * RS;
constexpr int O_LPF = O_VH + VH_H * RS, O_PD = O_LPF + LPF_H * RH, O_QD = O_PD + PD_H * RH, O_GRB = O_QD + PD_H * RH;
constexpr int O_PQ = O_GRB + GRB_H * RH, O_CRB = O_PQ + PQD_H * RH, O_END = O_CRB + CRB_H * RH;
constexpr int O_NEW = O_END;                          // the next iteration's 16 CFA rows, unpacked
constexpr int O_RAW = (O_NEW + R * RS + 31) / 32 * 32;  // 128-byte aligned (tensor-map destination)
constexpr int RAWB = PW * 4;                          // bytes per raw row: 88 floats, or <= 160 packed bytes
constexpr int O_BAR = O_RAW + R * RAWB / 4;
constexpr int SMEM_FLOATS = O_BAR + 4;
static_assert(O_CFA % 2 == 0 && O_VD % 2 == 0 && O_VH % 2 == 0 && O_LPF % 2 == 0 && O_GRB % 2 == 0 && O_CRB % 2 == 0 && RS % 2 == 0 && RH % 2 == 0,
              "row shifts move float2");

struct StripArgs {
  CfaSource src;
  float *rgb;
  int width, height;
  uint32_t filters;
  int x_origin;          // first output column of strip 0
  int nstrips, nseg;     // strips across, segments down
  int y_start, n_iter;   // the strips cover rows [y_start, y_start + 16 n_iter)
  int n_strip_jobs;      // nstrips * nseg
  int frame_first;       // the 32 x 32 frame tiles occupy the first blocks of the grid instead of the last
  int n_frame;
};

// rows [16, 16 + kRows) of a plane -> rows [0, kRows): the part of its window that the next iteration still reads
template <int kRows, int kStride>
__device__ __forceinline__ void shift_rows(float *plane, int tid) {
  static_assert(kRows <= R, "source and destination rows must not overlap");
  float2 *p = reinterpret_cast<float2 *>(plane);
  constexpr int n = kRows * kStride / 2, off = R * kStride / 2;
  for (int i = tid; i < n; i += NT) p[i] = p[i + off];
}

// four packed pairs (12 bytes, 8 pixels) starting `sh` bytes into the aligned words w0 .. w3
template <bool kIds>
__device__ __forceinline__ void unpack_group(const CfaSource &s, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t sh, int gy, int gx,
                                             float *dst) {
  const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh), x2 = __funnelshift_r(w2, w3, sh);
  const uint32_t pr[4] = {x0 & 0xffffffu, (x0 >> 24) | ((x1 & 0xffffu) << 8), (x1 >> 16) | ((x2 & 0xffu) << 16), x2 >> 8};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t p0, p1;
    unpack_pair<kIds>(pr[k], p0, p1);
    dst[fo(0, 2 * k)] = fmaxf(finish_sample(s, p0, gy, gx + 2 * k), 0.0f);
    dst[fo(0, 2 * k + 1)] = fmaxf(finish_sample(s, p1, gy, gx + 2 * k + 1), 0.0f);
  }
}

#define TDB_RF(plane, rel) (((rel) - plane##_B) * RS)
#define TDB_RHF(plane, rel) (((rel) - plane##_B) * RH)

template <bool kG0>
__device__ __forceinline__ void rcd_strip(float *sm, const CUtensorMap *tmap, const StripArgs &a, int job) {
  float *cfa = sm + O_CFA, *vd = sm + O_VD, *hd = sm + O_HD, *vh = sm + O_VH, *lpf = sm + O_LPF;
  float *pd = sm + O_PD, *qd = sm + O_QD, *grb = sm + O_GRB, *pq = sm + O_PQ, *crb = sm + O_CRB;
  float *nxt = sm + O_NEW;  // 16 unpacked rows of the NEXT iteration, planar like the CFA rows
  uint8_t *raw = reinterpret_cast<uint8_t *>(sm + O_RAW);
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + O_BAR);
  CfaSource src = a.src;
  resolve_gains(src, a.filters);
  const int tid = threadIdx.x;
  const int sx = job % a.nstrips, sg = job / a.nstrips;
  const int it0 = (int)((int64_t)sg * a.n_iter / a.nseg), it1 = (int)((int64_t)(sg + 1) * a.n_iter / a.nseg);
  const int W = a.width;
  const int x0 = a.x_origin + sx * TW, gx0 = x0 - HX;  // image column of patch column 0: a multiple of 4
  const int rp = tid / NQ, qc = 1 + tid - rp * NQ;      // this thread's row pair and quad in every step but the last

  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();

  // the 16 CFA rows [Y + 13, Y + 29) of iteration `it` -> raw buffer, asynchronously
  auto fill = [&](int it) {
    const int gy = a.y_start + R * it + 13;
    if (src.cfa) {
      if (tid == 0) {
        mbar_arrive_expect_tx(bar, (uint32_t)(R * RAWB));
        tma_load_2d(raw, tmap, gx0, gy, bar);
      }
    } else if (tid < 32) {
      uint32_t bytes = 0;
      const uint8_t *g = src.packed;
      if (tid < R) {
        const int64_t off = ((((int64_t)(gy + tid)) * W + gx0) >> 1) * 3;  // byte offset of the row's first pair
        bytes = ((uint32_t)(off & 15) + (uint32_t)(PW * 3 / 2) + 15u) & ~15u;
        g += off & ~(int64_t)15;
      }
      uint32_t total = bytes;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      if (tid == 0) mbar_arrive_expect_tx(bar, total);
      __syncwarp();
      if (tid < R) bulk_copy_g2s(raw + tid * RAWB, g, bytes, bar);
    }
  };

  // the raw rows of iteration `it` (image rows [Y + 13, Y + 29)) -> the planar staging rows `nxt` (clamped at zero like the reference's
  // populate step).  Runs beside step 5.2: the warp without 5.2 work takes most of it.
  auto unpack = [&](int it) {
    const int Yn = a.y_start + R * it;
    if (src.cfa) {
      auto task = [&](int i) {
        const int r = i / PQ, q = i - r * PQ;
        const float4 v = *reinterpret_cast<const float4 *>(raw + r * RAWB + 16 * q);
        float *d = nxt + r * RS + q;
        d[fo(0, 0)] = fmaxf(v.x, 0.0f), d[fo(0, 1)] = fmaxf(v.y, 0.0f), d[fo(0, 2)] = fmaxf(v.z, 0.0f), d[fo(0, 3)] = fmaxf(v.w, 0.0f);
      };
      constexpr int total = R * PQ, w4 = total;  // a task is one float4: all of them fit beside one 5.2 block
      if (tid >= 128) for (int i = tid - 128; i < w4; i += 32) task(i);
      else for (int i = w4 + tid; i < total; i += 128) task(i);
    } else {
      constexpr int G = PW / 8;  // 12-byte groups per row
      auto task = [&](int i) {
        const int r = i / G, g = i - r * G;
        const int gy = Yn + 13 + r, gx = gx0 + 8 * g;
        const uint32_t boff = (uint32_t)((((((int64_t)gy) * W + gx0) >> 1) * 3) & 15) + 12u * g;  // byte offset inside the raw row
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(raw + r * RAWB + (boff & ~3u));
        const uint32_t sh = (boff & 3u) * 8u;
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = sh ? wp[3] : 0u;
        float *d = nxt + r * RS + 2 * g;
        if (src.ids) unpack_group<true>(src, w0, w1, w2, w3, sh, gy, gx, d);
        else unpack_group<false>(src, w0, w1, w2, w3, sh, gy, gx, d);
      };
      constexpr int total = R * G, w4 = 96;     // three unpack tasks beside one 5.2 block (five were measured slower: 0.168 -> 0.176 ms at 4K)
      if (tid >= 128) for (int i = tid - 128; i < w4; i += 32) task(i);
      else for (int i = w4 + tid; i < total; i += 128) task(i);
    }
  };

  // ---- fill: the first 16 CFA rows go straight into place
  uint32_t phase = 0;
  fill(it0 - 2);
  mbar_wait(bar, phase);
  phase ^= 1;
  unpack(it0 - 2);
  __syncthreads();
  fill(it0 - 1);
  {
    const float2 *n2 = reinterpret_cast<const float2 *>(nxt);
    float2 *c2 = reinterpret_cast<float2 *>(cfa) + R * RS / 2;
    for (int i = tid; i < R * RS / 2; i += NT) c2[i] = n2[i];
  }
  __syncthreads();

  for (int it = it0 - 2; it < it1; it++) {
    const int Y = a.y_start + R * it;
    const bool steady = it >= it0 - 1;  // the first fill iteration only needs step 1.1 (one row of v diff, for step 1.2 of the next)

    // ---- A: steps that read the CFA only.  1.1 squared vertical / horizontal high-pass (rcd.cu:63-75), rows [10, 26);
    //         2.1 low-pass at R/B sites (rcd.cu:93-104) and 4.1 P/Q diagonal high-pass on odd columns (rcd.cu:149-163), rows [8, 24)
    {
      const int rel0 = 10 + 2 * rp;
      const float *cf = cfa + TDB_RF(CFA, rel0) + qc;
      float c[8][4];
#pragma unroll
      for (int k = 0; k < 8; k++)
#pragma unroll
        for (int j = 0; j < 4; j++) c[k][j] = cf[fo(k - 3, j)];
#pragma unroll
      for (int rho = 0; rho < 2; rho++) {
        float w[12];
#pragma unroll
        for (int m = 1; m <= 3; m++) w[m] = cf[fo(rho, m - 4)];
#pragma unroll
        for (int j = 0; j < 4; j++) w[4 + j] = c[rho + 3][j];
#pragma unroll
        for (int m = 8; m <= 10; m++) w[m] = cf[fo(rho, m - 4)];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          vd[TDB_RF(VD, rel0) + qc + fo(rho, j)] = hp7(c[rho][j], c[rho + 1][j], c[rho + 2][j], c[rho + 3][j], c[rho + 4][j], c[rho + 5][j], c[rho + 6][j]);
          hd[TDB_RF(VD, rel0) + qc + fo(rho, j)] = hp7(w[1 + j], w[2 + j], w[3 + j], w[4 + j], w[5 + j], w[6 + j], w[7 + j]);
        }
      }
    }
    if (steady) {
      {
        const int rel0 = 8 + 2 * rp;
        const float *cf = cfa + TDB_RF(CFA, rel0) + qc;
        float w[4][6];  // rows rel0-1 .. rel0+2, columns -1 .. 4
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
          for (int m = 0; m < 6; m++) w[k][m] = cf[fo(k - 1, m - 1)];
#pragma unroll
        for (int rho = 0; rho < 2; rho++) {
          const int e = (kG0 ? 1 : 0) ^ rho;  // column parity of this row's R/B sites
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const int s = 1 + e + 2 * t;
            const float *up = w[rho], *ce = w[rho + 1], *dn = w[rho + 2];
            lpf[TDB_RHF(LPF, rel0) + qc + ho(rho, t)] =
                ce[s] + 0.5f * (up[s] + dn[s] + ce[s - 1] + ce[s + 1]) + 0.25f * (up[s - 1] + up[s + 1] + dn[s - 1] + dn[s + 1]);
          }
        }
      }
      {
        const int rel0 = 8 + 2 * rp;
        const float *cf = cfa + TDB_RF(CFA, rel0) + qc;
#pragma unroll
        for (int rho = 0; rho < 2; rho++) {
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const int m = 1 + 2 * t;  // the odd column of the quad
#define C_(dr, dc) cf[fo(rho + (dr), m + (dc))]
            const float p = sqr((C_(-3, -3) - C_(-1, -1) - C_(1, 1) + C_(3, 3)) - 3.0f * (C_(-2, -2) + C_(2, 2)) + 6.0f * C_(0, 0));
            const float q = sqr((C_(-3, 3) - C_(-1, 1) - C_(1, -1) + C_(3, -3)) - 3.0f * (C_(-2, 2) + C_(2, -2)) + 6.0f * C_(0, 0));
#undef C_
            pd[TDB_RHF(PD, rel0) + qc + ho(rho, t)] = p;
            qd[TDB_RHF(PD, rel0) + qc + ho(rho, t)] = q;
          }
        }
      }
    }
    __syncthreads();

    if (steady) {
      // ---- B: step 1.2 VH_dir (rcd.cu:78-90), rows [8, 24), and step 4.2 PQ_dir at R/B sites (rcd.cu:166-182), rows [6, 22)
      {
        const int rel0 = 8 + 2 * rp;
        const float *vdp = vd + TDB_RF(VD, rel0) + qc, *hdp = hd + TDB_RF(VD, rel0) + qc;
        float v[4][4];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
          for (int j = 0; j < 4; j++) v[k][j] = vdp[fo(k - 1, j)];
#pragma unroll
        for (int rho = 0; rho < 2; rho++) {
          float h[6];  // columns -1 .. 4
#pragma unroll
          for (int m = 0; m < 6; m++) h[m] = hdp[fo(rho, m - 1)];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float V = fmaxf(1e-10f, v[rho][j] + v[rho + 1][j] + v[rho + 2][j]);
            const float Hs = fmaxf(1e-10f, h[j] + h[j + 1] + h[j + 2]);
            vh[TDB_RF(VH, rel0) + qc + fo(rho, j)] = V / (V + Hs);
          }
        }
      }
      {
        const int rel0 = 6 + 2 * rp;
        const float *pdp = pd + TDB_RHF(PD, rel0) + qc, *qdp = qd + TDB_RHF(PD, rel0) + qc;
#pragma unroll
        for (int rho = 0; rho < 2; rho++) {
          const int e = (kG0 ? 1 : 0) ^ rho;
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const float Ps = fmaxf(1e-10f, pdp[ho(rho - 1, t - 1 + e)] + pdp[ho(rho, t)] + pdp[ho(rho + 1, t + e)]);
            const float Qs = fmaxf(1e-10f, qdp[ho(rho - 1, t + e)] + qdp[ho(rho, t)] + qdp[ho(rho + 1, t - 1 + e)]);
            pq[TDB_RHF(PQD, rel0) + qc + ho(rho, t)] = Ps / (Ps + Qs);
          }
        }
      }
      __syncthreads();

      // ---- C: step 3.1 green at R/B sites (rcd.cu:107-146), rows [6, 22); v/h and p/q diff have had their last readers
      {
        const int rel0 = 6 + 2 * rp;
        const float *cf = cfa + TDB_RF(CFA, rel0) + qc;
        const float *vhp = vh + TDB_RF(VH, rel0) + qc;
        const float *lp = lpf + TDB_RHF(LPF, rel0) + qc;
#pragma unroll
        for (int rho = 0; rho < 2; rho++) {
          const int e = (kG0 ? 1 : 0) ^ rho;
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const int s = e + 2 * t;  // site offset inside the quad
            const float eps = 1e-5f;
            float cv[9], ch[9];  // the site's column, rows -4 .. +4, and its row, columns -4 .. +4
#pragma unroll
            for (int k = 0; k < 9; k++) cv[k] = cf[fo(rho + k - 4, s)], ch[k] = cf[fo(rho, s + k - 4)];
            const float c0 = vhp[fo(rho, s)];
            const float nb = 0.25f * (vhp[fo(rho - 1, s - 1)] + vhp[fo(rho - 1, s + 1)] + vhp[fo(rho + 1, s - 1)] + vhp[fo(rho + 1, s + 1)]);
            const float disc = (fabsf(0.5f - c0) < fabsf(0.5f - nb)) ? nb : c0;
            const float ci = cv[4];
            const float Ng = eps + fabsf(cv[3] - cv[5]) + fabsf(ci - cv[2]) + fabsf(cv[3] - cv[1]) + fabsf(cv[2] - cv[0]);
            const float Sg = eps + fabsf(cv[5] - cv[3]) + fabsf(ci - cv[6]) + fabsf(cv[5] - cv[7]) + fabsf(cv[6] - cv[8]);
            const float Wg = eps + fabsf(ch[3] - ch[5]) + fabsf(ci - ch[2]) + fabsf(ch[3] - ch[1]) + fabsf(ch[2] - ch[0]);
            const float Eg = eps + fabsf(ch[5] - ch[3]) + fabsf(ci - ch[6]) + fabsf(ch[5] - ch[7]) + fabsf(ch[6] - ch[8]);
            const float li = lp[ho(rho, t)];
            const float Ne = cv[3] * (li + li) / (eps + li + lp[ho(rho - 2, t)]);
            const float Se = cv[5] * (li + li) / (eps + li + lp[ho(rho + 2, t)]);
            const float We = ch[3] * (li + li) / (eps + li + lp[ho(rho, t - 1)]);
            const float Ee = ch[5] * (li + li) / (eps + li + lp[ho(rho, t + 1)]);
            const float Ve = (Sg * Ne + Ng * Se) / (Ng + Sg);
            const float He = (Wg * Ee + Eg * We) / (Eg + Wg);
            grb[TDB_RHF(GRB, rel0) + qc + ho(rho, t)] = mixf(Ve, He, disc);
          }
        }
      }
      shift_rows<VD_H - R, RS>(vd, tid);
      shift_rows<VD_H - R, RS>(hd, tid);
      shift_rows<PD_H - R, RH>(pd, tid);
      shift_rows<PD_H - R, RH>(qd, tid);
      __syncthreads();

      // ---- D: step 5.1, the opposite colour at R/B sites along the diagonals (rcd.cu:185-224), rows [4, 20); lpf is done
      {
        const int rel0 = 4 + 2 * rp;
        const float *cf = cfa + TDB_RF(CFA, rel0) + qc;
        const float *pqp = pq + TDB_RHF(PQD, rel0) + qc, *gp = grb + TDB_RHF(GRB, rel0) + qc;
#pragma unroll
        for (int rho = 0; rho < 2; rho++) {
          const int e = (kG0 ? 1 : 0) ^ rho;
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const float eps = 1e-5f;
            const int s = e + 2 * t;
            const float c0 = pqp[ho(rho, t)];
            const float nb = 0.25f * (pqp[ho(rho - 1, t - 1 + e)] + pqp[ho(rho - 1, t + e)] + pqp[ho(rho + 1, t - 1 + e)] + pqp[ho(rho + 1, t + e)]);
            const float disc = (fabsf(0.5f - c0) < fabsf(0.5f - nb)) ? nb : c0;
            const float g0 = gp[ho(rho, t)];
            const int kw = ((s + 3) >> 1) - 2, ke = (s + 1) >> 1;
            const float gNW = gp[ho(rho - 1, kw)], gNE = gp[ho(rho - 1, ke)];
            const float gSW = gp[ho(rho + 1, kw)], gSE = gp[ho(rho + 1, ke)];
            const float gNW2 = gp[ho(rho - 2, t - 1)], gNE2 = gp[ho(rho - 2, t + 1)];
            const float gSW2 = gp[ho(rho + 2, t - 1)], gSE2 = gp[ho(rho + 2, t + 1)];
#define C_(dr, dc) cf[fo(rho + (dr), s + (dc))]
            const float cNW = C_(-1, -1), cNE = C_(-1, 1), cSW = C_(1, -1), cSE = C_(1, 1);
            const float NWg = eps + fabsf(cNW - cSE) + fabsf(cNW - C_(-3, -3)) + fabsf(g0 - gNW2);
            const float NEg = eps + fabsf(cNE - cSW) + fabsf(cNE - C_(-3, 3)) + fabsf(g0 - gNE2);
            const float SWg = eps + fabsf(cNE - cSW) + fabsf(cSW - C_(3, -3)) + fabsf(g0 - gSW2);
            const float SEg = eps + fabsf(cNW - cSE) + fabsf(cSE - C_(3, 3)) + fabsf(g0 - gSE2);
#undef C_
            const float NWe = cNW - gNW, NEe = cNE - gNE, SWe = cSW - gSW, SEe = cSE - gSE;
            const float Pe = (NWg * SEe + SEg * NWe) / (NWg + SEg);
            const float Qe = (NEg * SWe + SWg * NEe) / (NEg + SWg);
            crb[TDB_RHF(CRB, rel0) + qc + ho(rho, t)] = g0 + mixf(Pe, Qe, disc);
          }
        }
      }
      shift_rows<LPF_H - R, RH>(lpf, tid);
      __syncthreads();
    } else {
      shift_rows<VD_H - R, RS>(vd, tid);
      shift_rows<VD_H - R, RS>(hd, tid);
      __syncthreads();
    }

    // ---- E: step 5.2 at green sites + output (rcd.cu:227-282, :49-60), rows [0, 16), quads 3 .. 18.  The two half-warps of a warp take
    //         row pairs that are 4 apart (bank plan of rcd_planar.cuh).  Beside it the next iteration's CFA rows are unpacked.
    if (it >= it0 && tid < 128) {
      const int warp = tid >> 5, half = (tid >> 4) & 1;
      const int rp2 = warp + 4 * half, q2 = 3 + (tid & 15);
      const int rel0 = 2 * rp2;
      const float *cf = cfa + TDB_RF(CFA, rel0) + q2, *vhp = vh + TDB_RF(VH, rel0) + q2;
      const float *gp = grb + TDB_RHF(GRB, rel0) + q2, *cp = crb + TDB_RHF(CRB, rel0) + q2;
#pragma unroll
      for (int rho = 0; rho < 2; rho++) {
        const int e = (kG0 ? 1 : 0) ^ rho;
        const int gy = Y + rel0 + rho;
        const bool row_red = fc(gy & 1, e, a.filters) == 0;  // colour of this row's R/B sites
        float Rv[4], Gv[4], Bv[4];
#pragma unroll
        for (int t = 0; t < 2; t++) {  // R/B sites
          const int s = e + 2 * t;
          const float own = cf[fo(rho, s)], g = gp[ho(rho, t)], opp = cp[ho(rho, t)];
          Gv[s] = g;
          Rv[s] = row_red ? own : opp;
          Bv[s] = row_red ? opp : own;
        }
#pragma unroll
        for (int t = 0; t < 2; t++) {  // green sites
          const int s = (1 - e) + 2 * t;
          const float eps = 1e-5f;
#define C_(dr, dc) cf[fo(rho + (dr), s + (dc))]
          const int k = s >> 1, kw = ((s + 3) >> 1) - 2, ke = (s + 1) >> 1, kw3 = ((s + 1) >> 1) - 2, ke3 = (s + 3) >> 1;
          const float c0 = vhp[fo(rho, s)];
          const float nb = 0.25f * (vhp[fo(rho - 1, s - 1)] + vhp[fo(rho - 1, s + 1)] + vhp[fo(rho + 1, s - 1)] + vhp[fo(rho + 1, s + 1)]);
          const float disc = (fabsf(0.5f - c0) < fabsf(0.5f - nb)) ? nb : c0;
          const float g = C_(0, 0);
          const float N1 = eps + fabsf(g - C_(-2, 0)), S1 = eps + fabsf(g - C_(2, 0));
          const float W1 = eps + fabsf(g - C_(0, -2)), E1 = eps + fabsf(g - C_(0, 2));
          const float gN = gp[ho(rho - 1, k)], gS = gp[ho(rho + 1, k)], gW = gp[ho(rho, kw)], gE = gp[ho(rho, ke)];
          float res[2];
#pragma unroll
          for (int pass = 0; pass < 2; pass++) {
            float n1, s1, w1, e1, n3, s3, w3, e3;
            if (pass == 0) {  // the colour of this row's R/B sites: native left/right, interpolated (step 5.1) above/below
              w1 = C_(0, -1), e1 = C_(0, 1), w3 = C_(0, -3), e3 = C_(0, 3);
              n1 = cp[ho(rho - 1, k)], s1 = cp[ho(rho + 1, k)], n3 = cp[ho(rho - 3, k)], s3 = cp[ho(rho + 3, k)];
            } else {          // the other colour: native above/below, interpolated left/right
              n1 = C_(-1, 0), s1 = C_(1, 0), n3 = C_(-3, 0), s3 = C_(3, 0);
              w1 = cp[ho(rho, kw)], e1 = cp[ho(rho, ke)], w3 = cp[ho(rho, kw3)], e3 = cp[ho(rho, ke3)];
            }
            const float SN = fabsf(n1 - s1), EW = fabsf(w1 - e1);
            const float Ng = N1 + SN + fabsf(n1 - n3), Sg = S1 + SN + fabsf(s1 - s3);
            const float Wg = W1 + EW + fabsf(w1 - w3), Eg = E1 + EW + fabsf(e1 - e3);
            const float Ne = n1 - gN, Se = s1 - gS, We = w1 - gW, Ee = e1 - gE;
            const float Ve = (Ng * Se + Sg * Ne) / (Ng + Sg);
            const float He = (Eg * We + Wg * Ee) / (Eg + Wg);
            res[pass] = g + mixf(Ve, He, disc);
          }
#undef C_
          Gv[s] = g;
          Rv[s] = row_red ? res[0] : res[1];
          Bv[s] = row_red ? res[1] : res[0];
        }
        float4 *o = reinterpret_cast<float4 *>(a.rgb + 3 * ((int64_t)gy * W + gx0 + 4 * q2));
#define TDB_Z(v) fmaxf(v, 0.0f)
        st_stream(o, make_float4(TDB_Z(Rv[0]), TDB_Z(Gv[0]), TDB_Z(Bv[0]), TDB_Z(Rv[1])));
        st_stream(o + 1, make_float4(TDB_Z(Gv[1]), TDB_Z(Bv[1]), TDB_Z(Rv[2]), TDB_Z(Gv[2])));
        st_stream(o + 2, make_float4(TDB_Z(Bv[2]), TDB_Z(Rv[3]), TDB_Z(Gv[3]), TDB_Z(Bv[3])));
#undef TDB_Z
      }
    }
    if (it + 1 < it1) {
      mbar_wait(bar, phase);
      phase ^= 1;
      unpack(it + 1);
    }
    shift_rows<PQD_H - R, RH>(pq, tid);
    __syncthreads();

    // ---- F: move the windows of the planes that step 5.2 reads; the unpacked rows take the place the CFA window vacates
    {
      const float2 *n2 = reinterpret_cast<const float2 *>(nxt);
      float2 *c2 = reinterpret_cast<float2 *>(cfa);
      constexpr int off = R * RS / 2;
      for (int i = tid; i < off; i += NT) {
        c2[i] = c2[i + off];
        c2[i + off] = n2[i];
      }
    }
    shift_rows<VH_H - R, RS>(vh, tid);
    shift_rows<GRB_H - R, RH>(grb, tid);
    shift_rows<CRB_H - R, RH>(crb, tid);
    __syncthreads();
    if (it + 2 < it1) fill(it + 2);  // the raw rows are consumed: the load for the iteration after next runs under a whole iteration
  }
}

#undef TDB_RF
#undef TDB_RHF

}  // namespace v4
}  // namespace tdb
