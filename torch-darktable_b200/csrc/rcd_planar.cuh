// RCD interior tiles: a CTA produces a 64 x 32 tile from an 88 x 56 patch (halo 12: every patch row and the output tile start on a
// 16-byte boundary); every thread owns 2 rows x 4 columns of pixels per step, so both Bayer row parities sit in one thread (the
// R/B-site steps are branch-free, the parity is a template argument), all offsets are immediates, and the result leaves as three
// 128-bit global stores per pixel quad.  The arithmetic expressions are those of the 32 x 32 kernel in rcd.cu, term by term (the
// selects of RCD amplify any re-association into visible differences).
//
// Every shared-memory plane is stored PHASE-PLANAR.  The previous version of this kernel kept the planes row-major: ncu showed it
// bound by shared-memory wavefronts (77 % of peak) with HALF of its 55 M wavefronts per 4K frame being bank-conflict replays -- a
// thread owns a quad of 4 columns, so the lanes of a warp are 4 floats apart and every scalar neighbourhood load (the diagonals of
// steps 4.1 / 5.1, the VH_dir crosses, the +-2 / +-3 taps of step 5.2) hit each bank four times.
// Here column c of a row lives at  row * RS + (c & 3) * PQ + (c >> 2):  four planes of 22 quads per row, so that "column 4 q + m
// of the quad q I own" is the contiguous word q + const for every m, i.e. consecutive lanes read consecutive banks whatever the
// offset.  The half-resolution planes (cells k = c >> 1) are stored as two planes of 22 the same way.  Row strides satisfy
// 2 * stride = 20 (mod 32): when a warp runs over the end of a row pair (20 quads) its remaining lanes continue in the banks
// where the first ones stopped.  Every step therefore iterates 20 quads per row pair (two more than it needs in steps 4.2 / 5.1:
// results nobody reads), and the last step, which needs 16, pairs row pairs that are 4 apart (8 * stride = 16 mod 32).
// Vector loads became four scalar loads (same wavefronts, more instructions: the issue slots were 38 % busy).  Result: 27 M bank
// conflicts -> 2.8 M, 0.243 -> 0.19 ms per 4K frame, bit-identical output (profiles/r01_ncu_full_v4_summary.csv -> v5).
#pragma once

#include "cfa_tile.cuh"

namespace tdb {
namespace v3 {

constexpr int TW = 64, TH = 32, HX = 12, HY = 12;
constexpr int PW = TW + 2 * HX, PH = TH + 2 * HY;  // 88 x 56
constexpr int PQ = PW / 4;                          // 22 quads per row = length of one phase plane
constexpr int RS = 90, RH = 58;                     // row strides of the full / half planes: 2 * stride = 20 (mod 32)
constexpr int H0 = 4, HR = 48;                      // the half planes hold patch rows 4 .. 51
constexpr int FULL = RS * PH, HALF = RH * HR;
constexpr int O_CFA = 0, O_VH = FULL, O_LPF = 2 * FULL, O_CRB = O_LPF + HALF, O_U = O_CRB + HALF;
constexpr int O_VD = O_U, O_HD = O_U + FULL;                                                   // phase 1
constexpr int O_PD = O_U, O_QD = O_U + HALF, O_PQ = O_U + 2 * HALF, O_GRB = O_U + 3 * HALF;    // later phases
constexpr int SMEM_FLOATS = O_U + 4 * HALF;
constexpr int kThreads3 = 256;  // (512 threads per CTA = 32 warps per SM was measured: no faster, and the frame tiles need a launch of their own)
static_assert(4 * HALF >= 2 * FULL, "union region");
static_assert((2 * RS) % 32 == 20 && (2 * RH) % 32 == 20 && 4 * PQ <= RS && 2 * PQ <= RH, "bank plan");

// offset of column m (relative to the first column of the owned quad) in the row dr rows below, full plane
__host__ __device__ constexpr int fo(int dr, int m) {
  const int ph = ((m % 4) + 4) % 4;
  return dr * RS + ph * PQ + (m - ph) / 4;
}
// offset of cell k (relative to the first cell 2 q of the owned quad) in the row dr rows below, half plane
__host__ __device__ constexpr int ho(int dr, int k) {
  const int ph = ((k % 2) + 2) % 2;
  return dr * RH + ph * PQ + (k - ph) / 2;
}

__device__ __forceinline__ float hp7(float a, float b, float c, float d, float e, float f, float g) {
  return sqr(a - 3.0f * b - c + 6.0f * d - e - 3.0f * f + g);
}

// stage one 12-byte group (4 packed pairs = 8 pixels = quads 2 g and 2 g + 1) of a patch row
template <bool kIds>
__device__ __forceinline__ void stage_group(const CfaSource &s, const uint8_t *row_bytes, int gy, int gx, float *dst /* row + 2 g */) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(row_bytes);
  const uint32_t *wp = reinterpret_cast<const uint32_t *>(addr & ~uintptr_t(3));
  const uint32_t sh = (uint32_t)(addr & 3) * 8;
  const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = sh ? __ldg(wp + 3) : 0u;
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

// kG0: the CFA site (even row, even column) is green, i.e. the R/B sites of even rows sit on odd columns
template <bool kG0>
__device__ __forceinline__ void rcd3_tile(float *sm, const CfaSource &src_in, float *__restrict__ rgb, int width, int height, uint32_t filters,
                                          int x_origin, int by_lo, int bx, int by) {
  CfaSource src = src_in;
  float *cfa = sm + O_CFA, *vh = sm + O_VH, *lpf = sm + O_LPF, *crb = sm + O_CRB;
  float *vd = sm + O_VD, *hd = sm + O_HD, *pd = sm + O_PD, *qd = sm + O_QD, *pq = sm + O_PQ, *grb = sm + O_GRB;
  resolve_gains(src, filters);
  const int tid = threadIdx.x;
  const int x0 = x_origin + bx * TW, y0 = (by + by_lo) * TH;
  const int gx0 = x0 - HX, gy0 = y0 - HY;  // image coordinates of patch cell (0, 0): both even, gx0 a multiple of 4

  // ---- stage the CFA patch (clamped at zero like the reference's populate step)
  if (src.cfa) {
    for (int i = tid; i < PH * PQ; i += kThreads3) {
      const int r = i / PQ, q = i - r * PQ;
      const float4 v = __ldg(reinterpret_cast<const float4 *>(src.cfa + (int64_t)(gy0 + r) * width + gx0 + 4 * q));
      float *d = cfa + r * RS + q;
      d[fo(0, 0)] = fmaxf(v.x, 0.0f), d[fo(0, 1)] = fmaxf(v.y, 0.0f), d[fo(0, 2)] = fmaxf(v.z, 0.0f), d[fo(0, 3)] = fmaxf(v.w, 0.0f);
    }
  } else {
    constexpr int G = PW / 8;  // 12-byte groups per patch row
    for (int i = tid; i < PH * G; i += kThreads3) {
      const int r = i / G, g = i - r * G;
      const int gy = gy0 + r, gx = gx0 + 8 * g;
      const uint8_t *b = src.packed + (((int64_t)gy * width + gx) >> 1) * 3;
      if (src.ids) stage_group<true>(src, b, gy, gx, cfa + r * RS + 2 * g);
      else stage_group<false>(src, b, gy, gx, cfa + r * RS + 2 * g);
    }
  }
  __syncthreads();

  constexpr int NQ = 20;  // quads 1 .. 20 of a row pair, in every step but the last

  // ---- step 1.1: squared vertical / horizontal high-pass (rcd.cu:63-75).  rows 4..51
  for (int i = tid; i < 24 * NQ; i += kThreads3) {
    const int rp = i / NQ, qc = 1 + i - rp * NQ;
    const int v0 = 4 + 2 * rp;
    const float *cf = cfa + v0 * RS + qc;
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
        vd[v0 * RS + qc + fo(rho, j)] = hp7(c[rho][j], c[rho + 1][j], c[rho + 2][j], c[rho + 3][j], c[rho + 4][j], c[rho + 5][j], c[rho + 6][j]);
        hd[v0 * RS + qc + fo(rho, j)] = hp7(w[1 + j], w[2 + j], w[3 + j], w[4 + j], w[5 + j], w[6 + j], w[7 + j]);
      }
    }
  }
  __syncthreads();

  // ---- step 1.2: VH_dir (rcd.cu:78-90), rows 6..49, and step 2.1: low-pass at R/B sites (rcd.cu:93-104), rows 4..51
  {
    constexpr int NB1 = 22 * NQ, NB2 = 24 * NQ;
    for (int i = tid; i < NB1 + NB2; i += kThreads3) {
      if (i < NB1) {
        const int rp = i / NQ, qc = 1 + i - rp * NQ;
        const int v0 = 6 + 2 * rp;
        const float *vdp = vd + v0 * RS + qc, *hdp = hd + v0 * RS + qc;
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
            vh[v0 * RS + qc + fo(rho, j)] = V / (V + Hs);
          }
        }
      } else {
        const int j2 = i - NB1;
        const int rp = j2 / NQ, qc = 1 + j2 - rp * NQ;
        const int v0 = 4 + 2 * rp;
        const float *cf = cfa + v0 * RS + qc;
        float w[4][6];  // rows v0-1 .. v0+2, columns -1 .. 4
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
          for (int m = 0; m < 6; m++) w[k][m] = cf[fo(k - 1, m - 1)];
#pragma unroll
        for (int rho = 0; rho < 2; rho++) {
          const int e = (kG0 ? 1 : 0) ^ rho;  // column parity of this row's R/B sites (compile time after unrolling)
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const int s = 1 + e + 2 * t;  // index of the site in w[][]
            const float *up = w[rho], *ce = w[rho + 1], *dn = w[rho + 2];
            lpf[(v0 - H0) * RH + qc + ho(rho, t)] =
                ce[s] + 0.5f * (up[s] + dn[s] + ce[s - 1] + ce[s + 1]) + 0.25f * (up[s - 1] + up[s + 1] + dn[s - 1] + dn[s + 1]);
          }
        }
      }
    }
  }
  __syncthreads();  // vd/hd are dead from here on: the union region is reused for pd/qd/pq/grb

  // ---- step 4.1: squared P/Q diagonal high-pass on odd columns (rcd.cu:149-163), rows 6..49
  // ---- step 3.1: green at R/B sites (rcd.cu:107-146), rows 6..49
  {
    constexpr int NB = 22 * NQ;
    for (int i = tid; i < 2 * NB; i += kThreads3) {
      const int j2 = i < NB ? i : i - NB;
      const int rp = j2 / NQ, qc = 1 + j2 - rp * NQ;
      const int v0 = 6 + 2 * rp;
      const float *cf = cfa + v0 * RS + qc;
      if (i < NB) {
#pragma unroll
        for (int rho = 0; rho < 2; rho++) {
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const int m = 1 + 2 * t;  // the odd column of the quad
#define C_(dr, dc) cf[fo(rho + (dr), m + (dc))]
            const float p = sqr((C_(-3, -3) - C_(-1, -1) - C_(1, 1) + C_(3, 3)) - 3.0f * (C_(-2, -2) + C_(2, 2)) + 6.0f * C_(0, 0));
            const float q = sqr((C_(-3, 3) - C_(-1, 1) - C_(1, -1) + C_(3, -3)) - 3.0f * (C_(-2, 2) + C_(2, -2)) + 6.0f * C_(0, 0));
#undef C_
            pd[(v0 - H0) * RH + qc + ho(rho, t)] = p;
            qd[(v0 - H0) * RH + qc + ho(rho, t)] = q;
          }
        }
      } else {
        const float *vhp = vh + v0 * RS + qc;
        const float *lp = lpf + (v0 - H0) * RH + qc;
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
            grb[(v0 - H0) * RH + qc + ho(rho, t)] = mixf(Ve, He, disc);
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- step 4.2: PQ_dir at R/B sites (rcd.cu:166-182), rows 8..47 (quads 2..19 are needed, 1..20 are computed: see the header)
  for (int i = tid; i < 20 * NQ; i += kThreads3) {
    const int rp = i / NQ, qc = 1 + i - rp * NQ;
    const int v0 = 8 + 2 * rp;
    const float *pdp = pd + (v0 - H0) * RH + qc, *qdp = qd + (v0 - H0) * RH + qc;
#pragma unroll
    for (int rho = 0; rho < 2; rho++) {
      const int e = (kG0 ? 1 : 0) ^ rho;
#pragma unroll
      for (int t = 0; t < 2; t++) {
        // row-major cell indices of the reference: i2 = row * HS + 2 qc + t, i3 = i2 - HS - 1 + e, i4 = i2 + HS - 1 + e
        const float Ps = fmaxf(1e-10f, pdp[ho(rho - 1, t - 1 + e)] + pdp[ho(rho, t)] + pdp[ho(rho + 1, t + e)]);
        const float Qs = fmaxf(1e-10f, qdp[ho(rho - 1, t + e)] + qdp[ho(rho, t)] + qdp[ho(rho + 1, t - 1 + e)]);
        pq[(v0 - H0) * RH + qc + ho(rho, t)] = Ps / (Ps + Qs);
      }
    }
  }
  __syncthreads();

  // ---- step 5.1: the opposite colour at R/B sites along the diagonals (rcd.cu:185-224), rows 8..47
  for (int i = tid; i < 20 * NQ; i += kThreads3) {
    const int rp = i / NQ, qc = 1 + i - rp * NQ;
    const int v0 = 8 + 2 * rp;
    const float *cf = cfa + v0 * RS + qc;
    const float *pqp = pq + (v0 - H0) * RH + qc, *gp = grb + (v0 - H0) * RH + qc;
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
        // cells (u - 1) >> 1 and (u + 1) >> 1 of the neighbouring rows, u = 4 qc + s
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
        crb[(v0 - H0) * RH + qc + ho(rho, t)] = g0 + mixf(Pe, Qe, disc);
      }
    }
  }
  __syncthreads();

  // ---- step 5.2 at green sites + output (rcd.cu:227-282, :49-60): rows 12..43, quads 3..18, one block per thread.
  // The two half-warps of a warp take row pairs that are 4 apart (bank plan in the header).
  {
    const int warp = tid >> 5, half = (tid >> 4) & 1;
    const int rp = (warp & 3) + 8 * (warp >> 2) + 4 * half, qc = 3 + (tid & 15);
    const int v0 = HY + 2 * rp;
    const float *cf = cfa + v0 * RS + qc, *vhp = vh + v0 * RS + qc;
    const float *gp = grb + (v0 - H0) * RH + qc, *cp = crb + (v0 - H0) * RH + qc;
#pragma unroll
    for (int rho = 0; rho < 2; rho++) {
      const int e = (kG0 ? 1 : 0) ^ rho;
      const int gy = gy0 + v0 + rho;
      // colour of this row's R/B sites: 0 = red, 2 = blue
      const bool row_red = fc(gy & 1, e, filters) == 0;
      float R[4], G[4], B[4];
#pragma unroll
      for (int t = 0; t < 2; t++) {  // R/B sites
        const int s = e + 2 * t;
        const float own = cf[fo(rho, s)], g = gp[ho(rho, t)], opp = cp[ho(rho, t)];
        G[s] = g;
        R[s] = row_red ? own : opp;
        B[s] = row_red ? opp : own;
      }
#pragma unroll
      for (int t = 0; t < 2; t++) {  // green sites
        const int s = (1 - e) + 2 * t;
        const float eps = 1e-5f;
#define C_(dr, dc) cf[fo(rho + (dr), s + (dc))]
        // cells u >> 1, (u - 1) >> 1, (u + 1) >> 1, (u - 3) >> 1, (u + 3) >> 1 of a row, u = 4 qc + s
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
        G[s] = g;
        R[s] = row_red ? res[0] : res[1];
        B[s] = row_red ? res[1] : res[0];
      }
      float4 *o = reinterpret_cast<float4 *>(rgb + 3 * ((int64_t)gy * width + gx0 + 4 * qc));
#define TDB_Z(v) fmaxf(v, 0.0f)
      st_stream(o, make_float4(TDB_Z(R[0]), TDB_Z(G[0]), TDB_Z(B[0]), TDB_Z(R[1])));
      st_stream(o + 1, make_float4(TDB_Z(G[1]), TDB_Z(B[1]), TDB_Z(R[2]), TDB_Z(G[2])));
      st_stream(o + 2, make_float4(TDB_Z(B[2]), TDB_Z(R[3]), TDB_Z(G[3]), TDB_Z(B[3])));
#undef TDB_Z
    }
  }
}

}  // namespace v3
}  // namespace tdb
