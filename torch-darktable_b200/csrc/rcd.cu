// RCD (ratio corrected demosaicing) as ONE fused shared-memory kernel.
//
// The reference (csrc/debayer/rcd.cu:601-671) runs 13 launches over 9 full-frame scratch planes (about 145 B of HBM
// traffic per pixel).  Here a CTA stages a CFA patch with a 10-pixel halo (the dependency radius of the algorithm) from
// the float plane or straight from the 12-bit packed bytes, keeps every intermediate plane (v/h high-pass, VH_dir, lpf,
// P/Q diff, PQ_dir, G at R/B, opposite colour at R/B) in shared memory, and writes the RGB tile with 128-bit stores:
// 4 (or 1.5) B in + 12 B out per pixel, no scratch in HBM, no state between calls.
//
// Parity notes.  The reference addresses its half-resolution planes by flat idx/2 and
// re-uses VP_diff/HQ_diff first for v/h_diff (full-res index) and then for p/q_diff (idx/2) without clearing.  Step 4.2
// therefore reads a few cells that step 4.1 never wrote; they hold v/h_diff values of the CURRENT frame from other image
// positions (or zero).  `stale_cell` reproduces exactly that for a FRESH reference workspace, so the output matches a
// freshly constructed reference RCD object everywhere, including the band just inside the 7-px margin.  What the
// reference leaks from the PREVIOUS frame (rows 2-3 of VH_dir) is deliberately not reproduced.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "cfa_tile.cuh"
#include "rcd_planar.cuh"
#include "rcd_strip.cuh"

namespace tdb {
namespace {

constexpr int kThreads256 = 256;
constexpr int T = 32;          // output tile edge
constexpr int HALO = 10;
constexpr int P = T + 2 * HALO;  // patch edge (52)
constexpr int SP = P;            // full-plane row stride
constexpr int PH = P / 2;        // half-plane row stride (cells)
constexpr int PT = T + 2;        // PPG border tmp patch

// smem layout (floats)
constexpr int OFF_CFA = 0;
constexpr int OFF_VH = OFF_CFA + P * SP;
constexpr int OFF_LPF = OFF_VH + P * SP;
constexpr int OFF_CRB = OFF_LPF + P * PH;
constexpr int OFF_U = OFF_CRB + P * PH;     // union region: {vdiff, hdiff} then {pd, qd, pq, grb}
constexpr int OFF_VD = OFF_U, OFF_HD = OFF_U + P * SP;
constexpr int OFF_PD = OFF_U, OFF_QD = OFF_U + P * PH, OFF_PQ = OFF_U + 2 * P * PH, OFF_GRB = OFF_U + 3 * P * PH;
constexpr int OFF_OUT = OFF_U + 2 * P * SP;
constexpr int SMEM_FLOATS = OFF_OUT + T * T * 3;
static_assert(4 * P * PH <= 2 * P * SP, "union region too small");
static_assert(PT * PT * 3 <= 2 * P * SP, "border tmp must fit the union region");
static_assert((OFF_OUT % 4) == 0, "output tile must be 16-byte aligned");

__device__ __forceinline__ float cfa_clamped(const CfaSource &s, int x, int y, int width) { return fmaxf(0.0f, cfa_at(s, x, y, width)); }

// content of the reference's VP_diff (is_p) / HQ_diff plane at half-cell (row r, cell K) when step 4.1 did not write it
__device__ float stale_cell(const CfaSource &s, bool is_p, int r, int K, int width, int height) {
  const int64_t flat = (int64_t)r * (width >> 1) + K;
  const int fr = (int)(flat / width), fx = (int)(flat - (int64_t)fr * width);
  if (fr < 3 || fr > height - 4 || fx < 3 || fx > width - 4) return 0.0f;
  float c[7];
#pragma unroll
  for (int k = -3; k <= 3; k++) c[k + 3] = is_p ? cfa_clamped(s, fx, fr + k, width) : cfa_clamped(s, fx + k, fr, width);
  return sqr(c[0] - 3.0f * c[1] - c[2] + 6.0f * c[3] - c[4] - 3.0f * c[5] + c[6]);
}

// which 32x32 tiles a launch covers: up to four rectangles of tiles (the frame around the interior of rcd_planar.cuh), or everything
struct TileRects {
  int n;                 // 0: plain 2-D grid over the whole image
  int tx0[4], ty0[4], ntx[4];
  int start[5];          // prefix sums of the tile counts
};

// one 32 x 32 tile; `tile` = blockIdx.x of a 1-D launch over `rects`, (bx, by) = tile coordinates of a plain 2-D launch.
// kThreads = threads of the CTA that runs it (256 on its own and inside the tile kernel, 160 inside the strip kernel)
template <int kThreads>
__device__ __forceinline__ void rcd_tile(float *sm, const CfaSource &src_in, float *__restrict__ rgb, int width, int height, uint32_t filters,
                                         const TileRects &rects, int tile, int bx, int by) {
  CfaSource src = src_in;
  float *cfa = sm + OFF_CFA, *vh = sm + OFF_VH, *lpf = sm + OFF_LPF, *crb = sm + OFF_CRB;
  float *vd = sm + OFF_VD, *hd = sm + OFF_HD;
  float *pd = sm + OFF_PD, *qd = sm + OFF_QD, *pq = sm + OFF_PQ, *grb = sm + OFF_GRB;
  float *outt = sm + OFF_OUT;

  resolve_gains(src, filters);
  const int tid = threadIdx.x;
  if (rects.n > 0) {
    int r = 0;
#pragma unroll
    for (int k = 1; k < 4; k++) r += (k < rects.n && tile >= rects.start[k]) ? 1 : 0;
    const int local = tile - rects.start[r];
    by = rects.ty0[r] + local / rects.ntx[r], bx = rects.tx0[r] + local % rects.ntx[r];
  }
  const int x0 = bx * T, y0 = by * T;
  const int gx0 = x0 - HALO, gy0 = y0 - HALO;  // image coordinates of patch cell (0,0); both even

  stage_patch<Oob::kZero, true>(cfa, SP, gx0, gy0, P, P, src, width, height);
  __syncthreads();

  // ---- step 1.1: squared vertical / horizontal high-pass (rcd.cu:63-75); zero where the reference never writes
  {
    constexpr int N = P - 6;
    for (int i = tid; i < N * N; i += kThreads) {
      const int v = 3 + i / N, u = 3 + i % N;
      const int gx = gx0 + u, gy = gy0 + v;
      float a = 0.0f, b = 0.0f;
      if (gy >= 3 && gy <= height - 4 && gx >= 3 && gx <= width - 4) {
        const float *c = cfa + v * SP + u;
        a = sqr(c[-3 * SP] - 3.0f * c[-2 * SP] - c[-SP] + 6.0f * c[0] - c[SP] - 3.0f * c[2 * SP] + c[3 * SP]);
        b = sqr(c[-3] - 3.0f * c[-2] - c[-1] + 6.0f * c[0] - c[1] - 3.0f * c[2] + c[3]);
      }
      vd[v * SP + u] = a, hd[v * SP + u] = b;
    }
  }
  __syncthreads();
  // ---- step 1.2: VH_dir (rcd.cu:78-90)
  {
    constexpr int N = P - 8;
    for (int i = tid; i < N * N; i += kThreads) {
      const int v = 4 + i / N, u = 4 + i % N;
      const int gx = gx0 + u, gy = gy0 + v;
      float r = 0.0f;
      if (gy >= 2 && gy <= height - 3 && gx >= 2 && gx <= width - 3) {
        const float V = fmaxf(1e-10f, vd[(v - 1) * SP + u] + vd[v * SP + u] + vd[(v + 1) * SP + u]);
        const float Hs = fmaxf(1e-10f, hd[v * SP + u - 1] + hd[v * SP + u] + hd[v * SP + u + 1]);
        r = V / (V + Hs);
      }
      vh[v * SP + u] = r;
    }
  }
  // ---- step 2.1: low-pass at R/B sites -> half plane (rcd.cu:93-104); reads cfa only, no barrier needed before it
  {
    constexpr int NR = P - 6, NC = PH - 2;  // rows 3..P-4, cells 1..PH-2
    for (int i = tid; i < NR * NC; i += kThreads) {
      const int v = 3 + i / NC, k = 1 + i % NC;
      const int gy = gy0 + v;
      const int u = 2 * k + (fc(gy & 1, 0, filters) & 1);  // R/B sites of this row (patch and image columns share parity)
      const int gx = gx0 + u;
      float r = 0.0f;
      if (gy >= 2 && gy <= height - 2 && gx >= 2 && gx <= width - 2) {
        const float *c = cfa + v * SP + u;
        r = c[0] + 0.5f * (c[-SP] + c[SP] + c[-1] + c[1]) + 0.25f * (c[-SP - 1] + c[-SP + 1] + c[SP - 1] + c[SP + 1]);
      }
      lpf[v * PH + k] = r;
    }
  }
  __syncthreads();  // vd/hd are dead from here on: the union region is reused for pd/qd/pq/grb

  // ---- step 4.1: squared P/Q diagonal high-pass on odd columns of every row (rcd.cu:149-163), with the stale cells
  {
    constexpr int NR = P - 8, NC = PH - 4;  // rows 4..P-5, cells 2..PH-3 (pixels 5..P-5)
    for (int i = tid; i < NR * NC; i += kThreads) {
      const int v = 4 + i / NC, k = 2 + i % NC;
      const int u = 2 * k + 1;
      const int gx = gx0 + u, gy = gy0 + v;
      float p = 0.0f, q = 0.0f;
      if (gy >= 3 && gy <= height - 4 && gx >= 3 && gx <= width - 4) {
        const float *c = cfa + v * SP + u;
        p = sqr((c[-3 * SP - 3] - c[-SP - 1] - c[SP + 1] + c[3 * SP + 3]) - 3.0f * (c[-2 * SP - 2] + c[2 * SP + 2]) + 6.0f * c[0]);
        q = sqr((c[-3 * SP + 3] - c[-SP + 1] - c[SP - 1] + c[3 * SP - 3]) - 3.0f * (c[-2 * SP + 2] + c[2 * SP - 2]) + 6.0f * c[0]);
      } else if (gy >= 0 && gy < height && gx >= 0 && gx < width) {
        p = stale_cell(src, true, gy, gx >> 1, width, height);
        q = stale_cell(src, false, gy, gx >> 1, width, height);
      }
      pd[v * PH + k] = p, qd[v * PH + k] = q;
    }
  }
  // ---- step 3.1: green at R/B sites (rcd.cu:107-146); independent of step 4.1
  {
    constexpr int NR = P - 10, NC = PH - 4;  // rows 5..P-6, cells 2..PH-3  (sites 4/5 .. P-6/P-5; guarded below)
    for (int i = tid; i < NR * NC; i += kThreads) {
      const int v = 5 + i / NC, k = 2 + i % NC;
      const int gy = gy0 + v;
      const int u = 2 * k + (fc(gy & 1, 0, filters) & 1);
      const int gx = gx0 + u;
      float r = 0.0f;
      if (u >= 5 && u <= P - 6 && gy >= 4 && gy <= height - 5 && gx >= 4 && gx <= width - 5) {
        const float eps = 1e-5f;
        const float *c = cfa + v * SP + u;
        const float *d = vh + v * SP + u;
        const float c0 = d[0];
        const float nb = 0.25f * (d[-SP - 1] + d[-SP + 1] + d[SP - 1] + d[SP + 1]);
        const float disc = (fabsf(0.5f - c0) < fabsf(0.5f - nb)) ? nb : c0;
        const float ci = c[0];
        const float Ng = eps + fabsf(c[-SP] - c[SP]) + fabsf(ci - c[-2 * SP]) + fabsf(c[-SP] - c[-3 * SP]) + fabsf(c[-2 * SP] - c[-4 * SP]);
        const float Sg = eps + fabsf(c[SP] - c[-SP]) + fabsf(ci - c[2 * SP]) + fabsf(c[SP] - c[3 * SP]) + fabsf(c[2 * SP] - c[4 * SP]);
        const float Wg = eps + fabsf(c[-1] - c[1]) + fabsf(ci - c[-2]) + fabsf(c[-1] - c[-3]) + fabsf(c[-2] - c[-4]);
        const float Eg = eps + fabsf(c[1] - c[-1]) + fabsf(ci - c[2]) + fabsf(c[1] - c[3]) + fabsf(c[2] - c[4]);
        const float *l = lpf + v * PH + k;
        const float li = l[0];
        const float Ne = c[-SP] * (li + li) / (eps + li + l[-2 * PH]);
        const float Se = c[SP] * (li + li) / (eps + li + l[2 * PH]);
        const float We = c[-1] * (li + li) / (eps + li + l[-1]);
        const float Ee = c[1] * (li + li) / (eps + li + l[1]);
        const float Ve = (Sg * Ne + Ng * Se) / (Ng + Sg);
        const float He = (Wg * Ee + Eg * We) / (Eg + Wg);
        r = mixf(Ve, He, disc);
      }
      grb[v * PH + k] = r;
    }
  }
  __syncthreads();
  // ---- step 4.2: PQ_dir at R/B sites (rcd.cu:166-182), literal half-cell arithmetic
  {
    constexpr int NR = P - 12, NC = PH - 6;  // rows 6..P-7, cells 3..PH-4
    for (int i = tid; i < NR * NC; i += kThreads) {
      const int v = 6 + i / NC, k = 3 + i % NC;
      const int gy = gy0 + v;
      const int e = fc(gy & 1, 0, filters) & 1;
      const int gx = gx0 + 2 * k + e;
      float r = 0.0f;
      if (gy >= 2 && gy <= height - 3 && gx >= 2 && gx <= width - 3) {
        const int i2 = v * PH + k, i3 = (v - 1) * PH + k - 1 + e, i4 = (v + 1) * PH + k - 1 + e;
        const float Ps = fmaxf(1e-10f, pd[i3] + pd[i2] + pd[i4 + 1]);
        const float Qs = fmaxf(1e-10f, qd[i3 + 1] + qd[i2] + qd[i4]);
        r = Ps / (Ps + Qs);
      }
      pq[v * PH + k] = r;
    }
  }
  __syncthreads();
  // ---- step 5.1: the opposite colour at R/B sites along the diagonals (rcd.cu:185-224)
  {
    constexpr int NR = P - 14, NC = PH - 6;  // rows 7..P-8, cells 3..PH-4 (sites guarded to 7..P-8)
    for (int i = tid; i < NR * NC; i += kThreads) {
      const int v = 7 + i / NC, k = 3 + i % NC;
      const int gy = gy0 + v;
      const int e = fc(gy & 1, 0, filters) & 1;
      const int u = 2 * k + e;
      const int gx = gx0 + u;
      float r = 0.0f;
      if (u >= 7 && u <= P - 8 && gy >= 4 && gy <= height - 4 && gx >= 4 && gx <= width - 4) {
        const float eps = 1e-5f;
        const int q1 = v * PH + k, q2 = (v - 1) * PH + k - 1 + e, q3 = (v + 1) * PH + k - 1 + e;
        const float c0 = pq[q1];
        const float nb = 0.25f * (pq[q2] + pq[q2 + 1] + pq[q3] + pq[q3 + 1]);
        const float disc = (fabsf(0.5f - c0) < fabsf(0.5f - nb)) ? nb : c0;
        const float *c = cfa + v * SP + u;  // the opposite colour is native at every diagonal neighbour
        // G at R/B sites: centre, diagonal +-1 (cells of the adjacent rows) and diagonal +-2 (same parity rows)
        const float g0 = grb[q1];
        const float gNW = grb[(v - 1) * PH + ((u - 1) >> 1)], gNE = grb[(v - 1) * PH + ((u + 1) >> 1)];
        const float gSW = grb[(v + 1) * PH + ((u - 1) >> 1)], gSE = grb[(v + 1) * PH + ((u + 1) >> 1)];
        const float gNW2 = grb[(v - 2) * PH + k - 1], gNE2 = grb[(v - 2) * PH + k + 1];
        const float gSW2 = grb[(v + 2) * PH + k - 1], gSE2 = grb[(v + 2) * PH + k + 1];
        const float cNW = c[-SP - 1], cNE = c[-SP + 1], cSW = c[SP - 1], cSE = c[SP + 1];
        const float NWg = eps + fabsf(cNW - cSE) + fabsf(cNW - c[-3 * SP - 3]) + fabsf(g0 - gNW2);
        const float NEg = eps + fabsf(cNE - cSW) + fabsf(cNE - c[-3 * SP + 3]) + fabsf(g0 - gNE2);
        const float SWg = eps + fabsf(cNE - cSW) + fabsf(cSW - c[3 * SP - 3]) + fabsf(g0 - gSW2);
        const float SEg = eps + fabsf(cNW - cSE) + fabsf(cSE - c[3 * SP + 3]) + fabsf(g0 - gSE2);
        const float NWe = cNW - gNW, NEe = cNE - gNE, SWe = cSW - gSW, SEe = cSE - gSE;
        const float Pe = (NWg * SEe + SEg * NWe) / (NWg + SEg);
        const float Qe = (NEg * SWe + SWg * NEe) / (NEg + SWg);
        r = g0 + mixf(Pe, Qe, disc);
      }
      crb[v * PH + k] = r;
    }
  }
  __syncthreads();
  // ---- step 5.2 at green sites + output assembly inside the 7-px margin (rcd.cu:227-282, :49-60)
  const bool edge_tile = (x0 < 7) || (y0 < 7) || (x0 + T > width - 7) || (y0 + T > height - 7);
  for (int i = tid; i < T * T; i += kThreads) {
    const int ly = i / T, lx = i - ly * T;
    const int v = HALO + ly, u = HALO + lx;
    const int gx = x0 + lx, gy = y0 + ly;
    float R = 0.0f, G = 0.0f, B = 0.0f;
    if (gx >= 7 && gx < width - 7 && gy >= 7 && gy < height - 7) {
      const int col = fc(gy & 1, gx & 1, filters);
      const float *c = cfa + v * SP + u;
      const int k = u >> 1;
      if (col != 1) {
        const float own = c[0], g = grb[v * PH + k], opp = crb[v * PH + k];
        G = g;
        if (col == 0) R = own, B = opp; else B = own, R = opp;
      } else {
        const float eps = 1e-5f;
        const float *d = vh + v * SP + u;
        const float c0 = d[0];
        const float nb = 0.25f * (d[-SP - 1] + d[-SP + 1] + d[SP - 1] + d[SP + 1]);
        const float disc = (fabsf(0.5f - c0) < fabsf(0.5f - nb)) ? nb : c0;
        const float g = c[0];
        const float N1 = eps + fabsf(g - c[-2 * SP]), S1 = eps + fabsf(g - c[2 * SP]);
        const float W1 = eps + fabsf(g - c[-2]), E1 = eps + fabsf(g - c[2]);
        const float gN = grb[(v - 1) * PH + k], gS = grb[(v + 1) * PH + k];
        const float gW = grb[v * PH + ((u - 1) >> 1)], gE = grb[v * PH + ((u + 1) >> 1)];
        // colour of the horizontal neighbours; the vertical ones carry the other colour
        const int hcol = fc(gy & 1, (gx + 1) & 1, filters);
        float res[2];
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
          const int want = pass == 0 ? 0 : 2;  // plane: 0 = red, 2 = blue
          float n1, s1, w1, e1, n3, s3, w3, e3;
          if (hcol == want) {  // native left/right, interpolated (step 5.1) above/below
            w1 = c[-1], e1 = c[1], w3 = c[-3], e3 = c[3];
            n1 = crb[(v - 1) * PH + k], s1 = crb[(v + 1) * PH + k], n3 = crb[(v - 3) * PH + k], s3 = crb[(v + 3) * PH + k];
          } else {
            n1 = c[-SP], s1 = c[SP], n3 = c[-3 * SP], s3 = c[3 * SP];
            w1 = crb[v * PH + ((u - 1) >> 1)], e1 = crb[v * PH + ((u + 1) >> 1)];
            w3 = crb[v * PH + ((u - 3) >> 1)], e3 = crb[v * PH + ((u + 3) >> 1)];
          }
          const float SN = fabsf(n1 - s1), EW = fabsf(w1 - e1);
          const float Ng = N1 + SN + fabsf(n1 - n3), Sg = S1 + SN + fabsf(s1 - s3);
          const float Wg = W1 + EW + fabsf(w1 - w3), Eg = E1 + EW + fabsf(e1 - e3);
          const float Ne = n1 - gN, Se = s1 - gS, We = w1 - gW, Ee = e1 - gE;
          const float Ve = (Ng * Se + Sg * Ne) / (Ng + Sg);
          const float He = (Eg * We + Wg * Ee) / (Eg + Wg);
          res[pass] = g + mixf(Ve, He, disc);
        }
        R = res[0], G = g, B = res[1];
      }
      R = fmaxf(R, 0.0f), G = fmaxf(G, 0.0f), B = fmaxf(B, 0.0f);
    }
    outt[3 * i] = R, outt[3 * i + 1] = G, outt[3 * i + 2] = B;
  }

  if (edge_tile) {
    // ---- the 7-px frame: PPG-style border (rcd.cu:616-631 = border_interpolate(3), border green, border red/blue)
    __syncthreads();         // pd/qd/pq/grb are dead; the union region now holds the per-pixel RGB tmp plane
    float *tmp = sm + OFF_U;  // PT*PT*3
    for (int i = tid; i < PT * PT; i += kThreads) {
      const int ty = i / PT, tx = i - ty * PT;
      const int x = x0 - 1 + tx, y = y0 - 1 + ty;
      float r = 0.0f, g = 0.0f, b = 0.0f;
      if (x >= 0 && y >= 0 && x < width && y < height) {
        const int c = fc(y & 1, x & 1, filters);
        const float *p = cfa + (ty - 1 + HALO) * SP + (tx - 1 + HALO);
        if (x < 3 || y < 3 || x >= width - 3 || y >= height - 3) {
          float sum[3] = {0, 0, 0};
          int cnt[3] = {0, 0, 0};
#pragma unroll
          for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
              const int xx = x + dx, yy = y + dy;
              if (xx >= 0 && yy >= 0 && xx < width && yy < height) {
                const int f = fc(yy & 1, xx & 1, filters);
                const float v = p[dy * SP + dx];
                sum[0] += f == 0 ? v : 0.0f, sum[1] += f == 1 ? v : 0.0f, sum[2] += f == 2 ? v : 0.0f;
                cnt[0] += f == 0, cnt[1] += f == 1, cnt[2] += f == 2;
              }
            }
          const float v = p[0];
          r = cnt[0] > 0 ? sum[0] / cnt[0] : v;
          g = cnt[1] > 0 ? sum[1] / cnt[1] : v;
          b = cnt[2] > 0 ? sum[2] / cnt[2] : v;
          if (c == 0) r = v; else if (c == 2) b = v; else g = v;
        } else {
          const float pc = p[0];
          if (c == 0) r = pc; else if (c == 2) b = pc; else g = pc;
          if (c != 1) {
            const float pym = p[-SP], pym2 = p[-2 * SP], pym3 = p[-3 * SP], pyM = p[SP], pyM2 = p[2 * SP], pyM3 = p[3 * SP];
            const float pxm = p[-1], pxm2 = p[-2], pxm3 = p[-3], pxM = p[1], pxM2 = p[2], pxM3 = p[3];
            const float guessx = (pxm + pc + pxM) * 2.0f - pxM2 - pxm2;
            const float diffx = (fabsf(pxm2 - pc) + fabsf(pxM2 - pc) + fabsf(pxm - pxM)) * 3.0f + (fabsf(pxM3 - pxM) + fabsf(pxm3 - pxm)) * 2.0f;
            const float guessy = (pym + pc + pyM) * 2.0f - pyM2 - pym2;
            const float diffy = (fabsf(pym2 - pc) + fabsf(pyM2 - pc) + fabsf(pym - pyM)) * 3.0f + (fabsf(pyM3 - pyM) + fabsf(pym3 - pym)) * 2.0f;
            if (diffx > diffy) g = fmaxf(fminf(guessy * 0.25f, fmaxf(pym, pyM)), fminf(pym, pyM));
            else g = fmaxf(fminf(guessx * 0.25f, fmaxf(pxm, pxM)), fminf(pxm, pxM));
          }
          r = fmaxf(r, 0.0f), g = fmaxf(g, 0.0f), b = fmaxf(b, 0.0f);
        }
      }
      tmp[3 * i] = r, tmp[3 * i + 1] = g, tmp[3 * i + 2] = b;
    }
    __syncthreads();
    for (int i = tid; i < T * T; i += kThreads) {
      const int ly = i / T, lx = i - ly * T;
      const int x = x0 + lx, y = y0 + ly;
      if (x >= width || y >= height) continue;
      if (x >= 7 && x < width - 7 && y >= 7 && y < height - 7) continue;  // RCD proper owns the interior
      const float *p = tmp + 3 * ((ly + 1) * PT + lx + 1);
      float r = p[0], g = p[1], b = p[2];
      if (!(x == 0 || y == 0 || x == width - 1 || y == height - 1)) {
        const int c = fc(y & 1, x & 1, filters);
        constexpr int R = 3 * PT;
        if (c == 1) {
          const float *nt = p - R, *nb = p + R, *nl = p - 3, *nr = p + 3;
          if (fc(y & 1, (x + 1) & 1, filters) == 0) {
            b = (nt[2] + nb[2] + 2.0f * g - nt[1] - nb[1]) * 0.5f;
            r = (nl[0] + nr[0] + 2.0f * g - nl[1] - nr[1]) * 0.5f;
          } else {
            r = (nt[0] + nb[0] + 2.0f * g - nt[1] - nb[1]) * 0.5f;
            b = (nl[2] + nr[2] + 2.0f * g - nl[1] - nr[1]) * 0.5f;
          }
        } else {
          const float *ntl = p - R - 3, *ntr = p - R + 3, *nbl = p + R - 3, *nbr = p + R + 3;
          const int k = (c == 0) ? 2 : 0;
          const float diff1 = fabsf(ntl[k] - nbr[k]) + fabsf(ntl[1] - g) + fabsf(nbr[1] - g);
          const float guess1 = ntl[k] + nbr[k] + 2.0f * g - ntl[1] - nbr[1];
          const float diff2 = fabsf(ntr[k] - nbl[k]) + fabsf(ntr[1] - g) + fabsf(nbl[1] - g);
          const float guess2 = ntr[k] + nbl[k] + 2.0f * g - ntr[1] - nbl[1];
          const float val = diff1 > diff2 ? guess2 * 0.5f : (diff1 < diff2 ? guess1 * 0.5f : (guess1 + guess2) * 0.25f);
          if (c == 0) b = val; else r = val;
        }
      }
      outt[3 * i] = fmaxf(r, 0.0f), outt[3 * i + 1] = fmaxf(g, 0.0f), outt[3 * i + 2] = fmaxf(b, 0.0f);
    }
  }
  __syncthreads();
  store_rgb_tile(outt, T * 3, rgb, x0, y0, T, T, width, height);
}

__global__ void __launch_bounds__(kThreads256) rcd_kernel(CfaSource src, float *__restrict__ rgb, int width, int height, uint32_t filters,
                                                       TileRects rects) {
  extern __shared__ __align__(128) float sm[];
  rcd_tile<kThreads256>(sm, src, rgb, width, height, filters, rects, blockIdx.x, blockIdx.x, blockIdx.y);
}


// ======================================================================================================================
// Interior tiles: rcd_planar.cuh (64 x 32 tiles, 2 x 4-pixel register blocks, phase-planar shared memory).  Only tiles whose 88 x 56
// patch lies inside the image take that path, which removes every bounds test; the frame of 32 x 32 tiles around them (and the
// PPG-style 7-px border) stays with the kernel above and rides at the end of the same grid: as a launch of its own the frame is
// latency-bound (504 CTAs at 4K), at the end of the interior grid its CTAs fill the slots that drain.
template <bool kG0>
__global__ void __launch_bounds__(v3::kThreads3, 2) rcd3_kernel(CfaSource src, float *__restrict__ rgb, int width, int height,
                                                                uint32_t filters, int x_origin, int by_lo, int nbx, int n_interior,
                                                                TileRects rects) {
  extern __shared__ __align__(128) float sm[];
  const int b = blockIdx.x;
  if (b < n_interior) v3::rcd3_tile<kG0>(sm, src, rgb, width, height, filters, x_origin, by_lo, b % nbx, b / nbx);
  else rcd_tile<kThreads256>(sm, src, rgb, width, height, filters, rects, b - n_interior, 0, 0);
}

// ======================================================================================================================
// Interior as column strips: rcd_strip.cuh.  A strip job = (strip, segment of its rows); the 32 x 32 frame tiles ride in the same
// grid (its tail by default), run by the same 160-thread CTAs.
template <bool kG0>
__global__ void __launch_bounds__(v4::NT, 3) rcd_strip_kernel(const __grid_constant__ CUtensorMap tmap, const v4::StripArgs a, const TileRects rects) {
  extern __shared__ __align__(128) float sm[];
  int b = blockIdx.x;
  const bool frame = a.frame_first ? b < a.n_frame : b >= a.n_strip_jobs;
  if (frame) {
    rcd_tile<v4::NT>(sm, a.src, a.rgb, a.width, a.height, a.filters, rects, a.frame_first ? b : b - a.n_strip_jobs, 0, 0);
    return;
  }
  if (a.frame_first) b -= a.n_frame;
  v4::rcd_strip<kG0>(sm, &tmap, a, b);
}

int env_int(const char *name, int fallback) {
  const char *e = getenv(name);
  return e ? atoi(e) : fallback;
}
}  // namespace

int launch_rcd(const CfaSource &src, float *rgb, int width, int height, uint32_t filters, cudaStream_t s) {
  static DeviceOnce attr;
  constexpr size_t bytes = SMEM_FLOATS * sizeof(float);
  constexpr size_t bytes3 = v3::SMEM_FLOATS * sizeof(float);
  static_assert(v3::SMEM_FLOATS >= SMEM_FLOATS && v3::kThreads3 == kThreads256, "the frame tiles run inside the interior kernel's CTAs");
  attr.run([&] {
    cudaFuncSetAttribute(rcd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    cudaFuncSetAttribute(rcd3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes3);
    cudaFuncSetAttribute(rcd3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes3);
  });
  // interior tiles (64 x 32) whose 88 x 56 patch lies inside the image; needs 16-byte aligned rows on both sides.  The tiling starts
  // at x = 32 so that the frame left to the 32 x 32 kernel is a ring of single tiles
  const int x_origin = T;
  const int nbx = (width - x_origin - (v3::TW + v3::HX)) / v3::TW + 1;   // x_origin + 64 bx + 76 <= width
  const int by_lo = 1, by_hi = (height - (v3::TH + v3::HY)) / v3::TH;    // 32 by + 44 <= height
  const bool aligned = (width % 4 == 0) && (reinterpret_cast<uintptr_t>(rgb) % 16 == 0) &&
                       (src.cfa ? reinterpret_cast<uintptr_t>(src.cfa) % 16 == 0
                                : (reinterpret_cast<uintptr_t>(src.packed) % 4 == 0 && ((int64_t)width * height * 3 / 2) % 4 == 0));
  TileRects rects{};
  const int ntx = div_up(width, T), nty = div_up(height, T);

  // ---- strips (rcd_strip.cuh): rows [32, y_end) of the columns [32, 32 + 64 nbx), y_end a multiple of 32 with 16 rows to spare below
  // (the staging reads 13 rows ahead of the output, and a packed row copy may be rounded up to the next 16-byte boundary)
  static const int use_strips = env_int("TDB_RCD_STRIPS", 1);
  const int y_end = (height - 16) / 32 * 32;
  const bool strip_align = src.cfa ? true : reinterpret_cast<uintptr_t>(src.packed) % 16 == 0;
  CUtensorMap tmap;
  const bool have_map = make_tensor_map_f32(&tmap, src.cfa, width, height, v4::PW, v4::R);  // zeroed (and unused) for a packed source
  if (use_strips && aligned && strip_align && nbx >= 1 && y_end >= 64 && (!src.cfa || have_map)) {
    v4::StripArgs a{};
    a.src = src, a.rgb = rgb, a.width = width, a.height = height, a.filters = filters, a.x_origin = x_origin;
    a.nstrips = nbx, a.y_start = 32, a.n_iter = (y_end - 32) / v4::R;
    // frame tiles: top, bottom, left, right of the strip area
    const int ix0 = x_origin / T, ix1 = (x_origin + nbx * v4::TW) / T, iy0 = 1, iy1 = y_end / T;
    const int rx0[4] = {0, 0, 0, ix1}, ry0[4] = {0, iy1, iy0, iy0};
    const int rnx[4] = {ntx, ntx, ix0, ntx - ix1}, rny[4] = {iy0, nty - iy1, iy1 - iy0, iy1 - iy0};
    int n = 0, total = 0;
    for (int k = 0; k < 4; k++) {
      if (rnx[k] <= 0 || rny[k] <= 0) continue;
      rects.tx0[n] = rx0[k], rects.ty0[n] = ry0[k], rects.ntx[n] = rnx[k], rects.start[n] = total;
      total += rnx[k] * rny[k], n++;
    }
    for (int k = n; k < 5; k++) rects.start[k] = total;
    for (int k = n; k < 4; k++) rects.ntx[k] = 1;
    rects.n = n;
    a.n_frame = total;
    // segments per strip: every segment pays ~1.1 iterations of pipeline fill; 3 CTAs x 148 SMs run at a time; the frame tiles (about
    // 2.2 iterations of work each) fill the slots the strips leave free and whatever remains of them runs after the last wave
    static const int forced = env_int("TDB_RCD_SEGMENTS", 0);
    int best = 1;
    double best_cost = 1e30;
    const double slots = 3.0 * kNumSMs, frame_work = 2.2 * total;
    for (int sgm = 1; sgm <= 64 && sgm <= a.n_iter; sgm++) {
      const double jobs = (double)nbx * sgm, len = (double)a.n_iter / sgm + 1.1;
      const double waves = ceil(jobs / slots);
      const double free_work = (waves * slots - jobs) * len;
      const double cost = waves * len + (frame_work > free_work ? 2.2 + (frame_work - free_work) / slots : 0.0);
      if (cost < best_cost) best_cost = cost, best = sgm;
    }
    a.nseg = forced > 0 ? (forced < a.n_iter ? forced : a.n_iter) : best;
    a.n_strip_jobs = a.nstrips * a.nseg;
    static const int frame_first = env_int("TDB_RCD_FRAME_FIRST", 0);
    a.frame_first = frame_first;
    constexpr size_t bytes4 = v4::SMEM_FLOATS * sizeof(float);
    static_assert(v4::SMEM_FLOATS >= SMEM_FLOATS, "the frame tiles run inside the strip kernel's CTAs");
    static DeviceOnce attr4;
    attr4.run([&] {
      cudaFuncSetAttribute(rcd_strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes4);
      cudaFuncSetAttribute(rcd_strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes4);
      // three CTAs of 74.6 KB (+ 1 KB each for the system) need 227 of the SM's 228 KB as shared memory
      cudaFuncSetAttribute(rcd_strip_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(rcd_strip_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    });
    const int grid4 = a.n_strip_jobs + a.n_frame;
    if (fc(0, 0, filters) == 1) rcd_strip_kernel<true><<<grid4, v4::NT, bytes4, s>>>(tmap, a, rects);
    else rcd_strip_kernel<false><<<grid4, v4::NT, bytes4, s>>>(tmap, a, rects);
    return check_launch("rcd_demosaic");
  }

  if (aligned && width >= x_origin + v3::TW + v3::HX && nbx >= 1 && by_hi >= by_lo) {
    // the frame of 32 x 32 tiles around the interior: top, bottom, left, right
    const int ix0 = x_origin / T, ix1 = (x_origin + nbx * v3::TW) / T, iy0 = by_lo * v3::TH / T, iy1 = (by_hi + 1) * v3::TH / T;
    const int rx0[4] = {0, 0, 0, ix1}, ry0[4] = {0, iy1, iy0, iy0};
    const int rnx[4] = {ntx, ntx, ix0, ntx - ix1}, rny[4] = {iy0, nty - iy1, iy1 - iy0, iy1 - iy0};
    int n = 0, total = 0;
    for (int k = 0; k < 4; k++) {
      if (rnx[k] <= 0 || rny[k] <= 0) continue;
      rects.tx0[n] = rx0[k], rects.ty0[n] = ry0[k], rects.ntx[n] = rnx[k], rects.start[n] = total;
      total += rnx[k] * rny[k], n++;
    }
    for (int k = n; k < 5; k++) rects.start[k] = total;
    for (int k = n; k < 4; k++) rects.ntx[k] = 1;
    rects.n = n;
    const int n_interior = nbx * (by_hi - by_lo + 1);
    const int grid2 = n_interior + total;
    if (fc(0, 0, filters) == 1)
      rcd3_kernel<true><<<grid2, kThreads256, bytes3, s>>>(src, rgb, width, height, filters, x_origin, by_lo, nbx, n_interior, rects);
    else
      rcd3_kernel<false><<<grid2, kThreads256, bytes3, s>>>(src, rgb, width, height, filters, x_origin, by_lo, nbx, n_interior, rects);
    return check_launch("rcd_demosaic");
  }
  dim3 grid(ntx, nty);
  rcd_kernel<<<grid, kThreads256, bytes, s>>>(src, rgb, width, height, filters, rects);
  return check_launch("rcd_demosaic");
}

}  // namespace tdb

using namespace tdb;

extern "C" int tdb_rcd(const float *cfa, float *rgb, int width, int height, uint32_t filters, tdb_stream_t stream) {
  TDB_REQUIRE(cfa && rgb, "RCD: null pointer");
  TDB_REQUIRE(width >= 16 && height >= 16 && !(width & 1) && !(height & 1), "RCD: width and height must be even and >= 16 (got %dx%d)", width, height);
  CfaSource src{};
  src.cfa = cfa;
  return launch_rcd(src, rgb, width, height, filters, as_stream(stream));
}
