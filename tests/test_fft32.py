"""The register FFT templates of csrc/fft32.cuh, compiled as host code with g++ and checked against a direct DFT.

fft_fwd leaves frequency k in position brev(k); fft_inv undoes it (unscaled).  CPU-only test: the same header is what
the Wiener kernel instantiates on the device."""

from pathlib import Path
import subprocess

ROOT = Path(__file__).resolve().parent.parent
HARNESS = r'''
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include "fft_quad.cuh"
using namespace tdb::fft;
template <int N> double check() {
  double err = 0;
  for (int trial = 0; trial < 20; trial++) {
    float re[N], im[N], r0[N], i0[N];
    for (int i = 0; i < N; i++) re[i] = r0[i] = rand() / (float)RAND_MAX - 0.5f, im[i] = i0[i] = rand() / (float)RAND_MAX - 0.5f;
    fft_fwd<N>(re, im);
    for (int k = 0; k < N; k++) {
      double sr = 0, si = 0;
      for (int n = 0; n < N; n++) {
        const double a = -2 * M_PI * k * n / N;
        sr += r0[n] * cos(a) - i0[n] * sin(a), si += r0[n] * sin(a) + i0[n] * cos(a);
      }
      const int p = brev<N>(k);
      err = fmax(err, fmax(fabs(sr - re[p]), fabs(si - im[p])));
    }
    fft_inv<N>(re, im);
    for (int n = 0; n < N; n++) err = fmax(err, fmax(fabs(re[n] / N - r0[n]), fabs(im[n] / N - i0[n])));
  }
  return err;
}
// the four-lane 32-point transform of fft_quad.cuh on four emulated lanes
double check_quad() {
  cpx tw[32];
  make_quad_twiddles(tw);
  double err = 0;
  for (int trial = 0; trial < 20; trial++) {
    float x_re[32], x_im[32], re[4][8], im[4][8];
    for (int n = 0; n < 32; n++) x_re[n] = rand() / (float)RAND_MAX - 0.5f, x_im[n] = rand() / (float)RAND_MAX - 0.5f;
    for (int j = 0; j < 4; j++)
      for (int m = 0; m < 8; m++) re[j][m] = x_re[4 * m + j], im[j][m] = x_im[4 * m + j];
    auto exchange = [&](int mask, void (*fn)(float &, float &, float, float, int)) {
      float pr[4][8], pi[4][8];
      for (int j = 0; j < 4; j++)
        for (int p = 0; p < 8; p++) pr[j][p] = re[j ^ mask][p], pi[j][p] = im[j ^ mask][p];
      for (int j = 0; j < 4; j++)
        for (int p = 0; p < 8; p++) fn(re[j][p], im[j][p], pr[j][p], pi[j][p], j);
    };
    for (int j = 0; j < 4; j++) quad_fwd_local(re[j], im[j], tw + 8 * j);
    exchange(2, quad_fwd_a);
    exchange(1, quad_fwd_b);
    for (int k = 0; k < 32; k++) {
      double sr = 0, si = 0;
      for (int n = 0; n < 32; n++) {
        const double a = -2 * M_PI * k * n / 32;
        sr += x_re[n] * cos(a) - x_im[n] * sin(a), si += x_re[n] * sin(a) + x_im[n] * cos(a);
      }
      const int p = quad_reg_of(k), j = quad_lane_of(k);
      if (quad_freq(p, j) != k) return 1.0;
      err = fmax(err, fmax(fabs(sr - re[j][p]), fabs(si - im[j][p])));
    }
    exchange(1, quad_inv_b);
    exchange(2, quad_inv_a);
    for (int j = 0; j < 4; j++) quad_inv_local(re[j], im[j], tw + 8 * j);
    for (int j = 0; j < 4; j++)
      for (int m = 0; m < 8; m++)
        err = fmax(err, fmax(fabs(re[j][m] / 32 - x_re[4 * m + j]), fabs(im[j][m] / 32 - x_im[4 * m + j])));
  }
  return err;
}
int main() { printf("%.3e %.3e %.3e %.3e\n", check<32>(), check<16>(), check<8>(), check_quad()); return 0; }
'''


def test_fft32_matches_dft(tmp_path):
  src = tmp_path / 'fft_harness.cpp'
  src.write_text(HARNESS)
  exe = tmp_path / 'fft_harness'
  subprocess.run(['g++', '-std=c++17', '-O2', '-ffp-contract=off', '-I', str(ROOT / 'torch-darktable_b200' / 'csrc'), str(src), '-o', str(exe)],
                 check=True)
  e32, e16, e8, equad = map(float, subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split())
  assert e32 < 2e-6 and e16 < 1e-6 and e8 < 1e-6, (e32, e16, e8)
  assert equad < 2e-6, equad  # fft_quad.cuh: 8-point register FFTs + two exchange stages across four lanes
