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
#include "fft32.cuh"
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
int main() { printf("%.3e %.3e\n", check<32>(), check<16>()); return 0; }
'''


def test_fft32_matches_dft(tmp_path):
  src = tmp_path / 'fft_harness.cpp'
  src.write_text(HARNESS)
  exe = tmp_path / 'fft_harness'
  subprocess.run(['g++', '-std=c++17', '-O2', '-ffp-contract=off', '-I', str(ROOT / 'torch-darktable_b200' / 'csrc'), str(src), '-o', str(exe)],
                 check=True)
  e32, e16 = map(float, subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split())
  assert e32 < 2e-6 and e16 < 1e-6, (e32, e16)
