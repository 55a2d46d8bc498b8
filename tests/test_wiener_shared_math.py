"""The algebra and the bookkeeping of the shared-column Wiener kernel (csrc/wiener.cu, namespace shr), restated in numpy and held to
the CPU oracle -- no GPU needed.

1. `shared_columns`: tiles paired vertically (tile rows oy and oy + 8 as real and imaginary part), one windowed column transform per
   image column and tile-row pair, the tile mean removed in the frequency domain, the four tiles covering a column summed in the
   frequency domain before one inverse column transform.  Must equal the oracle's tile-by-tile filter (reference
   csrc/denoise/denoise.cu:134-242) up to rounding.
2. `shared_columns_in_steps`: the same through the kernel's control flow -- the linearised (tile-row pair, step) sequence split
   evenly over "CTAs", an 88-column buffer per step, 24 columns of spectra and partial accumulators carried to the next step, full
   recomputation and flush where a CTA's range starts or ends in the middle of a tile row.  Any split must give the same image.
"""

import numpy as np
import pytest

import oracle
import synth

K, ST = 32, 8
TPS, NEWC, CARRY, BUFC = 8, 64, 24, 88


def window():
  half = K / 2.0
  r = np.arange(K, dtype=np.float32) - np.float32(half) + np.float32(0.5)
  w = np.exp(-(r * r) / np.float32(0.3 * half * half)).astype(np.float32)
  return (w / np.sqrt((w * w).sum(dtype=np.float32))).astype(np.float64)


def reflect(x, n):
  x = np.where(x < 0, -x, x)
  x = np.where(x >= n, 2 * n - x - 1, x)
  return np.clip(x, 0, n - 1)


def shrink_pair(z, sigma):
  """z = FFT2 of (tile_a + i tile_b): separate the two real tiles' spectra, apply gain = max(P - sigma^2, 0) / P, merge again."""
  zm = np.conj(np.roll(np.roll(z[::-1, ::-1], 1, axis=0), 1, axis=1))
  za, zb = (z + zm) / 2, (z - zm) / 2j

  def gain(x):
    p = np.abs(x) ** 2 + 1e-15
    return np.maximum(p - sigma * sigma, 0) / p

  return gain(za) * za + 1j * gain(zb) * zb


def column_spectra(img, oy, xs, w):
  h, wd = img.shape
  xr = reflect(xs, wd)
  z = img[np.ix_(reflect(oy + np.arange(K), h), xr)] + 1j * img[np.ix_(reflect(oy + ST + np.arange(K), h), xr)]
  return np.fft.fft(z * w[:, None], axis=0), z.sum(axis=0)


def tile_pair(spectra, sums, w, what, sigma):
  """One tile pair from the spectra / sums of its 32 columns: contribution to the accumulators of those columns."""
  mu = sums.sum() / (K * K)
  q = mu * what
  z = np.fft.fft((spectra - q[:, None]) * w[None, :], axis=1)
  b = np.fft.ifft(shrink_pair(z, sigma), axis=1) * K / (K * K)  # unnormalised inverse over kx, 1 / K^2 folded into the gains
  return w[None, :] * (b + q[:, None] * w[None, :] / 32)


def write_columns(acc, accu, oy, xs, w):
  h, wd = acc.shape
  y = np.fft.ifft(accu, axis=0) * K * w[:, None]
  ok = (xs >= 0) & (xs < wd)
  for r in range(K):
    if 0 <= oy + r < h:
      acc[oy + r, xs[ok]] += y[r, ok].real
    if 0 <= oy + ST + r < h:
      acc[oy + ST + r, xs[ok]] += y[r, ok].imag


def normalise(acc, w):
  h, wd = acc.shape
  m1 = np.array([sum(w[j] ** 2 for j in range(ph, K, ST)) for ph in range(ST)])
  return acc / (m1[np.arange(h) % ST][:, None] * m1[np.arange(wd) % ST][None, :] + 1e-15)


def shared_columns(img, sigma):
  h, wd = img.shape
  w = window()
  what = np.fft.fft(w)
  acc = np.zeros((h, wd))
  n_ty, n_tx = (h - 1 + CARRY) // ST + 1, (wd - 1 + CARRY) // ST + 1
  xs = np.arange(-CARRY, -CARRY + ST * (n_tx - 1) + K)
  for p in range((n_ty + 1) // 2):
    oy = -CARRY + 2 * ST * p
    spectra, sums = column_spectra(img, oy, xs, w)
    accu = np.zeros_like(spectra)
    for t in range(n_tx):
      accu[:, ST * t: ST * t + K] += tile_pair(spectra[:, ST * t: ST * t + K], sums[ST * t: ST * t + K], w, what, sigma)
    write_columns(acc, accu, oy, xs, w)
  return normalise(acc, w)


def shared_columns_in_steps(img, sigma, n_ctas):
  """The kernel's loop (wiener32_shared_kernel), one "CTA" after the other."""
  h, wd = img.shape
  w = window()
  what = np.fft.fft(w)
  acc = np.zeros((h, wd))
  n_ty, n_tx = (h - 1 + CARRY) // ST + 1, (wd - 1 + CARRY) // ST + 1
  steps_per_row = (n_tx + TPS - 1) // TPS
  total = steps_per_row * ((n_ty + 1) // 2)
  n_ctas = min(n_ctas, total)
  for cta in range(n_ctas):
    l0, l1 = cta * total // n_ctas, (cta + 1) * total // n_ctas
    spec = np.zeros((K, BUFC), complex)
    sums = np.zeros(BUFC, complex)
    accu = np.zeros((K, BUFC), complex)
    for l in range(l0, l1):
      p, k = divmod(l, steps_per_row)
      first, last = l == l0 or k == 0, l == l1 - 1 or k == steps_per_row - 1
      oy, cb = -CARRY + 2 * ST * p, -CARRY + NEWC * k
      xs = cb + np.arange(BUFC)
      new = slice(0 if first else CARRY, BUFC)
      if first:
        accu[:] = 0
      spec[:, new], sums[new] = column_spectra(img, oy, xs[new], w)
      for warp in range(TPS):
        if TPS * k + warp < n_tx:
          cols = slice(ST * warp, ST * warp + K)
          accu[:, cols] += tile_pair(spec[:, cols], sums[cols], w, what, sigma)
      done = slice(0, BUFC if last else NEWC)
      write_columns(acc, accu[:, done], oy, xs[done], w)
      if not last:  # carry the 24 columns the next step shares, clear the rest of the accumulators
        spec[:, :CARRY], sums[:CARRY], accu[:, :CARRY] = spec[:, NEWC:].copy(), sums[NEWC:].copy(), accu[:, NEWC:].copy()
        accu[:, CARRY:] = 0
  return normalise(acc, w)


def noisy(h, w, seed):
  rgb = synth.scene_rgb(h, w, seed)
  return (np.log(np.maximum(rgb[..., 1], 1e-4)) + np.random.default_rng(seed).normal(0, 0.05, (h, w))).astype(np.float32)


@pytest.mark.parametrize('h,w', [(72, 88), (33, 70), (41, 200)])
def test_shared_column_algebra_equals_the_tile_filter(h, w):
  x = noisy(h, w, 3)
  want = oracle.wiener(x[..., None], [0.075])[..., 0]
  got = shared_columns(x.astype(np.float64), 0.075)
  assert np.abs(got - want).max() < 3e-6  # the oracle works in float32


@pytest.mark.parametrize('h,w,n_ctas', [(72, 88, 1), (72, 88, 3), (41, 200, 2), (41, 200, 5), (100, 620, 7), (100, 620, 296)])
def test_step_sequence_and_carries_give_the_same_image(h, w, n_ctas):
  x = noisy(h, w, 4).astype(np.float64)
  want = shared_columns(x, 0.075)
  got = shared_columns_in_steps(x, 0.075, n_ctas)
  assert np.abs(got - want).max() < 1e-9
