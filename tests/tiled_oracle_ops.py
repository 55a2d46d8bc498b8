"""CPU stand-in for the stage kernels behind pipeline/tiled.py, built on the oracle (test infrastructure only).

Lets the host logic of the row-tile split -- partitioning, halo exchange, the three reductions, cropping -- run on CPU tensors
under gloo or between threads, with the oracle's untiled Pipeline as the expected result."""

import ctypes as C

import numpy as np
import torch

import oracle
import synth


def _np(t):
  return t.detach().cpu().numpy()


def _t(a):
  return torch.from_numpy(np.ascontiguousarray(a))


class OracleOps:
  def demosaic(self, packed, size, pattern, fmt, settings, white_balance):
    w, h = size
    cfa = oracle.decode12(_np(packed), np.float32, fmt.name.endswith('IDS')).reshape(h, w)
    if white_balance is not None:
      cfa = oracle.white_balance(cfa, _np(white_balance), pattern.name)
    method = settings.debayer.name
    rgb = {'rcd': lambda: oracle.rcd(cfa, pattern.name), 'ppg': lambda: oracle.ppg(cfa, pattern.name, settings.ppg_median_threshold),
           'bilinear': lambda: oracle.bilinear5x5(cfa, pattern.name)}[method]()
    return _t(rgb)

  def smooth(self, rgb, size, pattern, passes):
    return _t(oracle.postprocess(_np(rgb), pattern.name, passes, False, False, 0.04))

  def green_sums(self, rgb, pattern):
    a = _np(rgb)
    h, w = a.shape[0] & ~1, a.shape[1] & ~1
    rows, cols = np.mgrid[0:h, 0:w]
    green = synth.fc(rows, cols, synth.PATTERNS[pattern.name]) == 1
    g = a[:h, :w, 1].astype(np.float64)
    return _t(np.array([g[green & (rows % 2 == 0)].sum(), g[green & (rows % 2 == 1)].sum()], np.float32))

  def green_eq_apply(self, rgb, ratio, pattern):
    a = _np(rgb).copy()
    rows, cols = np.mgrid[0:a.shape[0], 0:a.shape[1]]
    g1 = (synth.fc(rows, cols, synth.PATTERNS[pattern.name]) == 1) & (rows % 2 == 0)
    a[..., 1] = np.where(g1, a[..., 1] * np.float32(_np(ratio)[0]), a[..., 1])
    return _t(np.maximum(a, np.float32(0)).astype(np.float32))

  def bounds(self, rgb, stride): return _t(oracle.compute_image_bounds([_np(rgb)], stride))

  def normalize(self, rgb, bounds):
    b = _np(bounds)
    return _t(((_np(rgb) - b[0]) / (b[1] - b[0])).astype(np.float32))

  def denoise(self, rgb, size, noise): return _t(oracle.wiener_log_luminance(_np(rgb), noise))
  def bilateral(self, rgb, size, sigma_s, sigma_r, detail): return _t(oracle.bilateral_rgb(_np(rgb), sigma_s, sigma_r, detail))

  def metric_sums(self, rgb, stride):
    a = np.ascontiguousarray(_np(rgb), np.float32)
    b = (C.c_float * 2)(0.0, 1.0)
    sums = (C.c_double * 6)()
    oracle.lib().orc_metrics_accumulate(a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], stride, C.c_float(1e-4), b, sums)
    return _t(np.array(list(sums), np.float32))

  def metrics_from_sums(self, sums):
    s = _np(sums)
    return _t((s[:5] * (np.float32(1.0) / np.float32(max(s[5], 1.0)))).astype(np.float32))

  def tonemap(self, rgb, op, metrics, params):
    return _t(oracle.tonemap(_np(rgb), op, None if metrics is None else _np(metrics), params.gamma, params.intensity, params.light_adapt,
                             params.vibrance))
