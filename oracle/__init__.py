"""numpy front-end of the CPU oracle (oracle/*.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs.  The product package (torch-darktable_b200/torch_darktable) never imports it.
Parity pin: outputs of the reference extension itself, tests/golden/*.npz (see oracle.h).
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
import subprocess

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / 'liboracle.so'

PATTERNS = {'RGGB': 0x94949494, 'BGGR': 0x16161616, 'GRBG': 0x61616161, 'GBRG': 0x49494949}


def build(force: bool = False) -> Path:
  srcs = [_HERE / n for n in ('packed.c', 'demosaic.c', 'color.c', 'filters.c', 'oracle.h', 'Makefile')]
  stale = not _LIB_PATH.exists() or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
  if force or stale:
    subprocess.run(['make', '-s', '-C', str(_HERE)], check=True, env={**os.environ, 'CC': ''})
  return _LIB_PATH


_lib = None


def lib():
  global _lib
  if _lib is None:
    build()
    _lib = C.CDLL(str(_LIB_PATH))
    _lib.orc_half_round.restype = C.c_float
    _lib.orc_half_round.argtypes = [C.c_float]
  return _lib


def _filters(pattern) -> int:
  if isinstance(pattern, str):
    return PATTERNS[pattern]
  return int(getattr(pattern, 'value', pattern))


def _f32(a):
  return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
  return a.ctypes.data_as(C.c_void_p)


# ---- packed codec -----------------------------------------------------------------------------
def decode12(packed: np.ndarray, dtype=np.float32, ids: bool = False, scaled: bool = True) -> np.ndarray:
  packed = np.ascontiguousarray(packed, dtype=np.uint8)
  npairs = packed.size // 3
  if dtype == np.float32:
    out = np.empty(2 * npairs, np.float32)
    lib().orc_decode12_f32(_p(packed), _p(out), C.c_long(npairs), int(ids), int(scaled))
  elif dtype == np.float16:
    bits = np.empty(2 * npairs, np.uint16)
    lib().orc_decode12_f16(_p(packed), _p(bits), C.c_long(npairs), int(ids), int(scaled))
    out = bits.view(np.float16)
  elif dtype == np.uint16:
    out = np.empty(2 * npairs, np.uint16)
    lib().orc_decode12_u16(_p(packed), _p(out), C.c_long(npairs), int(ids))
  else:
    raise ValueError(dtype)
  return out


def encode12(values: np.ndarray, ids: bool = False, scaled: bool = True) -> np.ndarray:
  npairs = values.size // 2
  out = np.empty(3 * npairs, np.uint8)
  if values.dtype == np.uint16:
    v = np.ascontiguousarray(values)
    lib().orc_encode12_u16(_p(v), _p(out), C.c_long(npairs), int(ids))
  else:
    v = _f32(values)
    lib().orc_encode12_f32(_p(v), _p(out), C.c_long(npairs), int(ids), int(scaled))
  return out


# ---- CFA ops ----------------------------------------------------------------------------------
def white_balance(bayer: np.ndarray, gains, pattern) -> np.ndarray:
  b = _f32(bayer)
  h, w = b.shape[:2]
  out = np.empty_like(b)
  g = (C.c_float * 3)(*[float(x) for x in gains])
  lib().orc_white_balance(_p(b), _p(out), w, h, C.c_uint32(_filters(pattern)), g)
  return out


def estimate_white_balance(images, pattern, quantile: float = 0.95, stride: int = 8) -> np.ndarray:
  ims = [_f32(i) for i in images]
  h, w = ims[0].shape[:2]
  ptrs = (C.c_void_p * len(ims))(*[i.ctypes.data for i in ims])
  g = (C.c_float * 3)()
  lib().orc_estimate_white_balance(ptrs, len(ims), w, h, C.c_uint32(_filters(pattern)), C.c_float(quantile), stride, g)
  return np.array(list(g), np.float32)


def bilinear5x5(cfa: np.ndarray, pattern) -> np.ndarray:
  c = _f32(cfa).reshape(cfa.shape[0], cfa.shape[1])
  h, w = c.shape
  out = np.empty((h, w, 3), np.float32)
  lib().orc_bilinear5x5(_p(c), _p(out), w, h, C.c_uint32(_filters(pattern)))
  return out


def ppg(cfa: np.ndarray, pattern, median_threshold: float = 0.0) -> np.ndarray:
  c = _f32(cfa).reshape(cfa.shape[0], cfa.shape[1])
  h, w = c.shape
  out = np.empty((h, w, 3), np.float32)
  lib().orc_ppg(_p(c), _p(out), w, h, C.c_uint32(_filters(pattern)), C.c_float(median_threshold))
  return out


class RCDWorkspace:
  """Mirrors the reference workspace: scratch persists between calls (rcd.cu:585-594)."""

  def __init__(self, width: int, height: int, pattern):
    self.width, self.height, self.filters = width, height, _filters(pattern)
    self.scratch = np.zeros(8 * width * height, np.float32)
    self.out = np.zeros((height, width, 3), np.float32)

  def process(self, cfa: np.ndarray) -> np.ndarray:
    c = _f32(cfa).reshape(self.height, self.width)
    lib().orc_rcd(_p(c), _p(self.out), self.width, self.height, C.c_uint32(self.filters), _p(self.scratch))
    return self.out.copy()


def rcd(cfa: np.ndarray, pattern) -> np.ndarray:
  return RCDWorkspace(cfa.shape[1], cfa.shape[0], pattern).process(cfa)


def postprocess(rgb: np.ndarray, pattern, color_smoothing_passes=0, green_eq_local=False, green_eq_global=False,
                green_eq_threshold=0.04) -> np.ndarray:
  a = _f32(rgb)
  h, w, _ = a.shape
  out = np.empty_like(a)
  lib().orc_postprocess(_p(a), _p(out), w, h, C.c_uint32(_filters(pattern)), int(color_smoothing_passes),
                        int(green_eq_local), int(green_eq_global), C.c_float(green_eq_threshold))
  return out


# ---- colour -----------------------------------------------------------------------------------
_COLOR_OPS = {'rgb_to_xyz': 0, 'xyz_to_lab': 1, 'lab_to_xyz': 2, 'xyz_to_rgb': 3, 'rgb_to_lab': 4, 'lab_to_rgb': 5,
              'modify_hsl': 6, 'modify_vibrance': 7, 'color_transform_3x3': 8}


def color_convert(x: np.ndarray, op: str, params=()) -> np.ndarray:
  a = _f32(x)
  out = np.empty_like(a)
  p = _f32(np.asarray(list(params) + [0.0] * 9, np.float32)[:9])
  lib().orc_color_convert(_p(a), _p(out), C.c_long(a.size // 3), _COLOR_OPS[op], _p(p))
  return out


def compute_luminance(rgb):
  a = _f32(rgb)
  out = np.empty(a.shape[:-1], np.float32)
  lib().orc_compute_luminance(_p(a), _p(out), C.c_long(out.size))
  return out


def compute_log_luminance(rgb, eps):
  a = _f32(rgb)
  out = np.empty(a.shape[:-1], np.float32)
  lib().orc_compute_log_luminance(_p(a), _p(out), C.c_long(out.size), C.c_float(eps))
  return out


def modify_luminance(rgb, lum):
  a, l = _f32(rgb), _f32(lum)
  out = np.empty_like(a)
  lib().orc_modify_luminance(_p(a), _p(l), _p(out), C.c_long(l.size))
  return out


def modify_log_luminance(rgb, loglum, eps):
  a, l = _f32(rgb), _f32(loglum)
  out = np.empty_like(a)
  lib().orc_modify_log_luminance(_p(a), _p(l), _p(out), C.c_long(l.size), C.c_float(eps))
  return out


# ---- statistics + tonemap ---------------------------------------------------------------------
def compute_image_bounds(images, stride: int = 8) -> np.ndarray:
  b = (C.c_float * 2)(np.finfo(np.float32).max, -np.finfo(np.float32).max)
  for im in images:
    a = _f32(im)
    lib().orc_bounds_accumulate(_p(a), a.shape[1], a.shape[0], stride, b)
  return np.array([b[0], b[1]], np.float32)


def compute_image_metrics(images, stride: int = 8, min_gray: float = 1e-4, rescale: bool = False) -> np.ndarray:
  bounds = compute_image_bounds(images, stride) if rescale else np.array([0.0, 1.0], np.float32)
  b = (C.c_float * 2)(float(bounds[0]), float(bounds[1]))
  sums = (C.c_double * 6)()
  for im in images:
    a = _f32(im)
    lib().orc_metrics_accumulate(_p(a), a.shape[1], a.shape[0], stride, C.c_float(min_gray), b, sums)
  norm = np.float32(1.0) / np.float32(max(sums[5], 1.0))  # color_adaption.cu:161-165
  return (np.array(list(sums)[:5], np.float32) * norm).astype(np.float32)


_TONEMAPS = {'reinhard': 0, 'aces': 1, 'adaptive_aces': 2, 'linear': 3}


def tonemap(rgb, op: str, metrics=None, gamma=1.0, intensity=0.0, light_adapt=0.8, vibrance=0.0) -> np.ndarray:
  a = _f32(rgb)
  out = np.empty(a.shape, np.uint8)
  m = _f32(metrics if metrics is not None else np.zeros(5))
  lib().orc_tonemap(_p(a), _p(out), C.c_long(a.size // 3), _TONEMAPS[op], _p(m), C.c_float(gamma), C.c_float(intensity),
                    C.c_float(light_adapt), C.c_float(vibrance))
  return out


# ---- filters ----------------------------------------------------------------------------------
def wiener(x: np.ndarray, sigmas, tile_size: int = 32, overlap_factor: int = 4) -> np.ndarray:
  a = _f32(x)
  h, w, c = a.shape
  out = np.empty_like(a)
  s = _f32(np.asarray(sigmas, np.float32).reshape(-1))
  assert s.size == c
  lib().orc_wiener(_p(a), _p(out), w, h, c, tile_size, overlap_factor, _p(s))
  return out


def wiener_log_luminance(rgb, noise: float, eps: float = 1e-4, tile_size=32, overlap_factor=4):
  ll = compute_log_luminance(rgb, eps)
  den = wiener(ll[..., None], [noise], tile_size, overlap_factor)[..., 0]
  return modify_log_luminance(rgb, den, eps)


def bilateral_grid_size(width, height, sigma_s, sigma_r):
  s = (C.c_int * 3)()
  lib().orc_bilateral_grid_size(width, height, C.c_float(sigma_s), C.c_float(sigma_r), s)
  return tuple(s)


def bilateral(lum, sigma_s: float, sigma_r: float, detail: float) -> np.ndarray:
  a = _f32(lum)
  out = np.empty_like(a)
  lib().orc_bilateral(_p(a), _p(out), a.shape[1], a.shape[0], C.c_float(sigma_s), C.c_float(sigma_r), C.c_float(detail))
  return out


def bilateral_rgb(rgb, sigma_s, sigma_r, detail):
  return modify_luminance(rgb, bilateral(compute_luminance(rgb), sigma_s, sigma_r, detail))


def laplacian(lum, sigma=0.2, shadows=1.0, highlights=1.0, clarity=0.0) -> np.ndarray:
  a = _f32(lum)
  out = np.empty_like(a)
  lib().orc_laplacian(_p(a), _p(out), a.shape[1], a.shape[0], C.c_float(sigma), C.c_float(shadows), C.c_float(highlights),
                      C.c_float(clarity))
  return out


# ---- the pipeline composite (pipeline/image_processor.py:236-319) ------------------------------
_TRANSFORMS = {
  'none': lambda a: a,
  'rotate_90': lambda a: np.rot90(a, 1, (0, 1)),
  'rotate_180': lambda a: np.rot90(a, 2, (0, 1)),
  'rotate_270': lambda a: np.rot90(a, 3, (0, 1)),
  'flip_horiz': lambda a: a[:, ::-1],
  'flip_vert': lambda a: a[::-1],
  'transverse': lambda a: a[::-1, ::-1],
  'transpose': lambda a: np.swapaxes(a, 0, 1),
}


class Pipeline:
  """CPU restatement of ImageProcessor.process_image_set with its EMA state and persistent RCD scratch."""

  def __init__(self, width, height, pattern='RGGB', ids=False, white_balance=None, debayer='rcd', postprocess=True,
               color_smoothing_passes=3, green_eq_threshold=0.04, ppg_median_threshold=0.0, enable_denoise=True,
               denoise=0.075, enable_bilateral=True, bilateral=0.4, bil_sigma_spatial=2.0, bil_sigma_luminance=0.2,
               tone_mapping='adaptive_aces', tone_gamma=1.5, tone_intensity=2.0, light_adapt=0.8, vibrance=0.5,
               moving_average=1.0, transform='none'):
    self.__dict__.update(locals())
    del self.__dict__['self']
    self.rcd_ws = RCDWorkspace(width, height, pattern)
    self.bounds = None
    self.metrics = None

  def debayer_frame(self, packed: np.ndarray) -> np.ndarray:
    cfa = decode12(packed, np.float32, self.ids).reshape(self.height, self.width)
    if self.white_balance is not None:
      cfa = white_balance(cfa, self.white_balance, self.pattern)
    if self.debayer == 'bilinear':
      rgb = bilinear5x5(cfa, self.pattern)
    elif self.debayer == 'ppg':
      rgb = ppg(cfa, self.pattern, self.ppg_median_threshold)
    else:
      rgb = self.rcd_ws.process(cfa)
    if self.postprocess:
      rgb = postprocess(rgb, self.pattern, self.color_smoothing_passes, False, True, self.green_eq_threshold)
    return rgb

  def process_rgb(self, rgb, bounds):
    rgb = ((rgb - bounds[0]) / (bounds[1] - bounds[0])).astype(np.float32)
    if self.enable_denoise:
      rgb = wiener_log_luminance(rgb, self.denoise)
    if self.enable_bilateral:
      rgb = bilateral_rgb(rgb, self.bil_sigma_spatial, self.bil_sigma_luminance, self.bilateral)
    return rgb

  def process_image_set(self, frames: list[np.ndarray]) -> list[np.ndarray]:
    raw = [self.debayer_frame(f) for f in frames]
    bounds = compute_image_bounds(raw, 8)
    prev = self.bounds if self.bounds is not None else bounds
    self.bounds = (prev + (bounds - prev) * np.float32(self.moving_average)).astype(np.float32)
    rgb = [self.process_rgb(r, self.bounds) for r in raw]
    metrics = compute_image_metrics(rgb, 8)
    prev = self.metrics if self.metrics is not None else metrics
    self.metrics = (prev + (metrics - prev) * np.float32(self.moving_average)).astype(np.float32)
    out = [tonemap(r, self.tone_mapping, self.metrics, self.tone_gamma, self.tone_intensity, self.light_adapt, self.vibrance)
           for r in rgb]
    return [np.ascontiguousarray(_TRANSFORMS[self.transform](o)) for o in out]
