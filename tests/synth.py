"""Seeded synthetic inputs shared by the golden generator, the parity tests and bench.py.

Everything here is plain numpy on the host so the very same bytes can be produced in this
container, on the GPU box and inside the golden-vector generator (SURVEY.md section 8d describes the
scene: gradient + two gratings + checker + noise, clipped, ~2 % of pixels saturated).
"""

from __future__ import annotations

import numpy as np

# darktable "filters" words, reference torch_darktable/csrc/debayer/demosaic.h:7-12
PATTERNS = {'RGGB': 0x94949494, 'BGGR': 0x16161616, 'GRBG': 0x61616161, 'GBRG': 0x49494949}


def fc(row, col, filters: int):
  """Colour of the CFA site (0=R, 1=G, 2=B); reference csrc/debayer/bayer_device.h:9-11."""
  row = np.asarray(row)
  col = np.asarray(col)
  shift = ((((row << 1) & 14) + (col & 1)) << 1).astype(np.uint64)
  return ((np.uint64(filters) >> shift) & np.uint64(3)).astype(np.int32)


def scene_rgb(height: int, width: int, seed: int = 1234) -> np.ndarray:
  """(H, W, 3) float32 scene in [0.02, ~1.0] with texture, edges and noise."""
  rng = np.random.default_rng(seed)
  y, x = np.mgrid[0:height, 0:width].astype(np.float32)
  out = np.empty((height, width, 3), np.float32)
  for c in range(3):
    grad = 0.15 + 0.5 * (x / max(width - 1, 1) * (0.6 + 0.2 * c) + y / max(height - 1, 1) * (0.4 - 0.1 * c))
    g1 = 0.12 * np.sin(2 * np.pi * (x + 0.5 * y) / 37.0 + 0.7 * c)
    g2 = 0.10 * np.sin(2 * np.pi * (y - 0.3 * x) / 211.0 + 1.3 * c)
    checker = 0.1 * ((((x // 64) + (y // 64)) % 2) * 2 - 1)
    noise = rng.normal(0.0, 0.01, size=(height, width)).astype(np.float32)
    out[..., c] = grad + g1 + g2 + checker + noise
  out *= 1.25  # pushes about 2 % of the samples over 1.0 so that clipping paths are exercised
  return np.clip(out, 0.02, 1.0).astype(np.float32)


def mosaic(rgb: np.ndarray, pattern: str = 'RGGB') -> np.ndarray:
  """(H, W, 3) -> (H, W) CFA samples for one of the four Bayer patterns."""
  h, w, _ = rgb.shape
  rows, cols = np.mgrid[0:h, 0:w]
  ch = fc(rows, cols, PATTERNS[pattern])
  return np.take_along_axis(rgb, ch[..., None], axis=2)[..., 0].astype(np.float32)


def pack12(values_u16: np.ndarray, ids: bool = False) -> np.ndarray:
  """Host-side packing used only to build inputs (standard / IDS byte layouts of a *decoder*)."""
  v = values_u16.reshape(-1, 2).astype(np.uint16)
  p0, p1 = v[:, 0], v[:, 1]
  out = np.empty((v.shape[0], 3), np.uint8)
  if ids:
    out[:, 0] = p0 >> 4
    out[:, 1] = p1 >> 4
    out[:, 2] = (p0 & 0xF) | ((p1 & 0xF) << 4)
  else:
    out[:, 0] = p0 & 0xFF
    out[:, 1] = ((p1 & 0xF) << 4) | (p0 >> 8)
    out[:, 2] = p1 >> 4
  return out.reshape(-1)


def packed_frame(height: int, width: int, seed: int = 1234, pattern: str = 'RGGB', ids: bool = False) -> np.ndarray:
  """A 12-bit packed Bayer frame (uint8, 1.5*H*W bytes) of the synthetic scene."""
  cfa = mosaic(scene_rgb(height, width, seed), pattern)
  q = np.floor(cfa * 4095.0 + 0.5).clip(0, 4095).astype(np.uint16)
  return pack12(q, ids)
