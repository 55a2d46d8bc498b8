"""Bayer pattern helpers: enums, RGB -> CFA mosaicking and 2x2 plane (un)stacking.

Mirrors the public names of the reference's torch_darktable/bayer.py (pure torch, device agnostic there too).
"""

from enum import Enum
from pathlib import Path

from beartype import beartype
import torch

from .extension import extension


class BayerPattern(Enum):
  RGGB = extension.BayerPattern.RGGB
  BGGR = extension.BayerPattern.BGGR
  GRBG = extension.BayerPattern.GRBG
  GBRG = extension.BayerPattern.GBRG


class PackedFormat(Enum):
  Packed12 = 0
  Packed12_IDS = 1


# colour channel (0=R, 1=G, 2=B) sampled at the four quad positions (0,0) (0,1) (1,0) (1,1)
_QUAD_CHANNELS = {
  BayerPattern.RGGB: (0, 1, 1, 2),
  BayerPattern.BGGR: (2, 1, 1, 0),
  BayerPattern.GRBG: (1, 0, 1, 2),
  BayerPattern.GBRG: (1, 2, 1, 0),
}
# position of each quad sample in R, G1, G2, B order
_QUAD_ORDER = {
  BayerPattern.RGGB: (0, 1, 2, 3),
  BayerPattern.BGGR: (3, 1, 2, 0),
  BayerPattern.GRBG: (1, 0, 3, 2),
  BayerPattern.GBRG: (1, 3, 0, 2),
}


def channels(pattern: BayerPattern) -> tuple[int, int, int, int]:
  if pattern not in _QUAD_CHANNELS:
    raise ValueError(f'Invalid bayer pattern: {pattern}')
  return _QUAD_CHANNELS[pattern]


def pixel_order(pattern: BayerPattern) -> tuple[int, int, int, int]:
  if pattern not in _QUAD_ORDER:
    raise ValueError(f'Invalid bayer pattern: {pattern}')
  return _QUAD_ORDER[pattern]


def stack_bayer(bayer_image: torch.Tensor) -> torch.Tensor:
  """(H, W) CFA -> (H/2, W/2, 4) planes in quad order."""
  quads = [bayer_image[dy::2, dx::2] for dy in (0, 1) for dx in (0, 1)]
  return torch.stack(quads, dim=-1)


def expand_bayer(x: torch.Tensor) -> torch.Tensor:
  """(H/2, W/2, 4) planes in quad order -> (H, W, 1) CFA."""
  h2, w2 = x.shape[0], x.shape[1]
  out = torch.zeros(h2 * 2, w2 * 2, device=x.device, dtype=x.dtype)
  for i, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
    out[dy::2, dx::2] = x[..., i]
  return out.unsqueeze(-1)


@beartype
def rgb_to_bayer(rgb_tensor: torch.Tensor, pattern: BayerPattern = BayerPattern.RGGB) -> torch.Tensor:
  """(H, W, 3) RGB -> (H, W, 1) CFA samples of `pattern`."""
  ch = channels(pattern)
  quads = [rgb_tensor[dy::2, dx::2, ch[2 * dy + dx]] for dy in (0, 1) for dx in (0, 1)]
  return expand_bayer(torch.stack(quads, dim=-1))


@beartype
def load_as_bayer(image_path: Path, pattern: BayerPattern = BayerPattern.RGGB,
                  device: torch.device = torch.device('cuda')) -> torch.Tensor:
  """Load an RGB image file and mosaic it (needs OpenCV, imported lazily)."""
  if not image_path.exists():
    raise FileNotFoundError(f'Image not found: {image_path}')
  import cv2
  import numpy as np

  bgr = cv2.imread(str(image_path), cv2.IMREAD_COLOR)
  rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0
  return rgb_to_bayer(torch.from_numpy(rgb).to(device), pattern)
