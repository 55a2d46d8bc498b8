"""White balance on CFA data (public names of the reference's torch_darktable/white_balance.py)."""

from beartype import beartype
import torch

from .bayer import BayerPattern
from .extension import extension


@beartype
def apply_white_balance(bayer_image: torch.Tensor, gains: torch.Tensor, pattern: BayerPattern) -> torch.Tensor:
  """clamp(cfa * gain[colour of the site], 0, 1); `gains` = (R, G, B) on the device."""
  return extension.apply_white_balance(bayer_image, gains, pattern.value)


@beartype
def estimate_white_balance(bayer_images: list[torch.Tensor], pattern: BayerPattern, quantile: float = 0.98,
                           stride: int = 8) -> torch.Tensor:
  """Gains (green = 1) from the mean chromaticity of the brightest unsaturated 2x2 patches."""
  return extension.estimate_white_balance(bayer_images, pattern.value, quantile, stride)


__all__ = ['apply_white_balance', 'estimate_white_balance']
