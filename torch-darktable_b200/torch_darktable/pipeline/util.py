"""Small helpers of the pipeline.  normalize_image runs a libtdb200 kernel (the reference uses a torch.compile'd
Triton kernel here; there is no Triton in this build)."""

import torch

from ..extension import extension


def lerp(a: torch.Tensor, b: torch.Tensor, t: float) -> torch.Tensor:
  if a.is_cuda:
    return extension.lerp(a, b, float(t))
  return a + (b - a) * t


def normalize_image(rgb_raw: torch.Tensor, bounds: torch.Tensor) -> torch.Tensor:
  return extension.normalize(rgb_raw, bounds)


def resize(image: torch.Tensor, size: tuple[int, int]) -> torch.Tensor:
  chw = image.unsqueeze(0).permute(0, 3, 1, 2)
  out = torch.nn.functional.interpolate(chw, size=size, mode='bilinear', align_corners=False)
  return out.permute(0, 2, 3, 1).squeeze(0).contiguous()


def resize_longest_edge(size: tuple[int, int], longest: int) -> tuple[int, int]:
  if longest == 0:
    return size
  w, h = size
  return (longest, h * longest // w) if w > h else (w * longest // h, longest)


def resize_image(image: torch.Tensor, longest: int) -> torch.Tensor:
  h, w = image.shape[:2]
  return resize(image, resize_longest_edge((w, h), longest))
