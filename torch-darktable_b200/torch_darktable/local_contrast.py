"""Local contrast: bilateral grid and local Laplacian (public names of the reference's local_contrast.py)."""

from dataclasses import dataclass

from beartype import beartype
import torch

from .extension import extension


@beartype
@dataclass
class LaplacianParams:
  num_gamma: int = 6
  sigma: float = 0.2
  shadows: float = 1.0
  highlights: float = 1.0
  clarity: float = 0.0


class Laplacian:
  """Local Laplacian workspace on (H, W) luminance; `image_size` is (width, height)."""

  @beartype
  def __init__(self, device: torch.device, image_size: tuple[int, int], params: LaplacianParams):
    self._laplacian = extension.Laplacian(device, image_size[0], image_size[1], params.num_gamma, params.sigma, params.shadows,
                                          params.highlights, params.clarity)

  def process(self, input_tensor: torch.Tensor) -> torch.Tensor:
    expected = (self._laplacian.height, self._laplacian.width)
    if tuple(input_tensor.shape) != expected:
      raise RuntimeError(f'Laplacian input shape {input_tensor.shape} != expected {expected}')
    return self._laplacian.process(input_tensor)

  @beartype
  def process_rgb(self, input_image: torch.Tensor) -> torch.Tensor:
    return extension.modify_luminance(input_image, self.process(extension.compute_luminance(input_image)))

  @property
  def image_size(self) -> tuple[int, int]:
    return (self._laplacian.width, self._laplacian.height)

  sigma = property(lambda self: self._laplacian.sigma)
  shadows = property(lambda self: self._laplacian.shadows)
  highlights = property(lambda self: self._laplacian.highlights)
  clarity = property(lambda self: self._laplacian.clarity)


class Bilateral:
  """Bilateral-grid local contrast workspace; `image_size` is (width, height)."""

  @beartype
  def __init__(self, device: torch.device, image_size: tuple[int, int], *, sigma_s: float, sigma_r: float):
    self._bilateral = extension.Bilateral(device, image_size[0], image_size[1], sigma_s, sigma_r)

  def process(self, luminance: torch.Tensor, detail: float) -> torch.Tensor:
    expected = (self._bilateral.height, self._bilateral.width)
    if tuple(luminance.shape) != expected:
      raise RuntimeError(f'Bilateral input shape {luminance.shape} != expected {expected}')
    return self._bilateral.process(luminance, detail)

  @beartype
  def process_rgb(self, input_image: torch.Tensor, detail: float) -> torch.Tensor:
    assert input_image.dim() == 3, f'image must have 3 dimensions, got {input_image.shape}'
    return self._bilateral.process_rgb(input_image, float(detail))  # luminance extract / replace fused into splat / slice

  @beartype
  def process_log_rgb(self, input_image: torch.Tensor, detail: float, eps: float = 1e-6) -> torch.Tensor:
    log_luminance = extension.compute_log_luminance(input_image, eps)
    return extension.modify_log_luminance(input_image, self.process(log_luminance, float(detail)), eps)

  @property
  def image_size(self) -> tuple[int, int]:
    return (self._bilateral.width, self._bilateral.height)

  sigma_s = property(lambda self: self._bilateral.sigma_s)
  sigma_r = property(lambda self: self._bilateral.sigma_r)


__all__ = ['Bilateral', 'Laplacian', 'LaplacianParams']
