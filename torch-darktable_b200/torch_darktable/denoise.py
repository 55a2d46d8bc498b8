"""Wiener tile denoiser front-end (public names of the reference's torch_darktable/denoise.py)."""

from beartype import beartype
import torch

from .extension import extension

_OVERLAPS = (2, 4, 8)
_TILES = (16, 32)


def check_overlap_factor(overlap_factor: int):
  if overlap_factor not in _OVERLAPS:
    raise ValueError('overlap_factor must be 2, 4, or 8')


class Wiener:
  """Overlapped-tile FFT Wiener filter for (H, W, 1) or (H, W, 3) float32 CUDA images."""

  @beartype
  def __init__(self, device: torch.device, image_size: tuple[int, int], overlap_factor: int = 4, tile_size: int = 32):
    width, height = image_size
    if device.type != 'cuda':
      raise ValueError(f'Device must be CUDA, got {device}')
    if width <= 0 or height <= 0:
      raise ValueError(f'Image dimensions must be positive, got {width}x{height}')
    check_overlap_factor(overlap_factor)
    if tile_size not in _TILES:
      raise ValueError(f'tile_size must be 16 or 32, got {tile_size}')
    try:
      self._wiener = extension.Wiener(device, width, height, overlap_factor, tile_size)
    except Exception as e:
      raise RuntimeError(f'Failed to create Wiener extension: {e}') from e
    self._tile_size = tile_size
    self._device = device

  def __repr__(self):
    return (f'Wiener({self._wiener.width}x{self._wiener.height},'
            f'overlap_factor={self.overlap_factor}, tile_size={self._tile_size})')

  @property
  def overlap_factor(self) -> int:
    return self._wiener.overlap_factor

  def _sigmas(self, noise, channels: int) -> torch.Tensor:
    if isinstance(noise, float):
      return torch.full((channels,), noise, dtype=torch.float32, device=self._device)
    if isinstance(noise, torch.Tensor):
      if noise.shape != (channels,):
        raise ValueError(f'noise tensor must have {channels} elements for {channels}-channel image')
      return noise.to(dtype=torch.float32, device=self._device)
    raise ValueError(f'noise must be float, or Tensor[{channels}]')

  @beartype
  def process(self, image: torch.Tensor, noise: float | torch.Tensor) -> torch.Tensor:
    """Denoise `image`; `noise` is one sigma for all channels or a per-channel tensor."""
    assert image.dim() == 3, f'image must have 3 dimensions, got {image.shape}'
    expected = (self._wiener.height, self._wiener.width, image.size(2))
    if tuple(image.shape) != expected:
      raise RuntimeError(f'Wiener input shape {image.shape} != expected {expected}')
    channels = image.size(2)
    if channels not in {1, 3}:
      raise ValueError(f'image channels must be 1 or 3, got {channels}')
    return self._wiener.process(image, self._sigmas(noise, channels))

  def process_luminance(self, image: torch.Tensor, noise: float | torch.Tensor) -> torch.Tensor:
    luminance = extension.compute_luminance(image)
    return extension.modify_luminance(image, self.process(luminance.unsqueeze(2), noise).squeeze(2))

  def process_log_luminance(self, image: torch.Tensor, noise: float | torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    if isinstance(noise, float):  # fused path: log-luminance, tiles and the Lab write-back without intermediate planes
      # the reference fails in process() on the (H, W, 1) log-luminance plane (denoise.py:87-89): same exception, same text
      got, expected = torch.Size((*image.shape[:2], 1)), (self._wiener.height, self._wiener.width, 1)
      if image.dim() != 3 or tuple(got) != expected:
        raise RuntimeError(f'Wiener input shape {got} != expected {expected}')
      return self._wiener.process_log_luminance(image, noise, eps)
    log_luminance = extension.compute_log_luminance(image, eps=eps)
    return extension.modify_log_luminance(image, self.process(log_luminance.unsqueeze(2), noise).squeeze(2), eps=eps)

  def process_log(self, image: torch.Tensor, noise: float | torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    return self.process((image + eps).log(), noise).exp()


@beartype
def create_wiener(device: torch.device, image_size: tuple[int, int], *, overlap: int = 4, tile_size: int = 32) -> Wiener:
  return Wiener(device, image_size, overlap_factor=overlap, tile_size=tile_size)


@beartype
def estimate_channel_noise(image: torch.Tensor, stride: int = 8) -> torch.Tensor:
  """Per-channel sigma from the MAD of a 4-neighbour Laplacian response, sampled every `stride` pixels.

  CUDA float32 images take the fused path (csrc/noise.cu: the response is evaluated at the sampled pixels only and the two medians
  are exact on-device selections; the (3,) result stays on the device and can be handed to `Wiener.process` as its noise).  Other
  tensors run the reference's device-agnostic torch code (denoise.py:131-158) as it is."""
  if image.is_cuda and image.dtype == torch.float32 and image.dim() == 3 and image.size(2) == 3:
    return extension.channel_noise(image.contiguous(), stride)
  kernel = torch.tensor([[0, -1, 0], [-1, 4, -1], [0, -1, 0]], dtype=image.dtype, device=image.device)
  chw = image.permute(2, 0, 1).unsqueeze(0)
  response = torch.conv2d(chw, kernel.expand(3, 1, 3, 3).contiguous(), groups=3, padding=1)[0, :, ::stride, ::stride].flatten(1)
  median = response.median(dim=1).values
  mad = (response - median.unsqueeze(1)).abs().median(dim=1).values
  return mad / 0.6745
