"""Demosaic algorithms and the 12-bit packed codec (public names of the reference's torch_darktable/debayer.py)."""

from beartype import beartype
import torch

from .bayer import BayerPattern, PackedFormat
from .extension import extension


def _expect_shape(name: str, tensor: torch.Tensor, expected: tuple) -> None:
  if tuple(tensor.shape) != expected:
    raise RuntimeError(f'{name} input shape {tensor.shape} != expected {expected}')


class Bilinear5x5:
  def __init__(self, bayer_pattern: BayerPattern):
    self.bayer_pattern = bayer_pattern

  def process(self, image: torch.Tensor) -> torch.Tensor:
    return bilinear5x5_demosaic(image, self.bayer_pattern)


class PPG:
  """PPG demosaic workspace; `image_size` is (width, height)."""

  @beartype
  def __init__(self, device: torch.device, image_size: tuple[int, int], bayer_pattern: BayerPattern, *,
               median_threshold: float = 0.0):
    self._ppg = extension.PPG(device, image_size[0], image_size[1], bayer_pattern.value, median_threshold)

  def process(self, input_tensor: torch.Tensor) -> torch.Tensor:
    _expect_shape('PPG', input_tensor, (self._ppg.height, self._ppg.width, 1))
    return self._ppg.process(input_tensor)

  @property
  def image_size(self) -> tuple[int, int]:
    return (self._ppg.width, self._ppg.height)

  @property
  def median_threshold(self) -> float:
    return self._ppg.median_threshold


class RCD:
  """RCD demosaic workspace; `image_size` is (width, height)."""

  @beartype
  def __init__(self, device: torch.device, image_size: tuple[int, int], bayer_pattern: BayerPattern):
    self._rcd = extension.RCD(device, image_size[0], image_size[1], bayer_pattern.value)

  def process(self, input_tensor: torch.Tensor) -> torch.Tensor:
    _expect_shape('RCD', input_tensor, (self._rcd.height, self._rcd.width, 1))
    return self._rcd.process(input_tensor)

  @property
  def image_size(self) -> tuple[int, int]:
    return (self._rcd.width, self._rcd.height)


class PostProcess:
  """Colour smoothing + green equilibration workspace; `image_size` is (width, height)."""

  @beartype
  def __init__(self, device: torch.device, image_size: tuple[int, int], bayer_pattern: BayerPattern, *,
               color_smoothing_passes: int = 0, green_eq_local: bool = False, green_eq_global: bool = False,
               green_eq_threshold: float = 0.04):
    self._postprocess = extension.PostProcess(device, image_size[0], image_size[1], bayer_pattern.value,
                                              color_smoothing_passes, green_eq_local, green_eq_global, green_eq_threshold)

  def process(self, input_tensor: torch.Tensor) -> torch.Tensor:
    _expect_shape('PostProcess', input_tensor, (self._postprocess.height, self._postprocess.width, 3))
    return self._postprocess.process(input_tensor)

  @property
  def image_size(self) -> tuple[int, int]:
    return (self._postprocess.width, self._postprocess.height)

  @property
  def color_smoothing_passes(self) -> int:
    return self._postprocess.color_smoothing_passes

  @property
  def green_eq_threshold(self) -> float:
    return self._postprocess.green_eq_threshold


@beartype
def encode(image: torch.Tensor, format_type: PackedFormat = PackedFormat.Packed12,
           dtype: torch.dtype = torch.float32) -> torch.Tensor:
  """uint16 or float32 samples -> 12-bit packed bytes."""
  assert dtype in {torch.float32, torch.uint16}
  ids = format_type is PackedFormat.Packed12_IDS
  if image.dtype == torch.uint16:
    return extension.encode12_u16(image, ids_format=ids)
  if image.dtype == torch.float32:
    return extension.encode12_float(image, ids_format=ids)
  raise ValueError(f'Unsupported input dtype: {image.dtype}')


@beartype
def decode12(packed_data: torch.Tensor, output_dtype: torch.dtype = torch.float32,
             format_type: PackedFormat = PackedFormat.Packed12) -> torch.Tensor:
  """12-bit packed bytes -> float32 / float16 (scaled to [0,1]) or uint16 samples."""
  ids = format_type is PackedFormat.Packed12_IDS
  decoders = {torch.float32: extension.decode12_float, torch.float16: extension.decode12_half, torch.uint16: extension.decode12_u16}
  if output_dtype not in decoders:
    raise ValueError(f'Unsupported output dtype: {output_dtype}')
  return decoders[output_dtype](packed_data, ids_format=ids)


encode12_u16 = beartype(extension.encode12_u16)
encode12_float = beartype(extension.encode12_float)
decode12_float = beartype(extension.decode12_float)
decode12_half = beartype(extension.decode12_half)
decode12_u16 = beartype(extension.decode12_u16)


@beartype
def bilinear5x5_demosaic(image: torch.Tensor, bayer_pattern: BayerPattern) -> torch.Tensor:
  """(H, W, 1) CFA -> (H, W, 3) RGB with the 13-tap linear kernel."""
  return extension.bilinear5x5_demosaic(image, bayer_pattern.value)


@beartype
def demosaic_packed(packed: torch.Tensor, image_size: tuple[int, int], bayer_pattern: BayerPattern, *, method: str = 'rcd',
                    format_type: PackedFormat = PackedFormat.Packed12, black: float = 0.0,
                    white_balance: torch.Tensor | None = None, ppg_median_threshold: float = 0.0) -> torch.Tensor:
  """B200 addition: packed frame -> RGB in one kernel (unpack + black level + white balance + demosaic)."""
  return extension.demosaic_packed(packed, image_size[0], image_size[1], bayer_pattern.value, method,
                                   format_type is PackedFormat.Packed12_IDS, black, white_balance, ppg_median_threshold)


__all__ = ['PPG', 'RCD', 'BayerPattern', 'Bilinear5x5', 'PackedFormat', 'PostProcess', 'bilinear5x5_demosaic', 'decode12',
           'decode12_float', 'decode12_half', 'decode12_u16', 'demosaic_packed', 'encode', 'encode12_float', 'encode12_u16']
