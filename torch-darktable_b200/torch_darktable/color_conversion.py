"""Colour space conversions on (H, W, 3) float32 CUDA tensors (names of the reference's color_conversion.py)."""

from beartype import beartype
import torch

from .extension import extension


def _unary(name: str, doc: str):
  fn = getattr(extension, name)

  @beartype
  def op(image: torch.Tensor) -> torch.Tensor:
    return fn(image)

  op.__name__ = op.__qualname__ = name
  op.__doc__ = doc
  return op


rgb_to_xyz = _unary('rgb_to_xyz', 'sRGB -> CIE XYZ (D65).')
xyz_to_rgb = _unary('xyz_to_rgb', 'CIE XYZ (D65) -> sRGB.')
xyz_to_lab = _unary('xyz_to_lab', 'CIE XYZ -> normalised Lab (L in [0,1], a,b / 128).')
lab_to_xyz = _unary('lab_to_xyz', 'Normalised Lab -> CIE XYZ.')
rgb_to_lab = _unary('rgb_to_lab', 'sRGB -> normalised Lab.')
lab_to_rgb = _unary('lab_to_rgb', 'Normalised Lab -> sRGB.')
compute_luminance = _unary('compute_luminance', 'Lab L of the clipped colour, (H, W).')


@beartype
def modify_luminance(rgb_image: torch.Tensor, luminance_multiplier: torch.Tensor) -> torch.Tensor:
  """Replace Lab L by the given (H, W) plane (clamped to [0,1]), keep a/b, clip the result."""
  return extension.modify_luminance(rgb_image, luminance_multiplier)


@beartype
def compute_log_luminance(rgb_image: torch.Tensor, eps: float) -> torch.Tensor:
  """log(max(eps, L))."""
  return extension.compute_log_luminance(rgb_image, eps)


@beartype
def modify_log_luminance(rgb_image: torch.Tensor, log_luminance: torch.Tensor, eps: float) -> torch.Tensor:
  """Replace Lab L by exp(log_luminance)."""
  return extension.modify_log_luminance(rgb_image, log_luminance, eps)


@beartype
def modify_hsl(rgb_image: torch.Tensor, hue_adjust: float = 0.0, sat_adjust: float = 0.0, lum_adjust: float = 0.0) -> torch.Tensor:
  """Hue shift plus power-law saturation / lightness adjustment in HSL."""
  return extension.modify_hsl(rgb_image, hue_adjust, sat_adjust, lum_adjust)


@beartype
def modify_vibrance(rgb_image: torch.Tensor, amount: float = 0.0) -> torch.Tensor:
  """darktable-style vibrance in Lab."""
  return extension.modify_vibrance(rgb_image, amount)


@beartype
def color_transform_3x3(image: torch.Tensor, matrix_3x3: torch.Tensor) -> torch.Tensor:
  """clip(M @ rgb) per pixel; `matrix_3x3` is a (3, 3) float32 CUDA tensor."""
  return extension.color_transform_3x3(image, matrix_3x3)


__all__ = ['color_transform_3x3', 'compute_log_luminance', 'compute_luminance', 'lab_to_rgb', 'lab_to_xyz', 'modify_hsl',
           'modify_log_luminance', 'modify_luminance', 'modify_vibrance', 'rgb_to_lab', 'rgb_to_xyz', 'xyz_to_lab', 'xyz_to_rgb']
