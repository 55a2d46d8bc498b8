"""Camera orientation transforms of the output image."""

from enum import Enum

from beartype import beartype
import torch


class ImageTransform(Enum):
  none = 0
  rotate_90 = 1
  rotate_180 = 2
  rotate_270 = 3
  transpose = 4
  flip_horiz = 5
  flip_vert = 6
  transverse = 7

  def next_rotation(self) -> 'ImageTransform':
    """Cycle through the four rotations, or through the four mirrored variants."""
    rotations = [ImageTransform.none, ImageTransform.rotate_90, ImageTransform.rotate_180, ImageTransform.rotate_270]
    mirrored = [ImageTransform.transpose, ImageTransform.flip_horiz, ImageTransform.flip_vert, ImageTransform.transverse]
    for ring in (rotations, mirrored):
      if self in ring:
        return ring[(ring.index(self) + 1) % 4]
    return ImageTransform.rotate_90


_SWAPS_AXES = {ImageTransform.rotate_90, ImageTransform.rotate_270, ImageTransform.transpose}


@beartype
def transformed_size(original_size: tuple[int, int], transform: ImageTransform) -> tuple[int, int]:
  return (original_size[1], original_size[0]) if transform in _SWAPS_AXES else original_size


def transform(image: torch.Tensor, transform: ImageTransform):
  """torch implementation (used when the transform is not fused into the tone-map store)."""
  t = transform
  if t is ImageTransform.none:
    return image
  if t in (ImageTransform.rotate_90, ImageTransform.rotate_180, ImageTransform.rotate_270):
    return torch.rot90(image, {ImageTransform.rotate_90: 1, ImageTransform.rotate_180: 2, ImageTransform.rotate_270: 3}[t], (0, 1)).contiguous()
  if t is ImageTransform.flip_horiz:
    return torch.flip(image, (1,)).contiguous()
  if t is ImageTransform.flip_vert:
    return torch.flip(image, (0,)).contiguous()
  if t is ImageTransform.transverse:
    return torch.flip(image, (0, 1)).contiguous()
  if t is ImageTransform.transpose:
    return torch.transpose(image, 0, 1).contiguous()
  raise ValueError(f'Unknown transform: {transform}')
