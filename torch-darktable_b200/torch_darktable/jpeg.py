"""JPEG output of the uint8 sRGB result.

API of the reference's torch_darktable/jpeg.py (`Jpeg().encode(image, quality, input_format, subsampling, progressive)` -> 1-D uint8
CPU tensor with the stream; `InputFormat`, `Subsampling`, `JpegException`); the coder behind it is libtdb200's nvJPEG binding
(csrc/jpeg.cu): stream-ordered on torch's current stream, reading pitched images in place."""

import enum

from .extension import extension as _ext

JpegException = _ext.JpegException
# the wrapper-level enums carry the values of the binding's (reference csrc/jpeg_encoder.h:6-17)
InputFormat = enum.IntEnum('InputFormat', {member.name: int(member) for member in _ext.JpegInputFormat})
Subsampling = enum.IntEnum('Subsampling', {member.name: int(member) for member in _ext.JpegSubsampling})


class Jpeg:
  """One nvJPEG handle + encoder state; not thread-safe, one object per stream of work."""

  def __init__(self):
    self.jpeg = _ext.Jpeg()

  def encode(self, image, quality=94, input_format=InputFormat.RGBI, subsampling=Subsampling.CSS_422, progressive=False):
    """(H, W, 3) interleaved or (3, H, W) planar uint8 CUDA tensor -> JPEG stream (optimised Huffman tables, baseline unless
    `progressive`)."""
    return self.jpeg.encode(image, quality, int(input_format), int(subsampling), progressive)

  def __repr__(self):
    return 'Jpeg'


__all__ = ['InputFormat', 'Jpeg', 'JpegException', 'Subsampling']
