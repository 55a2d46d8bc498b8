"""JPEG output.  nvJPEG encoding sits after the sRGB result and is outside the B200 hot path;
the names exist so code importing them keeps working, `Jpeg.encode` raises JpegException."""

from enum import Enum

from .extension import extension

JpegException = extension.JpegException


class InputFormat(Enum):
  BGR = extension.BGR
  RGB = extension.RGB
  BGRI = extension.BGRI
  RGBI = extension.RGBI


class Subsampling(Enum):
  CSS_444 = extension.CSS_444
  CSS_422 = extension.CSS_422
  CSS_GRAY = extension.CSS_GRAY


class Jpeg:
  def __init__(self):
    self._coder = extension.Jpeg()

  def encode(self, image, quality: int = 94, input_format: InputFormat = InputFormat.RGBI,
             subsampling: Subsampling = Subsampling.CSS_422, progressive: bool = False):
    return self._coder.encode(image, quality, input_format.value, subsampling.value, progressive)
