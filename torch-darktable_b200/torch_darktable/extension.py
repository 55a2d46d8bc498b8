"""`extension`: the object the rest of the package calls into.

In the reference this is the pybind11 module `torch_darktable.torch_darktable_extension`
(torch_darktable/csrc/extension.cpp:50-248, typed in torch_darktable_extension.pyi).  Here it is a thin Python
namespace with the same attribute surface (7 classes, 3 enums, TonemapParams, JpegException, the free functions) whose
bodies marshal torch tensors into raw device pointers for the C ABI of libtdb200.so (include/tdb200.h).

Differences from the reference binding, all deliberate:
  * `pattern` arguments accept both `BayerPattern` members and plain ints (the .pyi says int, pybind rejects ints);
  * no call synchronises the device; scalars the reference reads back with `.item()` (gains, sigmas, green ratio,
    valid-pixel count) stay on the GPU;
  * everything launches on torch's CURRENT stream of the tensor's device, under a device guard;
  * `RCD.process` returns a fresh tensor and does not depend on the previous frame (see csrc/rcd.cu);
  * `color_transform_3x3` works on ordinary device tensors (the reference dereferences the device pointer on the host);
  * extra fused entry points used by the pipeline: unpack12_wb, demosaic_packed, wiener_log_luminance, bilateral_rgb,
    normalize, lerp, tonemap (with optional 3x3 matrix and output transform).
`Jpeg` drives nvJPEG (a vendor library, loaded lazily by libtdb200) stream-ordered on the uint8 result of the tone map.
"""

from __future__ import annotations

import ctypes as C
import enum
from types import SimpleNamespace

import torch

from . import _lib
from ._lib import check, lib


class BayerPattern(enum.Enum):
  """darktable CFA `filters` words (reference csrc/debayer/demosaic.h:7-12)."""

  RGGB = 0x94949494
  BGGR = 0x16161616
  GRBG = 0x61616161
  GBRG = 0x49494949


class JpegInputFormat(enum.IntEnum):
  """Values of the reference's `enum class JpegInputFormat` (csrc/jpeg_encoder.h:6-11), not nvJPEG's."""

  BGR = 0
  RGB = 1
  BGRI = 2
  RGBI = 3


class JpegSubsampling(enum.IntEnum):
  """csrc/jpeg_encoder.h:13-17."""

  CSS_444 = 0
  CSS_422 = 1
  CSS_GRAY = 2


class JpegException(Exception):
  """nvJPEG reported an error: "<call>, nvjpeg error <code>: <text>" (csrc/jpeg_encoder.cu:84-88)."""


class Jpeg:
  """nvJPEG encoder object (extension.cpp:228-233, csrc/jpeg_encoder.cu:104-180): `encode` returns the JPEG stream as a CPU
  uint8 tensor.  The encoder runs on torch's current stream of the image's device and reads the image in place through its
  strides (rows need not be dense), so the transformed uint8 result of the tone map feeds it without a copy."""

  def __init__(self):
    handle = C.c_void_p()
    self._handle = None
    status = lib.tdb_jpeg_create(C.byref(handle))
    if status in (3, 4):  # TDB_EUNSUPPORTED: libnvjpeg missing; TDB_EJPEG: nvJPEG refused
      raise JpegException(_lib.last_error())
    check(status)
    self._handle = handle

  def __del__(self):
    handle = getattr(self, '_handle', None)
    if handle:
      lib.tdb_jpeg_destroy(handle)
      self._handle = None

  def encode(self, image: torch.Tensor, quality: int, input_format: int, subsampling: int, progressive: bool) -> torch.Tensor:
    input_format, subsampling = int(input_format), int(subsampling)
    _require(image.is_cuda, 'Input image should be on CUDA device')
    _require(image.dtype == torch.uint8, 'Input image should be uint8')
    if input_format not in tuple(JpegInputFormat):
      raise RuntimeError('Invalid input format')
    if subsampling not in tuple(JpegSubsampling):
      raise RuntimeError('Invalid subsampling')
    if input_format in (JpegInputFormat.BGRI, JpegInputFormat.RGBI):
      _require(image.dim() == 3 and image.size(2) == 3, 'for interleaved (BGRI, RGBI) expected 3D tensor (H, W, C)')
      if image.stride(2) != 1 or image.stride(1) != 3 or image.stride(0) < 3 * image.size(1):
        image = image.contiguous()
      height, width, pitch, plane = image.size(0), image.size(1), image.stride(0), 0
    else:
      _require(image.dim() == 3 and image.size(0) == 3, 'for planar (BGR, RGB) expected 3D tensor (C, H, W)')
      if image.stride(2) != 1 or image.stride(1) < image.size(2) or image.stride(0) < image.stride(1) * image.size(1):
        image = image.contiguous()
      height, width, pitch, plane = image.size(1), image.size(2), image.stride(1), image.stride(0)
    length = C.c_size_t(0)
    with torch.cuda.device(image.device):
      stream = _stream(image.device)
      self._check(lib.tdb_jpeg_encode(self._handle, _ptr(image), width, height, pitch, plane, input_format, int(quality), subsampling,
                                      1 if progressive else 0, C.byref(length), stream))
      out = torch.empty(length.value, dtype=torch.uint8)
      self._check(lib.tdb_jpeg_retrieve(self._handle, C.c_void_p(out.data_ptr()), length.value, C.byref(length), stream))
    return out[: length.value]

  @staticmethod
  def _check(status: int):
    if status == 4:  # TDB_EJPEG
      raise JpegException(_lib.last_error())
    check(status)

  def __repr__(self):
    return 'Jpeg'


# ---------------------------------------------------------------------------------------------------------------
# marshalling helpers
def _filters(pattern) -> int:
  if isinstance(pattern, BayerPattern):
    return pattern.value
  value = getattr(pattern, 'value', pattern)
  if isinstance(value, BayerPattern):
    value = value.value
  value = int(value)
  if value not in (0x94949494, 0x16161616, 0x61616161, 0x49494949):
    raise ValueError(f'Invalid bayer pattern: {pattern!r}')
  return value


def _ptr(t: torch.Tensor | None):
  return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device: torch.device):
  return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require(cond: bool, message: str):
  if not cond:
    raise RuntimeError(message)


def _cuda_f32(t: torch.Tensor, what: str):
  _require(t.is_cuda, f'{what} must be on CUDA device')
  _require(t.dtype == torch.float32, f'{what} must be float32')


def _rgb_image(t: torch.Tensor, what: str = 'Input'):
  _require(t.dtype == torch.float32, f'{what} must be float32')
  _require(t.dim() == 3 and t.size(2) == 3, f'{what} must be (H, W, 3)')
  _require(t.is_cuda, f'{what} must be on CUDA device')
  _require(t.is_contiguous(), f'{what} tensor must be contiguous')


def _device_scalars(values, device) -> torch.Tensor:
  return torch.tensor(list(values), dtype=torch.float32, device=device)


# ---------------------------------------------------------------------------------------------------------------
# packed codec (extension.cpp:159-169)
def _check_packed(t: torch.Tensor):
  _require(t.is_cuda, 'Input must be on CUDA device')
  _require(t.dtype == torch.uint8, 'Input must be uint8')
  _require(t.dim() == 1, 'Input must be 1D tensor')
  _require(t.size(0) % 3 == 0, 'Input length must be multiple of 3')


def decode12_float(input: torch.Tensor, ids_format: bool = False, scaled: bool = True) -> torch.Tensor:
  _check_packed(input)
  src = input.contiguous()
  npairs = src.numel() // 3
  out = torch.empty(npairs * 2, dtype=torch.float32, device=src.device)
  with torch.cuda.device(src.device):
    check(lib.tdb_decode12_f32(_ptr(src), _ptr(out), npairs, int(ids_format), int(scaled), _stream(src.device)))
  return out


def decode12_half(input: torch.Tensor, ids_format: bool = False, scaled: bool = True) -> torch.Tensor:
  _check_packed(input)
  src = input.contiguous()
  npairs = src.numel() // 3
  out = torch.empty(npairs * 2, dtype=torch.float16, device=src.device)
  with torch.cuda.device(src.device):
    check(lib.tdb_decode12_f16(_ptr(src), _ptr(out), npairs, int(ids_format), int(scaled), _stream(src.device)))
  return out


def decode12_u16(input: torch.Tensor, ids_format: bool = False) -> torch.Tensor:
  _check_packed(input)
  src = input.contiguous()
  npairs = src.numel() // 3
  out = torch.empty(npairs * 2, dtype=torch.uint16, device=src.device)
  with torch.cuda.device(src.device):
    check(lib.tdb_decode12_u16(_ptr(src), _ptr(out), npairs, int(ids_format), _stream(src.device)))
  return out


def _check_values(t: torch.Tensor, dtype, name):
  _require(t.is_cuda, 'Input must be on CUDA device')
  _require(t.dtype == dtype, f'Input must be {name}')
  _require(t.dim() == 1, 'Input must be 1D tensor')
  _require(t.size(0) % 2 == 0, 'Input length must be even')


def encode12_u16(input: torch.Tensor, ids_format: bool = False) -> torch.Tensor:
  _check_values(input, torch.uint16, 'uint16')
  src = input.contiguous()
  npairs = src.numel() // 2
  out = torch.empty(npairs * 3, dtype=torch.uint8, device=src.device)
  with torch.cuda.device(src.device):
    check(lib.tdb_encode12_u16(_ptr(src), _ptr(out), npairs, int(ids_format), _stream(src.device)))
  return out


def encode12_float(input: torch.Tensor, ids_format: bool = False, scaled: bool = True) -> torch.Tensor:
  _check_values(input, torch.float32, 'float32')
  src = input.contiguous()
  npairs = src.numel() // 2
  out = torch.empty(npairs * 3, dtype=torch.uint8, device=src.device)
  with torch.cuda.device(src.device):
    check(lib.tdb_encode12_f32(_ptr(src), _ptr(out), npairs, int(ids_format), int(scaled), _stream(src.device)))
  return out


def unpack12_wb(packed: torch.Tensor, width: int, height: int, pattern, ids_format: bool = False, black: float = 0.0,
                gains: torch.Tensor | None = None) -> torch.Tensor:
  """Fused decode12_float + black level + apply_white_balance -> (H, W) float32.  black=0, gains=None == decode12_float."""
  _check_packed(packed)
  _require(packed.numel() * 2 == width * height * 3, f'packed size {packed.numel()} does not match {width}x{height}')
  src = packed.contiguous()
  if gains is not None:
    gains = gains.to(device=src.device, dtype=torch.float32).contiguous()
  out = torch.empty((height, width), dtype=torch.float32, device=src.device)
  with torch.cuda.device(src.device):
    check(lib.tdb_unpack12_wb(_ptr(src), _ptr(out), width, height, int(ids_format), _filters(pattern), float(black), _ptr(gains),
                              _stream(src.device)))
  return out


# ---------------------------------------------------------------------------------------------------------------
# white balance (extension.cpp:209-212)
def apply_white_balance(bayer_image: torch.Tensor, gains: torch.Tensor, pattern) -> torch.Tensor:
  _cuda_f32(bayer_image, 'bayer_image')
  _require(bayer_image.dim() >= 2, 'bayer_image must be (H, W)')
  src = bayer_image.contiguous()
  g = gains.to(device=src.device, dtype=torch.float32).contiguous()
  _require(g.numel() == 3, 'gains must have 3 elements')
  out = torch.empty_like(src)
  with torch.cuda.device(src.device):
    check(lib.tdb_white_balance(_ptr(src), _ptr(out), src.size(1), src.size(0), _filters(pattern), _ptr(g), _stream(src.device)))
  return out


def estimate_white_balance(bayer_images, pattern, quantile: float = 0.95, stride: int = 8) -> torch.Tensor:
  if len(bayer_images) == 0:
    raise RuntimeError('No images provided')
  first = bayer_images[0]
  height, width = first.size(0), first.size(1)
  sh, sw = height // stride, width // stride
  device = first.device
  n = len(bayer_images) * sh * sw
  if n == 0:  # no sample site at all (an image smaller than the stride): the reference's empty-selection answer
    return torch.ones(3, dtype=torch.float32, device=device)
  chroma = torch.empty((n, 2), dtype=torch.float32, device=device)
  intensity = torch.empty(n, dtype=torch.float32, device=device)
  valid = torch.empty(n, dtype=torch.uint8, device=device)
  with torch.cuda.device(device):
    for i, img in enumerate(bayer_images):
      _cuda_f32(img, 'bayer image')
      src = img.contiguous()
      o = i * sh * sw
      check(lib.tdb_wb_collect_samples(_ptr(src), width, height, _filters(pattern), stride, _ptr(chroma[o:]), _ptr(intensity[o:]),
                                       _ptr(valid[o:]), _stream(device)))
    gains = torch.empty(3, dtype=torch.float32, device=device)
    check(lib.tdb_wb_estimate_gains(_ptr(chroma), _ptr(intensity), _ptr(valid), n, float(quantile), _ptr(gains), _stream(device)))
  return gains


# ---------------------------------------------------------------------------------------------------------------
# demosaic (extension.cpp:57-90, :205)
def _check_cfa(t: torch.Tensor):
  _require(t.is_cuda, 'Input tensor must be on CUDA device')
  _require(t.dtype == torch.float32, 'Input tensor must be float32')
  _require(t.dim() == 3, 'Input tensor must be 3D (H, W, 1)')
  _require(t.size(2) == 1, 'Input must have single channel (raw Bayer)')


def bilinear5x5_demosaic(input: torch.Tensor, pattern) -> torch.Tensor:
  _check_cfa(input)
  src = input.contiguous()
  h, w = src.size(0), src.size(1)
  out = torch.empty((h, w, 3), dtype=torch.float32, device=src.device)
  with torch.cuda.device(src.device):
    check(lib.tdb_bilinear5x5(_ptr(src), _ptr(out), w, h, _filters(pattern), _stream(src.device)))
  return out


class _Workspace:
  def __init__(self, device, width: int, height: int):
    self._device = torch.device(device)
    self._width, self._height = int(width), int(height)
    self._scratch: torch.Tensor | None = None

  @property
  def width(self) -> int:
    return self._width

  @property
  def height(self) -> int:
    return self._height

  def _check_size(self, t: torch.Tensor):
    _require(t.size(0) == self._height and t.size(1) == self._width, 'Input dimensions must match workspace size')

  def _ensure_scratch(self, nbytes: int) -> torch.Tensor:
    if self._scratch is None or self._scratch.numel() < nbytes:
      self._scratch = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self._device)
    return self._scratch


class PPG(_Workspace):
  def __init__(self, device, width, height, pattern, median_threshold: float = 0.0):
    super().__init__(device, width, height)
    self._filters = _filters(pattern)
    self.median_threshold = float(median_threshold)

  def process(self, input: torch.Tensor) -> torch.Tensor:
    _check_cfa(input)
    self._check_size(input)
    src = input.contiguous()
    out = torch.empty((self._height, self._width, 3), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
      check(lib.tdb_ppg(_ptr(src), _ptr(out), self._width, self._height, self._filters, float(self.median_threshold),
                        _stream(src.device)))
    return out


class RCD(_Workspace):
  def __init__(self, device, width, height, pattern):
    super().__init__(device, width, height)
    self._filters = _filters(pattern)

  def process(self, input: torch.Tensor) -> torch.Tensor:
    _check_cfa(input)
    self._check_size(input)
    src = input.contiguous()
    out = torch.empty((self._height, self._width, 3), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
      check(lib.tdb_rcd(_ptr(src), _ptr(out), self._width, self._height, self._filters, _stream(src.device)))
    return out


_DEMOSAIC_METHODS = {'bilinear': 0, 'ppg': 1, 'rcd': 2}


def demosaic_packed(packed: torch.Tensor, width: int, height: int, pattern, method: str = 'rcd', ids_format: bool = False,
                    black: float = 0.0, gains: torch.Tensor | None = None, ppg_median_threshold: float = 0.0) -> torch.Tensor:
  """12-bit packed frame -> (H, W, 3) RGB in one kernel: unpack + black level + white balance + demosaic."""
  _check_packed(packed)
  _require(packed.numel() * 2 == width * height * 3, f'packed size {packed.numel()} does not match {width}x{height}')
  src = packed.contiguous()
  if gains is not None:
    gains = gains.to(device=src.device, dtype=torch.float32).contiguous()
  out = torch.empty((height, width, 3), dtype=torch.float32, device=src.device)
  with torch.cuda.device(src.device):
    check(lib.tdb_demosaic_packed(_ptr(src), _ptr(out), width, height, int(ids_format), _filters(pattern), _DEMOSAIC_METHODS[method],
                                  float(black), _ptr(gains), float(ppg_median_threshold), _stream(src.device)))
  return out


class PostProcess(_Workspace):
  def __init__(self, device, width, height, pattern, color_smoothing_passes: int = 0, green_eq_local: bool = False,
               green_eq_global: bool = False, green_eq_threshold: float = 0.04):
    super().__init__(device, width, height)
    self._filters = _filters(pattern)
    self.color_smoothing_passes = int(color_smoothing_passes)
    self.green_eq_local = bool(green_eq_local)
    self.green_eq_global = bool(green_eq_global)
    self.green_eq_threshold = float(green_eq_threshold)

  def process(self, input: torch.Tensor) -> torch.Tensor:
    _require(input.is_cuda, 'Input tensor must be on CUDA device')
    _require(input.dtype == torch.float32, 'Input tensor must be float32')
    _require(input.dim() == 3, 'Input tensor must be 3D (H, W, 3)')
    _require(input.size(2) == 3, 'Input must have 3 channels (RGB)')
    _require(input.size(0) == self._height and input.size(1) == self._width,
             f'Input size {input.size(0)}x{input.size(1)} does not match expected {self._height}x{self._width}')
    src = input.contiguous()
    out = torch.empty_like(src)
    scratch = self._ensure_scratch(lib.tdb_postprocess_scratch_bytes(self._width, self._height))
    with torch.cuda.device(src.device):
      check(lib.tdb_postprocess(_ptr(src), _ptr(out), _ptr(scratch), self._width, self._height, self._filters,
                                self.color_smoothing_passes, int(self.green_eq_local), int(self.green_eq_global),
                                float(self.green_eq_threshold), _stream(src.device)))
    return out


# ---------------------------------------------------------------------------------------------------------------
# colour ops (extension.cpp:127-156)
_COLOR_OPS = {'rgb_to_xyz': 0, 'xyz_to_lab': 1, 'lab_to_xyz': 2, 'xyz_to_rgb': 3, 'rgb_to_lab': 4, 'lab_to_rgb': 5,
              'modify_hsl': 6, 'modify_vibrance': 7}


def _convert(t: torch.Tensor, op: str, p0=0.0, p1=0.0, p2=0.0) -> torch.Tensor:
  _rgb_image(t)
  out = torch.empty_like(t)
  with torch.cuda.device(t.device):
    check(lib.tdb_color_convert(_ptr(t), _ptr(out), t.size(0) * t.size(1), _COLOR_OPS[op], float(p0), float(p1), float(p2),
                                _stream(t.device)))
  return out


def rgb_to_xyz(rgb): return _convert(rgb, 'rgb_to_xyz')
def xyz_to_lab(xyz): return _convert(xyz, 'xyz_to_lab')
def lab_to_xyz(lab): return _convert(lab, 'lab_to_xyz')
def xyz_to_rgb(xyz): return _convert(xyz, 'xyz_to_rgb')
def rgb_to_lab(rgb): return _convert(rgb, 'rgb_to_lab')
def lab_to_rgb(lab): return _convert(lab, 'lab_to_rgb')


def modify_hsl(rgb, hue_adjust: float = 0.0, sat_adjust: float = 0.0, lum_adjust: float = 0.0):
  return _convert(rgb, 'modify_hsl', hue_adjust, sat_adjust, lum_adjust)


def modify_vibrance(rgb, amount: float = 0.0):
  return _convert(rgb, 'modify_vibrance', amount)


def color_transform_3x3(input: torch.Tensor, matrix_3x3: torch.Tensor) -> torch.Tensor:
  _require(matrix_3x3.dtype == torch.float32, 'Matrix must be float32')
  _require(matrix_3x3.dim() == 2 and matrix_3x3.size(0) == 3 and matrix_3x3.size(1) == 3, 'Matrix must be (3, 3)')
  _require(matrix_3x3.is_cuda and matrix_3x3.is_contiguous(), 'Matrix tensor must be contiguous CUDA tensor')
  _rgb_image(input)
  out = torch.empty_like(input)
  with torch.cuda.device(input.device):
    check(lib.tdb_color_transform_3x3(_ptr(input), _ptr(out), input.size(0) * input.size(1), _ptr(matrix_3x3), _stream(input.device)))
  return out


def compute_luminance(rgb: torch.Tensor) -> torch.Tensor:
  _rgb_image(rgb)
  out = torch.empty(rgb.shape[:2], dtype=torch.float32, device=rgb.device)
  with torch.cuda.device(rgb.device):
    check(lib.tdb_compute_luminance(_ptr(rgb), _ptr(out), out.numel(), _stream(rgb.device)))
  return out


def compute_log_luminance(rgb: torch.Tensor, eps: float) -> torch.Tensor:
  _require(eps > 0.0, 'Epsilon must be positive')
  _rgb_image(rgb)
  out = torch.empty(rgb.shape[:2], dtype=torch.float32, device=rgb.device)
  with torch.cuda.device(rgb.device):
    check(lib.tdb_compute_log_luminance(_ptr(rgb), _ptr(out), out.numel(), float(eps), _stream(rgb.device)))
  return out


def _check_modify(rgb: torch.Tensor, lum: torch.Tensor):
  _require(rgb.dtype == torch.float32, 'Input1 must be float32')
  _require(lum.dtype == torch.float32, 'Input2 must be float32')
  _require(rgb.dim() == 3 and rgb.size(2) == 3, 'Input1 must be (H, W, 3)')
  _require(lum.dim() == 2, 'Input2 must be (H, W)')
  _require(rgb.is_cuda and lum.is_cuda, 'Inputs must be on CUDA device')
  _require(rgb.is_contiguous() and lum.is_contiguous(), 'Input tensors must be contiguous')
  _require(lum.size(0) == rgb.size(0) and lum.size(1) == rgb.size(1), 'Input dimensions must match')


def modify_luminance(rgb: torch.Tensor, new_luminance: torch.Tensor) -> torch.Tensor:
  _check_modify(rgb, new_luminance)
  out = torch.empty_like(rgb)
  with torch.cuda.device(rgb.device):
    check(lib.tdb_modify_luminance(_ptr(rgb), _ptr(new_luminance), _ptr(out), new_luminance.numel(), _stream(rgb.device)))
  return out


def modify_log_luminance(rgb: torch.Tensor, log_luminance: torch.Tensor, eps: float) -> torch.Tensor:
  _require(eps > 0.0, 'Epsilon must be positive')
  _check_modify(rgb, log_luminance)
  out = torch.empty_like(rgb)
  with torch.cuda.device(rgb.device):
    check(lib.tdb_modify_log_luminance(_ptr(rgb), _ptr(log_luminance), _ptr(out), log_luminance.numel(), float(eps),
                                       _stream(rgb.device)))
  return out


def normalize(rgb: torch.Tensor, bounds: torch.Tensor) -> torch.Tensor:
  """(rgb - bounds[0]) / (bounds[1] - bounds[0]) with device-resident bounds (pipeline/util.py:8-10)."""
  _cuda_f32(rgb, 'image')
  src = rgb.contiguous()
  b = bounds.to(device=src.device, dtype=torch.float32).contiguous()
  out = torch.empty_like(src)
  with torch.cuda.device(src.device):
    check(lib.tdb_normalize(_ptr(src), _ptr(out), src.numel(), _ptr(b), _stream(src.device)))
  return out


def lerp(a: torch.Tensor, b: torch.Tensor, t: float) -> torch.Tensor:
  """a + (b - a) * t on small device vectors (the EMA of bounds / metrics, pipeline/util.py:4)."""
  _cuda_f32(a, 'a')
  _cuda_f32(b, 'b')
  a, b = a.contiguous(), b.contiguous()
  out = torch.empty_like(a)
  with torch.cuda.device(a.device):
    check(lib.tdb_lerp(_ptr(a), _ptr(b), float(t), _ptr(out), a.numel(), _stream(a.device)))
  return out


# ---------------------------------------------------------------------------------------------------------------
# statistics + tone mapping (extension.cpp:172-195)
class TonemapParams:
  def __init__(self, gamma: float = 1.0, intensity: float = 0.0, light_adapt: float = 0.8, vibrance: float = 0.0):
    self.gamma, self.intensity, self.light_adapt, self.vibrance = float(gamma), float(intensity), float(light_adapt), float(vibrance)


def _check_image(t: torch.Tensor):
  _require(t.is_cuda, 'image must be CUDA')
  _require(t.dtype == torch.float32, 'image must be float32')
  _require(t.dim() == 3 and t.size(2) == 3, 'image must be (H,W,3)')


def compute_image_bounds(images, stride: int = 8) -> torch.Tensor:
  _require(len(images) > 0, 'images must be non-empty')
  device = images[0].device
  bounds = torch.empty(2, dtype=torch.float32, device=device)
  with torch.cuda.device(device):
    s = _stream(device)
    check(lib.tdb_bounds_init(_ptr(bounds), s))
    for img in images:
      _check_image(img)
      src = img.contiguous()
      check(lib.tdb_bounds_accumulate(_ptr(src), src.size(1), src.size(0), int(stride), _ptr(bounds), s))
  return bounds


def compute_image_metrics(images, stride: int = 8, min_gray: float = 1e-4, rescale: bool = False) -> torch.Tensor:
  _require(len(images) > 0, 'images must be non-empty')
  device = images[0].device
  bounds = compute_image_bounds(images, stride) if rescale else None
  sums = torch.empty(6, dtype=torch.float32, device=device)
  metrics = torch.empty(5, dtype=torch.float32, device=device)
  with torch.cuda.device(device):
    s = _stream(device)
    check(lib.tdb_metrics_init(_ptr(sums), s))
    for img in images:
      _check_image(img)
      src = img.contiguous()
      check(lib.tdb_metrics_accumulate(_ptr(src), src.size(1), src.size(0), int(stride), float(min_gray), _ptr(bounds), _ptr(sums), s))
    check(lib.tdb_metrics_finalize(_ptr(sums), _ptr(metrics), s))
  return metrics


def image_metric_sums(images, stride: int = 8, min_gray: float = 1e-4) -> torch.Tensor:
  """The six raw sums behind compute_image_metrics (log-gray, gray, r, g, b, count) -- what ranks all-reduce when one frame is
  split across GPUs; `metrics_from_sums` finishes them."""
  _require(len(images) > 0, 'images must be non-empty')
  device = images[0].device
  sums = torch.empty(6, dtype=torch.float32, device=device)
  with torch.cuda.device(device):
    s = _stream(device)
    check(lib.tdb_metrics_init(_ptr(sums), s))
    for img in images:
      _check_image(img)
      src = img.contiguous()
      check(lib.tdb_metrics_accumulate(_ptr(src), src.size(1), src.size(0), int(stride), float(min_gray), None, _ptr(sums), s))
  return sums


def metrics_from_sums(sums: torch.Tensor) -> torch.Tensor:
  metrics = torch.empty(5, dtype=torch.float32, device=sums.device)
  with torch.cuda.device(sums.device):
    check(lib.tdb_metrics_finalize(_ptr(sums.contiguous()), _ptr(metrics), _stream(sums.device)))
  return metrics


def green_sums(rgb: torch.Tensor, pattern) -> torch.Tensor:
  """(G1 sum, G2 sum) of an (H, W, 3) image whose first row is an even row of the CFA (postprocess.cu:195-203)."""
  _check_image(rgb)
  src = rgb.contiguous()
  h, w = src.size(0), src.size(1)
  sums = torch.empty(2, dtype=torch.float32, device=src.device)
  scratch = torch.empty(lib.tdb_postprocess_scratch_bytes(w, h), dtype=torch.uint8, device=src.device)
  with torch.cuda.device(src.device):
    check(lib.tdb_green_sums(_ptr(src), w, h, _filters(pattern), _ptr(scratch), _ptr(sums), _stream(src.device)))
  return sums


def green_eq_apply(rgb: torch.Tensor, ratio: torch.Tensor, pattern, green_eq_local: bool = False, green_eq_threshold: float = 0.04):
  """Global green equilibration with a given G2/G1 ratio (one-element device tensor), optionally followed by the local one."""
  _check_image(rgb)
  src = rgb.contiguous()
  out = torch.empty_like(src)
  with torch.cuda.device(src.device):
    check(lib.tdb_green_eq_apply(_ptr(src), _ptr(out), src.size(1), src.size(0), _filters(pattern), int(green_eq_local),
                                 float(green_eq_threshold), _ptr(ratio.to(torch.float32).contiguous()), _stream(src.device)))
  return out


_TM = {'reinhard': 0, 'aces': 1, 'adaptive_aces': 2, 'linear': 3}
_TF = {'none': 0, 'rotate_90': 1, 'rotate_180': 2, 'rotate_270': 3, 'transpose': 4, 'flip_horiz': 5, 'flip_vert': 6, 'transverse': 7}


def _tonemap_out(out: torch.Tensor | None, h: int, w: int, transform: str, device) -> torch.Tensor:
  shape = (w, h, 3) if transform in ('rotate_90', 'rotate_270', 'transpose') else (h, w, 3)
  if out is None:
    return torch.empty(shape, dtype=torch.uint8, device=device)
  _require(out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == shape and out.is_contiguous(),
           f'out must be a contiguous uint8 CUDA tensor of shape {shape}')
  return out


def tonemap(image: torch.Tensor, op: str, metrics: torch.Tensor | None, params, matrix: torch.Tensor | None = None,
            transform: str = 'none', out: torch.Tensor | None = None) -> torch.Tensor:
  """Fused [3x3 matrix] -> tone map -> gamma -> vibrance -> uint8 [-> rotate/flip].  Returns the transformed (H', W', 3);
  `out` (optional) receives it in place (a slot of a batch buffer, see ImageProcessor.process_batch)."""
  _check_image(image)
  src = image.contiguous()
  h, w = src.size(0), src.size(1)
  if op != 'aces':
    _require(metrics is not None and metrics.dtype == torch.float32 and metrics.numel() == 5, 'metrics must be 5 float32 values')
    metrics = metrics.to(src.device).contiguous()
  else:
    metrics = None
  if matrix is not None:
    matrix = matrix.to(device=src.device, dtype=torch.float32).contiguous()
  out = _tonemap_out(out, h, w, transform, src.device)
  with torch.cuda.device(src.device):
    check(lib.tdb_tonemap(_ptr(src), _ptr(out), w, h, _TM[op], _ptr(metrics), params.gamma, params.intensity, params.light_adapt,
                          params.vibrance, _ptr(matrix), _TF[transform], _stream(src.device)))
  return out


def reinhard_tonemap(image, metrics, params): return tonemap(image, 'reinhard', metrics, params)
def aces_tonemap(image, params): return tonemap(image, 'aces', None, params)
def adaptive_aces_tonemap(image, metrics, params): return tonemap(image, 'adaptive_aces', metrics, params)
def linear_tonemap(image, metrics, params): return tonemap(image, 'linear', metrics, params)


# ---------------------------------------------------------------------------------------------------------------
# Wiener (extension.cpp:215-223)
class Wiener(_Workspace):
  def __init__(self, device, width, height, overlap_factor: int = 4, tile_size: int = 32):
    super().__init__(device, width, height)
    # the reference silently builds the 32-pixel implementation for any tile_size other than 16
    self._tile = 16 if int(tile_size) == 16 else 32
    self._overlap = int(overlap_factor)

  @property
  def overlap_factor(self) -> int:
    return self._overlap

  def process(self, input: torch.Tensor, noise_sigmas: torch.Tensor) -> torch.Tensor:
    _require(input.dim() == 3, 'expected HWC tensor')
    _require(input.device == self._device, 'input device mismatch')
    channels = input.size(2)
    _require(channels in (1, 3), f'input channels must be 1 or 3, got {channels}')
    _require(noise_sigmas.numel() == channels, 'noise_sigmas must have C elements')
    _cuda_f32(input, 'input')
    src = input.contiguous()
    h, w = src.size(0), src.size(1)
    sig = noise_sigmas.to(device=src.device, dtype=torch.float32).contiguous()
    out = torch.empty_like(src)
    scratch = self._ensure_scratch(lib.tdb_wiener_scratch_bytes(w, h, channels, self._tile))
    with torch.cuda.device(src.device):
      check(lib.tdb_wiener(_ptr(src), _ptr(out), _ptr(scratch), w, h, channels, self._tile, self._overlap, _ptr(sig), _stream(src.device)))
    return out

  def process_log_luminance(self, rgb: torch.Tensor, noise: float, eps: float = 1e-4) -> torch.Tensor:
    """compute_log_luminance -> process -> modify_log_luminance (denoise.py:54-58) without the intermediate planes."""
    _rgb_image(rgb)
    self._check_size(rgb)
    h, w = rgb.size(0), rgb.size(1)
    out = torch.empty_like(rgb)
    scratch = self._ensure_scratch(lib.tdb_wiener_scratch_bytes(w, h, 1, self._tile))
    with torch.cuda.device(rgb.device):
      check(lib.tdb_wiener_log_luminance(_ptr(rgb), _ptr(out), _ptr(scratch), w, h, self._tile, self._overlap, float(noise), float(eps),
                                         _stream(rgb.device)))
    return out


def channel_noise(image: torch.Tensor, stride: int = 8) -> torch.Tensor:
  """Per-channel noise sigma (3,) of an (H, W, 3) float32 CUDA image, on the device (csrc/noise.cu)."""
  _rgb_image(image)
  _require(stride >= 1, 'stride must be positive')
  h, w = image.size(0), image.size(1)
  scratch = torch.empty(lib.tdb_channel_noise_scratch_bytes(w, h, int(stride)), dtype=torch.uint8, device=image.device)
  sigma = torch.empty(3, dtype=torch.float32, device=image.device)
  with torch.cuda.device(image.device):
    check(lib.tdb_channel_noise(_ptr(image), w, h, int(stride), _ptr(scratch), _ptr(sigma), _stream(image.device)))
  return sigma


# ---------------------------------------------------------------------------------------------------------------
# local contrast (extension.cpp:94-121)
class Bilateral(_Workspace):
  def __init__(self, device, width, height, sigma_s: float = 8.0, sigma_r: float = 0.1):
    _require(width > 0 and height > 0, 'Invalid dimensions')
    super().__init__(device, width, height)
    self._sigma_s, self._sigma_r = float(sigma_s), float(sigma_r)

  @property
  def sigma_s(self) -> float:
    return self._sigma_s

  @sigma_s.setter
  def sigma_s(self, v: float):
    self._sigma_s, self._scratch = float(v), None

  @property
  def sigma_r(self) -> float:
    return self._sigma_r

  @sigma_r.setter
  def sigma_r(self, v: float):
    self._sigma_r, self._scratch = float(v), None

  def grid_size(self) -> tuple[int, int, int]:
    size = (C.c_int * 3)()
    check(lib.tdb_bilateral_grid_size(self._width, self._height, self._sigma_s, self._sigma_r, size))
    return tuple(size)

  def _grid_scratch(self) -> torch.Tensor:
    return self._ensure_scratch(lib.tdb_bilateral_scratch_bytes(self._width, self._height, self._sigma_s, self._sigma_r))

  def process(self, luminance: torch.Tensor, detail: float) -> torch.Tensor:
    _require(luminance.dtype == torch.float32, 'Input must be float32')
    _require(luminance.dim() == 2, 'Input must be 2D (H,W)')
    _require(luminance.size(0) == self._height and luminance.size(1) == self._width, 'Input shape must match (H,W)')
    _require(luminance.is_cuda, 'Input must be CUDA tensor')
    src = luminance.contiguous()
    out = torch.empty_like(src)
    scratch = self._grid_scratch()
    with torch.cuda.device(src.device):
      check(lib.tdb_bilateral(_ptr(src), _ptr(out), _ptr(scratch), self._width, self._height, self._sigma_s, self._sigma_r,
                              float(detail), _stream(src.device)))
    return out

  def process_rgb(self, rgb: torch.Tensor, detail: float) -> torch.Tensor:
    """compute_luminance -> process -> modify_luminance (local_contrast.py:110-114) with both colour passes fused in."""
    _rgb_image(rgb)
    _require(rgb.size(0) == self._height and rgb.size(1) == self._width, 'Input shape must match (H,W)')
    out = torch.empty_like(rgb)
    scratch = self._grid_scratch()
    with torch.cuda.device(rgb.device):
      check(lib.tdb_bilateral_rgb(_ptr(rgb), _ptr(out), _ptr(scratch), self._width, self._height, self._sigma_s, self._sigma_r,
                                  float(detail), _stream(rgb.device)))
    return out


class Laplacian(_Workspace):
  def __init__(self, device, width, height, num_gamma: int = 6, sigma: float = 0.2, shadows: float = 1.0, highlights: float = 1.0,
               clarity: float = 0.0):
    if int(num_gamma) != 6:
      raise RuntimeError(f'Unsupported gamma count: {num_gamma}')
    super().__init__(device, width, height)
    self.sigma, self.shadows, self.highlights, self.clarity = float(sigma), float(shadows), float(highlights), float(clarity)

  def process(self, input: torch.Tensor) -> torch.Tensor:
    _require(input.dtype == torch.float32, 'Input tensor must be float32')
    _require(input.dim() == 2, 'Input tensor must be 2D')
    _require(input.size(0) == self._height and input.size(1) == self._width, 'Input tensor dimensions must match workspace dimensions')
    _require(input.is_cuda, 'Input tensor must be on CUDA device')
    src = input.contiguous()
    out = torch.empty_like(src)
    scratch = self._ensure_scratch(lib.tdb_laplacian_scratch_bytes(self._width, self._height))
    with torch.cuda.device(src.device):
      check(lib.tdb_laplacian(_ptr(src), _ptr(out), _ptr(scratch), self._width, self._height, self.sigma, self.shadows, self.highlights,
                              self.clarity, _stream(src.device)))
    return out


# ---------------------------------------------------------------------------------------------------------------
# fused frame pipeline (include/tdb200.h "Fused frame pipeline"; what ImageProcessor.process_image_set runs)
class FramePipeline:
  """Stage calls of ImageProcessor in their fused form.  Holds the device-resident statistics state of one processor; the
  workspaces (PostProcess, Wiener, Bilateral of this module) are passed in so that their scratch buffers are shared with the
  stage-by-stage API.  An image set is a run of frames bracketed by first / last."""

  def __init__(self, device: torch.device, width: int, height: int, pattern):
    self._device, self._width, self._height, self._filters = device, int(width), int(height), _filters(pattern)
    self._state = torch.zeros(lib.tdb_frame_state_bytes(), dtype=torch.uint8, device=device)

  def _s(self):
    return _stream(self._device)

  def smooth_deferred(self, post: 'PostProcess', rgb: torch.Tensor, first: bool, last: bool, prev_bounds: torch.Tensor | None,
                      moving_average: float, bounds_out: torch.Tensor):
    """Smoothing passes of `post`; the global green equilibration is deferred to `prepare`.  Returns (smoothed, ratio)."""
    _rgb_image(rgb)
    _require(rgb.size(0) == self._height and rgb.size(1) == self._width, 'image size mismatch')
    out = torch.empty_like(rgb)
    ratio = torch.empty(1, dtype=torch.float32, device=rgb.device)
    scratch = post._ensure_scratch(lib.tdb_postprocess_scratch_bytes(self._width, self._height))
    with torch.cuda.device(self._device):
      check(lib.tdb_postprocess_deferred(_ptr(rgb), _ptr(out), _ptr(scratch), self._width, self._height, self._filters,
                                         post.color_smoothing_passes, 8, _ptr(self._state), int(first), int(last), _ptr(prev_bounds),
                                         float(moving_average), _ptr(bounds_out), _ptr(ratio), self._s()))
    return out, ratio

  def smooth_band(self, post: 'PostProcess', rgb: torch.Tensor, row_lo: int, row_hi: int):
    """smooth_deferred for a rank's band of a frame that is split across GPUs: returns (smoothed, raw) where raw[6] = G1 sum,
    G2 sum, min / max of the sampled G1 greens, min / max of every other sampled value over the band rows [row_lo, row_hi) only."""
    _rgb_image(rgb)
    out = torch.empty_like(rgb)
    raw = torch.empty(6, dtype=torch.float32, device=rgb.device)
    scratch = post._ensure_scratch(lib.tdb_postprocess_scratch_bytes(self._width, self._height))
    with torch.cuda.device(self._device):
      check(lib.tdb_postprocess_deferred_band(_ptr(rgb), _ptr(out), _ptr(scratch), self._width, self._height, self._filters,
                                              post.color_smoothing_passes, 8, int(row_lo), int(row_hi), _ptr(raw), self._s()))
    return out, raw

  def metric_sums_band(self, rgb: torch.Tensor, bilateral: 'Bilateral | None', detail: float, row_lo: int, row_hi: int,
                       lab_input: bool = False, stride: int = 8, min_gray: float = 1e-4) -> torch.Tensor:
    """The six raw metric sums of the (sliced) band rows [row_lo, row_hi): what the ranks all-reduce."""
    sums = torch.empty(6, dtype=torch.float32, device=rgb.device)
    grid = bilateral._grid_scratch() if bilateral is not None else None
    ss, sr = (bilateral._sigma_s, bilateral._sigma_r) if bilateral is not None else (1.0, 1.0)
    with torch.cuda.device(self._device):
      check(lib.tdb_metrics_sliced_band(_ptr(rgb), int(lab_input), _ptr(grid), self._width, self._height, ss, sr, float(detail), int(stride),
                                        float(min_gray), _ptr(self._state), int(row_lo), int(row_hi), _ptr(sums), self._s()))
    return sums

  def prepare(self, rgb: torch.Tensor, ratio: torch.Tensor | None, bounds: torch.Tensor, wiener: 'Wiener | None', eps: float = 1e-4):
    """green_eq_global (ratio) + normalize (bounds).  With `wiener` the result is the (H, W, 2) Lab (a, b) plane of the normalised
    colour, and its log-luminance and a cleared accumulator sit in the Wiener scratch: the input of `denoise(..., prepared=True)`."""
    _rgb_image(rgb)
    out = torch.empty_like(rgb) if wiener is None else torch.empty((self._height, self._width, 2), dtype=torch.float32, device=rgb.device)
    scratch = wiener._ensure_scratch(lib.tdb_wiener_scratch_bytes(self._width, self._height, 1, wiener._tile)) if wiener is not None else None
    with torch.cuda.device(self._device):
      check(lib.tdb_frame_prepare(_ptr(rgb), _ptr(out), _ptr(scratch), self._width, self._height, self._filters, _ptr(ratio), _ptr(bounds),
                                  float(eps), self._s()))
    return out

  def denoise(self, wiener: 'Wiener', rgb: torch.Tensor, noise: float, prepared: bool, bilateral: 'Bilateral | None', eps: float = 1e-4):
    """Wiener.process_log_luminance; with `bilateral` the blurred grid of the result is built in its scratch on the way and the
    result is returned as Lab (rgb_to_lab), to be passed on with lab_input=True.
    prepared: `rgb` is what `prepare(..., wiener)` returned (the Lab (a, b) plane)."""
    if prepared:
      _require(rgb.is_cuda and rgb.dtype == torch.float32 and rgb.is_contiguous() and rgb.shape == (self._height, self._width, 2),
               'prepared input must be the (H, W, 2) plane returned by prepare()')
    else:
      _rgb_image(rgb)
    out = torch.empty((self._height, self._width, 3), dtype=torch.float32, device=rgb.device)
    scratch = wiener._ensure_scratch(lib.tdb_wiener_scratch_bytes(self._width, self._height, 1, wiener._tile))
    grid = bilateral._grid_scratch() if bilateral is not None else None
    ss, sr = (bilateral._sigma_s, bilateral._sigma_r) if bilateral is not None else (0.0, 0.0)
    with torch.cuda.device(self._device):
      check(lib.tdb_wiener_log_luminance_fused(_ptr(rgb), _ptr(out), _ptr(scratch), self._width, self._height, wiener._tile,
                                               wiener._overlap, float(noise), float(eps), 2 if prepared else 0, _ptr(grid), ss, sr, self._s()))
    return out

  def bilateral_grid(self, bilateral: 'Bilateral', rgb: torch.Tensor):
    _rgb_image(rgb)
    with torch.cuda.device(self._device):
      check(lib.tdb_bilateral_grid_rgb(_ptr(rgb), _ptr(bilateral._grid_scratch()), self._width, self._height, bilateral._sigma_s,
                                       bilateral._sigma_r, self._s()))

  def metrics(self, rgb: torch.Tensor, bilateral: 'Bilateral | None', detail: float, first: bool, last: bool,
              prev_metrics: torch.Tensor | None, moving_average: float, metrics_out: torch.Tensor, stride: int = 8, min_gray: float = 1e-4,
              lab_input: bool = False):
    _rgb_image(rgb)
    grid = bilateral._grid_scratch() if bilateral is not None else None
    ss, sr = (bilateral._sigma_s, bilateral._sigma_r) if bilateral is not None else (1.0, 1.0)
    with torch.cuda.device(self._device):
      check(lib.tdb_metrics_sliced(_ptr(rgb), int(lab_input), _ptr(grid), self._width, self._height, ss, sr, float(detail), int(stride), float(min_gray),
                                   _ptr(self._state), int(first), int(last), _ptr(prev_metrics), float(moving_average), _ptr(metrics_out),
                                   self._s()))

  def slice_tonemap(self, rgb: torch.Tensor, bilateral: 'Bilateral', detail: float, op: str, metrics: torch.Tensor | None, params,
                    matrix: torch.Tensor | None = None, transform: str = 'none', lab_input: bool = False,
                    out: torch.Tensor | None = None) -> torch.Tensor:
    _rgb_image(rgb)
    h, w = self._height, self._width
    out = _tonemap_out(out, h, w, transform, rgb.device)
    with torch.cuda.device(self._device):
      check(lib.tdb_bilateral_slice_tonemap(_ptr(rgb), int(lab_input), _ptr(bilateral._grid_scratch()), _ptr(out), w, h, bilateral._sigma_s,
                                            bilateral._sigma_r, float(detail), _TM[op], _ptr(None if op == 'aces' else metrics), params.gamma,
                                            params.intensity, params.light_adapt, params.vibrance, _ptr(matrix), _TF[transform], self._s()))
    return out


def band_stats_finish(gathered: torch.Tensor, prev_bounds: torch.Tensor | None, moving_average: float, bounds_out: torch.Tensor):
  """(world, 6) gathered band statistics -> green ratio (returned) and the EMA of the bounds (written into bounds_out, which may be
  prev_bounds itself); one kernel (csrc/fused.cu)."""
  _cuda_f32(gathered, 'gathered')
  g = gathered.contiguous()
  ratio = torch.empty(1, dtype=torch.float32, device=g.device)
  with torch.cuda.device(g.device):
    check(lib.tdb_band_stats_finish(_ptr(g), g.size(0), _ptr(prev_bounds), float(moving_average), _ptr(bounds_out), _ptr(ratio), _stream(g.device)))
  return ratio


def band_metrics_finish(sums: torch.Tensor, prev_metrics: torch.Tensor | None, moving_average: float, metrics_out: torch.Tensor):
  _cuda_f32(sums, 'sums')
  with torch.cuda.device(sums.device):
    check(lib.tdb_band_metrics_finish(_ptr(sums.contiguous()), _ptr(prev_metrics), float(moving_average), _ptr(metrics_out), _stream(sums.device)))


def launch_count() -> int:
  """Kernels launched through libtdb200 by this process so far."""
  return _lib.launch_count()


extension = SimpleNamespace(
  BayerPattern=BayerPattern, JpegInputFormat=JpegInputFormat, JpegSubsampling=JpegSubsampling, JpegException=JpegException,
  Jpeg=Jpeg, PPG=PPG, RCD=RCD, PostProcess=PostProcess, Laplacian=Laplacian, Bilateral=Bilateral, Wiener=Wiener,
  TonemapParams=TonemapParams,
  BGR=JpegInputFormat.BGR, RGB=JpegInputFormat.RGB, BGRI=JpegInputFormat.BGRI, RGBI=JpegInputFormat.RGBI,
  CSS_444=JpegSubsampling.CSS_444, CSS_422=JpegSubsampling.CSS_422, CSS_GRAY=JpegSubsampling.CSS_GRAY,
  compute_luminance=compute_luminance, modify_luminance=modify_luminance, compute_log_luminance=compute_log_luminance,
  modify_log_luminance=modify_log_luminance, modify_hsl=modify_hsl, modify_vibrance=modify_vibrance, rgb_to_xyz=rgb_to_xyz,
  xyz_to_lab=xyz_to_lab, lab_to_xyz=lab_to_xyz, xyz_to_rgb=xyz_to_rgb, rgb_to_lab=rgb_to_lab, lab_to_rgb=lab_to_rgb,
  color_transform_3x3=color_transform_3x3, encode12_u16=encode12_u16, encode12_float=encode12_float,
  decode12_float=decode12_float, decode12_half=decode12_half, decode12_u16=decode12_u16,
  compute_image_bounds=compute_image_bounds, compute_image_metrics=compute_image_metrics, reinhard_tonemap=reinhard_tonemap,
  aces_tonemap=aces_tonemap, adaptive_aces_tonemap=adaptive_aces_tonemap, linear_tonemap=linear_tonemap,
  bilinear5x5_demosaic=bilinear5x5_demosaic, apply_white_balance=apply_white_balance, estimate_white_balance=estimate_white_balance,
  # fused additions (not in the reference binding)
  unpack12_wb=unpack12_wb, demosaic_packed=demosaic_packed, normalize=normalize, lerp=lerp, tonemap=tonemap,
  FramePipeline=FramePipeline, image_metric_sums=image_metric_sums, metrics_from_sums=metrics_from_sums, green_sums=green_sums, green_eq_apply=green_eq_apply,
  launch_count=launch_count, channel_noise=channel_noise, band_stats_finish=band_stats_finish, band_metrics_finish=band_metrics_finish,
)

__all__ = ['extension']
