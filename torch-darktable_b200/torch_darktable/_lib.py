"""ctypes binding of libtdb200.so (the C ABI declared in include/tdb200.h).

There is no CPU fallback: if the library is missing this module raises at import, and every entry point raises
RuntimeError when the CUDA call fails.  Build the library with `python torch-darktable_b200/build.py`.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / 'lib' / 'libtdb200.so'

if not _LIB_PATH.exists():
  raise ImportError(
    f'{_LIB_PATH} not found: the B200 CUDA library has not been built (run `python torch-darktable_b200/build.py`). '
    'torch_darktable has no CPU fallback.'
  )

lib = C.CDLL(str(_LIB_PATH))

_P, _I, _I64, _F, _U32, _SZ = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32, C.c_size_t

_SIGNATURES = {
  # name: (restype, argtypes)
  'tdb_version': (_I, []),
  'tdb_last_error': (C.c_char_p, []),
  'tdb_launch_count': (C.c_uint64, []),
  'tdb_set_concurrency_hint': (None, [_I]),
  'tdb_timing_begin': (None, [_P]),
  'tdb_timing_end': (_SZ, [C.c_char_p, _SZ]),
  'tdb_decode12_f32': (_I, [_P, _P, _I64, _I, _I, _P]),
  'tdb_decode12_f16': (_I, [_P, _P, _I64, _I, _I, _P]),
  'tdb_decode12_u16': (_I, [_P, _P, _I64, _I, _P]),
  'tdb_encode12_u16': (_I, [_P, _P, _I64, _I, _P]),
  'tdb_encode12_f32': (_I, [_P, _P, _I64, _I, _I, _P]),
  'tdb_unpack12_wb': (_I, [_P, _P, _I, _I, _I, _U32, _F, _P, _P]),
  'tdb_white_balance': (_I, [_P, _P, _I, _I, _U32, _P, _P]),
  'tdb_wb_collect_samples': (_I, [_P, _I, _I, _U32, _I, _P, _P, _P, _P]),
  'tdb_wb_estimate_gains': (_I, [_P, _P, _P, _I64, _F, _P, _P]),
  'tdb_bilinear5x5': (_I, [_P, _P, _I, _I, _U32, _P]),
  'tdb_ppg': (_I, [_P, _P, _I, _I, _U32, _F, _P]),
  'tdb_rcd': (_I, [_P, _P, _I, _I, _U32, _P]),
  'tdb_demosaic_packed': (_I, [_P, _P, _I, _I, _I, _U32, _I, _F, _P, _F, _P]),
  'tdb_postprocess_scratch_bytes': (_SZ, [_I, _I]),
  'tdb_postprocess': (_I, [_P, _P, _P, _I, _I, _U32, _I, _I, _I, _F, _P]),
  'tdb_green_sums': (_I, [_P, _I, _I, _U32, _P, _P, _P]),
  'tdb_green_eq_apply': (_I, [_P, _P, _I, _I, _U32, _I, _F, _P, _P]),
  'tdb_color_convert': (_I, [_P, _P, _I64, _I, _F, _F, _F, _P]),
  'tdb_color_transform_3x3': (_I, [_P, _P, _I64, _P, _P]),
  'tdb_compute_luminance': (_I, [_P, _P, _I64, _P]),
  'tdb_compute_log_luminance': (_I, [_P, _P, _I64, _F, _P]),
  'tdb_modify_luminance': (_I, [_P, _P, _P, _I64, _P]),
  'tdb_modify_log_luminance': (_I, [_P, _P, _P, _I64, _F, _P]),
  'tdb_normalize': (_I, [_P, _P, _I64, _P, _P]),
  'tdb_bounds_init': (_I, [_P, _P]),
  'tdb_bounds_accumulate': (_I, [_P, _I, _I, _I, _P, _P]),
  'tdb_metrics_init': (_I, [_P, _P]),
  'tdb_metrics_accumulate': (_I, [_P, _I, _I, _I, _F, _P, _P, _P]),
  'tdb_metrics_finalize': (_I, [_P, _P, _P]),
  'tdb_lerp': (_I, [_P, _P, _F, _P, _I, _P]),
  'tdb_tonemap': (_I, [_P, _P, _I, _I, _I, _P, _F, _F, _F, _F, _P, _I, _P]),
  'tdb_wiener_scratch_bytes': (_SZ, [_I, _I, _I, _I]),
  'tdb_wiener': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
  'tdb_wiener_log_luminance': (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _F, _P]),
  'tdb_bilateral_grid_size': (_I, [_I, _I, _F, _F, C.POINTER(C.c_int)]),
  'tdb_bilateral_scratch_bytes': (_SZ, [_I, _I, _F, _F]),
  'tdb_bilateral': (_I, [_P, _P, _P, _I, _I, _F, _F, _F, _P]),
  'tdb_bilateral_rgb': (_I, [_P, _P, _P, _I, _I, _F, _F, _F, _P]),
  'tdb_laplacian_scratch_bytes': (_SZ, [_I, _I]),
  'tdb_laplacian': (_I, [_P, _P, _P, _I, _I, _F, _F, _F, _F, _P]),
  'tdb_frame_state_bytes': (_SZ, []),
  'tdb_postprocess_deferred': (_I, [_P, _P, _P, _I, _I, _U32, _I, _I, _P, _I, _I, _P, _F, _P, _P, _P]),
  'tdb_frame_prepare': (_I, [_P, _P, _P, _I, _I, _U32, _P, _P, _F, _P]),
  'tdb_wiener_log_luminance_fused': (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _F, _I, _P, _F, _F, _P]),
  'tdb_metrics_sliced': (_I, [_P, _I, _P, _I, _I, _F, _F, _F, _I, _F, _P, _I, _I, _P, _F, _P, _P]),
  'tdb_bilateral_slice_tonemap': (_I, [_P, _I, _P, _P, _I, _I, _F, _F, _F, _I, _P, _F, _F, _F, _F, _P, _I, _P]),
  'tdb_postprocess_deferred_band': (_I, [_P, _P, _P, _I, _I, _U32, _I, _I, _I, _I, _P, _P]),
  'tdb_metrics_sliced_band': (_I, [_P, _I, _P, _I, _I, _F, _F, _F, _I, _F, _P, _I, _I, _P, _P]),
  'tdb_band_stats_finish': (_I, [_P, _I, _P, _F, _P, _P, _P]),
  'tdb_band_metrics_finish': (_I, [_P, _P, _F, _P, _P]),
  'tdb_bilateral_grid_rgb': (_I, [_P, _P, _I, _I, _F, _F, _P]),
  'tdb_channel_noise_scratch_bytes': (_SZ, [_I, _I, _I]),
  'tdb_channel_noise': (_I, [_P, _I, _I, _I, _P, _P, _P]),
  'tdb_probe_fp32': (_I, [_P, _I, C.POINTER(C.c_double), _P]),
  'tdb_probe_mufu': (_I, [_P, _I, C.POINTER(C.c_double), _P]),
  'tdb_jpeg_available': (_I, []),
  'tdb_jpeg_create': (_I, [C.POINTER(C.c_void_p)]),
  'tdb_jpeg_destroy': (_I, [_P]),
  'tdb_jpeg_encode': (_I, [_P, _P, _I, _I, _I64, _I64, _I, _I, _I, _I, C.POINTER(C.c_size_t), _P]),
  'tdb_jpeg_retrieve': (_I, [_P, _P, _SZ, C.POINTER(C.c_size_t), _P]),
}

for _name, (_res, _args) in _SIGNATURES.items():
  _fn = getattr(lib, _name)  # AttributeError here = header and library out of sync
  _fn.restype = _res
  _fn.argtypes = _args

EXPORTED = tuple(_SIGNATURES)


def last_error() -> str:
  msg = lib.tdb_last_error()
  return msg.decode('utf-8', 'replace') if msg else ''


def check(status: int) -> None:
  """Map a TDB_E* status to the exception type the reference raises (TORCH_CHECK -> RuntimeError)."""
  if status != 0:
    raise RuntimeError(last_error() or f'libtdb200 call failed with status {status}')


def launch_count() -> int:
  return int(lib.tdb_launch_count())


def timing_begin(stream_handle: int) -> None:
  lib.tdb_timing_begin(C.c_void_p(stream_handle))


def timing_end() -> dict[str, tuple[int, float]]:
  """{kernel name: (launches, total milliseconds)} since timing_begin."""
  buf = C.create_string_buffer(1 << 16)
  lib.tdb_timing_end(buf, len(buf))
  table = {}
  for line in buf.value.decode().splitlines():
    name, count, ms = line.rsplit(',', 2)
    table[name] = (int(count), float(ms))
  return table
