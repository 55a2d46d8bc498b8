"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the Python surface mirrors
the reference's, settings round-trip through JSON, input validation raises the reference's exception types."""

import ctypes
from pathlib import Path
import re

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / 'include' / 'tdb200.h'
LIB = ROOT / 'torch-darktable_b200' / 'torch_darktable' / 'lib' / 'libtdb200.so'


def declared_symbols():
  text = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
  return sorted(set(re.findall(r'\b(tdb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
  assert LIB.exists(), 'build the library first: python torch-darktable_b200/build.py'
  lib = ctypes.CDLL(str(LIB))
  names = declared_symbols()
  assert len(names) >= 40
  for name in names:
    assert hasattr(lib, name), f'{name} declared in include/tdb200.h but not exported'
  lib.tdb_version.restype = ctypes.c_int
  assert lib.tdb_version() >= 100


def test_binding_table_matches_header():
  from torch_darktable import _lib
  assert sorted(_lib.EXPORTED) == declared_symbols()


def test_scratch_size_queries_need_no_gpu():
  from torch_darktable._lib import lib
  assert lib.tdb_postprocess_scratch_bytes(3840, 2160) >= 2 * 3840 * 2160 * 12
  assert 0 <= lib.tdb_wiener_scratch_bytes(3840, 2160, 1, 32) - 3840 * 2160 * 2 * 4 <= 4096
  size = (ctypes.c_int * 3)()
  assert lib.tdb_bilateral_grid_size(4096, 3000, 2.0, 0.2, size) == 0
  assert tuple(size) == (2049, 1501, 6)
  assert lib.tdb_bilateral_grid_size(8192, 6144, 2.0, 0.2, size) == 0
  assert tuple(size) == (3001, 2251, 6)  # saturating case
  assert lib.tdb_laplacian_scratch_bytes(4096, 3000) > 0


def test_public_surface_matches_reference_names():
  import torch_darktable as td
  ext = td.extension.extension
  expected = ['BayerPattern', 'Bilateral', 'Jpeg', 'JpegException', 'JpegInputFormat', 'JpegSubsampling', 'Laplacian', 'PPG',
              'PostProcess', 'RCD', 'TonemapParams', 'Wiener', 'aces_tonemap', 'adaptive_aces_tonemap', 'apply_white_balance',
              'bilinear5x5_demosaic', 'color_transform_3x3', 'compute_image_bounds', 'compute_image_metrics', 'compute_log_luminance',
              'compute_luminance', 'decode12_float', 'decode12_half', 'decode12_u16', 'encode12_float', 'encode12_u16',
              'estimate_white_balance', 'lab_to_rgb', 'lab_to_xyz', 'linear_tonemap', 'modify_hsl', 'modify_log_luminance',
              'modify_luminance', 'modify_vibrance', 'reinhard_tonemap', 'rgb_to_lab', 'rgb_to_xyz', 'xyz_to_lab', 'xyz_to_rgb']
  for name in expected:
    assert hasattr(ext, name), name
  assert td.BayerPattern.RGGB.value.value == 0x94949494
  for name in ('decode12', 'encode', 'PPG', 'RCD', 'PostProcess', 'Wiener', 'Bilateral', 'Laplacian', 'TonemapParameters'):
    assert hasattr(td, name)


def test_camera_settings_roundtrip():
  """The reference's only automated test (tests/test_camera_settings_serialization.py)."""
  from torch_darktable.pipeline.camera_settings import load_camera_settings_from_dir
  settings = load_camera_settings_from_dir()
  assert set(settings) == {'artichoke', 'beetroot', 'carrot', 'pfr'}
  for s in settings.values():
    assert s == s.__class__.model_validate_json(s.model_dump_json())
  assert settings['artichoke'].image_size == (4096, 3000)
  assert settings['artichoke'].bytes == 4096 * 3000 * 3 // 2


def test_settings_validation():
  from pydantic import ValidationError
  from torch_darktable.pipeline import ImageProcessingSettings, get_preset
  with pytest.raises(ValidationError):
    ImageProcessingSettings(tone_gamma=9.0)
  assert get_preset('adaptive_aces').tone_mapping.name == 'adaptive_aces'
  with pytest.raises(ValueError):
    get_preset('nope')


def test_cpu_tensors_are_rejected_loudly():
  import torch
  import torch_darktable as td
  with pytest.raises(RuntimeError, match='CUDA'):
    td.bilinear5x5_demosaic(torch.zeros(16, 16, 1), td.BayerPattern.RGGB)
  with pytest.raises(RuntimeError, match='CUDA'):
    td.rgb_to_lab(torch.zeros(4, 4, 3))
  with pytest.raises(RuntimeError):
    td.decode12_float(torch.zeros(6, dtype=torch.uint8))
  with pytest.raises(ValueError):
    td.Wiener(torch.device('cpu'), (64, 64))


def test_bayer_helpers_roundtrip():
  import torch
  import torch_darktable as td
  rgb = torch.rand(8, 12, 3)
  for p in td.BayerPattern:
    cfa = td.rgb_to_bayer(rgb, p)
    assert cfa.shape == (8, 12, 1)
    assert torch.equal(td.bayer.expand_bayer(td.bayer.stack_bayer(cfa[..., 0])), cfa)
