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


def test_raw_ingest_finds_the_camera_by_directory_then_by_size(tmp_path):
  """pipeline/camera_settings.py:55-132 of the reference: headerless packed files, camera from the directory name or the byte count."""
  import numpy as np
  import torch
  from torch_darktable.pipeline.camera_settings import (load_camera_settings_from_dir, load_raw_bytes, load_raw_bytes_stripped,
                                                        settings_for_file)
  known = load_camera_settings_from_dir()
  cam = known['artichoke']
  (tmp_path / 'artichoke').mkdir()
  by_dir = tmp_path / 'artichoke' / 'frame0.raw'
  by_dir.write_bytes(b'\x00' * 10)
  assert settings_for_file(by_dir).name == 'artichoke'
  (tmp_path / 'unknown').mkdir()
  by_size = tmp_path / 'unknown' / 'frame1.raw'
  payload = np.random.default_rng(3).integers(0, 256, cam.bytes, dtype=np.uint8)
  by_size.write_bytes(payload.tobytes())
  found = settings_for_file(by_size)
  assert found.bytes == cam.bytes
  raw = load_raw_bytes(by_size, torch.device('cpu'))
  assert raw.dtype == torch.uint8 and np.array_equal(raw.numpy(), payload)
  stripped = load_raw_bytes_stripped(by_size, found, torch.device('cpu'))
  assert stripped.numel() == found.bytes - found.padding
  bad = tmp_path / 'unknown' / 'frame2.raw'
  bad.write_bytes(b'\x00' * 12345)
  with pytest.raises(ValueError, match='Could not find camera settings'):
    settings_for_file(bad)


def test_estimate_channel_noise_matches_a_numpy_restatement():
  """denoise.py:131-158 of the reference: 4-neighbour Laplacian, every stride-th response, MAD / 0.6745 per channel (device-agnostic
  torch code in both packages; checked on the CPU against plain numpy, and against the sigma it is meant to recover)."""
  import numpy as np
  import torch
  import torch_darktable as td
  rng = np.random.default_rng(11)
  sig = np.array([0.01, 0.02, 0.04], np.float32)
  img = (0.5 + rng.normal(0.0, 1.0, (160, 200, 3)) * sig).astype(np.float32)
  got = td.estimate_channel_noise(torch.from_numpy(img), stride=4).numpy()
  pad = np.pad(img, ((1, 1), (1, 1), (0, 0)))
  resp = 4 * pad[1:-1, 1:-1] - pad[:-2, 1:-1] - pad[2:, 1:-1] - pad[1:-1, :-2] - pad[1:-1, 2:]
  resp = resp[::4, ::4].reshape(-1, 3)

  def lower_median(a):  # torch.median returns the lower of the two middle values
    return np.sort(a, axis=0)[(a.shape[0] - 1) // 2]

  med = lower_median(resp)
  want = lower_median(np.abs(resp - med)) / 0.6745
  np.testing.assert_allclose(got, want, rtol=1e-4)
  # the Laplacian of white noise has sigma * sqrt(20): the estimate scales with the true sigma
  np.testing.assert_allclose(got / np.sqrt(20.0), sig, rtol=0.15)
