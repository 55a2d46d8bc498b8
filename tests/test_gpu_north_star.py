"""GPU parity, part 5: north_star's two additions to the reference surface -- the black level fused into the unpack, and the 3x3
colour matrix fused into the tone-map epilogue -- plus `color_transform_3x3`, whose reference op faults on ordinary device memory
(it dereferences the device matrix on the host, csrc/color_conversions.cu:158-159), so its oracle is the formula of
csrc/device_conversions.h:209-211.

The reference has no fused form of either, so parity means: fused == the same arithmetic written with the stage calls the reference
does have (decode12_float, a float subtraction, apply_white_balance, the demosaics, the tone mappers), evaluated by the CPU oracle.
  unpack12_wb                   bit-exact (integer unpack, one multiply, one subtract, one multiply + clamp: no re-association)
  demosaic_packed(black=...)    the demosaic tolerances of tests/cases.py against the oracle on the oracle-unpacked CFA
  tonemap(matrix=...)           <= 1 LSB on <= 1e-3 of the uint8 samples against oracle.tonemap(M . rgb)
  color_transform_3x3           2e-5 (cases.py's colour-op tolerance)
"""

import numpy as np
import pytest

import cases
import synth

pytestmark = pytest.mark.gpu

PATTERNS = ['RGGB', 'BGGR', 'GRBG', 'GBRG']
MATRIX = np.array([[1.6, -0.4, -0.2], [-0.3, 1.5, -0.2], [0.05, -0.5, 1.45]], np.float32)  # sRGB-ish camera matrix, rows sum to 1


@pytest.fixture(scope='module')
def td():
  import torch
  assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
  import torch_darktable
  return torch_darktable


@pytest.fixture(scope='module')
def oracle():
  import oracle
  return oracle


def dev(a):
  import torch
  return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
  return t.detach().cpu().numpy()


def oracle_unpack(oracle, packed, h, w, pattern, ids, black, gains):
  """decode12_float -> minus black (float32) -> apply_white_balance: the stage-by-stage definition of the fused unpack."""
  cfa = oracle.decode12(packed, np.float32, ids, True).reshape(h, w)
  cfa = (cfa - np.float32(black)).astype(np.float32)
  return oracle.white_balance(cfa, np.asarray(gains, np.float32), pattern) if gains is not None else cfa


@pytest.mark.parametrize('pattern', PATTERNS)
@pytest.mark.parametrize('ids', [False, True])
@pytest.mark.parametrize('black,gains', [(0.0, None), (0.0625, None), (0.0625, (1.8, 1.0, 2.1)), (0.015, (2.2, 1.0, 1.3)), (0.0, (1.8, 1.0, 2.1))])
def test_unpack12_wb_bit_exact(td, oracle, pattern, ids, black, gains):
  """Uniform random bytes (every 12-bit value occurs), ragged size (tail path of the kernel included): bit-exact."""
  h, w = 38, 116  # 4408 pixels: 275 full 16-pixel groups + a tail of 4 pairs
  packed = np.random.default_rng(5).integers(0, 256, h * w * 3 // 2, dtype=np.uint8)
  want = oracle_unpack(oracle, packed, h, w, pattern, ids, black, gains)
  got = host(td.extension.extension.unpack12_wb(dev(packed), w, h, td.BayerPattern[pattern], ids, black,
                                                 dev(np.asarray(gains, np.float32)) if gains is not None else None))
  assert got.shape == (h, w) and got.dtype == np.float32
  assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f'{int((got != want).sum())} of {got.size} samples differ'
  if black > 0 and gains is None:
    assert got.min() < 0.0  # the black level is a plain subtraction: no clamp without white balance


def test_unpack12_wb_24mp_equals_stages(td):
  """At 6000 x 4000 (configs[1]) the fused unpack equals decode12_float -> subtract -> apply_white_balance of the package, bit for bit."""
  import torch
  h, w = 4000, 6000
  packed = torch.randint(0, 256, (h * w * 3 // 2,), dtype=torch.uint8, device='cuda', generator=torch.Generator(device='cuda').manual_seed(3))
  gains = torch.tensor([1.8, 1.0, 2.1], device='cuda')
  fused = td.extension.extension.unpack12_wb(packed, w, h, td.BayerPattern.RGGB, False, 0.0625, gains)
  staged = td.apply_white_balance(td.decode12(packed, torch.float32).view(h, w) - 0.0625, gains, td.BayerPattern.RGGB)
  assert torch.equal(fused, staged)


@pytest.mark.parametrize('method', ['bilinear', 'ppg', 'rcd'])
@pytest.mark.parametrize('pattern', ['RGGB', 'GBRG'])
@pytest.mark.parametrize('h,w', [(130, 204), (250, 372)])
def test_demosaic_packed_with_black_level(td, oracle, method, pattern, h, w):
  """Packed scene with a pedestal of 256 counts -> RGB with black = 256 / 4095 and white balance, against the oracle demosaic of
  the oracle-unpacked CFA (the interior tiles stage the packed bytes with funnel-shifted 32-bit words, the frame tiles bytewise)."""
  cfa = synth.mosaic(synth.scene_rgb(h, w, 7), pattern)
  q = np.floor(np.clip(cfa * 0.9 * 4095.0 + 256.0, 0, 4095) + 0.5).astype(np.uint16)
  packed = synth.pack12(q)
  black, gains = 256.0 / 4095.0, (1.8, 1.0, 2.1)
  cfa_o = oracle_unpack(oracle, packed, h, w, pattern, False, black, gains)
  op = {'bilinear': 'bilinear5x5_demosaic', 'ppg': 'ppg', 'rcd': 'rcd'}[method]
  want = {'bilinear': lambda: oracle.bilinear5x5(cfa_o, pattern), 'ppg': lambda: oracle.ppg(cfa_o, pattern, 0.0),
          'rcd': lambda: oracle.rcd(cfa_o, pattern)}[method]()
  got = host(td.demosaic_packed(dev(packed), (w, h), td.BayerPattern[pattern], method=method, black=black,
                                white_balance=dev(np.asarray(gains, np.float32))))
  msg = cases.compare(op, got, want, cases.ORACLE_TOLERANCE.get(op))
  assert msg is None, f'{method} {pattern} {h}x{w} with black level: {msg}'
  assert abs(float(got.mean()) - float(np.clip(0.9 * synth.scene_rgb(h, w, 7) * np.asarray(gains), 0, 1).mean())) < 0.02  # pedestal removed


@pytest.mark.parametrize('h,w', [(48, 64), (130, 202)])
def test_color_transform_3x3(td, oracle, h, w):
  """clip(M . rgb) (device_conversions.h:209-211) on values inside and outside [0, 1]; the matrix is an ordinary device tensor."""
  rng = np.random.default_rng(9)
  x = (synth.scene_rgb(h, w, 23) * 1.3 - 0.1 + rng.normal(0, 0.02, (h, w, 3))).astype(np.float32)
  want = oracle.color_convert(x, 'color_transform_3x3', list(MATRIX.reshape(-1)))
  got = host(td.color_transform_3x3(dev(x), dev(MATRIX)))
  assert got.min() >= 0.0 and got.max() <= 1.0
  assert np.abs(got - want).max() <= 2e-5
  ident = host(td.color_transform_3x3(dev(x), dev(np.eye(3, dtype=np.float32))))
  assert np.array_equal(ident, np.clip(x, 0, 1))


@pytest.mark.parametrize('op', ['reinhard', 'aces', 'adaptive_aces', 'linear'])
@pytest.mark.parametrize('transform', ['none', 'rotate_270'])
def test_tonemap_with_matrix_epilogue(td, oracle, op, transform):
  """tone map of M . rgb in ONE kernel == oracle.tonemap applied to the separately multiplied image (no clip in between: the tone
  curves take max(., 0) themselves, reference reinhard.cu:39-42)."""
  h, w = 96, 132
  x = synth.scene_rgb(h, w, 29)
  metrics = oracle.compute_image_metrics([x], 8, 1e-4, False)
  mx = np.einsum('ij,hwj->hwi', MATRIX, x).astype(np.float32)
  want = oracle.tonemap(mx, op, metrics, 1.5, 2.0, 0.8, 0.5)
  if transform == 'rotate_270':
    want = np.ascontiguousarray(np.rot90(want, k=-1))
  ext = td.extension.extension
  params = td.TonemapParameters(1.5, 2.0, 0.8, 0.5).to_cpp()
  got = host(ext.tonemap(dev(x), op, dev(metrics), params, dev(MATRIX), transform))
  msg = cases.compare('reinhard_tonemap', got, want)
  assert msg is None, f'{op} with matrix, {transform}: {msg}'
  plain = host(ext.tonemap(dev(x), op, dev(metrics), params, None, transform))
  assert (plain != got).mean() > 0.2, 'the matrix must change the picture'  # (the linear curve saturates most of this scene)
  same = host(ext.tonemap(dev(x), op, dev(metrics), params, dev(np.eye(3, dtype=np.float32)), transform))
  assert np.array_equal(same, plain), 'identity matrix == no matrix'


def test_slice_tonemap_with_matrix_equals_stages(td):
  """The fused bilateral slice + tone map with a matrix == Bilateral.process_rgb -> color matrix -> tone map through the stage calls
  (<= 1 LSB on <= 1e-3 of the samples)."""
  import torch
  h, w = 250, 372
  ext = td.extension.extension
  x = dev(synth.scene_rgb(h, w, 19))
  bil = td.Bilateral(torch.device('cuda:0'), (w, h), sigma_s=2.0, sigma_r=0.2)
  sliced = bil.process_rgb(x, 0.4)
  metrics = td.compute_image_metrics([sliced], stride=8)
  params = td.TonemapParameters(1.5, 2.0, 0.8, 0.5).to_cpp()
  staged = ext.tonemap(sliced, 'adaptive_aces', metrics, params, dev(MATRIX), 'rotate_270')
  frame = ext.FramePipeline(torch.device('cuda:0'), w, h, td.BayerPattern.RGGB.value)
  frame.bilateral_grid(bil._bilateral, x)
  fused = frame.slice_tonemap(x, bil._bilateral, 0.4, 'adaptive_aces', metrics, params, dev(MATRIX), 'rotate_270')
  msg = cases.compare('pipeline', host(fused), host(staged))
  assert msg is None, msg
