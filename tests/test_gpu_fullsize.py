"""GPU parity, part 3: the BASELINE.json shapes.

The golden vectors and the oracle comparisons of the other GPU tests run on small frames (the oracle needs seconds there).  At the
full sizes of BASELINE.json -- 24 MP packed Bayer, 4K pipeline frames, 50 MP local contrast, a 20 MP frame of the sharded batch --
the CUDA path is held to properties that do not depend on the size: codec round trips, constant images, fused = unfused, row bands
= whole frame, neutral parameters = identity, and (where the oracle still finishes in seconds) the oracle itself on one frame.
Generated on the device, seeded; every test states its tolerance.
"""

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu

MP24 = (4000, 6000)   # configs[1]
UHD = (2160, 3840)    # configs[2]
MP50 = (6144, 8192)   # configs[3]
MP20 = (3648, 5472)   # configs[4]
MP200 = (12288, 16384)  # configs[4], the oversize frame


@pytest.fixture(scope='module')
def td():
  import torch
  assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
  import torch_darktable
  return torch_darktable


def device_scene(h, w, seed):
  """(h, w, 3) float32 scene on the device: gradient + gratings + noise in [0.02, 1] (the recipe of tests/synth.py, in torch)."""
  import torch
  g = torch.Generator(device='cuda').manual_seed(seed)
  y = torch.arange(h, device='cuda', dtype=torch.float32)[:, None]
  x = torch.arange(w, device='cuda', dtype=torch.float32)[None, :]
  planes = []
  for c in range(3):
    v = 0.15 + 0.5 * (x / (w - 1) * (0.6 + 0.2 * c) + y / (h - 1) * (0.4 - 0.1 * c))
    v = v + 0.12 * torch.sin(2 * np.pi * (x + 0.5 * y) / 37.0 + 0.7 * c) + 0.10 * torch.sin(2 * np.pi * (y - 0.3 * x) / 211.0 + 1.3 * c)
    v = v + 0.1 * ((((x // 64) + (y // 64)) % 2) * 2 - 1) + 0.01 * torch.randn((h, w), device='cuda', generator=g)
    planes.append(v)
  return (torch.stack(planes, dim=2) * 1.25).clamp_(0.02, 1.0)


def to_u16(t):
  """int tensor with values < 32768 -> torch.uint16 (via int16: uint16 has no conversion kernels of its own)."""
  import torch
  return t.to(torch.int16).view(torch.uint16)


def device_packed(td, h, w, seed):
  """12-bit packed RGGB frame of the scene, built with the package's own encoder (its bit-exactness is the first test)."""
  import torch
  rgb = device_scene(h, w, seed)
  cfa = torch.empty((h, w), device='cuda')
  cfa[0::2, 0::2], cfa[0::2, 1::2], cfa[1::2, 0::2], cfa[1::2, 1::2] = rgb[0::2, 0::2, 0], rgb[0::2, 1::2, 1], rgb[1::2, 0::2, 1], rgb[1::2, 1::2, 2]
  u16 = to_u16(torch.floor(cfa * 4095.0 + 0.5).clamp_(0, 4095).to(torch.int32))
  return td.encode(u16.reshape(-1), td.PackedFormat.Packed12), u16


def test_codec_round_trip_24mp(td):
  """encode12 . decode12 = identity on all 24 M samples (bit-exact), the float decode is u16 * (1/4095) exactly, and the packed
  bytes follow the bit layout of csrc/packed.cu:8-18 (checked on the host for the first 3 MB)."""
  import torch
  h, w = MP24
  g = torch.Generator(device='cuda').manual_seed(7)
  u16 = to_u16(torch.randint(0, 4096, (h * w,), device='cuda', generator=g, dtype=torch.int32))
  packed = td.encode(u16, td.PackedFormat.Packed12)
  assert packed.dtype == torch.uint8 and packed.numel() == h * w * 3 // 2
  assert torch.equal(td.decode12(packed, torch.uint16).view(torch.int16), u16.view(torch.int16))
  f32 = td.decode12(packed, torch.float32)
  assert torch.equal(f32, u16.view(torch.int16).to(torch.float32) * float(np.float32(1.0 / 4095.0)))
  assert torch.equal(td.decode12(packed, torch.float16), f32.to(torch.float16))
  n = 2 * 1024 * 1024
  assert np.array_equal(packed[: n * 3 // 2].cpu().numpy(), synth.pack12(u16[:n].view(torch.int16).cpu().numpy().view(np.uint16)).reshape(-1))
  # IDS: the reference's encoder and decoder disagree about the low nibbles; two round trips are the identity
  ids = td.PackedFormat.Packed12_IDS
  once = td.decode12(td.encode(u16, ids), torch.uint16, ids)
  twice = td.decode12(td.encode(once, ids), torch.uint16, ids)
  assert torch.equal(twice.view(torch.int16), u16.view(torch.int16))
  assert torch.equal(once.view(torch.int16) >> 4, u16.view(torch.int16) >> 4)  # the high eight bits always survive


@pytest.mark.parametrize('method', ['bilinear', 'ppg', 'rcd'])
def test_demosaic_24mp_constant_and_fused(td, method):
  """A constant CFA demosaics to the constant (all three methods, every pixel); the demosaic straight from the packed bytes
  equals decode12 -> apply_white_balance -> demosaic (bit-exact for bilinear / PPG, 5e-6 for RCD: tests/cases.py)."""
  import torch
  h, w = MP24
  pat = td.BayerPattern.RGGB
  stage = {'bilinear': lambda x: td.bilinear5x5_demosaic(x, pat), 'ppg': td.PPG(torch.device('cuda'), (w, h), pat).process,
           'rcd': td.RCD(torch.device('cuda'), (w, h), pat).process}[method]
  const = torch.full((h, w, 1), 0.37, device='cuda')
  out = stage(const)
  assert out.shape == (h, w, 3)
  # bilinear and PPG reproduce a constant exactly; RCD's ratio estimates carry an epsilon in the denominator (1.25e-6 in the oracle)
  assert float((out - 0.37).abs().max()) <= (5e-6 if method == 'rcd' else 0.0)
  packed, _ = device_packed(td, h, w, 11)
  wb = torch.tensor([1.8, 1.0, 2.1], device='cuda')
  fused = td.demosaic_packed(packed, (w, h), pat, method=method, white_balance=wb)
  cfa = td.apply_white_balance(td.decode12(packed, torch.float32).view(h, w), wb, pat)
  unfused = stage(cfa.unsqueeze(-1))
  tol = 5e-6 if method == 'rcd' else 0.0
  assert float((fused - unfused).abs().max()) <= tol
  del out, fused, unfused, cfa
  torch.cuda.empty_cache()


def make_processor(td, w, h, transform='rotate_270', **kw):
  import torch
  from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
  from torch_darktable.pipeline.config import Debayer, ToneMapper
  settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, enable_denoise=True, enable_bilateral=True,
                                     postprocess=True, tone_gamma=1.5, tone_intensity=2.0, light_adapt=0.8, vibrance=0.5, moving_average=1.0, **kw)
  return ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, torch.device('cuda:0'), (1.8, 1.0, 2.1),
                        ImageTransform[transform])


def assert_u8_close(got, want, frac, beyond=0.0):
  """uint8 images: at most `frac` of the samples different and at most `beyond` of them more than 1 LSB apart."""
  import torch
  assert got.shape == want.shape and got.dtype == torch.uint8 and want.dtype == torch.uint8
  d = (got.cuda().to(torch.int16) - want.cuda().to(torch.int16)).abs()
  far = d > 1
  where = ''
  if bool(far.any()):
    idx = far.nonzero()
    where = f'; first at {idx[0].tolist()}, rows {int(idx[:, 0].min())}..{int(idx[:, 0].max())}, columns {int(idx[:, 1].min())}..{int(idx[:, 1].max())}'
  assert float(far.float().mean()) <= beyond, f'max difference {int(d.max())} LSB, {int(far.sum())} of {d.numel()} samples beyond 1 LSB{where}'
  assert float((d > 0).float().mean()) <= frac, f'{float((d > 0).float().mean()):.2e} of the samples differ'


def test_pipeline_4k_fused_equals_stages_and_oracle(td):
  """One 3840 x 2160 frame of the bench workload: the fused frame pipeline, the same composite through the public stage calls, and
  the CPU oracle (1.5 s at this size).  The two CUDA paths: <= 1 LSB everywhere, <= 1e-4 of the uint8 samples different.  Against the
  oracle the tolerance is the one of tests/cases.py (ORACLE_TOLERANCE['pipeline']): <= 1e-3 of the samples different, <= 2e-4 of them
  by more than 1 LSB -- RCD chooses directions with hard selects, and CPU arithmetic flips a few of them that the reference's own
  GPU arithmetic and ours take alike (profiles/r01_three_way_mid.log: cuda-ref 2e-7, oracle-ref 2e-2 on 133 samples of 516 x 1100)."""
  import torch
  import oracle
  h, w = UHD
  frame = synth.packed_frame(h, w, seed=1234)
  dev_frame = torch.from_numpy(frame).cuda()
  fused = make_processor(td, w, h).process_image_set({'cam': dev_frame})['cam']
  staged = make_processor(td, w, h).process_image_set_by_stage({'cam': dev_frame})['cam']
  assert fused.shape == (w, h, 3) and fused.dtype == torch.uint8
  assert_u8_close(fused, staged, 1e-4)
  ref = oracle.Pipeline(w, h, white_balance=(1.8, 1.0, 2.1), debayer='rcd', tone_mapping='adaptive_aces', moving_average=1.0,
                        transform='rotate_270').process_image_set([frame])[0]
  assert_u8_close(fused, torch.from_numpy(ref), 1e-3, beyond=2e-4)


def test_pipeline_20mp_row_bands_equal_whole_frame(td):
  """A 5472 x 3648 frame (the shape of the sharded batch) split into four row bands on one GPU -- halo exchange and the three
  statistics reductions through a thread collective -- equals the whole-frame pipeline: <= 5e-4 of the samples different, <= 1e-6 of
  them by more than 1 LSB (the reference's RCD leaves a position-dependent stale-cell band just inside its 7-px margin, SURVEY.md 8a6:
  one sample at row H - 8 of this frame)."""
  import torch
  from test_tiled import WB, make_settings, run_threads, split_rows
  from torch_darktable.pipeline import ImageProcessor, ImageTransform
  from torch_darktable.pipeline.tiled import TiledFrameProcessor
  h, w = MP20
  frame = synth.packed_frame(h, w, seed=77)
  dev = torch.device('cuda:0')
  settings = make_settings('rcd', 'adaptive_aces', 1.0)
  whole = ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, WB, ImageTransform.none)
  want = whole.process_image_set({'a': torch.from_numpy(frame).to(dev)})['a']
  rows = [[torch.from_numpy(r.copy()).to(dev) for r in split_rows(frame, w, h, 4, 32)]]
  got, _ = run_threads(4, lambda col: TiledFrameProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, WB, col), rows)
  assert_u8_close(torch.from_numpy(np.ascontiguousarray(got[0])), want, 5e-4, beyond=1e-6)


def test_pipeline_200mp_eight_row_bands_equal_whole_frame(td):
  """The 16384 x 12288 frame of BASELINE.json configs[4] split into EIGHT row bands of 1536 rows (+ 96-row halos), on one GPU through a
  thread collective (the NCCL form of the same code runs in tests/test_tiled.py when the box has several GPUs), against the untiled
  ImageProcessor on the same 201 MP frame.  sigma_s = 8: at sigma_s = 2 the full frame's grid saturates in y and the split refuses
  it (pipeline/tiled.py).  <= 5e-4 of the samples different, <= 1e-6 of them by more than 1 LSB (the stale-cell band of the
  reference's RCD just inside its margin depends on absolute positions)."""
  import torch
  from test_tiled import WB, make_settings, run_threads
  from torch_darktable.pipeline import ImageProcessor, ImageTransform
  from torch_darktable.pipeline.tiled import TiledFrameProcessor, partition_rows
  h, w = MP200
  dev = torch.device('cuda:0')
  packed, _ = device_packed(td, h, w, 201)
  settings = make_settings('rcd', 'adaptive_aces', 1.0, bil_sigma_spatial=8.0, bil_sigma_luminance=0.1)
  whole = ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, WB, ImageTransform.none)
  want = whole.process_image_set({'a': packed})['a']
  del whole
  torch.cuda.empty_cache()
  rb = w * 3 // 2
  rows = [[packed[y0 * rb: y1 * rb] for (y0, y1) in partition_rows(h, 8, 32)]]
  got, proc = run_threads(8, lambda col: TiledFrameProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, WB, col), rows)
  assert proc.band.y1 - proc.band.y0 == 1536
  assert_u8_close(torch.from_numpy(np.ascontiguousarray(got[0])), want, 5e-4, beyond=1e-6)


def test_wiener_4k_zero_noise_is_identity(td):
  """sigma = 0 makes every gain 1 - eps/P: the overlap-add of 135 k windowed tiles must give the image back (every pixel, borders
  included, 2e-5) -- a size-independent check of the tile geometry, the window normalisation and the shared column transforms."""
  import torch
  h, w = UHD
  x = torch.log(device_scene(h, w, 5)[..., 1:2].contiguous())
  out = td.Wiener(torch.device('cuda:0'), (w, h)).process(x, torch.zeros(1, device='cuda'))
  assert float((out - x).abs().max()) < 2e-5


def test_local_contrast_50mp_neutral_parameters(td):
  """8192 x 6144: bilateral with detail 0 returns max(0, L) exactly (the slice adds 0 * grid); the local Laplacian with
  shadows = highlights = 1, clarity = 0 is the identity up to its fp16 storage (2e-3, tests/cases.py)."""
  import torch
  h, w = MP50
  lum = device_scene(h, w, 9)[..., 1].contiguous()
  out = td.Bilateral(torch.device('cuda:0'), (w, h), sigma_s=8.0, sigma_r=0.1).process(lum, 0.0)
  assert torch.equal(out, lum.clamp_min(0.0))
  del out
  lap = td.Laplacian(torch.device('cuda:0'), (w, h), td.LaplacianParams(sigma=0.2, shadows=1.0, highlights=1.0, clarity=0.0)).process(lum)
  assert float((lap - lum).abs().max()) < 2e-3
