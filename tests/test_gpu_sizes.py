"""GPU parity, part 2: the CUDA path against the CPU oracle at sizes that span many CTA tiles.

The golden vectors of the reference (tests/golden) are 64x96 frames, i.e. a handful of tiles per kernel.  The kernels are
tiled (32x32, 56x32, 64x64 pixel tiles, 8-pixel Wiener strides, ...), so every stage is run here on seeded synthetic
frames whose sides are NOT multiples of the tile sizes, with widths both divisible by four (128-bit paths) and not
(scalar fallbacks), and compared with the oracle under the per-op tolerances of tests/cases.py.
"""

import numpy as np
import pytest

import cases
import synth

pytestmark = pytest.mark.gpu

SIZES = [(250, 372), (130, 202), (516, 1100)]  # (H, W): multi-tile + ragged, width % 4 != 0, larger


@pytest.fixture(scope='module')
def impl():
  import torch
  assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
  from cuda_impl import CudaImpl
  return CudaImpl()


@pytest.fixture(scope='module')
def oracle():
  return cases.OracleImpl()


def check(op, impl, oracle, params, ins):
  got = cases.run_case(impl, op, params, ins)
  want = cases.run_case(oracle, op, params, ins)
  problems = cases.check_outputs(op, got, want, oracle=True)
  assert not problems, f'{op} {params}: ' + '; '.join(problems)


def cfa_of(h, w, pattern, seed):
  return synth.mosaic(synth.scene_rgb(h, w, seed), pattern)


@pytest.mark.parametrize('h,w', SIZES)
@pytest.mark.parametrize('pattern', ['RGGB', 'GBRG'])
@pytest.mark.parametrize('op', ['rcd', 'ppg', 'bilinear5x5_demosaic'])
def test_demosaic(impl, oracle, op, pattern, h, w):
  h, w = h & ~1, w & ~1
  params = {'pattern': pattern}
  if op == 'ppg':
    params['median_threshold'] = 0.0
  check(op, impl, oracle, params, {'cfa': cfa_of(h, w, pattern, 7)})


@pytest.mark.parametrize('h,w', SIZES)
@pytest.mark.parametrize('passes,glob,loc', [(1, False, False), (3, True, False), (4, True, True), (5, False, False), (9, True, False),
                                             (0, True, True)])
def test_postprocess(impl, oracle, passes, glob, loc, h, w):
  rng = np.random.default_rng(3)
  rgb = synth.scene_rgb(h, w, 11) + rng.normal(0, 0.02, size=(h, w, 3)).astype(np.float32)  # some negatives: the clamps matter
  rgb[0::2, 1::2, 1] *= 1.04
  params = {'pattern': 'RGGB', 'color_smoothing_passes': passes, 'green_eq_local': loc, 'green_eq_global': glob,
            'green_eq_threshold': 4.0}
  check('postprocess', impl, oracle, params, {'rgb': rgb.astype(np.float32)})


@pytest.mark.parametrize('h,w', SIZES[:2])
@pytest.mark.parametrize('k,ov,c', [(32, 4, 1), (32, 4, 3), (32, 2, 1), (32, 8, 1), (16, 4, 1)])
def test_wiener(impl, oracle, k, ov, c, h, w):
  rng = np.random.default_rng(5)
  x = (synth.scene_rgb(h, w, 13)[..., :c] + rng.normal(0, 0.03, size=(h, w, c))).astype(np.float32)
  sigmas = [0.03, 0.05, 0.02][:c]
  check('wiener', impl, oracle, {'sigmas': sigmas, 'tile_size': k, 'overlap_factor': ov}, {'x': x})


# The shared-column kernel (K = 32, overlap 4; csrc/wiener.cu namespace shr) walks tile-row PAIRS in steps of eight tile pairs with
# a static split of the step sequence over the CTAs: sides below one tile / one step, odd numbers of tile rows, sides that are
# not multiples of the stride, frames narrower than the 88-column step buffer, and a frame large enough that CTAs start and stop
# in the middle of a tile row (1104 steps over 296 CTAs).
@pytest.mark.parametrize('h,w,c', [(20, 44, 1), (33, 70, 1), (33, 70, 3), (40, 64, 1), (47, 130, 3), (96, 56, 1), (121, 333, 1),
                                   (700, 1500, 1)])
def test_wiener_shared_columns_geometry(impl, oracle, h, w, c):
  rng = np.random.default_rng(7)
  x = (synth.scene_rgb(h, w, 23)[..., :c] + rng.normal(0, 0.03, size=(h, w, c))).astype(np.float32)
  sigmas = [0.03, 0.05, 0.02][:c]
  check('wiener', impl, oracle, {'sigmas': sigmas, 'tile_size': 32, 'overlap_factor': 4}, {'x': x})


def test_wiener_shared_columns_repeatable():
  """The CTAs hand spectra and accumulators over through shared memory between barriers: a missing one would show up as a
  sporadic, large difference between runs (the only legitimate run-to-run variation is the order of the global float atomics)."""
  import torch
  import torch_darktable as td
  h, w = 516, 1100
  x = torch.from_numpy((synth.scene_rgb(h, w, 29)[..., :1] + np.random.default_rng(8).normal(0, 0.03, size=(h, w, 1))).astype(np.float32)).cuda()
  wiener = td.Wiener(torch.device('cuda:0'), (w, h))
  sig = torch.tensor([0.04], device='cuda')
  first = wiener.process(x, sig).clone()
  for _ in range(8):
    again = wiener.process(x, sig)
    assert float((again - first).abs().max()) < 2e-6


@pytest.mark.parametrize('h,w', SIZES)
def test_wiener_log_luminance(impl, oracle, h, w):
  rng = np.random.default_rng(6)
  x = np.clip(synth.scene_rgb(h, w, 17) + rng.normal(0, 0.02, size=(h, w, 3)), 0, 1).astype(np.float32)
  check('wiener_log_luminance', impl, oracle, {'noise': 0.075, 'eps': 1e-4}, {'x': x})


@pytest.mark.parametrize('h,w', SIZES)
@pytest.mark.parametrize('ss,sr', [(2.0, 0.2), (8.0, 0.1)])
def test_bilateral(impl, oracle, ss, sr, h, w):
  x = synth.scene_rgb(h, w, 19)
  check('bilateral_rgb', impl, oracle, {'sigma_s': ss, 'sigma_r': sr, 'detail': 0.4}, {'x': x})
  check('bilateral', impl, oracle, {'sigma_s': ss, 'sigma_r': sr, 'detail': 0.2}, {'lum': np.ascontiguousarray(x[..., 1])})


@pytest.mark.parametrize('h,w', [(250, 372), (130, 204)])
@pytest.mark.parametrize('tm,deb,tf', [('adaptive_aces', 'rcd', 'rotate_270'), ('reinhard', 'ppg', 'rotate_90'),
                                       ('aces', 'bilinear', 'transpose'), ('linear', 'rcd', 'rotate_180'),
                                       ('adaptive_aces', 'rcd', 'flip_vert')])
def test_pipeline(impl, oracle, tm, deb, tf, h, w):
  frames = [synth.packed_frame(h, w, seed=40 + i) for i in range(3)]
  params = {'width': w, 'height': h, 'white_balance': [1.8, 1.0, 2.1], 'debayer': deb, 'tone_mapping': tm,
            'moving_average': 0.5, 'transform': tf}
  check('pipeline', impl, oracle, params, {'frame0': frames[0], 'frame1': frames[1], 'frame2': frames[2]})


@pytest.mark.parametrize('post,den,bil', [(True, True, True), (True, False, True), (True, True, False), (False, True, True),
                                          (False, False, False), (True, False, False)])
@pytest.mark.parametrize('tm,deb,pattern,h,w', [('adaptive_aces', 'rcd', 'RGGB', 200, 328), ('reinhard', 'ppg', 'GRBG', 200, 328),
                                                 ('aces', 'bilinear', 'BGGR', 200, 328), ('adaptive_aces', 'rcd', 'GBRG', 202, 330),
                                                 ('linear', 'rcd', 'RGGB', 300, 520)])
def test_fused_image_set_equals_stage_by_stage(post, den, bil, tm, deb, pattern, h, w):
  """process_image_set (fused kernels, device-resident statistics) against the same composite written with the public stage
  calls; three image sets of 2, 1 and 3 frames so that the set merge and the moving average are exercised.  202 x 330 has a
  width that is not a multiple of four (every scalar fall-back: no 128-bit paths, no bulk copies, no planar RCD tiles), 300 x 520
  is large enough for interior RCD / smoothing tiles."""
  import torch
  import torch_darktable as td
  from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
  from torch_darktable.pipeline.config import Debayer, ToneMapper
  dev = torch.device('cuda:0')
  settings = ImageProcessingSettings(enable_denoise=den, enable_bilateral=bil, postprocess=post, tone_gamma=1.5, tone_intensity=2.0,
                                     light_adapt=0.8, tone_mapping=ToneMapper[tm], vibrance=0.5, debayer=Debayer[deb], moving_average=0.3)
  tfs = {'a': ImageTransform.rotate_90, 'b': ImageTransform.none, 'c': ImageTransform.flip_horiz}
  procs = [ImageProcessor((w, h), td.BayerPattern[pattern], td.PackedFormat.Packed12, settings, dev, (1.8, 1.0, 2.1), tfs) for _ in range(2)]
  seed = 300
  for names in (('a', 'b'), ('c',), ('b', 'c', 'a')):
    frames = {}
    for name in names:
      frames[name] = torch.from_numpy(synth.packed_frame(h, w, seed=seed, pattern=pattern)).to(dev)
      seed += 1
    fused = procs[0].process_image_set(frames)
    staged = procs[1].process_image_set_by_stage(frames)
    np.testing.assert_allclose(procs[0].bounds.cpu().numpy(), procs[1].bounds.cpu().numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(procs[0].metrics.cpu().numpy(), procs[1].metrics.cpu().numpy(), rtol=1e-5, atol=1e-7)
    for name in names:
      a, b = fused[name].cpu().numpy().astype(np.int16), staged[name].cpu().numpy().astype(np.int16)
      assert a.shape == b.shape
      diff = np.abs(a - b)
      assert diff.max() <= 1 and (diff > 0).mean() <= 1e-4, (name, diff.max(), (diff > 0).mean())


@pytest.mark.parametrize('pattern', ['RGGB', 'BGGR', 'GRBG', 'GBRG'])
def test_estimate_white_balance(pattern):
  """SURVEY.md 8f rank 1.  The reference leaves the last row / column of its sample arrays uninitialised (white_balance.cu:69,
  :107-109), so its own result is not reproducible; both the oracle and the CUDA path define those cells as invalid.  Checked
  against the oracle, and against the colour cast of a neutral scene shot through known channel gains."""
  import torch
  import torch_darktable as td
  import oracle
  h, w = 516, 1100
  gains = np.array([1.8, 1.0, 2.1], np.float32)
  rng = np.random.default_rng(5)
  grey = np.clip(synth.scene_rgb(h, w, 23)[..., 1:2] * 0.6 + rng.normal(0, 0.002, (h, w, 1)).astype(np.float32), 0.02, 0.95)
  rgb = np.repeat(grey, 3, axis=2) / gains  # a neutral scene seen through a sensor that needs `gains`
  images = [synth.mosaic(rgb.astype(np.float32), pattern), synth.mosaic(rgb[::-1].copy().astype(np.float32), pattern)]
  want = oracle.estimate_white_balance(images, pattern, 0.95, 8)
  dev = torch.device('cuda:0')
  got = td.estimate_white_balance([torch.from_numpy(i).to(dev) for i in images], td.BayerPattern[pattern], 0.95, 8).cpu().numpy()
  np.testing.assert_allclose(got, want, rtol=2e-4)
  np.testing.assert_allclose(got, 1.0 / gains, rtol=0.02)  # the estimate is the cast (R/G, 1, B/G) of the neutral scene


@pytest.mark.parametrize('quantile', [0.0, 0.5, 0.95, 0.98, 1.0])
def test_estimate_white_balance_matches_the_torch_formulation(quantile):
  """The selection kernel (tdb_wb_estimate_gains) against the reference's own composition of the second phase
  (white_balance.cu:135-161: masked gather, torch.quantile, masked gather, mean) on the sample arrays of the first phase, incl. the
  quantile's float32 rank arithmetic; saturated patches make about a fifth of the samples invalid."""
  import torch
  import torch_darktable as td
  from torch_darktable import _lib
  from torch_darktable.extension import _filters
  dev = torch.device('cuda:0')
  h, w, stride = 776, 1032, 8
  images = []
  for seed in (3, 4, 5):
    cfa = synth.mosaic(synth.scene_rgb(h, w, seed), 'RGGB') * 1.6  # pushes the highlights beyond 1: invalid samples
    images.append(torch.from_numpy(np.ascontiguousarray(cfa, np.float32)).to(dev))
  got = td.estimate_white_balance(images, td.BayerPattern.RGGB, quantile, stride)
  sh, sw = h // stride, w // stride
  n = sh * sw
  chroma = torch.empty((3 * n, 2), dtype=torch.float32, device=dev)
  inten = torch.empty(3 * n, dtype=torch.float32, device=dev)
  valid = torch.empty(3 * n, dtype=torch.uint8, device=dev)
  for i, img in enumerate(images):
    rc = _lib.lib.tdb_wb_collect_samples(img.data_ptr(), w, h, _filters(td.BayerPattern.RGGB.value),
                                         stride, chroma[i * n:].data_ptr(), inten[i * n:].data_ptr(), valid[i * n:].data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
  mask = valid.bool()
  assert 0.02 < 1.0 - mask.float().mean().item() < 0.9
  c, v = chroma[mask], inten[mask]
  bright = c[v >= torch.quantile(v, quantile)]
  mean = bright.mean(0)
  want = torch.stack((mean[0] / mean[1], torch.tensor(1.0, device=dev), (1.0 - mean[0] - mean[1]) / mean[1]))
  np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=2e-6)
  assert int((v >= torch.quantile(v, quantile)).sum()) >= 1


def test_estimate_white_balance_without_valid_samples():
  """every 2x2 patch saturated -> (1, 1, 1) (white_balance.cu:143-145); an image smaller than the stride has no sample site at all"""
  import torch
  import torch_darktable as td
  dev = torch.device('cuda:0')
  img = torch.full((64, 96, 1), 1.5, dtype=torch.float32, device=dev)
  assert td.estimate_white_balance([img], td.BayerPattern.RGGB, 0.95, 8).cpu().tolist() == [1.0, 1.0, 1.0]
  assert td.estimate_white_balance([img[:4, :4]], td.BayerPattern.RGGB, 0.95, 8).cpu().tolist() == [1.0, 1.0, 1.0]


@pytest.mark.parametrize('h,w', [(516, 1100), (1030, 700), (250, 372)])
@pytest.mark.parametrize('sigma,shadows,highlights,clarity', [(0.2, 1.0, 1.0, 0.0), (0.3, 1.4, 0.7, 0.25)])
def test_laplacian(impl, oracle, sigma, shadows, highlights, clarity, h, w):
  """Sizes whose replicate padding (2^(levels-1) = 256 / 512 px) is wider than a CTA's patch, so that the kernels' flat-tile paths
  (one row / column / pixel computed and replicated) run next to the ordinary ones, at every pyramid level."""
  lum = synth.scene_rgb(h, w, 31)[..., 1].copy()
  check('laplacian', impl, oracle, {'sigma': sigma, 'shadows': shadows, 'highlights': highlights, 'clarity': clarity}, {'lum': lum})


def test_host_frame_runner_streams_batches():
  """HostFrameRunner: pinned host frames in, pinned host results out, three streams; consecutive batches queued without a host
  wait in between (after_caller=False) hand their slots over by events.  Every result must equal ImageProcessor.process (up to the run-to-run noise of the float atomics)."""
  import torch
  import torch_darktable as td
  from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
  from torch_darktable.pipeline.batch import HostFrameRunner
  w, h = 256, 192
  dev = torch.device('cuda:0')
  settings = ImageProcessingSettings(moving_average=1.0)
  make = lambda: ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, (1.8, 1.0, 2.1), ImageTransform.rotate_90)
  frames = [torch.from_numpy(synth.packed_frame(h, w, seed=40 + i)).pin_memory() for i in range(11)]
  want = [make().process(f.to(dev), 'cam').cpu() for f in frames]
  runner = HostFrameRunner(make(), slots=3)
  outs = [torch.empty((w, h, 3), dtype=torch.uint8).pin_memory() for _ in frames]
  runner.run(frames[:5], outs[:5], after_caller=False)   # 5, 4 and 2 frames: the slot rotation continues across the calls
  runner.run(frames[5:9], outs[5:9], after_caller=False)
  runner.run(frames[9:], outs[9:])
  runner.wait()
  # two runs of the same frame are not bit-identical (the Wiener accumulator is filled with float atomics in arrival order): a frame
  # that went through the wrong slot or an unfinished copy would differ everywhere, a legitimate one by 1 LSB on a few samples
  for i, (got, ref) in enumerate(zip(outs, want)):
    d = (got.to(torch.int16) - ref.to(torch.int16)).abs()
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) <= 1e-4, f'frame {i}: max {int(d.max())} LSB, {float((d > 0).float().mean()):.2e} differ'


# ---- repeatability of the asynchronous / last-CTA paths (compute-sanitizer is closed on this pool: a race has to show up as a
# run-to-run difference) ------------------------------------------------------------------------------------------------------
def test_smoothing_bulk_copy_path_is_repeatable():
  """PostProcess with 3 smoothing passes stages its interior patches with cp.async.bulk + mbarrier (csrc/postprocess.cu) and reduces the
  green sums / bounds through a last-CTA ticket (csrc/frame_state.cuh): sixteen runs on the same 516 x 1100 input (several waves of
  CTAs) must be bit-identical -- medians are exact and the partials are summed in a fixed order."""
  import torch
  import torch_darktable as td
  h, w = 516, 1100
  rng = np.random.default_rng(3)
  rgb = torch.from_numpy((synth.scene_rgb(h, w, 11) + rng.normal(0, 0.02, (h, w, 3))).astype(np.float32)).cuda()
  pp = td.PostProcess(torch.device('cuda:0'), (w, h), td.BayerPattern.RGGB, color_smoothing_passes=3, green_eq_global=True)
  first = pp.process(rgb).clone()
  for _ in range(15):
    assert torch.equal(pp.process(rgb), first)


def test_frame_statistics_tickets_are_repeatable():
  """The fused frame pipeline's statistics (green ratio, bounds, metrics) come out of last-CTA ticket reductions: the same frame through
  a fresh processor sixteen times gives identical bounds and metrics (the metrics are sums of the Wiener output, whose float atomics
  may reorder: 1e-6), and the strip RCD (TMA-staged rows, mbarrier hand-over) gives bit-identical RGB."""
  import torch
  import torch_darktable as td
  from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
  from torch_darktable.pipeline.config import Debayer, ToneMapper
  h, w = 516, 1100
  frame = torch.from_numpy(synth.packed_frame(h, w, seed=21)).cuda()
  settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, enable_denoise=True, enable_bilateral=True,
                                     postprocess=True, tone_gamma=1.5, tone_intensity=2.0, light_adapt=0.8, vibrance=0.5, moving_average=1.0)
  ref_rgb = td.demosaic_packed(frame, (w, h), td.BayerPattern.RGGB, method='rcd')
  bounds, metrics = [], []
  for _ in range(16):
    assert torch.equal(td.demosaic_packed(frame, (w, h), td.BayerPattern.RGGB, method='rcd'), ref_rgb)
    proc = ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, torch.device('cuda:0'), None, ImageTransform.none)
    proc.process(frame, 'cam')
    bounds.append(proc.bounds.clone()), metrics.append(proc.metrics.clone())
  for b, m in zip(bounds[1:], metrics[1:]):
    assert torch.equal(b, bounds[0])
    assert torch.allclose(m, metrics[0], atol=1e-6, rtol=0)
