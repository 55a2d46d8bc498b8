"""Row-tile split of one oversize frame (torch_darktable.pipeline.tiled).

CPU part (`-m "not gpu"`): the host logic -- row partition, halo sizes, halo exchange, the three cross-rank reductions, cropping --
runs on CPU tensors with the oracle standing in for the stage kernels (tests/tiled_oracle_ops.py), between threads and between two
gloo processes, and must reproduce the oracle's untiled pipeline.
GPU part (`-m gpu`): the same split with the product kernels (libtdb200) as bands of one frame on one device, against the untiled
ImageProcessor, and -- when the box has two or more GPUs -- across NCCL ranks.
"""

import os
import socket
import threading

import numpy as np
import pytest
import torch

import synth
from torch_darktable.pipeline import ImageProcessingSettings
from torch_darktable.pipeline.config import Debayer, ToneMapper
from torch_darktable.pipeline.tiled import (DistCollective, ThreadCollective, TiledFrameProcessor, halo_rows, make_band,
                                            partition_rows)
import torch_darktable as td

WB = (1.8, 1.0, 2.1)


def make_settings(debayer='rcd', tone='adaptive_aces', ma=0.5, **kw):
  base = dict(enable_denoise=True, enable_bilateral=True, postprocess=True, tone_gamma=1.5, tone_intensity=2.0, light_adapt=0.8,
              tone_mapping=ToneMapper[tone], vibrance=0.5, debayer=Debayer[debayer], moving_average=ma)
  base.update(kw)
  return ImageProcessingSettings(**base)


def split_rows(packed: np.ndarray, width: int, height: int, world: int, align: int = 8):
  rb = width * 3 // 2
  return [packed[y0 * rb: y1 * rb] for (y0, y1) in partition_rows(height, world, align)]


def run_threads(world, make_proc, frames_rows):
  """frames_rows[f][r] = packed rows of frame f for rank r.  Returns out[f] = concatenated bands, plus rank 0's processor."""
  hub = ThreadCollective.Hub(world)
  results = [[None] * world for _ in frames_rows]
  procs = [None] * world
  errors = []

  def work(rank):
    try:
      proc = make_proc(ThreadCollective(hub, rank))
      procs[rank] = proc
      for f, rows in enumerate(frames_rows):
        results[f][rank] = proc.process(rows[rank])
    except BaseException as e:  # noqa: BLE001 - a dead rank would dead-lock the others on the barrier
      errors.append(e)
      hub.barrier.abort()

  threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
  for t in threads:
    t.start()
  for t in threads:
    t.join()
  if errors:
    raise errors[0]
  return [torch.cat(bands).cpu().numpy() for bands in results], procs[0]


def assert_srgb_close(got, want, frac=2e-4):
  assert got.shape == want.shape and got.dtype == np.uint8
  diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
  assert diff.max() <= 1, f'max diff {diff.max()}'
  assert (diff > 0).mean() <= frac, f'{(diff > 0).mean():.2e} of the samples differ'


# ---- host logic ----------------------------------------------------------------------------------------------------
def test_partition_rows_covers_the_frame_on_aligned_boundaries():
  for height, world in [(12288, 8), (3648, 8), (2160, 4), (250, 3), (64, 8), (3000, 7)]:
    bands = partition_rows(height, world)
    assert bands[0][0] == 0 and bands[-1][1] == height
    for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
      assert a1 == b0 and a1 % 8 == 0 and a1 > a0
  with pytest.raises(ValueError):
    partition_rows(31, 2)
  with pytest.raises(ValueError):
    partition_rows(32, 8)


def test_halo_rows_follow_the_enabled_stages():
  full = make_settings()
  assert halo_rows(full) == 64  # 10 + 3 + 31 + 6 -> 64
  assert halo_rows(make_settings(enable_denoise=False)) < halo_rows(full)
  assert halo_rows(make_settings(enable_denoise=False, enable_bilateral=False, postprocess=False)) == 16
  assert halo_rows(full) % 8 == 0
  band = make_band(12288, 3, 8, 64)
  assert (band.y0, band.y1, band.top, band.bottom) == (4608, 6144, 64, 64)
  assert make_band(12288, 0, 8, 64).top == 0 and make_band(12288, 7, 8, 64).bottom == 0
  with pytest.raises(ValueError):
    make_band(256, 0, 8, 64)


def oracle_untiled(h, w, frames, debayer, tone, ma):
  import oracle
  pipe = oracle.Pipeline(w, h, white_balance=WB, debayer=debayer, tone_mapping=tone, moving_average=ma)
  return [pipe.process_image_set([f])[0] for f in frames], pipe


@pytest.mark.parametrize('world,debayer,tone', [(2, 'rcd', 'adaptive_aces'), (3, 'ppg', 'reinhard')])
def test_thread_bands_match_untiled_oracle(world, debayer, tone):
  from tiled_oracle_ops import OracleOps
  h, w = 64 * 2 * world + 16, 200
  frames = [synth.packed_frame(h, w, seed=70 + i) for i in range(2 if debayer != 'rcd' else 1)]
  want, pipe = oracle_untiled(h, w, frames, debayer, tone, 0.5)
  settings = make_settings(debayer, tone)

  def make_proc(col):
    return TiledFrameProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, torch.device('cpu'), WB, col,
                               ops=OracleOps())

  rows = [[torch.from_numpy(r.copy()) for r in split_rows(f, w, h, world)] for f in frames]
  got, proc0 = run_threads(world, make_proc, rows)
  for g, e in zip(got, want):
    assert_srgb_close(g, e)
  np.testing.assert_allclose(proc0.bounds.numpy(), pipe.bounds, rtol=0, atol=0)
  np.testing.assert_allclose(proc0.metrics.numpy(), pipe.metrics, rtol=2e-5)


def _gloo_worker(rank, world, port, h, w, out_dir):
  import sys
  from pathlib import Path
  root = Path(__file__).resolve().parents[1]
  for p in (root, root / 'tests', root / 'torch-darktable_b200'):
    sys.path.insert(0, str(p))
  import torch.distributed as dist
  from tiled_oracle_ops import OracleOps
  os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    frame = synth.packed_frame(h, w, seed=91)
    proc = TiledFrameProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, make_settings('ppg', 'adaptive_aces', 1.0),
                               torch.device('cpu'), WB, DistCollective(), ops=OracleOps())
    own = torch.from_numpy(split_rows(frame, w, h, world)[rank].copy())
    np.save(os.path.join(out_dir, f'band{rank}.npy'), proc.process(own).numpy())
  finally:
    dist.destroy_process_group()


def test_gloo_two_ranks_match_untiled_oracle(tmp_path):
  import torch.multiprocessing as mp
  h, w, world = 272, 136, 2
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
  mp.spawn(_gloo_worker, args=(world, port, h, w, str(tmp_path)), nprocs=world, join=True)
  got = np.concatenate([np.load(tmp_path / f'band{r}.npy') for r in range(world)])
  want, _ = oracle_untiled(h, w, [synth.packed_frame(h, w, seed=91)], 'ppg', 'adaptive_aces', 1.0)
  assert_srgb_close(got, want[0])


# ---- product kernels -----------------------------------------------------------------------------------------------
def untiled_cuda(h, w, frames, settings, dev):
  from torch_darktable.pipeline import ImageProcessor, ImageTransform
  proc = ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, WB, ImageTransform.none)
  return [proc.process_image_set({'a': torch.from_numpy(f).to(dev)})['a'].cpu().numpy() for f in frames], proc


@pytest.mark.gpu
@pytest.mark.parametrize('fused', [True, False])
@pytest.mark.parametrize('world,h,w,debayer,tone', [(2, 512, 640, 'rcd', 'adaptive_aces'), (4, 1024, 328, 'rcd', 'reinhard'),
                                                     (3, 648, 200, 'ppg', 'linear'), (8, 2048, 256, 'bilinear', 'aces')])
def test_bands_on_one_gpu_match_untiled(world, h, w, debayer, tone, fused):
  """fused: the band runs through the fused frame kernels (band statistics + all-reduce); otherwise stage by stage (CudaOps)."""
  from torch_darktable.pipeline.tiled import CudaOps
  dev = torch.device('cuda:0')
  frames = [synth.packed_frame(h, w, seed=50 + i) for i in range(2)]
  settings = make_settings(debayer, tone)
  want, ref = untiled_cuda(h, w, frames, settings, dev)

  def make_proc(col):
    return TiledFrameProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, WB, col,
                               ops=None if fused else CudaOps())

  align = 32 if fused else 8
  rows = [[torch.from_numpy(r.copy()).to(dev) for r in split_rows(f, w, h, world, align)] for f in frames]
  got, proc0 = run_threads(world, make_proc, rows)
  for g, e in zip(got, want):
    assert_srgb_close(g, e, frac=5e-4)
  np.testing.assert_allclose(proc0.bounds.cpu().numpy(), ref.bounds.cpu().numpy(), rtol=0, atol=0)
  np.testing.assert_allclose(proc0.metrics.cpu().numpy(), ref.metrics.cpu().numpy(), rtol=2e-5)


def _nccl_worker(rank, world, port, h, w, out_dir):
  import sys
  from pathlib import Path
  root = Path(__file__).resolve().parents[1]
  for p in (root, root / 'tests', root / 'torch-darktable_b200'):
    sys.path.insert(0, str(p))
  import torch.distributed as dist
  os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
  torch.cuda.set_device(rank)
  dev = torch.device('cuda', rank)
  dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
  try:
    frame = synth.packed_frame(h, w, seed=93)
    proc = TiledFrameProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, make_settings('rcd', 'adaptive_aces', 1.0), dev, WB,
                               DistCollective())
    own = torch.from_numpy(split_rows(frame, w, h, world, 32)[rank].copy()).to(dev)
    np.save(os.path.join(out_dir, f'band{rank}.npy'), proc.process(own).cpu().numpy())
  finally:
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_ranks_match_untiled(tmp_path):
  world = min(torch.cuda.device_count(), 4)
  if world < 2:
    pytest.skip('needs two or more GPUs')
  import torch.multiprocessing as mp
  h, w = 256 * world, 512
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
  mp.spawn(_nccl_worker, args=(world, port, h, w, str(tmp_path)), nprocs=world, join=True)
  got = np.concatenate([np.load(tmp_path / f'band{r}.npy') for r in range(world)])
  want, _ = untiled_cuda(h, w, [synth.packed_frame(h, w, seed=93)], make_settings('rcd', 'adaptive_aces', 1.0), torch.device('cuda:0'))
  assert_srgb_close(got, want[0], frac=5e-4)
