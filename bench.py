"""Headline benchmark: MP/s of the full RAW -> sRGB pipeline on batch-32 4K frames (BASELINE.json configs[2]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU; frames are sharded, no collective on the data path)

One "step" = one pass of the hot path over the rank's batch of 32 synthetic 12-bit packed 3840x2160 frames
(RCD demosaic + post-process + Wiener log-luminance denoise + bilateral local contrast + adaptive ACES, rotate_270:
the `artichoke` camera settings).  Prints ONE JSON line on rank 0.

  value      frames already resident in HBM, CUDA-event time, max over ranks
  e2e        the same through pinned HOST buffers (H2D of every packed frame and D2H of every uint8 result inside the
             timed region), via torch_darktable.pipeline.batch.HostFrameRunner; the steps are streamed (the copy-in of a
             step overlaps the tail of the previous one), the host waits once, inside the timed region's closing sync
  roofline   the dominant kernel, timed per launch with CUDA events on its stream (libtdb200's timing hook), against
             the measured HBM copy bandwidth of MEASURED_PEAKS.json; `stages` lists every kernel the same way
  cpu_baseline  the CPU oracle (oracle/, OpenMP on all host cores) on a bounded sample (one frame), rank 0, N=1 only

--impl reference: the UNMODIFIED reference CUDA extension from baseline/_ref through its own ImageProcessor on the same
workload (the reference has no CPU implementation of any op); falls back to the CPU oracle port on a
bounded sample when baseline/_ref cannot be imported.
"""

from __future__ import annotations

import argparse
import json
import os
from pathlib import Path
import statistics
import subprocess
import sys
import threading
import time

ROOT = Path(__file__).resolve().parent
WIDTH, HEIGHT, FRAMES = 3840, 2160, 32
METRIC = 'MP/s RAW->sRGB pipeline (batch-32 4K frames)'

# algorithmic HBM bytes per PIXEL of every kernel (grid terms use the 4K bilateral grid, 6.0 B/px)
GRID_BPP = 4.0 * 1921 * 1081 * 6 / (WIDTH * HEIGHT)
ALG_BYTES = {
  'rcd_demosaic': 13.5, 'color_smoothing': 24.0, 'green_eq_ratio': 0.0, 'green_equilibration': 24.0,
  'bounds_init': 0.0, 'compute_image_bounds': 12.0 / 64, 'lerp': 0.0, 'normalize': 24.0,
  'wiener_log_luminance': 16.0, 'wiener_zero_accumulator': 4.0, 'wiener_tiles': 8.0, 'wiener_tiles_border': 0.0, 'rcd_demosaic_frame': 0.0, 'wiener_normalize': 28.0,
  'bilateral_zero_grid': GRID_BPP, 'bilateral_splat': 12.0 + GRID_BPP, 'bilateral_blur': 2 * GRID_BPP,
  'bilateral_slice': 24.0, 'metrics_init': 0.0, 'compute_image_metrics': 12.0 / 64, 'metrics_finalize': 0.0, 'tonemap': 15.0,
  # fused frame pipeline
  'frame_prepare': 12.0 + 8.0 + 4.0 + 4.0, 'wiener_normalize_lum': 24.0 + 4.0, 'bilateral_grid_build': 4.0 + GRID_BPP,
  'metrics_sliced': 12.0 / 64, 'frame_stats': 0.0,
  'bilateral_slice_tonemap': 15.0,
}


def measured_peak_gbs() -> tuple[float, str]:
  p = ROOT / 'MEASURED_PEAKS.json'
  if p.exists():
    return float(json.loads(p.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
  return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
  """nvidia-smi clocks / throttle reasons while the timed region runs."""

  QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, index: int):
    self.samples, self.reasons, self.max_mhz = [], set(), None
    self._stop = threading.Event()
    self._thread = threading.Thread(target=self._run, args=(index,), daemon=True)

  def _run(self, index):
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    while not self._stop.is_set():
      try:
        out = subprocess.run(['nvidia-smi', f'--id={index}', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits'],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(',')
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for n, v in zip(names, out[2:]):
          if v.strip().lower().startswith('active'):
            self.reasons.add(n)
      except Exception:
        pass
      self._stop.wait(0.1)

  def __enter__(self):
    self._thread.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    self._thread.join(timeout=10)

  def summary(self):
    return {'sm_mhz': statistics.median(self.samples) if self.samples else None, 'sm_max_mhz': self.max_mhz,
            'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def dist_env():
  return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def make_frames(n: int, rank: int):
  """n packed 4K frames (numpy uint8).  Frame i of the global batch is seeded by its global index, so sharding never
  changes content; four distinct scenes are cycled to keep host-side generation short."""
  sys.path.insert(0, str(ROOT / 'tests'))
  import synth
  scenes = {}
  frames = []
  for i in range(n):
    g = (rank * n + i) % 4
    if g not in scenes:
      scenes[g] = synth.packed_frame(HEIGHT, WIDTH, seed=1234 + g)
    frames.append(scenes[g])
  return frames


def timed_steps(torch, dist, fn, steps, warmup, world):
  for _ in range(warmup):
    fn()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  start.record()
  for _ in range(steps):
    fn()
  stop.record()
  torch.cuda.synchronize()
  ms = start.elapsed_time(stop)
  if world > 1:
    t = torch.tensor([ms], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    dist.barrier()
  return ms / steps


def settings_kwargs():
  return dict(enable_denoise=True, enable_bilateral=True, postprocess=True, tone_gamma=1.5, tone_intensity=2.0, light_adapt=0.8,
              vibrance=0.5, moving_average=1.0, bilateral=0.4, bil_sigma_spatial=2.0, bil_sigma_luminance=0.2, denoise=0.075,
              color_smoothing_passes=3)


def cpu_oracle_baseline():
  """The CPU oracle on ONE frame (bounded sample), all host cores through OpenMP."""
  sys.path.insert(0, str(ROOT))
  sys.path.insert(0, str(ROOT / 'tests'))
  import oracle
  import synth
  frame = synth.packed_frame(HEIGHT, WIDTH, seed=1234)
  pipe = oracle.Pipeline(WIDTH, HEIGHT, debayer='rcd', tone_mapping='adaptive_aces', moving_average=1.0, transform='rotate_270')
  t0 = time.perf_counter()
  pipe.process_image_set([frame])
  dt = time.perf_counter() - t0
  return {'value': round(WIDTH * HEIGHT / 1e6 / dt, 3), 'unit': 'MP/s', 'cores': os.cpu_count(), 'kind': 'port',
          'sample': f'1 frame {WIDTH}x{HEIGHT} of the same pipeline, CPU oracle (C + OpenMP), {dt:.2f} s wall'}


def run_ours(args):
  sys.path.insert(0, str(ROOT / 'torch-darktable_b200'))
  import torch
  import torch.distributed as dist

  import torch_darktable as td
  from torch_darktable import _lib
  from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
  from torch_darktable.pipeline.batch import HostFrameRunner
  from torch_darktable.pipeline.config import Debayer, ToneMapper

  rank, local_rank, world = dist_env()
  torch.cuda.set_device(local_rank)
  dev = torch.device(f'cuda:{local_rank}')
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)

  frames_np = make_frames(FRAMES, rank)
  host = [torch.from_numpy(f).pin_memory() for f in frames_np]
  resident = [h.to(dev) for h in host]
  settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, **settings_kwargs())
  proc = ImageProcessor((WIDTH, HEIGHT), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, None, ImageTransform.rotate_270)
  px_step = WIDTH * HEIGHT * FRAMES

  def step_resident():
    for f in resident:
      proc.process(f, 'cam')

  with ClockSampler(local_rank) as clocks:
    launches0 = _lib.launch_count()
    ms = timed_steps(torch, dist, step_resident, args.steps, args.warmup, world)
    launches = (_lib.launch_count() - launches0) // (args.steps + args.warmup) * args.steps

  # per-kernel CUDA-event timing of one more step (same stream, back-to-back launches)
  stream = torch.cuda.current_stream(dev)
  torch.cuda.synchronize()
  _lib.timing_begin(stream.cuda_stream)
  step_resident()
  table = _lib.timing_end()
  stages = []
  peak, peak_src = measured_peak_gbs()
  for name, (count, total_ms) in sorted(table.items(), key=lambda kv: -kv[1][1]):
    if name == '<begin>':
      continue
    per_launch_ms = total_ms / count
    bpp = ALG_BYTES.get(name)
    gbs = (bpp * WIDTH * HEIGHT / 1e9) / (per_launch_ms / 1e3) if bpp else None
    stages.append({'kernel': name, 'launches_per_step': count, 'ms_per_launch': round(per_launch_ms, 4),
                   'share': round(total_ms / sum(v[1] for v in table.values()), 4), 'alg_bytes_per_px': bpp,
                   'achieved_gbs': round(gbs, 1) if gbs else None, 'frac': round(gbs / peak, 4) if gbs else None})
  top = stages[0]
  traffic = None  # dram__bytes_read + dram__bytes_write of the dominant kernel per launch, from the committed ncu --set full capture
  try:
    t = json.loads((ROOT / 'profiles' / 'ncu_traffic.json').read_text())['kernels'].get(top['kernel'])
    traffic = (t['dram_read_bytes'] + t['dram_write_bytes']) if t else None
  except (OSError, ValueError, KeyError):
    pass
  roofline = {'kernel': top['kernel'], 'bound': 'hbm', 'achieved': top['achieved_gbs'], 'peak': peak, 'unit': 'GB/s',
              'frac': top['frac'], 'traffic': traffic, 'peak_source': peak_src, 'ms_per_launch': top['ms_per_launch'],
              'note': 'wiener_tiles is bound by the FP32 lanes (register FFTs of 16 covering tiles per pixel, column transforms shared between tiles), not by HBM; see DESIGN.md'}

  # end to end through host buffers
  out_shape = (WIDTH, HEIGHT, 3)  # rotate_270 swaps the axes
  host_out = [torch.empty(out_shape, dtype=torch.uint8).pin_memory() for _ in range(FRAMES)]
  runner = HostFrameRunner(proc)

  def step_e2e():  # batches are streamed: the copy-in of the next step overlaps the tail of this one, one host wait at the end
    runner.run(host, host_out, after_caller=False)

  e2e_ms = timed_steps(torch, dist, step_e2e, args.steps, max(args.warmup, 1), world)
  runner.wait()

  if rank == 0:
    line = {
      'metric': METRIC, 'value': round(px_step * world / 1e6 / (ms / 1e3), 1), 'unit': 'MP/s', 'n_gpus': world, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': round(ms, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': f'full pipeline RAW->sRGB, batch {FRAMES} x {WIDTH}x{HEIGHT} 12-bit packed RGGB per GPU (BASELINE.json configs[2])',
                 'settings': 'artichoke: RCD + postprocess(3 smoothing, global green-eq) + Wiener log-lum 0.075 (K=32, overlap 4) + '
                             'bilateral 0.4 @ sigma 2/0.2 + adaptive ACES gamma 1.5, rotate_270; one image set per frame',
                 'l2': f'inputs {FRAMES * WIDTH * HEIGHT * 3 // 2 / 1e6:.0f} MB per step > 126 MB L2, no explicit flush',
                 'parallelism': f'frames sharded over {world} GPU(s), no collective'},
      'e2e': {'value': round(px_step * world / 1e6 / (e2e_ms / 1e3), 1), 'unit': 'MP/s', 'ms_per_step': round(e2e_ms, 3),
              'h2d_bytes_per_step': FRAMES * WIDTH * HEIGHT * 3 // 2, 'd2h_bytes_per_step': FRAMES * WIDTH * HEIGHT * 3},
      'gpu_launches': int(launches), 'roofline': roofline, 'stages': stages, 'clocks': clocks.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
      line['cpu_baseline'] = cpu_oracle_baseline()
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def run_reference(args):
  rank, local_rank, world = dist_env()
  if rank != 0:
    return
  ref = ROOT / 'baseline' / '_ref'
  try:
    sys.path.insert(0, str(ref))
    import torch

    import torch_darktable as td
    assert str(ref) in td.__file__ and torch.cuda.is_available()
    from torch_darktable.pipeline.config import Debayer, ImageProcessingSettings, ToneMapper
    from torch_darktable.pipeline.image_processor import ImageProcessor
    from torch_darktable.pipeline.transform import ImageTransform
  except Exception as e:  # no GPU / reference not installed: time the CPU oracle port on a bounded sample
    base = cpu_oracle_baseline()
    print(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'MP/s', 'n_gpus': 1, 'steps': 1, 'warmup': 0,
                      'higher_is_better': True, 'data': 'synthetic', 'dtype': 'f32', 'cpu_baseline': base,
                      'config': {'workload': 'bounded sample: ' + base['sample'], 'why': f'reference extension unavailable: {e!r}'[:200]},
                      'e2e': {'value': base['value'], 'unit': 'MP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}), flush=True)
    return

  torch.cuda.set_device(local_rank)
  dev = torch.device(f'cuda:{local_rank}')
  frames_np = make_frames(FRAMES, 0)
  host = [torch.from_numpy(f).pin_memory() for f in frames_np]
  resident = [h.to(dev) for h in host]
  settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, **settings_kwargs())
  proc = ImageProcessor((WIDTH, HEIGHT), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, None, ImageTransform.rotate_270)
  px_step = WIDTH * HEIGHT * FRAMES

  def step_resident():
    for f in resident:
      proc.process(f, 'cam')

  def step_e2e():
    outs = [proc.process(h.to(dev, non_blocking=True), 'cam').to('cpu', non_blocking=True) for h in host]
    torch.cuda.synchronize()
    return outs

  with ClockSampler(local_rank) as clocks:
    ms = timed_steps(torch, None, step_resident, args.steps, args.warmup, 1)
  e2e_ms = timed_steps(torch, None, step_e2e, args.steps, max(args.warmup, 1), 1)
  value = round(px_step / 1e6 / (ms / 1e3), 1)
  print(json.dumps({
    'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'MP/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
    'ms_per_step': round(ms, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': {'workload': f'full pipeline RAW->sRGB, batch {FRAMES} x {WIDTH}x{HEIGHT} 12-bit packed RGGB (BASELINE.json configs[2])',
               'what': 'unmodified reference CUDA extension (baseline/_ref, sm_100a build) through its own ImageProcessor.process; '
                       'the reference has no CPU implementation of this path'},
    'cpu_baseline': {'value': value, 'unit': 'MP/s', 'cores': 0, 'kind': 'reference',
                     'sample': 'full workload on the GPU: the reference ops reject CPU tensors (SURVEY.md 8c)'},
    'e2e': {'value': round(px_step / 1e6 / (e2e_ms / 1e3), 1), 'unit': 'MP/s', 'ms_per_step': round(e2e_ms, 3),
            'h2d_bytes_per_step': FRAMES * WIDTH * HEIGHT * 3 // 2, 'd2h_bytes_per_step': FRAMES * WIDTH * HEIGHT * 3},
    'clocks': clocks.summary()}), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--frames', type=int, default=FRAMES, help='frames per GPU per step (32 = the headline config; smaller only for profiling)')
  ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU oracle leg (profiling runs)')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3)
  globals()['FRAMES'] = args.frames
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
