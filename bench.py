"""Headline benchmark: MP/s of the full RAW -> sRGB pipeline on batch-32 4K frames (BASELINE.json configs[2]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU; frames are sharded, no collective on the data path)

One "step" = one pass of the hot path over the rank's batch of 32 synthetic 12-bit packed 3840x2160 frames
(RCD demosaic + post-process + Wiener log-luminance denoise + bilateral local contrast + adaptive ACES, rotate_270:
the `artichoke` camera settings).  Prints ONE JSON line on rank 0.

  value      frames already resident in HBM, CUDA-event time, max over ranks; the step is ONE call of ImageProcessor.process_batch, i.e. one
             CUDA-graph replay of its 9 x 32 kernels (--no-graph: 32 per-frame process() calls, 1.6 % slower at 4K)
  e2e        the same through pinned HOST buffers (H2D of every packed frame and D2H of every uint8 result inside the
             timed region), via torch_darktable.pipeline.batch.HostFrameRunner; `copy_ceiling` = the same bytes copied with
             no kernel running, all ranks at once, measured in the same run; `frac_of_ceiling` = e2e / that ceiling
  roofline   the dominant kernel, timed per launch with CUDA events on its stream (libtdb200's timing hook), against the
             roofline that BINDS it: `bound` = hbm (algorithmic bytes vs the measured copy bandwidth of MEASURED_PEAKS.json),
             fp32 / mufu (FLOPs / MUFU ops counted by ncu vs the pipe peaks measured live by csrc/probe.cu) or issue (warp
             instructions vs 4 per clock per SM); `stages` lists every kernel of the frame the same way
  parity     one output frame of the TIMED loop compared with the CPU oracle in the same run (N = 1)
  cpu_baseline  the CPU oracle (oracle/, OpenMP on all host cores) on a bounded sample (one frame), rank 0, N = 1 only
  extra      BASELINE.json configs[4], bounded: `batch_20mp` = a 1024-frame 5472x3648 batch sharded over the N ranks (strong scaling,
             frames cycled from four resident scenes); `tiled_200mp` = one 16384x12288 frame split into N row bands with the
             halo rows exchanged GPU to GPU (pipeline/tiled.py)

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
NUM_SMS = 148


def measured_peak_gbs() -> tuple[float, str]:
  p = ROOT / 'MEASURED_PEAKS.json'
  if p.exists():
    return float(json.loads(p.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
  return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_kernel_table() -> tuple[dict, str]:
  """Per-launch ncu counters of the frame kernels (DRAM bytes, FLOPs, XU instructions, pipe utilisations) from the committed
  `ncu --set full` capture: profiles/ncu_frame_kernels.json, written by tools/ncu_frame_kernels.py.  ncu cannot run inside a timed
  benchmark, so these are the only numbers of the line that do not come from this very run; `source` says which capture."""
  try:
    doc = json.loads((ROOT / 'profiles' / 'ncu_frame_kernels.json').read_text())
    return doc['kernels'], doc.get('source', 'profiles/ncu_frame_kernels.json')
  except (OSError, ValueError, KeyError):
    return {}, 'absent'


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


def workload_config() -> dict:
  """The `config` object of the JSON line: ONE definition for both arms (the driver compares them)."""
  return {
    'workload': f'full pipeline RAW->sRGB, batch {FRAMES} x {WIDTH}x{HEIGHT} 12-bit packed RGGB per GPU (BASELINE.json configs[2])',
    'settings': 'artichoke: RCD + postprocess(3 smoothing, global green-eq) + Wiener log-lum 0.075 (K=32, overlap 4) + '
                'bilateral 0.4 @ sigma 2/0.2 + adaptive ACES gamma 1.5, rotate_270; one image set per frame',
    'frames': 'tests/synth.py scenes, seeds 1234..1237 cycled by global frame index',
    'l2': f'inputs {FRAMES * WIDTH * HEIGHT * 3 // 2 / 1e6:.0f} MB per step > 126 MB L2, no explicit flush',
    'parallelism': 'one rank per GPU, frames sharded over the ranks, no collective on the data path'}


def cpu_oracle_frame():
  """The CPU oracle on ONE frame (bounded sample), all host cores through OpenMP: (cpu_baseline dict, its uint8 output)."""
  sys.path.insert(0, str(ROOT))
  sys.path.insert(0, str(ROOT / 'tests'))
  import oracle
  import synth
  frame = synth.packed_frame(HEIGHT, WIDTH, seed=1234)
  pipe = oracle.Pipeline(WIDTH, HEIGHT, debayer='rcd', tone_mapping='adaptive_aces', moving_average=1.0, transform='rotate_270')
  t0 = time.perf_counter()
  out = pipe.process_image_set([frame])[0]
  dt = time.perf_counter() - t0
  return {'value': round(WIDTH * HEIGHT / 1e6 / dt, 3), 'unit': 'MP/s', 'cores': os.cpu_count(), 'kind': 'port',
          'sample': f'1 frame {WIDTH}x{HEIGHT} of the same pipeline, CPU oracle (C + OpenMP), {dt:.2f} s wall'}, out


# ---- pipe peaks, measured live ---------------------------------------------------------------------------------------------------
def measure_pipe_peaks(torch, _lib, dev) -> dict:
  """FP32 FMA and MUFU throughput of this GPU right now (csrc/probe.cu), CUDA events, best of 3."""
  import ctypes as C
  sink = torch.zeros(4, device=dev)
  stream = torch.cuda.current_stream(dev)
  out = {}
  for name, fn, iters in (('fp32_tflops', _lib.lib.tdb_probe_fp32, 4096), ('mufu_tops', _lib.lib.tdb_probe_mufu, 1024)):
    work = C.c_double(0.0)
    best = None
    for _ in range(4):
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record(stream)
      _lib.check(fn(C.c_void_p(sink.data_ptr()), iters, C.byref(work), C.c_void_p(stream.cuda_stream)))
      b.record(stream)
      b.synchronize()
      ms = a.elapsed_time(b)
      best = ms if best is None else min(best, ms)
    out[name] = round(work.value / (best / 1e3) / 1e12, 2)
  return out


def stage_table(table: dict, peak_gbs: float, pipe: dict, sm_mhz: float | None, ncu: dict) -> list[dict]:
  """One row per kernel of the frame: CUDA-event time of THIS run, algorithmic HBM fraction, and -- from the committed ncu capture
  -- which pipe binds the kernel and the fraction of THAT roofline it reaches."""
  total = sum(v[1] for k, v in table.items() if k != '<begin>')
  issue_peak = NUM_SMS * 4 * (sm_mhz or 1965.0) * 1e6  # warp instructions per second
  rows = []
  for name, (count, total_ms) in sorted(table.items(), key=lambda kv: -kv[1][1]):
    if name == '<begin>':
      continue
    ms = total_ms / count
    bpp = ALG_BYTES.get(name)
    gbs = (bpp * WIDTH * HEIGHT / 1e9) / (ms / 1e3) if bpp else None
    row = {'kernel': name, 'launches_per_step': count, 'ms_per_launch': round(ms, 4), 'share': round(total_ms / total, 4),
           'alg_bytes_per_px': bpp, 'achieved_gbs': round(gbs, 1) if gbs else None, 'frac': round(gbs / peak_gbs, 4) if gbs else None}
    k = ncu.get(name)
    if k:
      # utilisation of each candidate roofline as ncu saw it; the largest one names the bound
      pcts = {'hbm': k.get('dram_pct'), 'fp32': k.get('fma_pipe_pct'), 'mufu': k.get('xu_pipe_pct'), 'issue': k.get('issue_pct')}
      pipes = {b: p for b, p in pcts.items() if p is not None and b != 'issue'}
      bound = max(pipes, key=pipes.get) if pipes else 'issue'
      if pcts['issue'] is not None and pipes and pcts['issue'] > max(pipes.values()) and max(pipes.values()) < 40.0:
        bound = 'issue'  # no pipe is even 40 % busy: the kernel is limited by what the schedulers manage to issue (latency, dependencies)
      row['bound'], row['bound_pct_ncu'] = bound, pcts.get(bound)
      row['ncu_pcts'] = pcts
      sec = ms / 1e3
      if bound == 'hbm' and gbs:
        row['frac_of_bound'] = round(gbs / peak_gbs, 4)
      elif bound == 'fp32' and k.get('flops') and pipe.get('fp32_tflops'):
        row['achieved_tflops'] = round(k['flops'] / sec / 1e12, 2)
        row['frac_of_bound'] = round(k['flops'] / sec / 1e12 / pipe['fp32_tflops'], 4)
      elif bound == 'mufu' and k.get('xu_thread_ops') and pipe.get('mufu_tops'):
        row['achieved_mufu_tops'] = round(k['xu_thread_ops'] / sec / 1e12, 3)
        row['frac_of_bound'] = round(k['xu_thread_ops'] / sec / 1e12 / pipe['mufu_tops'], 4)
      elif k.get('warp_insts'):
        row['frac_of_bound'] = round(k['warp_insts'] / sec / issue_peak, 4)
      row['dram_bytes_per_launch'] = k.get('dram_bytes')
      if k.get('flops'):
        row['flops_per_launch'] = k['flops']
    rows.append(row)
  return rows


def roofline_of(top: dict, peak_gbs: float, peak_src: str, pipe: dict, ncu_src: str) -> dict:
  bound = top.get('bound', 'hbm')
  r = {'kernel': top['kernel'], 'bound': bound, 'ms_per_launch': top['ms_per_launch'], 'traffic': top.get('dram_bytes_per_launch'),
       'hbm_achieved_gbs': top['achieved_gbs'], 'hbm_frac': top['frac'], 'hbm_peak_gbs': peak_gbs, 'peak_source': peak_src,
       'ncu_source': ncu_src}
  if bound == 'fp32' and 'achieved_tflops' in top:
    r.update(achieved=top['achieved_tflops'], peak=pipe['fp32_tflops'], unit='TFLOP/s', frac=top['frac_of_bound'],
             peak_source='FP32 FMA pipe measured in this run (csrc/probe.cu); FLOPs per launch from the ncu capture')
  elif bound == 'mufu' and 'achieved_mufu_tops' in top:
    r.update(achieved=top['achieved_mufu_tops'], peak=pipe['mufu_tops'], unit='Tops/s', frac=top['frac_of_bound'],
             peak_source='MUFU unit measured in this run (csrc/probe.cu); XU instructions per launch from the ncu capture')
  elif bound == 'issue' and 'frac_of_bound' in top:
    r.update(achieved=top['frac_of_bound'], peak=1.0, unit='fraction of issue slots (4 warp instructions / clock / SM)', frac=top['frac_of_bound'])
  else:
    r.update(bound='hbm', achieved=top['achieved_gbs'], peak=peak_gbs, unit='GB/s', frac=top['frac'])
  return r


# ---- configs[4], bounded ---------------------------------------------------------------------------------------------------------
def device_packed_scene(torch, td, h, w, seed, dev, y0=0, y1=None, full_h=None):
  """Rows [y0, y1) of a packed synthetic scene generated on the device from GLOBAL pixel coordinates (so any row split sees the
  same image): gradient + gratings + coordinate-hash noise, RGGB mosaic, 12-bit packed."""
  y1 = h if y1 is None else y1
  full_h = h if full_h is None else full_h
  ys = torch.arange(y0, y1, device=dev, dtype=torch.float32).unsqueeze(1)
  xs = torch.arange(w, device=dev, dtype=torch.float32).unsqueeze(0)
  noise = torch.frac(torch.sin(xs * 12.9898 + ys * 78.233 + float(seed)) * 43758.5453)
  cfa = 0.15 + 0.5 * (xs / w * 0.6 + ys / full_h * 0.4) + 0.12 * torch.sin((xs + 0.5 * ys) * (6.2831853 / 37.0) + 0.1 * seed) \
      + 0.1 * torch.sin((ys - 0.3 * xs) * (6.2831853 / 211.0)) + 0.03 * noise
  return td.encode(cfa.clamp_(0.02, 1.0).reshape(-1))


def leg_batch_20mp(torch, dist, td, dev, rank, world, total_frames=1024):
  """BASELINE.json configs[4], first half: a 1024-frame batch of 5472x3648 frames sharded over the ranks (strong scaling: every
  rank processes 1024 / N frames, cycled from four resident scenes; one image set per frame; resident inputs)."""
  from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
  from torch_darktable.pipeline.config import Debayer, ToneMapper
  w, h = 5472, 3648
  settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, **settings_kwargs())
  proc = ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, (1.8, 1.0, 2.1), ImageTransform.rotate_270)
  scenes = [device_packed_scene(torch, td, h, w, 20 + g, dev) for g in range(4)]
  mine = range(rank, total_frames, world)  # global frame indices of this rank

  def run(indices):  # two frames in flight (ImageProcessor.submit), EMA chained frame to frame as by process()
    for i in indices:
      proc.submit(scenes[i % 4], 'cam', track=False)
    proc.join()

  run(list(mine)[:8])
  ms = timed_steps(torch, dist, lambda: run(mine), 1, 0, world)
  del proc, scenes
  torch.cuda.empty_cache()
  return {'workload': f'{total_frames} x {w}x{h} 12-bit packed frames sharded over {world} GPU(s), same pipeline and settings, white balance '
                      '(1.8, 1, 2.1), inputs resident in HBM', 'scaling': 'strong', 'frames_per_rank': len(mine), 'ms_total': round(ms, 2),
          'value': round(total_frames * w * h / 1e6 / (ms / 1e3), 1), 'unit': 'MP/s', 'ms_per_frame_per_gpu': round(ms / len(mine), 4)}


def leg_tiled_200mp(torch, dist, td, dev, rank, world, steps=3):
  """BASELINE.json configs[4], second half: ONE 16384x12288 frame split into N row bands, halo rows exchanged GPU to GPU over NCCL,
  three tiny all-reduces (pipeline/tiled.py).  sigma_s = 8: at sigma_s = 2 the full frame's bilateral grid saturates in y."""
  from torch_darktable.pipeline import ImageProcessingSettings
  from torch_darktable.pipeline.config import Debayer, ToneMapper
  from torch_darktable.pipeline.tiled import DistCollective, ThreadCollective, TiledFrameProcessor
  w, h = 16384, 12288
  kw = settings_kwargs()
  kw.update(bil_sigma_spatial=8.0, bil_sigma_luminance=0.1, moving_average=0.5)
  settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, **kw)
  col = DistCollective() if world > 1 else ThreadCollective(ThreadCollective.Hub(1), 0)
  proc = TiledFrameProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, (1.8, 1.0, 2.1), col)
  y0, y1 = proc.owned_rows
  own = proc.own_rows_buffer()  # the rank's rows live in the padded band buffer: no band-sized copy per frame
  own.copy_(device_packed_scene(torch, td, h, w, 7, dev, y0, y1, h))
  out = [None]

  def step():
    out[0] = proc.process(own)

  ms = timed_steps(torch, dist, step, steps, 2, world)
  checksum = torch.tensor([int(out[0].to(torch.int64).sum().item())], device=dev)
  if world > 1:
    dist.all_reduce(checksum)
  res = {'workload': f'one {w}x{h} ({w * h / 1e6:.0f} MP) 12-bit packed frame in {world} row band(s), RCD + postprocess + Wiener + bilateral '
                     '(sigma 8/0.1) + adaptive ACES', 'scaling': 'strong', 'halo_rows': proc.halo if world > 1 else 0,
         'halo_bytes_per_neighbour': proc.halo * w * 3 // 2 if world > 1 else 0, 'ms_per_frame': round(ms, 3),
         'value': round(w * h / 1e6 / (ms / 1e3), 1), 'unit': 'MP/s', 'checksum_u8_sum': int(checksum.item()),
         'collectives': 'packed halo rows by NCCL send/recv; one all-gather + one all-reduce of 6 floats' if world > 1 else 'none'}
  del proc, own, out
  torch.cuda.empty_cache()
  return res


def run_ours(args):
  sys.path.insert(0, str(ROOT / 'torch-darktable_b200'))
  import torch
  import torch.distributed as dist

  import torch_darktable as td
  from torch_darktable import _lib
  from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
  from torch_darktable.pipeline.batch import HostFrameRunner, bind_host_to_gpu
  from torch_darktable.pipeline.config import Debayer, ToneMapper

  rank, local_rank, world = dist_env()
  torch.cuda.set_device(local_rank)
  dev = torch.device(f'cuda:{local_rank}')
  # before any pinned allocation: this rank's host thread and its pinned pages go to the GPU's own NUMA node
  binding = bind_host_to_gpu(local_rank) if not args.no_bind else {'bound': False}
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)

  frames_np = make_frames(FRAMES, rank)
  host = [torch.from_numpy(f).pin_memory() for f in frames_np]
  resident = [h.to(dev) for h in host]
  settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, **settings_kwargs())
  proc = ImageProcessor((WIDTH, HEIGHT), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, None, ImageTransform.rotate_270)
  px_step = WIDTH * HEIGHT * FRAMES
  last = [None]
  use_graph = not args.no_graph
  batch_in = [torch.stack(resident)]  # (FRAMES, bytes); replaced by the graph's own input buffer after the capturing call

  def step_frames():  # frame by frame through ImageProcessor.process (the per-kernel timing pass, and the step itself with --no-graph)
    for i, f in enumerate(resident):
      out = proc.process(f, 'cam')
      if i == 0:
        last[0] = out

  def step_resident():
    # one call of the batch entry: every frame its own image set, exactly FRAMES calls of ImageProcessor.process; after the first call
    # (which runs eagerly and captures) the whole step is ONE CUDA-graph replay of the 9 x FRAMES kernels
    if not use_graph:
      return step_frames()
    last[0] = proc.process_batch(batch_in[0], 'cam', graph=True)[0]

  launches0 = _lib.launch_count()
  step_resident()  # with the graph: eager pass + capture pass, each issuing the step's launches once
  per_step_launches = (_lib.launch_count() - launches0) // (2 if use_graph else 1)
  if use_graph:
    batch_in[0] = proc.batch_input_buffer(FRAMES)  # already holds the frames: later steps skip the device-to-device copy
  with ClockSampler(local_rank) as clocks:
    ms = timed_steps(torch, dist, step_resident, args.steps, args.warmup, world)
  launches = per_step_launches * args.steps  # kernels of the timed region (graph nodes are launches too; the library counter only sees eager ones)
  timed_out0 = last[0].clone()
  clock_summary = clocks.summary()

  # per-kernel CUDA-event timing of one more step (same stream, back-to-back launches)
  stream = torch.cuda.current_stream(dev)
  torch.cuda.synchronize()
  _lib.timing_begin(stream.cuda_stream)
  step_frames()
  table = _lib.timing_end()
  peak, peak_src = measured_peak_gbs()
  pipe = measure_pipe_peaks(torch, _lib, dev)
  ncu, ncu_src = ncu_kernel_table()
  stages = stage_table(table, peak, pipe, clock_summary.get('sm_mhz'), ncu)
  roofline = roofline_of(stages[0], peak, peak_src, pipe, ncu_src)

  # end to end through host buffers
  out_shape = (WIDTH, HEIGHT, 3)  # rotate_270 swaps the axes
  host_out = [torch.empty(out_shape, dtype=torch.uint8).pin_memory() for _ in range(FRAMES)]
  runner = HostFrameRunner(proc)

  def step_e2e():  # batches are streamed: the copy-in of the next step overlaps the tail of this one, one host wait at the end
    runner.run(host, host_out, after_caller=False)

  allocs0 = torch.cuda.memory_stats(dev).get('num_device_alloc', 0)
  with ClockSampler(local_rank) as clocks_e2e:  # the end-to-end timed region is sampled too
    e2e_ms = timed_steps(torch, dist, step_e2e, args.steps, max(args.warmup, 1), world)
    runner.wait()
  e2e_clocks = clocks_e2e.summary()
  if e2e_clocks['samples']:
    clock_summary['e2e_leg'] = {k: e2e_clocks[k] for k in ('sm_mhz', 'reasons', 'samples')}
  e2e_device_allocs = torch.cuda.memory_stats(dev).get('num_device_alloc', 0) - allocs0  # cudaMalloc calls (they synchronise): should be 0 after warm-up

  # the same bytes with no kernel running, both directions at once, all ranks at once: the ceiling of the e2e figure on this box
  s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
  dev_out = [torch.empty(out_shape, dtype=torch.uint8, device=dev) for _ in range(2)]

  def step_copies():
    cur = torch.cuda.current_stream(dev)
    s_in.wait_stream(cur), s_out.wait_stream(cur)
    for i in range(FRAMES):
      with torch.cuda.stream(s_in):
        resident[i].copy_(host[i], non_blocking=True)
      with torch.cuda.stream(s_out):
        host_out[i].copy_(dev_out[i & 1], non_blocking=True)
    cur.wait_stream(s_in), cur.wait_stream(s_out)

  copy_ms = timed_steps(torch, dist, step_copies, 3, 1, world)
  del dev_out

  extra = {}
  if not args.no_extra:
    del runner, host_out, host, resident, proc
    torch.cuda.empty_cache()
    for name, leg in (('batch_20mp', leg_batch_20mp), ('tiled_200mp', leg_tiled_200mp)):
      try:
        extra[name] = leg(torch, dist, td, dev, rank, world)
      except Exception as e:  # noqa: BLE001 - the headline line must survive a failing extra leg
        extra[name] = {'error': repr(e)[:300]}

  if rank == 0:
    e2e_value = px_step * world / 1e6 / (e2e_ms / 1e3)
    ceiling = px_step * world / 1e6 / (copy_ms / 1e3)
    line = {
      'metric': METRIC, 'value': round(px_step * world / 1e6 / (ms / 1e3), 1), 'unit': 'MP/s', 'n_gpus': world, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': round(ms, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(),
      'e2e': {'value': round(e2e_value, 1), 'unit': 'MP/s', 'ms_per_step': round(e2e_ms, 3),
              'h2d_bytes_per_step': FRAMES * WIDTH * HEIGHT * 3 // 2, 'd2h_bytes_per_step': FRAMES * WIDTH * HEIGHT * 3,
              'device_allocs_incl_warmup': int(e2e_device_allocs), 'copy_ceiling': round(ceiling, 1), 'copy_ceiling_ms_per_step': round(copy_ms, 3), 'frac_of_ceiling': round(e2e_value / ceiling, 4),
              'copy_ceiling_gbs_per_gpu': round(FRAMES * WIDTH * HEIGHT * 4.5 / 1e6 / copy_ms, 1)},
      'e2e_method': 'HostFrameRunner: copy-in stream / two compute lanes (ImageProcessor.submit: two frames in flight) / copy-out stream over '
                    'three device slots, pinned host buffers, steps streamed into each other, one host wait inside the closing synchronise; '
                    'host thread + pinned pages bound to the GPU\'s NUMA node where the platform exposes one',
      'host_binding': binding,
      'resident_method': ('ImageProcessor.process_batch: one CUDA-graph replay per step' if use_graph else 'ImageProcessor.process per frame'),
      'gpu_launches': int(launches), 'roofline': roofline, 'pipe_peaks_measured': pipe, 'stages': stages, 'clocks': clock_summary,
    }
    if world == 1 and not args.no_cpu_baseline:
      import numpy as np
      base, want = cpu_oracle_frame()
      line['cpu_baseline'] = base
      d = np.abs(timed_out0.cpu().numpy().astype(np.int16) - want.astype(np.int16))
      line['parity'] = {'what': 'frame 0 of the last TIMED step vs the CPU oracle on the same packed bytes', 'max_lsb': int(d.max()),
                        'frac_different': float((d > 0).mean()), 'frac_beyond_1_lsb': float((d > 1).mean()),
                        'ok': bool((d > 0).mean() <= 1e-3 and (d > 1).mean() <= 2e-4),
                        'tolerance': 'tests/cases.py ORACLE_TOLERANCE[pipeline]: <= 1e-3 different, <= 2e-4 beyond 1 LSB (RCD direction '
                                     'flips of IEEE CPU arithmetic); against the reference itself the 20 MP frame test holds <= 1 LSB',
                        'checksum_u8_sum': int(timed_out0.to(torch.int64).sum().item())}
    if extra:
      line['extra'] = extra
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
    base, _ = cpu_oracle_frame()
    print(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'MP/s', 'n_gpus': 1, 'steps': 1, 'warmup': 0,
                      'higher_is_better': True, 'data': 'synthetic', 'dtype': 'f32', 'cpu_baseline': base, 'config': workload_config(),
                      'what': 'bounded sample: ' + base['sample'], 'why': f'reference extension unavailable: {e!r}'[:200],
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
    'config': workload_config(),
    'what': 'unmodified reference CUDA extension (baseline/_ref, sm_100a build) through its own ImageProcessor.process on ONE GPU '
            '(rank 0; the reference has neither a CPU implementation of this path nor a multi-GPU mode)',
    'cpu_baseline': {'value': value, 'unit': 'MP/s', 'cores': 0, 'kind': 'reference',
                     'sample': 'full workload on the GPU: the reference ops reject CPU tensors (SURVEY.md 8c)'},
    'e2e': {'value': round(px_step / 1e6 / (e2e_ms / 1e3), 1), 'unit': 'MP/s', 'ms_per_step': round(e2e_ms, 3),
            'h2d_bytes_per_step': FRAMES * WIDTH * HEIGHT * 3 // 2, 'd2h_bytes_per_step': FRAMES * WIDTH * HEIGHT * 3},
    'e2e_method': 'the caller\'s loop of the reference: per frame a non-blocking copy from a pinned host buffer, ImageProcessor.process (which '
                  'synchronises the device several times per frame), a non-blocking copy of the result to the host; one synchronise per step',
    'clocks': clocks.summary()}), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--frames', type=int, default=FRAMES, help='frames per GPU per step (32 = the headline config; smaller only for profiling)')
  ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU oracle leg (profiling runs)')
  ap.add_argument('--no-extra', action='store_true', help='skip the configs[4] legs (profiling runs)')
  ap.add_argument('--no-graph', action='store_true', help='resident leg through per-frame process() calls instead of the CUDA-graph batch entry')
  ap.add_argument('--no-bind', action='store_true', help='do not bind the host thread to the GPU\'s NUMA node (A/B of the e2e leg)')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3)
  globals()['FRAMES'] = args.frames
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
