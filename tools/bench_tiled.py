"""One oversize frame split into row bands across the GPUs of a box (BASELINE.json configs[4], second half).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_tiled.py \
      [--width 16384 --height 12288] [--steps 3] [--warmup 2]
  python tools/bench_tiled.py            # one GPU, no split: the baseline the split is compared with

Every rank owns a band of the packed frame, exchanges the halo rows with its neighbours over NCCL (NVLink P2P) and all-reduces the
three tiny statistics; timing is CUDA events around `steps` whole frames, barrier + synchronise on both sides, max over ranks.
Rank 0 prints one JSON line.  The frame content is generated on the device, band by band, from global pixel coordinates, so every
split processes the same image.
"""

from __future__ import annotations

import argparse
import json
import os
from pathlib import Path
import sys

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'torch-darktable_b200'))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--width', type=int, default=16384)
  ap.add_argument('--height', type=int, default=12288)
  ap.add_argument('--steps', type=int, default=3)
  ap.add_argument('--warmup', type=int, default=2)
  ap.add_argument('--sigma-s', type=float, default=8.0)
  ap.add_argument('--breakdown', action='store_true', help='one more frame with a CUDA event after every phase: per-phase ms, max over the ranks')
  args = ap.parse_args()
  rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))

  import torch
  import torch.distributed as dist

  import torch_darktable as td
  from torch_darktable.pipeline import ImageProcessingSettings
  from torch_darktable.pipeline.config import Debayer, ToneMapper
  from torch_darktable.pipeline.tiled import DistCollective, ThreadCollective, TiledFrameProcessor, partition_rows

  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
    col = DistCollective()
  else:
    col = ThreadCollective(ThreadCollective.Hub(1), 0)

  w, h = args.width, args.height
  settings = ImageProcessingSettings(enable_denoise=True, enable_bilateral=True, postprocess=True, tone_gamma=1.5, tone_intensity=2.0,
                                     light_adapt=0.8, tone_mapping=ToneMapper.adaptive_aces, vibrance=0.5, debayer=Debayer.rcd,
                                     moving_average=0.5, bil_sigma_spatial=args.sigma_s)
  proc = TiledFrameProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, (1.8, 1.0, 2.1), col)
  y0, y1 = proc.owned_rows

  # this rank's rows of a smooth synthetic CFA (gradient + gratings + noise), packed on the device
  ys = torch.arange(y0, y1, device=dev, dtype=torch.float32).unsqueeze(1)
  xs = torch.arange(w, device=dev, dtype=torch.float32).unsqueeze(0)
  noise = torch.frac(torch.sin(xs * 12.9898 + ys * 78.233) * 43758.5453)  # coordinate hash: the same in every split
  cfa = 0.15 + 0.5 * (xs / w * 0.6 + ys / h * 0.4) + 0.12 * torch.sin((xs + 0.5 * ys) * (6.2831853 / 37.0)) \
      + 0.1 * torch.sin((ys - 0.3 * xs) * (6.2831853 / 211.0)) + 0.03 * noise
  del noise
  own = proc.own_rows_buffer()  # the rank's rows go straight into the padded band buffer: no band-sized copy per frame
  own.copy_(td.encode(cfa.clamp_(0.02, 1.0).reshape(-1)))
  del cfa, ys, xs
  torch.cuda.empty_cache()

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  out = None
  for _ in range(args.warmup):
    out = proc.process(own)
  barrier()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(args.steps):
    out = proc.process(own)
  b.record()
  barrier()
  ms = a.elapsed_time(b) / args.steps
  if world > 1:
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
  phases = None
  if args.breakdown:
    proc.profile = True
    barrier()
    proc.process(own)
    local = proc.breakdown()
    proc.profile = False
    names = list(local)
    t = torch.tensor([local[k] for k in names], device=dev)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    phases = {k: round(float(v), 3) for k, v in zip(names, t.tolist())}
  checksum = int(out.to(torch.int64).sum().item())
  if world > 1:
    t = torch.tensor([checksum], device=dev)
    dist.all_reduce(t)
    checksum = int(t.item())
  if rank == 0:
    halo = proc.halo
    print(json.dumps({
      'metric': 'MP/s RAW->sRGB, one frame split into row bands', 'value': round(w * h / 1e6 / (ms / 1e3), 1), 'unit': 'MP/s',
      'n_gpus': world, 'ms_per_frame': round(ms, 3), 'steps': args.steps, 'warmup': args.warmup, 'scaling': 'strong',
      'config': {'workload': f'one {w}x{h} ({w * h / 1e6:.0f} MP) 12-bit packed RGGB frame, RCD + postprocess + Wiener + bilateral '
                             f'(sigma_s {args.sigma_s:g}) + adaptive ACES', 'bands': world, 'halo_rows': halo if world > 1 else 0,
                 'halo_bytes_per_neighbour': halo * w * 3 // 2 if world > 1 else 0,
                 'collectives': 'packed halo rows by NCCL send/recv; one all-gather + one all-reduce of 6 floats' if world > 1 else 'none'},
      'phases_ms_max_over_ranks': phases, 'checksum_u8_sum': checksum, 'peak_mem_gb': round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
