"""Pinned host<->device copy bandwidth with ALL ranks of a box copying at once: the ceiling of bench.py's `e2e` leg at N GPUs
(1.5 B/px in, 3 B/px out per frame; every rank moves one bench step: 398 MB in, 796 MB out).

  python tools/pcie_probe.py [--bind]                                                            # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py [--bind]

--bind: pin the rank's host thread (and therefore its pinned allocations, first touch) to the CPUs NVML reports as local to its
GPU (torch_darktable.pipeline.batch.bind_host_to_gpu) before anything is allocated.  Rank 0 prints one JSON line: per-rank and
aggregate GB/s for H2D alone, D2H alone and both directions at once (barrier + synchronise on both sides, max over ranks), the
NUMA node of every GPU and the CPU set of every rank, and the e2e ceiling in MP/s that follows from the bidirectional figure.
"""
import argparse
import json
import os
from pathlib import Path
import sys

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'torch-darktable_b200'))

ap = argparse.ArgumentParser()
ap.add_argument('--bind', action='store_true')
ap.add_argument('--reps', type=int, default=5)
args = ap.parse_args()
rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from torch_darktable.pipeline.batch import bind_host_to_gpu, gpu_locality  # noqa: E402

torch.cuda.set_device(local)
dev = torch.device('cuda', local)
binding = bind_host_to_gpu(local) if args.bind else {'bound': False, **gpu_locality(local)}
if world > 1:
  dist.init_process_group('nccl', device_id=dev)

n_in, n_out = 32 * 3840 * 2160 * 3 // 2, 32 * 3840 * 2160 * 3
h_in, h_out = torch.empty(n_in, dtype=torch.uint8).pin_memory(), torch.empty(n_out, dtype=torch.uint8).pin_memory()
h_in.fill_(1), h_out.fill_(1)  # touch every page from this (possibly bound) thread
d_in, d_out = torch.empty(n_in, dtype=torch.uint8, device=dev), torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def barrier():
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()


def timed(fn):
  fn()
  barrier()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(args.reps):
    fn()
  for s in (s1, s2):
    torch.cuda.current_stream(dev).wait_stream(s)
  b.record()
  barrier()
  ms = a.elapsed_time(b) / args.reps
  t = torch.tensor([ms], device=dev)
  if world > 1:
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    return [float(g.item()) for g in gathered]
  return [ms]


def h2d():
  with torch.cuda.stream(s1):
    s1.wait_stream(torch.cuda.current_stream(dev))
    d_in.copy_(h_in, non_blocking=True)


def d2h():
  with torch.cuda.stream(s2):
    s2.wait_stream(torch.cuda.current_stream(dev))
    h_out.copy_(d_out, non_blocking=True)


def both():
  h2d()
  d2h()


t_in, t_out, t_both = timed(h2d), timed(d2h), timed(both)
info = [None] * world
if world > 1:
  dist.all_gather_object(info, binding)
else:
  info = [binding]
if rank == 0:
  gbs = lambda n, ts: [round(n / t / 1e6, 1) for t in ts]  # noqa: E731
  worst = max(t_both)
  print(json.dumps({
    'n_gpus': world, 'bind': args.bind, 'host_cpus': os.cpu_count(), 'cpus_allowed': len(os.sched_getaffinity(0)),
    'h2d_gbs_per_rank': gbs(n_in, t_in), 'd2h_gbs_per_rank': gbs(n_out, t_out),
    'both_ms_per_rank': [round(t, 2) for t in t_both],
    'aggregate_both_gbs': round(world * (n_in + n_out) / worst / 1e6, 1),
    'aggregate_h2d_gbs': round(world * n_in / max(t_in) / 1e6, 1), 'aggregate_d2h_gbs': round(world * n_out / max(t_out) / 1e6, 1),
    'e2e_ceiling_mp_per_s': round(world * 32 * 3840 * 2160 / 1e6 / (worst / 1e3), 1),
    'ranks': info,
    'note': 'every rank copies one bench step (398 MB in, 796 MB out) from / to pinned host memory, all ranks at once; '
            'e2e_ceiling = pixels of those steps / slowest rank with both directions in flight'}), flush=True)
if world > 1:
  dist.destroy_process_group()
