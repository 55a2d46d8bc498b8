"""RCD from packed bytes at 4K / 24 MP / 20 MP: CUDA-event time per frame (A/B of the strip kernel: TDB_RCD_STRIPS=0|1, TDB_RCD_SEGMENTS=n,
TDB_RCD_FRAME_FIRST=0|1 are read once per process).  python tools/bench_rcd.py [tag]"""
import json
import os
from pathlib import Path
import sys

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'torch-darktable_b200'))
import torch  # noqa: E402

import torch_darktable as td  # noqa: E402

dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(1)
out = {'tag': sys.argv[1] if len(sys.argv) > 1 else '', 'env': {k: v for k, v in os.environ.items() if k.startswith('TDB_')}}
for name, (w, h) in (('4k', (3840, 2160)), ('24mp', (6000, 4000)), ('20mp', (5472, 3648))):
  frames = [torch.randint(0, 256, (w * h * 3 // 2,), dtype=torch.uint8, device=dev, generator=gen) for _ in range(8)]
  cfas = [torch.rand((h, w, 1), device=dev, generator=gen) for _ in range(4)]
  rcd = td.RCD(dev, (w, h), td.BayerPattern.RGGB)
  for what, fn, ins in (('packed', lambda p: td.demosaic_packed(p, (w, h), td.BayerPattern.RGGB, method='rcd'), frames),
                        ('f32', lambda c: rcd.process(c), cfas)):
    for i in range(4):
      fn(ins[i % len(ins)])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(24):
      fn(ins[i % len(ins)])
    b.record()
    torch.cuda.synchronize()
    out[f'{name}_{what}_ms'] = round(a.elapsed_time(b) / 24, 4)
  del frames, cfas
  torch.cuda.empty_cache()
print(json.dumps(out), flush=True)
