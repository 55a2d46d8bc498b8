"""A few RCD-from-packed launches at 4K (the target of `ncu -k regex:rcd`).  python tools/run_rcd_once.py [f32]"""
from pathlib import Path
import sys

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'torch-darktable_b200'))
import torch  # noqa: E402

import torch_darktable as td  # noqa: E402

dev = torch.device('cuda:0')
w, h = 3840, 2160
gen = torch.Generator(device=dev).manual_seed(1)
if len(sys.argv) > 1 and sys.argv[1] == 'f32':
  rcd = td.RCD(dev, (w, h), td.BayerPattern.RGGB)
  x = torch.rand((h, w, 1), device=dev, generator=gen)
  for _ in range(4):
    rcd.process(x)
else:
  p = torch.randint(0, 256, (w * h * 3 // 2,), dtype=torch.uint8, device=dev, generator=gen)
  for _ in range(4):
    td.demosaic_packed(p, (w, h), td.BayerPattern.RGGB, method='rcd')
torch.cuda.synchronize()
print('done')
