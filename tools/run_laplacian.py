"""One Laplacian call at the config-4 shape (for ncu): python tools/run_laplacian.py [width height]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / 'torch-darktable_b200'))
import torch
import torch_darktable as td
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 6144)
dev = torch.device('cuda:0')
lum = torch.rand((h, w), device=dev)
lap = td.Laplacian(dev, (w, h), td.LaplacianParams())
for _ in range(3):
  out = lap.process(lum)
torch.cuda.synchronize()
print(float(out.mean()))
