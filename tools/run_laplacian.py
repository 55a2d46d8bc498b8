"""Laplacian calls at the config-4 shape (for ncu and for timing): python tools/run_laplacian.py [width height]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / 'torch-darktable_b200'))
import torch
import torch_darktable as td
from torch_darktable import _lib
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 6144)
dev = torch.device('cuda:0')
lum = torch.rand((h, w), device=dev)
lap = td.Laplacian(dev, (w, h), td.LaplacianParams())
for _ in range(3):
  out = lap.process(lum)
torch.cuda.synchronize()
print(float(out.mean()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
  out = lap.process(lum)
b.record()
torch.cuda.synchronize()
print('ms per call', round(a.elapsed_time(b) / 5, 4))
_lib.timing_begin(torch.cuda.current_stream(dev).cuda_stream)
lap.process(lum)
print({k: (n, round(ms, 4)) for k, (n, ms) in _lib.timing_end().items()})
