"""How long does the HOST need to enqueue one frame of the pipeline (Python + ctypes + allocator), against the GPU time?
GPU box only:  python tools/host_overhead.py"""
from pathlib import Path
import sys
import time

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'torch-darktable_b200'))
sys.path.insert(0, str(ROOT / 'tests'))
import torch

import synth
import torch_darktable as td
from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
from torch_darktable.pipeline.config import Debayer, ToneMapper

W, H = 3840, 2160
dev = torch.device('cuda:0')
frame = torch.from_numpy(synth.packed_frame(H, W, seed=1234)).to(dev)
settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, enable_denoise=True, enable_bilateral=True,
                                   postprocess=True, tone_gamma=1.5, tone_intensity=2.0, light_adapt=0.8, vibrance=0.5, moving_average=1.0,
                                   bilateral=0.4, bil_sigma_spatial=2.0, bil_sigma_luminance=0.2, denoise=0.075, color_smoothing_passes=3)
proc = ImageProcessor((W, H), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, None, ImageTransform.rotate_270)
for _ in range(5):
  proc.process(frame, 'cam')
torch.cuda.synchronize()
n = 64
t0 = time.perf_counter()
for _ in range(n):
  proc.process(frame, 'cam')
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host enqueue {1e3 * (t1 - t0) / n:.3f} ms/frame, total {1e3 * (t2 - t0) / n:.3f} ms/frame')
