"""Upper bound of overlapping consecutive frames: two independent ImageProcessors on two streams against one processor on one stream
(same 32 resident 4K frames).  python tools/exp_two_streams.py"""
from pathlib import Path
import sys

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'torch-darktable_b200'))
sys.path.insert(0, str(ROOT / 'tests'))
import torch  # noqa: E402

import synth  # noqa: E402
import torch_darktable as td  # noqa: E402
from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform  # noqa: E402
from torch_darktable.pipeline.config import Debayer, ToneMapper  # noqa: E402

dev = torch.device('cuda:0')
w, h, n = 3840, 2160, 32
frames = [torch.from_numpy(synth.packed_frame(h, w, seed=1234 + g)).to(dev) for g in range(4)]
frames = [frames[i % 4] for i in range(n)]
settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, enable_denoise=True, enable_bilateral=True,
                                   postprocess=True, tone_gamma=1.5, tone_intensity=2.0, light_adapt=0.8, vibrance=0.5, moving_average=1.0)
mk = lambda: ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, None, ImageTransform.rotate_270)  # noqa: E731
procs = [mk(), mk(), mk()]
streams = [torch.cuda.Stream(dev) for _ in range(3)]


def run(k):
  cur = torch.cuda.current_stream(dev)
  for s in streams[:k]:
    s.wait_stream(cur)
  for i, f in enumerate(frames):
    with torch.cuda.stream(streams[i % k]):
      procs[i % k].process(f, 'cam')
  for s in streams[:k]:
    cur.wait_stream(s)


for k in (1, 2, 3):
  for _ in range(3):
    run(k)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(5):
    run(k)
  b.record()
  torch.cuda.synchronize()
  ms = a.elapsed_time(b) / 5
  print(f'{k} stream(s): {ms:.3f} ms per 32 frames = {w * h * n / 1e6 / (ms / 1e3):.1f} MP/s', flush=True)
