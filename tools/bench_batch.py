"""Frames per second of ImageProcessor on small to large frames: per-frame `process` calls, `process_batch` without and with CUDA-graph
replay.  python tools/bench_batch.py > profiles/rNN_batch_graph.jsonl"""
import json
from pathlib import Path
import sys

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'torch-darktable_b200'))
import torch  # noqa: E402

import torch_darktable as td  # noqa: E402
from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform  # noqa: E402
from torch_darktable.pipeline.config import Debayer, ToneMapper  # noqa: E402

dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(3)
for name, (w, h), n in (('256x192', (256, 192), 64), ('1280x720', (1280, 720), 32), ('1920x1080', (1920, 1080), 32), ('3840x2160', (3840, 2160), 16)):
  settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, enable_denoise=True, enable_bilateral=True,
                                     postprocess=True, tone_gamma=1.5, tone_intensity=2.0, light_adapt=0.8, vibrance=0.5, moving_average=0.5)
  frames = torch.randint(0, 256, (n, w * h * 3 // 2), dtype=torch.uint8, device=dev, generator=gen)
  row = {'frame': name, 'batch': n}
  for mode in ('process', 'batch', 'batch_graph'):
    proc = ImageProcessor((w, h), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, (1.8, 1.0, 2.1), ImageTransform.rotate_270)
    if mode == 'process':
      fn = lambda: [proc.process(f, 'cam') for f in frames]  # noqa: E731
    else:
      fn = lambda: proc.process_batch(frames, 'cam', graph=mode == 'batch_graph')  # noqa: E731
    for _ in range(3):
      fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
      fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps / n
    row[f'{mode}_ms_per_frame'] = round(ms, 4)
    row[f'{mode}_mp_per_s'] = round(w * h / 1e6 / (ms / 1e3), 1)
    del proc
  print(json.dumps(row), flush=True)
