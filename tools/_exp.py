import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / 'torch-darktable_b200'), str(ROOT / 'tests'), str(ROOT)]
import torch, synth
import torch_darktable as td
from torch_darktable import _lib
from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
from torch_darktable.pipeline.config import Debayer, ToneMapper
import bench
dev = torch.device('cuda:0')
frame = torch.from_numpy(synth.packed_frame(2160, 3840, seed=1234)).to(dev)
settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, **bench.settings_kwargs())
for tf in (ImageTransform.rotate_270, ImageTransform.none, ImageTransform.flip_horiz, ImageTransform.transpose):
  proc = ImageProcessor((3840, 2160), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, None, tf)
  for _ in range(4): proc.process(frame, 'cam')
  torch.cuda.synchronize()
  _lib.timing_begin(torch.cuda.current_stream(dev).cuda_stream)
  for _ in range(8): proc.process(frame, 'cam')
  t = _lib.timing_end()
  print(tf.name, {k: round(v[1] / v[0], 4) for k, v in t.items() if 'tonemap' in k})
