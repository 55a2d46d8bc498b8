"""Run the UNMODIFIED reference extension (baseline/_ref) on seeded mid-size inputs and save its outputs
(three-way arbitration reference / oracle / CUDA path, see tools/three_way.py).  GPU box only.
  python tools/ref_dump.py /tmp/ref_mid.npz"""
from pathlib import Path
import sys

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'baseline' / '_ref'))
sys.path.insert(0, str(ROOT / 'tests'))
import torch  # noqa: E402

import synth  # noqa: E402
import torch_darktable as td  # noqa: E402

assert 'baseline/_ref' in td.__file__
dev = torch.device('cuda:0')
out = {}
for (h, w) in [(250, 372), (516, 1100)]:
  for pattern in ['RGGB', 'GBRG']:
    cfa = synth.mosaic(synth.scene_rgb(h, w, 7), pattern)
    t = torch.from_numpy(cfa).to(dev).unsqueeze(-1)
    pat = td.BayerPattern[pattern]
    out[f'rcd/{pattern}/{h}x{w}'] = td.RCD(dev, (w, h), pat).process(t).clone().cpu().numpy()
    out[f'ppg/{pattern}/{h}x{w}'] = td.PPG(dev, (w, h), pat, median_threshold=0.0).process(t).clone().cpu().numpy()
    out[f'bilinear/{pattern}/{h}x{w}'] = td.bilinear5x5_demosaic(t, pat).cpu().numpy()
  rng = np.random.default_rng(3)
  rgb = (synth.scene_rgb(h, w, 11) + rng.normal(0, 0.02, size=(h, w, 3))).astype(np.float32)
  rgb[0::2, 1::2, 1] *= 1.04
  for passes, glob, loc in [(1, False, False), (3, True, False), (4, True, True), (5, False, False)]:
    pp = td.PostProcess(dev, (w, h), td.BayerPattern.RGGB, color_smoothing_passes=passes, green_eq_local=loc, green_eq_global=glob,
                        green_eq_threshold=4.0)
    out[f'pp/{passes}{int(glob)}{int(loc)}/{h}x{w}'] = pp.process(torch.from_numpy(rgb).to(dev)).clone().cpu().numpy()
  x = np.clip(synth.scene_rgb(h, w, 17) + np.random.default_rng(6).normal(0, 0.02, size=(h, w, 3)), 0, 1).astype(np.float32)
  out[f'wll/{h}x{w}'] = td.Wiener(dev, (w, h)).process_log_luminance(torch.from_numpy(x).to(dev), 0.075, 1e-4).cpu().numpy()
  x = synth.scene_rgb(h, w, 19)
  for ss, sr in [(2.0, 0.2), (8.0, 0.1)]:
    out[f'bil/{ss}/{h}x{w}'] = td.Bilateral(dev, (w, h), sigma_s=ss, sigma_r=sr).process_rgb(torch.from_numpy(x).to(dev), 0.4).cpu().numpy()
torch.cuda.synchronize()
np.savez(sys.argv[1], **out)
print('saved', len(out), 'reference outputs')
