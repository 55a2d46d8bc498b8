"""Three-way comparison on mid-size frames: reference outputs (tools/ref_dump.py) vs CPU oracle vs CUDA path.
  python tools/three_way.py /tmp/ref_mid.npz"""
from pathlib import Path
import sys

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / 'torch-darktable_b200', ROOT / 'tests'):
  sys.path.insert(0, str(p))
import cases  # noqa: E402
import synth  # noqa: E402
from cuda_impl import CudaImpl  # noqa: E402

ref = np.load(sys.argv[1])
cu, orc = CudaImpl(), cases.OracleImpl()


def diff(a, b):
  d = np.abs(a.astype(np.float64) - b.astype(np.float64))
  i = np.unravel_index(d.argmax(), d.shape)
  return f'{d.max():.3e}@{tuple(int(v) for v in i)} n>{1e-5:g}:{int((d > 1e-5).sum())}'


for key in ref.files:
  parts = key.split('/')
  h, w = map(int, parts[-1].split('x'))
  if parts[0] in ('rcd', 'ppg', 'bilinear'):
    op = {'rcd': 'rcd', 'ppg': 'ppg', 'bilinear': 'bilinear5x5_demosaic'}[parts[0]]
    params = {'pattern': parts[1], 'median_threshold': 0.0}
    ins = {'cfa': synth.mosaic(synth.scene_rgb(h, w, 7), parts[1])}
  elif parts[0] == 'pp':
    rng = np.random.default_rng(3)
    rgb = (synth.scene_rgb(h, w, 11) + rng.normal(0, 0.02, size=(h, w, 3))).astype(np.float32)
    rgb[0::2, 1::2, 1] *= 1.04
    op = 'postprocess'
    params = {'pattern': 'RGGB', 'color_smoothing_passes': int(parts[1][0]), 'green_eq_global': bool(int(parts[1][1])),
              'green_eq_local': bool(int(parts[1][2])), 'green_eq_threshold': 4.0}
    ins = {'rgb': rgb}
  elif parts[0] == 'wll':
    x = np.clip(synth.scene_rgb(h, w, 17) + np.random.default_rng(6).normal(0, 0.02, size=(h, w, 3)), 0, 1).astype(np.float32)
    op, params, ins = 'wiener_log_luminance', {'noise': 0.075, 'eps': 1e-4}, {'x': x}
  else:
    op, params, ins = 'bilateral_rgb', {'sigma_s': float(parts[1]), 'sigma_r': 0.2 if parts[1] == '2.0' else 0.1, 'detail': 0.4}, \
        {'x': synth.scene_rgb(h, w, 19)}
  r = ref[key]
  g = cases.run_case(cu, op, params, ins)['out'].reshape(r.shape)
  o = cases.run_case(orc, op, params, ins)['out'].reshape(r.shape)
  print(f'{key:28s} cuda-ref {diff(g, r):44s} oracle-ref {diff(o, r):44s} cuda-oracle {diff(g, o)}', flush=True)
