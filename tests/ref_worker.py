"""Runs the UNMODIFIED reference extension (baseline/_ref) on files handed over by a test, in a process of its own.

The reference package and the product package share the name `torch_darktable`, so they cannot live in one interpreter.  The live
parity tests (tests/test_gpu_reference_live.py) write their inputs as .npy files into a scratch directory together with a
jobs.json, start this script there, and read the outputs back:

  python tests/ref_worker.py <workdir>

jobs.json = [{"name": str, "op": str, "params": {...}, "inputs": {key: file}}, ...]; every output tensor of job `name` is written to
<workdir>/<name>.<key>.npy.  GPU box only (the reference has no CPU path); nothing here imports the product package or the oracle.
"""

from __future__ import annotations

import json
from pathlib import Path
import sys

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'baseline' / '_ref'))

import torch  # noqa: E402

import torch_darktable as td  # noqa: E402  (the reference)
from torch_darktable.pipeline.config import Debayer, ImageProcessingSettings, ToneMapper  # noqa: E402
from torch_darktable.pipeline.image_processor import ImageProcessor  # noqa: E402
from torch_darktable.pipeline.transform import ImageTransform  # noqa: E402

assert 'baseline/_ref' in td.__file__, td.__file__
dev = torch.device('cuda:0')


def cuda(a: np.ndarray) -> torch.Tensor:
  return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def demosaic_packed(p, i):
  """decode12 -> [apply_white_balance] -> demosaic on a FRESH workspace: what demosaic_packed of the product fuses."""
  w, h = p['width'], p['height']
  pat = td.BayerPattern[p['pattern']]
  cfa = td.decode12(cuda(i['packed']), output_dtype=torch.float32).view(h, w)
  if p.get('gains') is not None:
    cfa = td.apply_white_balance(cfa, torch.tensor(p['gains'], dtype=torch.float32, device=dev), pat)
  x = cfa.unsqueeze(-1)
  if p['method'] == 'bilinear':
    return {'out': td.bilinear5x5_demosaic(x, pat)}
  if p['method'] == 'ppg':
    return {'out': td.PPG(dev, (w, h), pat, median_threshold=float(p.get('median_threshold', 0.0))).process(x)}
  return {'out': td.RCD(dev, (w, h), pat).process(x).clone()}


def postprocess(p, i):
  x = cuda(i['rgb'])
  h, w, _ = x.shape
  pp = td.PostProcess(dev, (w, h), td.BayerPattern[p['pattern']], color_smoothing_passes=p['passes'], green_eq_local=p['local'],
                      green_eq_global=p['global'], green_eq_threshold=p['threshold'])
  return {'out': pp.process(x).clone()}


def bilateral(p, i):
  x = cuda(i['lum'])
  h, w = x.shape
  return {'out': td.Bilateral(dev, (w, h), sigma_s=p['sigma_s'], sigma_r=p['sigma_r']).process(x, float(p['detail']))}


def bilateral_rgb(p, i):
  x = cuda(i['rgb'])
  h, w, _ = x.shape
  return {'out': td.Bilateral(dev, (w, h), sigma_s=p['sigma_s'], sigma_r=p['sigma_r']).process_rgb(x, float(p['detail']))}


def laplacian(p, i):
  x = cuda(i['lum'])
  h, w = x.shape
  params = td.LaplacianParams(sigma=p['sigma'], shadows=p['shadows'], highlights=p['highlights'], clarity=p['clarity'])
  return {'out': td.Laplacian(dev, (w, h), params).process(x)}


def wiener_log_luminance(p, i):
  x = cuda(i['rgb'])
  h, w, _ = x.shape
  return {'out': td.Wiener(dev, (w, h)).process_log_luminance(x, float(p['noise']), float(p.get('eps', 1e-4)))}


def pipeline(p, i):
  settings = ImageProcessingSettings(debayer=Debayer[p['debayer']], tone_mapping=ToneMapper[p['tone_mapping']], **p['settings'])
  wb = tuple(p['white_balance']) if p.get('white_balance') is not None else None
  proc = ImageProcessor((p['width'], p['height']), td.BayerPattern[p['pattern']], td.PackedFormat.Packed12, settings, dev, wb,
                        ImageTransform[p['transform']])
  out = {}
  for s, names in enumerate(p['sets']):  # each set: list of input keys
    res = proc.process_image_set({n: cuda(i[n]) for n in names})
    for n in names:
      out[f'set{s}_{n}'] = res[n].clone()
    out[f'bounds{s}'], out[f'metrics{s}'] = proc.bounds.clone(), proc.metrics.clone()
  return out


OPS = {'demosaic_packed': demosaic_packed, 'postprocess': postprocess, 'bilateral': bilateral, 'bilateral_rgb': bilateral_rgb,
       'laplacian': laplacian, 'wiener_log_luminance': wiener_log_luminance, 'pipeline': pipeline}


def main():
  work = Path(sys.argv[1])
  for job in json.loads((work / 'jobs.json').read_text()):
    ins = {k: np.load(work / f) for k, f in job['inputs'].items()}
    outs = OPS[job['op']](job['params'], ins)
    torch.cuda.synchronize()
    for k, t in outs.items():
      np.save(work / f"{job['name']}.{k}.npy", t.detach().cpu().numpy())
    del ins, outs
    torch.cuda.empty_cache()
  print('ref_worker: done')


if __name__ == '__main__':
  main()
