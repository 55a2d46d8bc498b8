"""Time the unmodified reference extension (baseline/_ref) per op and end to end on a B200.

Method = the reference's own scripts/run_benchmark.py:16-39 (warm-up, CUDA events around back-to-back calls).
Writes one JSON object per line to stdout; used to fill BASELINE.md / DESIGN.md, not by bench.py.
"""

from __future__ import annotations

import json
from pathlib import Path
import sys

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'baseline' / '_ref'))
sys.path.insert(0, str(ROOT / 'tests'))

import torch  # noqa: E402

import synth  # noqa: E402
import torch_darktable as td  # noqa: E402
from torch_darktable.pipeline.config import Debayer, ImageProcessingSettings, ToneMapper  # noqa: E402
from torch_darktable.pipeline.image_processor import ImageProcessor  # noqa: E402

dev = torch.device('cuda:0')


def timed(fn, iters=10, warmup=3):
  for _ in range(warmup):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / iters


def report(name, ms, pixels, **kw):
  print(json.dumps({'op': name, 'ms': round(ms, 4), 'MP/s': round(pixels / ms / 1e3, 1), **kw}), flush=True)


def main():
  h, w = 4000, 6000
  packed = torch.from_numpy(synth.packed_frame(h, w, seed=1)).to(dev)
  px = h * w
  report('decode12_float 24MP', timed(lambda: td.decode12(packed)), px)
  cfa = td.decode12(packed).view(h, w, 1)
  pat = td.BayerPattern.RGGB
  ppg, rcd = td.PPG(dev, (w, h), pat), td.RCD(dev, (w, h), pat)
  report('bilinear 24MP', timed(lambda: td.bilinear5x5_demosaic(cfa, pat)), px)
  report('ppg 24MP', timed(lambda: ppg.process(cfa)), px)
  report('rcd 24MP', timed(lambda: rcd.process(cfa)), px)
  rgb = rcd.process(cfa).clone()
  pp = td.PostProcess(dev, (w, h), pat, color_smoothing_passes=3, green_eq_global=True)
  report('postprocess 24MP', timed(lambda: pp.process(rgb)), px)
  rgb = rgb.clamp(0, 1)
  wn = td.Wiener(dev, (w, h))
  report('wiener_log_luminance 24MP', timed(lambda: wn.process_log_luminance(rgb, 0.075), iters=5), px)
  bil = td.Bilateral(dev, (w, h), sigma_s=2.0, sigma_r=0.2)
  report('bilateral_rgb 24MP', timed(lambda: bil.process_rgb(rgb, 0.4)), px)
  lum = td.compute_luminance(rgb)
  lap = td.Laplacian(dev, (w, h), td.LaplacianParams())
  report('laplacian 24MP', timed(lambda: lap.process(lum), iters=5), px)
  metrics = td.compute_image_metrics([rgb])
  params = td.TonemapParameters(1.5, 2.0, 0.8, 0.5)
  report('adaptive_aces 24MP', timed(lambda: td.aces_tonemap(rgb, params, metrics)), px)
  del cfa, rgb, lum, ppg, rcd, pp, wn, bil, lap
  torch.cuda.empty_cache()

  # full pipeline, 4K frames, artichoke-style settings
  h, w = 2160, 3840
  n = 8
  host = [torch.from_numpy(synth.packed_frame(h, w, seed=100 + i % 2)).pin_memory() for i in range(n)]
  for denoise in (True, False):
    settings = ImageProcessingSettings(enable_denoise=denoise, enable_bilateral=True, postprocess=True, tone_gamma=1.5,
                                       tone_intensity=2.0, light_adapt=0.8, tone_mapping=ToneMapper.adaptive_aces,
                                       vibrance=0.5, debayer=Debayer.rcd, moving_average=1.0)
    proc = ImageProcessor((w, h), pat, td.PackedFormat.Packed12, settings, dev, None)
    frames = [f.to(dev) for f in host]

    def run_resident():
      for i, f in enumerate(frames):
        proc.process(f, 'cam')

    def run_e2e():
      for f in host:
        out = proc.process(f.to(dev, non_blocking=True), 'cam')
        out.cpu()

    report(f'pipeline 4K x{n} resident denoise={denoise}', timed(run_resident, iters=3, warmup=2), px_total := h * w * n)
    report(f'pipeline 4K x{n} e2e denoise={denoise}', timed(run_e2e, iters=3, warmup=1), px_total)


if __name__ == '__main__':
  main()
