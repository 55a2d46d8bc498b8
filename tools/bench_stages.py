"""Per-stage throughput of the public torch_darktable API at the BASELINE.json shapes (configs 1, 2 and 4), in ms, MP/s, GB/s of
ALGORITHMIC bytes (SURVEY.md 8d contract) and as a fraction of the measured HBM copy bandwidth.

  python tools/bench_stages.py [--impl ours|reference] [--iters 10] [--quick] > profiles/...jsonl

The same script times either package: `ours` imports torch-darktable_b200/torch_darktable, `reference` imports the unmodified reference
extension from baseline/_ref -- the stage calls below are the drop-in API both expose.  Fused entry points that only exist here
(demosaic_packed, Wiener/Bilateral composites) are skipped for the reference; for those the reference column is the sum of the stage
calls it needs instead.  Timing: CUDA events on the current stream around `iters` calls after 3 warm-up calls; inputs are rotated
over enough distinct buffers (> 256 MB in total) that no call finds its input in the 126 MB L2.
"""

from __future__ import annotations

import argparse
import json
from pathlib import Path
import sys

ROOT = Path(__file__).resolve().parents[1]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--iters', type=int, default=10)
  ap.add_argument('--kernels', action='store_true', help='also print the per-kernel times of one call of every stage (ours only)')
  ap.add_argument('--quick', action='store_true', help='small shapes (smoke test of the script itself)')
  ap.add_argument('--configs', default='2,4,1', help='comma-separated BASELINE.json config numbers to run (2 = 24 MP stages, 4 = 50 MP local contrast, 1 = 12 MP chain)')
  args = ap.parse_args()
  sys.path.insert(0, str(ROOT / ('torch-darktable_b200' if args.impl == 'ours' else 'baseline/_ref')))
  import torch

  import torch_darktable as td
  ours = args.impl == 'ours'
  peak = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text()).get('hbm_gbs', 6541.1) if (ROOT / 'MEASURED_PEAKS.json').exists() else 6541.1
  dev = torch.device('cuda:0')
  torch.cuda.set_device(dev)
  gen = torch.Generator(device=dev).manual_seed(1234)

  def rotate(make, nbytes):
    n = max(2, min(8, int(256e6 // max(nbytes, 1)) + 1))
    return [make() for _ in range(n)]

  def timeit(fn, inputs):
    for i in range(3):
      fn(inputs[i % len(inputs)])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.iters):
      fn(inputs[i % len(inputs)])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / args.iters

  def report(config, op, w, h, bpp, ms, note=''):
    mp = w * h / 1e6
    gbs = bpp * w * h / 1e9 / (ms / 1e3) if bpp else None
    print(json.dumps({'impl': args.impl, 'config': config, 'op': op, 'width': w, 'height': h, 'ms': round(ms, 4),
                      'mp_per_s': round(mp / (ms / 1e3), 1), 'alg_bytes_per_px': bpp, 'achieved_gbs': round(gbs, 1) if gbs else None,
                      'frac_of_measured_hbm': round(gbs / peak, 4) if gbs else None, 'note': note}), flush=True)

  def kernel_table(fn, inputs):
    """per-kernel CUDA-event times of ONE call (libtdb200's timing hook; ours only)"""
    from torch_darktable import _lib
    torch.cuda.synchronize()
    _lib.timing_begin(torch.cuda.current_stream(dev).cuda_stream)
    fn(inputs[-1])
    table = _lib.timing_end()
    return {k: [n, round(ms, 4)] for k, (n, ms) in table.items() if k != '<begin>'}

  def run(config, op, w, h, bpp, fn, inputs, note=''):
    try:
      report(config, op, w, h, bpp, timeit(fn, inputs), note)
      if ours and args.kernels:
        print(json.dumps({'impl': args.impl, 'config': config, 'kernels_of': op, 'launches_and_ms': kernel_table(fn, inputs)}), flush=True)
    except Exception as e:  # noqa: BLE001 - one missing op must not hide the rest of the table
      print(json.dumps({'impl': args.impl, 'config': config, 'op': op, 'error': repr(e)[:300]}), flush=True)
    torch.cuda.empty_cache()

  pat = td.BayerPattern.RGGB
  wb = torch.tensor([1.8, 1.0, 2.1], device=dev)
  wanted = set(args.configs.split(','))
  if '2' in wanted:
    config2(args, td, torch, dev, gen, ours, rotate, run, pat, wb)
  if '4' in wanted:
    config4(args, td, torch, dev, gen, ours, rotate, run)
  if '1' in wanted:
    config1(args, td, torch, dev, gen, ours, rotate, run, pat, wb)


def config2(args, td, torch, dev, gen, ours, rotate, run, pat, wb):

  # ---- config 2 (24 MP packed -> unpack + demosaic) and the pointwise stages at the same shape -------------------------------
  w, h = (1536, 1024) if args.quick else (6000, 4000)
  c = 'config2 24MP'
  packed = rotate(lambda: torch.randint(0, 256, (w * h * 3 // 2,), dtype=torch.uint8, device=dev, generator=gen), w * h * 1.5)
  run(c, 'decode12 -> f32', w, h, 5.5, lambda p: td.decode12(p, output_dtype=torch.float32), packed)
  run(c, 'decode12 -> f16', w, h, 3.5, lambda p: td.decode12(p, output_dtype=torch.float16), packed)
  run(c, 'decode12 -> u16', w, h, 3.5, lambda p: td.decode12(p, output_dtype=torch.uint16), packed)
  cfa = rotate(lambda: torch.rand((h, w), device=dev, generator=gen), w * h * 4)
  run(c, 'encode12 <- f32', w, h, 5.5, lambda x: td.encode(x.reshape(-1)), cfa)
  run(c, 'apply_white_balance', w, h, 8.0, lambda x: td.apply_white_balance(x, wb, pat), cfa)
  cfa1 = [x.unsqueeze(-1) for x in cfa]
  run(c, 'bilinear5x5 (f32 CFA)', w, h, 16.0, lambda x: td.bilinear5x5_demosaic(x, pat), cfa1)
  ppg = td.PPG(dev, (w, h), pat)
  run(c, 'PPG (f32 CFA)', w, h, 16.0, lambda x: ppg.process(x), cfa1)
  rcd = td.RCD(dev, (w, h), pat)
  run(c, 'RCD (f32 CFA)', w, h, 16.0, lambda x: rcd.process(x), cfa1)
  if ours:
    for m in ('bilinear', 'ppg', 'rcd'):
      run(c, f'{m} from packed (unpack + WB fused)', w, h, 13.5,
          lambda p, m=m: td.demosaic_packed(p, (w, h), pat, method=m, white_balance=wb), packed)
  else:  # what the reference needs for the same result: decode + white balance + demosaic
    for m, f in (('bilinear', lambda x: td.bilinear5x5_demosaic(x, pat)), ('ppg', ppg.process), ('rcd', rcd.process)):
      run(c, f'{m} from packed (decode12 + apply_white_balance + demosaic)', w, h, 13.5,
          lambda p, f=f: f(td.apply_white_balance(td.decode12(p, output_dtype=torch.float32).view(h, w), wb, pat).unsqueeze(-1)), packed)
  del ppg, rcd, cfa, cfa1, packed
  torch.cuda.empty_cache()
  rgb = rotate(lambda: torch.rand((h, w, 3), device=dev, generator=gen), w * h * 12)
  post = td.PostProcess(dev, (w, h), pat, color_smoothing_passes=3, green_eq_global=True)
  run(c, 'PostProcess (3 smoothing + global green-eq)', w, h, 24.0, lambda x: post.process(x), rgb)
  del post
  for name, bpp in (('rgb_to_lab', 24.0), ('lab_to_rgb', 24.0), ('rgb_to_xyz', 24.0), ('compute_luminance', 16.0)):
    run(c, name, w, h, bpp, getattr(td, name), rgb)
  lum = rotate(lambda: torch.rand((h, w), device=dev, generator=gen), w * h * 4)
  run(c, 'modify_luminance', w, h, 28.0, lambda x: td.modify_luminance(x, lum[0]), rgb)
  run(c, 'compute_image_bounds (stride 8)', w, h, 12.0 / 64, lambda x: td.compute_image_bounds([x], stride=8), rgb)
  run(c, 'compute_image_metrics (stride 8)', w, h, 12.0 / 64, lambda x: td.compute_image_metrics([x], stride=8), rgb)
  metrics = td.compute_image_metrics([rgb[0]], stride=8)
  params = td.TonemapParameters(1.5, 2.0, 0.8, 0.5)
  run(c, 'reinhard_tonemap', w, h, 15.0, lambda x: td.reinhard_tonemap(x, metrics, params), rgb)
  run(c, 'aces_tonemap', w, h, 15.0, lambda x: td.aces_tonemap(x, params), rgb)
  run(c, 'adaptive aces_tonemap', w, h, 15.0, lambda x: td.aces_tonemap(x, params, metrics), rgb)
  run(c, 'linear_tonemap', w, h, 15.0, lambda x: td.linear_tonemap(x, metrics, params), rgb)
  wiener = td.Wiener(dev, (w, h))
  run(c, 'Wiener log-luminance (K=32, overlap 4)', w, h, 24.0, lambda x: wiener.process_log_luminance(x, 0.075), rgb,
      'FP32-issue bound (16 covering tiles per pixel), not HBM bound')
  sig3 = torch.tensor([0.05, 0.05, 0.05], device=dev)
  run(c, 'Wiener RGB (K=32, overlap 4)', w, h, 24.0, lambda x: wiener.process(x, sig3), rgb, 'FP32-issue bound')
  del wiener, rgb, lum
  torch.cuda.empty_cache()


def config4(args, td, torch, dev, gen, ours, rotate, run):
  # ---- config 4 (50 MP local contrast) ---------------------------------------------------------------------------------------
  w, h = (2048, 1536) if args.quick else (8192, 6144)
  c = 'config4 50MP'
  lum = rotate(lambda: torch.rand((h, w), device=dev, generator=gen), w * h * 4)
  for ss, sr in ((8.0, 0.1), (8.0, 0.2), (4.0, 0.2), (2.0, 0.2)):  # 2 / 0.2 at 8192 px: the reference's saturating grid (3001 x 2251 x 6)
    bil = td.Bilateral(dev, (w, h), sigma_s=ss, sigma_r=sr)
    cells = 1
    for d in bil._bilateral.grid_size() if ours else ():
      cells *= d
    grid_bpp = 4.0 * cells / (w * h)
    run(c, f'Bilateral luminance (sigma_s {ss:g}, sigma_r {sr:g})', w, h, round(12.0 + 2 * grid_bpp, 3) if ours else 12.0,
        lambda x, bil=bil: bil.process(x, 0.4), lum, f'grid {grid_bpp:.2f} B/px each way' if ours else '')
    del bil
  rgb = rotate(lambda: torch.rand((h, w, 3), device=dev, generator=gen), w * h * 12)
  bil = td.Bilateral(dev, (w, h), sigma_s=8.0, sigma_r=0.2)
  run(c, 'Bilateral.process_rgb (sigma_s 8)', w, h, 36.0, lambda x: bil.process_rgb(x, 0.4), rgb,
      'fused: luminance inside splat and slice; the reference runs compute_luminance + process + modify_luminance (116 B/px)')
  del bil, rgb
  torch.cuda.empty_cache()
  lap = td.Laplacian(dev, (w, h), td.LaplacianParams())
  run(c, 'Laplacian (6 gammas)', w, h, 12.0, lambda x: lap.process(x), lum, 'fp16 pyramids of the padded frame')
  del lap, lum
  torch.cuda.empty_cache()


def config1(args, td, torch, dev, gen, ours, rotate, run, pat, wb):
  # ---- config 1 (12 MP, the chain the reference would run through its torch path) -------------------------------------------------
  w, h = (1024, 768) if args.quick else (4096, 3000)
  c = 'config1 12MP'
  cfa1 = rotate(lambda: torch.rand((h, w, 1), device=dev, generator=gen), w * h * 4)
  m = torch.tensor([[1.6, -0.4, -0.2], [-0.3, 1.5, -0.2], [0.0, -0.5, 1.5]], device=dev)

  def chain(x):
    rgb = td.bilinear5x5_demosaic(td.apply_white_balance(x[..., 0], wb, pat).unsqueeze(-1), pat)
    return td.color_transform_3x3(rgb, m) if ours else td.rgb_to_lab(rgb)  # the reference's 3x3 faults

  run(c, 'white_balance + bilinear5x5 + colour op', w, h, 8.0 + 16.0 + 24.0, chain, cfa1)


if __name__ == '__main__':
  main()
