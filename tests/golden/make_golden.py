"""Generate golden vectors by running the UNMODIFIED reference CUDA extension (baseline/_ref).

Run on a GPU box:   python tests/golden/make_golden.py gpurun_out/golden
The reference ships no golden vectors or known-answer tests, so these outputs of
the reference itself are what pins the CPU oracle (oracle/) and, through it, the CUDA kernels.
Inputs come from tests/synth.py (numpy, seeded) and are stored next to the outputs.

Layout of every .npz:  "<case>/in/<name>", "<case>/out/<name>" arrays plus "manifest" (JSON string:
case -> {"op": ..., "params": {...}}).
"""

from __future__ import annotations

import json
from pathlib import Path
import sys
import traceback

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / 'baseline' / '_ref'))
sys.path.insert(0, str(ROOT / 'tests'))

import torch  # noqa: E402

import synth  # noqa: E402
import torch_darktable as td  # noqa: E402  (the reference, from baseline/_ref)
from torch_darktable.pipeline.config import Debayer, ImageProcessingSettings, ToneMapper  # noqa: E402
from torch_darktable.pipeline.image_processor import ImageProcessor  # noqa: E402
from torch_darktable.pipeline.transform import ImageTransform  # noqa: E402

assert 'baseline/_ref' in td.__file__, td.__file__
ext = td.extension.extension
dev = torch.device('cuda:0')

H, W = 64, 96  # small frame for the stencil ops
HP, WP = 48, 64  # smaller still for pointwise ops


class Book:
  def __init__(self):
    self.arrays = {}
    self.manifest = {}
    self.errors = []

  def add(self, name, op, params, inputs, outputs):
    self.manifest[name] = {'op': op, 'params': params}
    for k, v in inputs.items():
      self.arrays[f'{name}/in/{k}'] = np.ascontiguousarray(v)
    for k, v in outputs.items():
      if isinstance(v, torch.Tensor):
        v = v.detach().cpu()
        v = v.view(torch.int16).numpy().view(np.uint16) if v.dtype == torch.uint16 else v.numpy()
      self.arrays[f'{name}/out/{k}'] = np.ascontiguousarray(v)

  def run(self, fn):
    try:
      fn(self)
      torch.cuda.synchronize()
    except Exception:  # keep going: one failing case must not lose the rest
      self.errors.append(f'{fn.__name__}: {traceback.format_exc()}')
      print(f'!! {fn.__name__} failed', file=sys.stderr)

  def save(self, path: Path):
    np.savez_compressed(path, manifest=np.array(json.dumps(self.manifest)), **self.arrays)
    print(f'wrote {path} ({path.stat().st_size / 1e6:.2f} MB, {len(self.manifest)} cases)')


def cuda(a):
  t = torch.from_numpy(np.ascontiguousarray(a))
  return t.to(dev)


def u16_tensor(a):
  return torch.from_numpy(a.view(np.int16)).view(torch.uint16).to(dev)


# ---------------------------------------------------------------------------------------------
def packed_cases(b: Book):
  rng = np.random.default_rng(7)
  raw = rng.integers(0, 256, size=3 * 2048, dtype=np.uint8)
  for ids in (False, True):
    tag = 'ids' if ids else 'std'
    b.add(f'decode12_float_{tag}', 'decode12_float', {'ids_format': ids, 'scaled': True}, {'packed': raw},
          {'out': ext.decode12_float(cuda(raw), ids, True)})
    b.add(f'decode12_float_{tag}_raw', 'decode12_float', {'ids_format': ids, 'scaled': False}, {'packed': raw},
          {'out': ext.decode12_float(cuda(raw), ids, False)})
    b.add(f'decode12_half_{tag}', 'decode12_half', {'ids_format': ids, 'scaled': True}, {'packed': raw},
          {'out': ext.decode12_half(cuda(raw), ids, True)})
    b.add(f'decode12_u16_{tag}', 'decode12_u16', {'ids_format': ids}, {'packed': raw},
          {'out': ext.decode12_u16(cuda(raw), ids)})
    vals = rng.integers(0, 65536, size=4096, dtype=np.uint16)
    vals[:64] = np.arange(4064, 4128, dtype=np.uint16)  # straddle the 4095 clamp
    b.add(f'encode12_u16_{tag}', 'encode12_u16', {'ids_format': ids}, {'values': vals},
          {'out': ext.encode12_u16(u16_tensor(vals), ids)})
    f = rng.uniform(-0.05, 1.05, size=4096).astype(np.float32)
    f[:8] = [0.0, 1.0, 0.5, 0.5 / 4095, 1.5 / 4095, 2.5 / 4095, 4094.5 / 4095, 1.2]
    b.add(f'encode12_float_{tag}', 'encode12_float', {'ids_format': ids, 'scaled': True}, {'values': f},
          {'out': ext.encode12_float(cuda(f), ids, True)})
    g = rng.uniform(0, 5000, size=4096).astype(np.float32)
    b.add(f'encode12_float_{tag}_raw', 'encode12_float', {'ids_format': ids, 'scaled': False}, {'values': g},
          {'out': ext.encode12_float(cuda(g), ids, False)})


def white_balance_cases(b: Book):
  rng = np.random.default_rng(11)
  bayer = rng.uniform(-0.05, 1.2, size=(H, W)).astype(np.float32)
  gains = np.array([1.8, 1.0, 2.1], np.float32)
  for name, pat in td.BayerPattern.__members__.items():
    out = td.apply_white_balance(cuda(bayer), cuda(gains), pat)
    b.add(f'white_balance_{name}', 'apply_white_balance', {'pattern': name}, {'bayer': bayer, 'gains': gains},
          {'out': out})


def demosaic_cases(b: Book):
  rgb = synth.scene_rgb(H, W, seed=21)
  rgb2 = synth.scene_rgb(H, W, seed=22)
  size = (W, H)
  for name, pat in td.BayerPattern.__members__.items():
    cfa = synth.mosaic(rgb, name)
    t = cuda(cfa).unsqueeze(-1)
    b.add(f'bilinear_{name}', 'bilinear5x5_demosaic', {'pattern': name}, {'cfa': cfa},
          {'out': td.bilinear5x5_demosaic(t, pat)})
    b.add(f'ppg_{name}', 'ppg', {'pattern': name, 'median_threshold': 0.0}, {'cfa': cfa},
          {'out': td.PPG(dev, size, pat).process(t)})
    b.add(f'rcd_{name}', 'rcd', {'pattern': name}, {'cfa': cfa}, {'out': td.RCD(dev, size, pat).process(t).clone()})
  cfa = synth.mosaic(rgb, 'RGGB')
  t = cuda(cfa).unsqueeze(-1)
  b.add('ppg_RGGB_median', 'ppg', {'pattern': 'RGGB', 'median_threshold': 5.0}, {'cfa': cfa},
        {'out': td.PPG(dev, size, td.BayerPattern.RGGB, median_threshold=5.0).process(t)})
  # negative / >1 samples exercise the clamps
  rng = np.random.default_rng(23)
  wild = (cfa + rng.normal(0, 0.2, size=cfa.shape)).astype(np.float32)
  tw = cuda(wild).unsqueeze(-1)
  b.add('ppg_GRBG_wild', 'ppg', {'pattern': 'GRBG', 'median_threshold': 0.0}, {'cfa': wild},
        {'out': td.PPG(dev, size, td.BayerPattern.GRBG).process(tw)})
  b.add('rcd_GRBG_wild', 'rcd', {'pattern': 'GRBG'}, {'cfa': wild},
        {'out': td.RCD(dev, size, td.BayerPattern.GRBG).process(tw).clone()})
  b.add('bilinear_GBRG_wild', 'bilinear5x5_demosaic', {'pattern': 'GBRG'}, {'cfa': wild},
        {'out': td.bilinear5x5_demosaic(tw, td.BayerPattern.GBRG)})
  # the reference keeps scratch between calls: second frame through a used workspace
  ws = td.RCD(dev, size, td.BayerPattern.RGGB)
  first = ws.process(t).clone()
  cfa2 = synth.mosaic(rgb2, 'RGGB')
  second = ws.process(cuda(cfa2).unsqueeze(-1)).clone()
  fresh2 = td.RCD(dev, size, td.BayerPattern.RGGB).process(cuda(cfa2).unsqueeze(-1)).clone()
  b.add('rcd_RGGB_reuse', 'rcd_reuse', {'pattern': 'RGGB'}, {'cfa_first': cfa, 'cfa': cfa2},
        {'first': first, 'out': second, 'fresh': fresh2})
  # a non-multiple-of-16 frame
  rgbo = synth.scene_rgb(54, 70, seed=24)
  cfao = synth.mosaic(rgbo, 'BGGR')
  to = cuda(cfao).unsqueeze(-1)
  b.add('rcd_BGGR_odd_tiles', 'rcd', {'pattern': 'BGGR'}, {'cfa': cfao},
        {'out': td.RCD(dev, (70, 54), td.BayerPattern.BGGR).process(to).clone()})
  b.add('ppg_BGGR_odd_tiles', 'ppg', {'pattern': 'BGGR', 'median_threshold': 0.0}, {'cfa': cfao},
        {'out': td.PPG(dev, (70, 54), td.BayerPattern.BGGR).process(to)})
  b.add('bilinear_BGGR_odd_tiles', 'bilinear5x5_demosaic', {'pattern': 'BGGR'}, {'cfa': cfao},
        {'out': td.bilinear5x5_demosaic(to, td.BayerPattern.BGGR)})


def postprocess_cases(b: Book):
  rgb = synth.scene_rgb(H, W, seed=31)
  cfa = synth.mosaic(rgb, 'RGGB')
  # G1/G2 imbalance so that green equilibration does something
  cfa[0::2, 1::2] *= 1.04
  size = (W, H)
  pat = td.BayerPattern.RGGB
  demosaiced = td.PPG(dev, size, pat).process(cuda(cfa).unsqueeze(-1))
  dm = demosaiced.cpu().numpy()
  configs = {
    'pp_smooth3_global': dict(color_smoothing_passes=3, green_eq_local=False, green_eq_global=True,
                              green_eq_threshold=0.04),
    'pp_smooth1': dict(color_smoothing_passes=1, green_eq_local=False, green_eq_global=False, green_eq_threshold=0.04),
    'pp_local': dict(color_smoothing_passes=0, green_eq_local=True, green_eq_global=False, green_eq_threshold=4.0),
    'pp_all': dict(color_smoothing_passes=2, green_eq_local=True, green_eq_global=True, green_eq_threshold=4.0),
    'pp_none': dict(color_smoothing_passes=0, green_eq_local=False, green_eq_global=False, green_eq_threshold=0.04),
  }
  for name, kw in configs.items():
    out = td.PostProcess(dev, size, pat, **kw).process(demosaiced)
    b.add(name, 'postprocess', {'pattern': 'RGGB', **kw}, {'rgb': dm}, {'out': out})
  out = td.PostProcess(dev, size, td.BayerPattern.GBRG, color_smoothing_passes=1, green_eq_local=True,
                       green_eq_global=True, green_eq_threshold=4.0).process(demosaiced)
  b.add('pp_all_GBRG', 'postprocess', {'pattern': 'GBRG', 'color_smoothing_passes': 1, 'green_eq_local': True,
                                       'green_eq_global': True, 'green_eq_threshold': 4.0}, {'rgb': dm}, {'out': out})


def color_cases(b: Book):
  rng = np.random.default_rng(41)
  rgb = rng.uniform(-0.05, 1.1, size=(HP, WP, 3)).astype(np.float32)
  rgb[0, :8] = [[0, 0, 0], [1, 1, 1], [0.04045, 0.0031308, 0.5], [0.5, 0.5, 0.5], [1, 0, 0], [0, 1, 0], [0, 0, 1],
                [0.2, 0.2, 0.2]]
  t = cuda(rgb)
  unit = np.clip(rgb, 0.0, 1.0)
  tu = cuda(unit)
  xyz = td.rgb_to_xyz(t)
  lab = td.rgb_to_lab(t)
  b.add('rgb_to_xyz', 'rgb_to_xyz', {}, {'x': rgb}, {'out': xyz})
  b.add('xyz_to_lab', 'xyz_to_lab', {}, {'x': xyz.cpu().numpy()}, {'out': td.xyz_to_lab(xyz)})
  b.add('rgb_to_lab', 'rgb_to_lab', {}, {'x': rgb}, {'out': lab})
  b.add('lab_to_xyz', 'lab_to_xyz', {}, {'x': lab.cpu().numpy()}, {'out': td.lab_to_xyz(lab)})
  b.add('lab_to_rgb', 'lab_to_rgb', {}, {'x': lab.cpu().numpy()}, {'out': td.lab_to_rgb(lab)})
  b.add('xyz_to_rgb', 'xyz_to_rgb', {}, {'x': xyz.cpu().numpy()}, {'out': td.xyz_to_rgb(xyz)})
  lum = td.compute_luminance(t)
  b.add('compute_luminance', 'compute_luminance', {}, {'x': rgb}, {'out': lum})
  loglum = td.compute_log_luminance(t, 1e-4)
  b.add('compute_log_luminance', 'compute_log_luminance', {'eps': 1e-4}, {'x': rgb}, {'out': loglum})
  newl = rng.uniform(-0.1, 1.1, size=(HP, WP)).astype(np.float32)
  b.add('modify_luminance', 'modify_luminance', {}, {'x': unit, 'lum': newl},
        {'out': td.modify_luminance(tu, cuda(newl))})
  newlog = np.log(np.clip(newl, 1e-4, None)).astype(np.float32) + 0.05
  b.add('modify_log_luminance', 'modify_log_luminance', {'eps': 1e-4}, {'x': unit, 'lum': newlog},
        {'out': td.modify_log_luminance(tu, cuda(newlog), 1e-4)})
  b.add('modify_hsl', 'modify_hsl', {'hue_adjust': 0.1, 'sat_adjust': 0.2, 'lum_adjust': -0.1}, {'x': unit},
        {'out': td.modify_hsl(tu, 0.1, 0.2, -0.1)})
  b.add('modify_vibrance', 'modify_vibrance', {'amount': 0.5}, {'x': unit}, {'out': td.modify_vibrance(tu, 0.5)})


def tonemap_cases(b: Book):
  img0 = synth.scene_rgb(HP, WP, seed=51) * 1.5
  img1 = synth.scene_rgb(HP, WP, seed=52) * 0.4
  imgs = [cuda(img0), cuda(img1)]
  ins = {'img0': img0, 'img1': img1}
  for stride in (8, 4, 3):
    b.add(f'bounds_s{stride}', 'compute_image_bounds', {'stride': stride}, ins,
          {'out': td.compute_image_bounds(imgs, stride)})
  for stride, rescale in ((8, False), (4, False), (4, True)):
    b.add(f'metrics_s{stride}_{int(rescale)}', 'compute_image_metrics',
          {'stride': stride, 'min_gray': 1e-4, 'rescale': rescale}, ins,
          {'out': td.compute_image_metrics(imgs, stride, 1e-4, rescale)})
  metrics = td.compute_image_metrics(imgs, 4, 1e-4, False)
  m = metrics.cpu().numpy()
  settings = {'a': (1.5, 2.0, 0.8, 0.5), 'b': (1.0, 0.0, 1.0, 0.0), 'c': (2.2, 1.0, 0.3, -0.4)}
  for tag, (gamma, intensity, la, vib) in settings.items():
    p = td.TonemapParameters(gamma, intensity, la, vib)
    pd = {'gamma': gamma, 'intensity': intensity, 'light_adapt': la, 'vibrance': vib}
    b.add(f'reinhard_{tag}', 'reinhard_tonemap', pd, {'img': img0, 'metrics': m},
          {'out': td.reinhard_tonemap(imgs[0], metrics, p)})
    b.add(f'linear_{tag}', 'linear_tonemap', pd, {'img': img0, 'metrics': m},
          {'out': td.linear_tonemap(imgs[0], metrics, p)})
    b.add(f'aces_{tag}', 'aces_tonemap', pd, {'img': img0}, {'out': td.aces_tonemap(imgs[0], p)})
    b.add(f'adaptive_aces_{tag}', 'adaptive_aces_tonemap', pd, {'img': img0, 'metrics': m},
          {'out': td.aces_tonemap(imgs[0], p, metrics)})


def wiener_cases(b: Book):
  rgb = synth.scene_rgb(H, W, seed=61)
  rng = np.random.default_rng(62)
  noisy = (rgb + rng.normal(0, 0.05, size=rgb.shape)).astype(np.float32)
  lum = noisy[..., 1:2].copy()
  size = (W, H)
  for k, ov, c, sig in ((32, 4, 1, [0.05]), (32, 4, 3, [0.02, 0.05, 0.1]), (16, 2, 1, [0.08]), (16, 8, 3, [0.05] * 3),
                        (32, 2, 3, [0.05] * 3), (32, 8, 1, [0.0])):
    ws = td.Wiener(dev, size, overlap_factor=ov, tile_size=k)
    x = lum if c == 1 else noisy
    out = ws.process(cuda(x), torch.tensor(sig, dtype=torch.float32, device=dev))
    b.add(f'wiener_k{k}_o{ov}_c{c}', 'wiener', {'tile_size': k, 'overlap_factor': ov, 'sigmas': sig}, {'x': x},
          {'out': out})
  ws = td.Wiener(dev, size)
  unit = np.clip(noisy, 0, 1)
  b.add('wiener_log_luminance', 'wiener_log_luminance', {'noise': 0.075, 'eps': 1e-4}, {'x': unit},
        {'out': ws.process_log_luminance(cuda(unit), 0.075)})
  # NB: the reference reads out of bounds (single reflection, denoise.cu:118-122) when a side is shorter
  # than 2*K-1; a 30x21 frame faulted the GPU context here, so the smallest golden frame is 70x63 with K=32.
  small = noisy[:63, :70, 1:2].copy()
  b.add('wiener_small', 'wiener', {'tile_size': 32, 'overlap_factor': 4, 'sigmas': [0.05]}, {'x': small},
        {'out': td.Wiener(dev, (70, 63)).process(cuda(small), torch.tensor([0.05], device=dev))})


def local_contrast_cases(b: Book):
  rgb = synth.scene_rgb(H, W, seed=71)
  t = cuda(rgb)
  lum = td.compute_luminance(t)
  ln = lum.cpu().numpy()
  size = (W, H)
  for ss, sr, detail in ((2.0, 0.2, 0.4), (8.0, 0.1, 0.2), (3.0, 0.05, -0.3)):
    ws = td.Bilateral(dev, size, sigma_s=ss, sigma_r=sr)
    b.add(f'bilateral_{ss:g}_{sr:g}', 'bilateral', {'sigma_s': ss, 'sigma_r': sr, 'detail': detail}, {'lum': ln},
          {'out': ws.process(lum, detail)})
  ws = td.Bilateral(dev, size, sigma_s=2.0, sigma_r=0.2)
  b.add('bilateral_rgb', 'bilateral_rgb', {'sigma_s': 2.0, 'sigma_r': 0.2, 'detail': 0.4}, {'x': rgb},
        {'out': ws.process_rgb(t, 0.4)})
  for tag, p in (('default', td.LaplacianParams()),
                 ('tuned', td.LaplacianParams(sigma=0.25, shadows=0.5, highlights=1.5, clarity=0.3))):
    out = td.Laplacian(dev, size, p).process(lum)
    b.add(f'laplacian_{tag}', 'laplacian',
          {'sigma': p.sigma, 'shadows': p.shadows, 'highlights': p.highlights, 'clarity': p.clarity}, {'lum': ln},
          {'out': out})
  odd = ln[:53, :75].copy()
  out = td.Laplacian(dev, (75, 53), td.LaplacianParams(clarity=0.2)).process(cuda(odd))
  b.add('laplacian_odd', 'laplacian', {'sigma': 0.2, 'shadows': 1.0, 'highlights': 1.0, 'clarity': 0.2}, {'lum': odd},
        {'out': out})


def pipeline_cases(b: Book):
  ph, pw = 96, 128
  frames = [synth.packed_frame(ph, pw, seed=81 + i) for i in range(3)]
  for tag, tm, deb, wb, tf in (('adaptive_aces_rcd', ToneMapper.adaptive_aces, Debayer.rcd, None, ImageTransform.rotate_270),
                               ('reinhard_ppg', ToneMapper.reinhard, Debayer.ppg, (1.8, 1.0, 2.1), ImageTransform.none),
                               ('aces_bilinear', ToneMapper.aces, Debayer.bilinear, (1.2, 1.0, 1.4), ImageTransform.flip_horiz)):
    settings = ImageProcessingSettings(enable_denoise=True, enable_bilateral=True, postprocess=True, tone_gamma=1.5,
                                       tone_intensity=2.0, light_adapt=0.8, tone_mapping=tm, vibrance=0.5, debayer=deb,
                                       moving_average=0.5)
    proc = ImageProcessor((pw, ph), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, wb, tf)
    outs = {}
    # two image sets: the second one sees the EMA-blended bounds/metrics and the re-used RCD scratch
    r0 = proc.process_image_set({'a': cuda(frames[0]), 'b': cuda(frames[1])})
    outs['set0_a'], outs['set0_b'] = r0['a'].clone(), r0['b'].clone()
    outs['bounds0'], outs['metrics0'] = proc.bounds.clone(), proc.metrics.clone()
    r1 = proc.process_image_set({'a': cuda(frames[2])})
    outs['set1_a'] = r1['a'].clone()
    outs['bounds1'], outs['metrics1'] = proc.bounds.clone(), proc.metrics.clone()
    b.add(f'pipeline_{tag}', 'pipeline',
          {'tone_mapping': tm.name, 'debayer': deb.name, 'white_balance': wb, 'transform': tf.name, 'width': pw,
           'height': ph, 'moving_average': 0.5},
          {'frame0': frames[0], 'frame1': frames[1], 'frame2': frames[2]}, outs)


def mid_cases(b: Book):
  """Frames that span many CTA tiles of the new kernels (the other groups are 64x96): multi-tile RCD for two patterns,
  the Wiener log-luminance composite and the bilateral composite.  Inputs are NOT stored (they are regenerated from
  tests/synth.py by seed) to keep the fixture small; the manifest carries the recipe."""
  h, w = 130, 372
  for pattern in ('RGGB', 'GBRG'):
    cfa = synth.mosaic(synth.scene_rgb(h, w, 7), pattern)
    out = td.RCD(dev, (w, h), td.BayerPattern[pattern]).process(cuda(cfa).unsqueeze(-1)).clone()
    b.add(f'rcd_mid_{pattern}', 'rcd', {'pattern': pattern, 'synth': {'kind': 'cfa', 'h': h, 'w': w, 'seed': 7}}, {}, {'out': out})
  h, w = 130, 204
  x = np.clip(synth.scene_rgb(h, w, 17) + np.random.default_rng(6).normal(0, 0.02, size=(h, w, 3)), 0, 1).astype(np.float32)
  out = td.Wiener(dev, (w, h)).process_log_luminance(cuda(x), 0.075, 1e-4)
  b.add('wiener_log_luminance_mid', 'wiener_log_luminance', {'noise': 0.075, 'eps': 1e-4,
        'synth': {'kind': 'noisy_rgb', 'h': h, 'w': w, 'seed': 17, 'noise_seed': 6, 'noise': 0.02}}, {}, {'out': out})
  x = synth.scene_rgb(h, w, 19)
  out = td.Bilateral(dev, (w, h), sigma_s=2.0, sigma_r=0.2).process_rgb(cuda(x), 0.4)
  b.add('bilateral_rgb_mid', 'bilateral_rgb', {'sigma_s': 2.0, 'sigma_r': 0.2, 'detail': 0.4,
        'synth': {'kind': 'rgb', 'h': h, 'w': w, 'seed': 19}}, {}, {'out': out})


def jpeg_image(h, w, seed):
  """uint8 sRGB-like test picture (regenerated by seed in tests/test_jpeg.py)."""
  return np.floor(synth.scene_rgb(h, w, seed) ** (1 / 2.2) * 255.0 + 0.5).clip(0, 255).astype(np.uint8)


JPEG_CASES = [  # (tag, input_format, quality, subsampling, progressive)
  ('rgbi_q94_422', 'RGBI', 94, 'CSS_422', False),
  ('rgbi_q80_444_prog', 'RGBI', 80, 'CSS_444', True),
  ('bgri_q94_422', 'BGRI', 94, 'CSS_422', False),
  ('rgb_planar_q90_gray', 'RGB', 90, 'CSS_GRAY', False),
  ('bgr_planar_q60_444', 'BGR', 60, 'CSS_444', False),
]


def jpeg_cases(b: Book):
  """Streams of the reference's `Jpeg.encode` (same nvJPEG library, so the new binding must reproduce them byte for byte)."""
  h, w = 120, 176
  img = jpeg_image(h, w, 23)
  coder = td.Jpeg()
  for tag, fmt, quality, css, prog in JPEG_CASES:
    x = cuda(img if fmt.endswith('I') else np.ascontiguousarray(img.transpose(2, 0, 1)))
    out = coder.encode(x, quality=quality, input_format=td.jpeg.InputFormat[fmt], subsampling=td.jpeg.Subsampling[css], progressive=prog)
    b.add(f'jpeg_{tag}', 'jpeg', {'input_format': fmt, 'quality': quality, 'subsampling': css, 'progressive': prog,
          'synth': {'kind': 'jpeg_image', 'h': h, 'w': w, 'seed': 23}}, {}, {'out': out})


def main():
  out_dir = Path(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/golden')
  out_dir.mkdir(parents=True, exist_ok=True)
  print('reference:', td.__file__, 'torch', torch.__version__, torch.cuda.get_device_name(0))
  groups = {
    'packed': [packed_cases, white_balance_cases],
    'demosaic': [demosaic_cases, postprocess_cases],
    'color': [color_cases, tonemap_cases],
    'filters': [wiener_cases, local_contrast_cases],
    'pipeline': [pipeline_cases],
    'mid': [mid_cases],
    'jpeg': [jpeg_cases],
  }
  only = sys.argv[2].split(',') if len(sys.argv) > 2 else list(groups)
  errors = []
  for gname, fns in groups.items():
    if gname not in only:
      continue
    book = Book()
    for fn in fns:
      book.run(fn)
    book.save(out_dir / f'{gname}.npz')
    errors += book.errors
  (out_dir / f'errors_{"_".join(only)}.txt').write_text('\n'.join(errors))
  print('errors:', len(errors))
  for e in errors:
    print(e)


if __name__ == '__main__':
  main()
