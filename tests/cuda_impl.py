"""The product package (torch_darktable on libtdb200.so) behind the golden-case table of tests/cases.py.

Everything goes through the reference-facing Python API, i.e. through the C ABI; nothing here touches oracle/.
"""

from __future__ import annotations

import numpy as np
import torch

import torch_darktable as td
from torch_darktable.pipeline import ImageProcessingSettings, ImageProcessor, ImageTransform
from torch_darktable.pipeline.config import Debayer, ToneMapper

DEV = torch.device('cuda:0')


def dev(a, dtype=None):
  a = np.ascontiguousarray(a)
  if a.dtype == np.uint16:
    return torch.from_numpy(a.view(np.int16)).view(torch.uint16).to(DEV)
  t = torch.from_numpy(a).to(DEV)
  return t if dtype is None else t.to(dtype)


def host(t: torch.Tensor) -> np.ndarray:
  t = t.detach().cpu()
  if t.dtype == torch.uint16:
    return t.view(torch.int16).numpy().view(np.uint16)
  return t.numpy()


def pat(name):
  return td.BayerPattern[name] if isinstance(name, str) else name


class CudaImpl:
  def decode12(self, packed, dtype, ids, scaled):
    fmt = {np.float32: td.decode12_float, np.float16: td.decode12_half, np.uint16: td.decode12_u16}[dtype]
    if dtype == np.uint16:
      return host(fmt(dev(packed), ids_format=ids))
    return host(fmt(dev(packed), ids_format=ids, scaled=scaled))

  def encode12(self, values, ids, scaled):
    if values.dtype == np.uint16:
      return host(td.encode12_u16(dev(values), ids_format=ids))
    return host(td.encode12_float(dev(values), ids_format=ids, scaled=scaled))

  def white_balance(self, bayer, gains, pattern):
    return host(td.apply_white_balance(dev(bayer), dev(gains), pat(pattern)))

  def bilinear5x5(self, cfa, pattern):
    return host(td.bilinear5x5_demosaic(dev(cfa).unsqueeze(-1), pat(pattern)))

  def ppg(self, cfa, pattern, thr):
    h, w = cfa.shape
    return host(td.PPG(DEV, (w, h), pat(pattern), median_threshold=float(thr)).process(dev(cfa).unsqueeze(-1)))

  def rcd(self, cfa, pattern):
    h, w = cfa.shape
    return host(td.RCD(DEV, (w, h), pat(pattern)).process(dev(cfa).unsqueeze(-1)))

  def rcd_sequence(self, cfas, pattern):
    h, w = cfas[0].shape
    ws = td.RCD(DEV, (w, h), pat(pattern))
    return [host(ws.process(dev(c).unsqueeze(-1))) for c in cfas]

  def postprocess(self, rgb, pattern, **kw):
    h, w, _ = rgb.shape
    return host(td.PostProcess(DEV, (w, h), pat(pattern), **kw).process(dev(rgb)))

  def color_convert(self, x, op, params):
    fn = getattr(td, op)
    return host(fn(dev(x), *[float(p) for p in params]))

  def compute_luminance(self, x): return host(td.compute_luminance(dev(x)))
  def compute_log_luminance(self, x, eps): return host(td.compute_log_luminance(dev(x), float(eps)))
  def modify_luminance(self, x, lum): return host(td.modify_luminance(dev(x), dev(lum)))
  def modify_log_luminance(self, x, lum, eps): return host(td.modify_log_luminance(dev(x), dev(lum), float(eps)))
  def compute_image_bounds(self, imgs, stride): return host(td.compute_image_bounds([dev(i) for i in imgs], stride))

  def compute_image_metrics(self, imgs, stride, mg, rs):
    return host(td.compute_image_metrics([dev(i) for i in imgs], stride, float(mg), bool(rs)))

  def tonemap(self, img, op, metrics, gamma, intensity, la, vib):
    p = td.TonemapParameters(float(gamma), float(intensity), float(la), float(vib))
    t = dev(img)
    if op == 'reinhard':
      return host(td.reinhard_tonemap(t, dev(metrics), p))
    if op == 'linear':
      return host(td.linear_tonemap(t, dev(metrics), p))
    if op == 'aces':
      return host(td.aces_tonemap(t, p))
    return host(td.aces_tonemap(t, p, dev(metrics)))

  def wiener(self, x, sigmas, k, ov):
    h, w, _ = x.shape
    ws = td.Wiener(DEV, (w, h), overlap_factor=ov, tile_size=k)
    return host(ws.process(dev(x), torch.tensor(sigmas, dtype=torch.float32, device=DEV)))

  def wiener_log_luminance(self, x, noise, eps):
    h, w, _ = x.shape
    return host(td.Wiener(DEV, (w, h)).process_log_luminance(dev(x), float(noise), float(eps)))

  def bilateral(self, lum, ss, sr, d):
    h, w = lum.shape
    return host(td.Bilateral(DEV, (w, h), sigma_s=float(ss), sigma_r=float(sr)).process(dev(lum), float(d)))

  def bilateral_rgb(self, x, ss, sr, d):
    h, w, _ = x.shape
    return host(td.Bilateral(DEV, (w, h), sigma_s=float(ss), sigma_r=float(sr)).process_rgb(dev(x), float(d)))

  def laplacian(self, lum, s, sh, hi, cl):
    h, w = lum.shape
    p = td.LaplacianParams(sigma=float(s), shadows=float(sh), highlights=float(hi), clarity=float(cl))
    return host(td.Laplacian(DEV, (w, h), p).process(dev(lum)))

  def pipeline(self, p, set0, set1):
    settings = ImageProcessingSettings(enable_denoise=True, enable_bilateral=True, postprocess=True, tone_gamma=1.5,
                                       tone_intensity=2.0, light_adapt=0.8, tone_mapping=ToneMapper[p['tone_mapping']], vibrance=0.5,
                                       debayer=Debayer[p['debayer']], moving_average=p['moving_average'])
    wb = tuple(p['white_balance']) if p['white_balance'] is not None else None
    proc = ImageProcessor((p['width'], p['height']), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, DEV, wb,
                          ImageTransform[p['transform']])
    r0 = proc.process_image_set({'a': dev(set0[0]), 'b': dev(set0[1])})
    out = {'set0_a': host(r0['a']), 'set0_b': host(r0['b']), 'bounds0': host(proc.bounds), 'metrics0': host(proc.metrics)}
    r1 = proc.process_image_set({'a': dev(set1[0])})
    out.update({'set1_a': host(r1['a']), 'bounds1': host(proc.bounds), 'metrics1': host(proc.metrics)})
    return out
