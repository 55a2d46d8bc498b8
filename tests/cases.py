"""Dispatch of the golden cases (tests/golden/*.npz, written by make_golden.py) onto an implementation.

`run_case(impl, op, params, inputs)` returns {output-name: ndarray}.  Two implementations exist:
  - OracleImpl : the CPU oracle (oracle/), numpy in / numpy out
  - CudaImpl   : the product package torch_darktable (B200 kernels through the C ABI), in tests/test_gpu_*.py
Both expose the same small method set so that the parity tests read like one table.
"""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np

GOLDEN_DIR = Path(__file__).resolve().parent / 'golden'
GROUPS = ('packed', 'demosaic', 'color', 'filters', 'pipeline', 'mid')


def synth_inputs(recipe: dict) -> dict:
  """Inputs of the 'mid' golden cases are regenerated from tests/synth.py instead of being stored."""
  import synth
  h, w = recipe['h'], recipe['w']
  if recipe['kind'] == 'cfa':
    return {'cfa': synth.mosaic(synth.scene_rgb(h, w, recipe['seed']), recipe['pattern'])}
  if recipe['kind'] == 'noisy_rgb':
    noise = np.random.default_rng(recipe['noise_seed']).normal(0, recipe['noise'], size=(h, w, 3))
    return {'x': np.clip(synth.scene_rgb(h, w, recipe['seed']) + noise, 0, 1).astype(np.float32)}
  return {'x': synth.scene_rgb(h, w, recipe['seed'])}


def load_group(group: str):
  path = GOLDEN_DIR / f'{group}.npz'
  if not path.exists():
    return []
  z = np.load(path)
  manifest = json.loads(str(z['manifest']))
  cases = []
  for name, info in manifest.items():
    ins = {k.split('/in/')[1]: z[k] for k in z.files if k.startswith(f'{name}/in/')}
    if 'synth' in info['params']:
      info['params'] = dict(info['params'])
      recipe = dict(info['params'].pop('synth'), pattern=info['params'].get('pattern'))
      ins = synth_inputs(recipe)
    outs = {k.split('/out/')[1]: z[k] for k in z.files if k.startswith(f'{name}/out/')}
    cases.append((name, info['op'], info['params'], ins, outs))
  return cases


def all_cases():
  for g in GROUPS:
    for c in load_group(g):
      yield (g, *c)


# Tolerances (max-abs unless stated) of an implementation against the reference's own outputs.  The
# reference is built with --use_fast_math; packed/WB/indexing work is exact by construction.
#   'exact'  : bit-identical
#   float    : max |a-b|
#   ('u8', f): uint8 images, at most 1 LSB apart and at most fraction f of samples different
TOLERANCE = {
  'decode12_float': 'exact', 'decode12_half': 'exact', 'decode12_u16': 'exact',
  'encode12_u16': 'exact', 'encode12_float': 'exact',
  'apply_white_balance': 'exact',
  'bilinear5x5_demosaic': 1e-6, 'ppg': 1e-6, 'rcd': 5e-6, 'rcd_reuse': 5e-6, 'postprocess': 1e-6,
  'rgb_to_xyz': 2e-5, 'xyz_to_lab': 2e-5, 'lab_to_xyz': 2e-5, 'xyz_to_rgb': 2e-5, 'rgb_to_lab': 2e-5,
  'lab_to_rgb': 2e-5, 'modify_hsl': 2e-5, 'modify_vibrance': 2e-5, 'compute_luminance': 2e-5,
  'compute_log_luminance': 2e-5, 'modify_luminance': 2e-5, 'modify_log_luminance': 2e-5,
  'compute_image_bounds': 'exact', 'compute_image_metrics': 2e-6,
  'reinhard_tonemap': ('u8', 1e-3), 'aces_tonemap': ('u8', 1e-3), 'adaptive_aces_tonemap': ('u8', 1e-3),
  'linear_tonemap': ('u8', 1e-3),
  # K=16/overlap 8 shows 4e-4 between the reference and ANY restatement: its block_mean zeroes a shared
  # accumulator without a barrier (denoise.cu:85-101), so the reference itself is not deterministic there
  'wiener': 1e-3, 'wiener_log_luminance': 2e-5,
  'bilateral': 2e-6, 'bilateral_rgb': 2e-5,
  'laplacian': 2e-3,  # fp16 storage: one half ulp at 1.0 is 4.9e-4; the oracle happens to be bit-exact
  'pipeline': ('u8', 1e-3),
}


def compare(op: str, got: np.ndarray, ref: np.ndarray, tol=None) -> str | None:
  """Returns None when `got` matches `ref` under the op's tolerance, else a message."""
  tol = TOLERANCE[op] if tol is None else tol
  got = np.asarray(got).reshape(ref.shape)
  if tol == 'exact':
    if got.dtype != ref.dtype:
      return f'dtype {got.dtype} != {ref.dtype}'
    same = np.array_equal(got.view(np.uint8) if got.dtype.kind == 'f' else got,
                          ref.view(np.uint8) if ref.dtype.kind == 'f' else ref)
    return None if same else f'not bit-exact: {np.sum(got != ref)} of {ref.size} differ'
  if isinstance(tol, tuple) and tol[0] == 'outliers':
    _, t, frac, hard = tol
    d = np.nan_to_num(np.abs(got.astype(np.float64) - ref.astype(np.float64)), nan=np.inf)
    if d.max() > hard:
      return f'max abs diff {d.max():.3e} > hard limit {hard:g}'
    if (d > t).mean() > frac:
      return f'fraction beyond {t:g}: {(d > t).mean():.2e} > {frac:g}'
    return None
  if isinstance(tol, tuple):
    frac = tol[1]
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    if len(tol) > 2:  # ('u8', frac, outlier_frac): a few samples may be further than 1 LSB apart
      if (d > 1).mean() > tol[2]:
        return f'uint8 fraction beyond 1 LSB {(d > 1).mean():.2e} > {tol[2]}'
    elif d.max() > 1:
      return f'uint8 max diff {d.max()} > 1 LSB'
    if (d > 0).mean() > frac:
      return f'uint8 fraction different {(d > 0).mean():.2e} > {frac}'
    return None
  g, r = got.astype(np.float64), ref.astype(np.float64)
  if np.any(np.isnan(g) != np.isnan(r)):
    return 'NaN pattern differs'
  d = np.nan_to_num(np.abs(g - r), nan=0.0)
  return None if d.max() <= tol else f'max abs diff {d.max():.3e} > {tol:g} (at {np.unravel_index(d.argmax(), d.shape)})'


def run_case(impl, op: str, p: dict, i: dict) -> dict:
  """Execute one golden case on `impl`; returns outputs keyed like the golden file."""
  if op in ('decode12_float', 'decode12_half', 'decode12_u16'):
    dt = {'decode12_float': np.float32, 'decode12_half': np.float16, 'decode12_u16': np.uint16}[op]
    return {'out': impl.decode12(i['packed'], dt, p['ids_format'], p.get('scaled', True))}
  if op in ('encode12_u16', 'encode12_float'):
    return {'out': impl.encode12(i['values'], p['ids_format'], p.get('scaled', True))}
  if op == 'apply_white_balance':
    return {'out': impl.white_balance(i['bayer'], i['gains'], p['pattern'])}
  if op == 'bilinear5x5_demosaic':
    return {'out': impl.bilinear5x5(i['cfa'], p['pattern'])}
  if op == 'ppg':
    return {'out': impl.ppg(i['cfa'], p['pattern'], p['median_threshold'])}
  if op == 'rcd':
    return {'out': impl.rcd(i['cfa'], p['pattern'])}
  if op == 'rcd_reuse':
    first, second = impl.rcd_sequence([i['cfa_first'], i['cfa']], p['pattern'])
    return {'first': first, 'out': second, 'fresh': impl.rcd(i['cfa'], p['pattern'])}
  if op == 'postprocess':
    q = dict(p)
    pattern = q.pop('pattern')
    return {'out': impl.postprocess(i['rgb'], pattern, **q)}
  if op in ('rgb_to_xyz', 'xyz_to_lab', 'lab_to_xyz', 'xyz_to_rgb', 'rgb_to_lab', 'lab_to_rgb'):
    return {'out': impl.color_convert(i['x'], op, [])}
  if op == 'modify_hsl':
    return {'out': impl.color_convert(i['x'], op, [p['hue_adjust'], p['sat_adjust'], p['lum_adjust']])}
  if op == 'modify_vibrance':
    return {'out': impl.color_convert(i['x'], op, [p['amount']])}
  if op == 'compute_luminance':
    return {'out': impl.compute_luminance(i['x'])}
  if op == 'compute_log_luminance':
    return {'out': impl.compute_log_luminance(i['x'], p['eps'])}
  if op == 'modify_luminance':
    return {'out': impl.modify_luminance(i['x'], i['lum'])}
  if op == 'modify_log_luminance':
    return {'out': impl.modify_log_luminance(i['x'], i['lum'], p['eps'])}
  if op == 'compute_image_bounds':
    return {'out': impl.compute_image_bounds([i['img0'], i['img1']], p['stride'])}
  if op == 'compute_image_metrics':
    return {'out': impl.compute_image_metrics([i['img0'], i['img1']], p['stride'], p['min_gray'], p['rescale'])}
  if op.endswith('_tonemap'):
    metrics = i.get('metrics')
    return {'out': impl.tonemap(i['img'], op[:-len('_tonemap')], metrics, p['gamma'], p['intensity'], p['light_adapt'],
                                p['vibrance'])}
  if op == 'wiener':
    return {'out': impl.wiener(i['x'], p['sigmas'], p['tile_size'], p['overlap_factor'])}
  if op == 'wiener_log_luminance':
    return {'out': impl.wiener_log_luminance(i['x'], p['noise'], p['eps'])}
  if op == 'bilateral':
    return {'out': impl.bilateral(i['lum'], p['sigma_s'], p['sigma_r'], p['detail'])}
  if op == 'bilateral_rgb':
    return {'out': impl.bilateral_rgb(i['x'], p['sigma_s'], p['sigma_r'], p['detail'])}
  if op == 'laplacian':
    return {'out': impl.laplacian(i['lum'], p['sigma'], p['shadows'], p['highlights'], p['clarity'])}
  if op == 'pipeline':
    return impl.pipeline(p, [i['frame0'], i['frame1']], [i['frame2']])
  raise KeyError(op)


# RCD: the reference's output inside the 7-px margin depends weakly on what the workspace held before
# (SURVEY 8a6); rows/cols 7-8 and H-9..H-8 of a *used* workspace may differ from a fresh one by ~1e-3.
RCD_REUSE_BAND_TOL = 2e-3


class OracleImpl:
  """The CPU oracle behind the case table."""

  def __init__(self):
    import oracle
    self.o = oracle

  def decode12(self, packed, dtype, ids, scaled): return self.o.decode12(packed, dtype, ids, scaled)
  def encode12(self, values, ids, scaled): return self.o.encode12(values, ids, scaled)
  def white_balance(self, bayer, gains, pattern): return self.o.white_balance(bayer, gains, pattern)
  def bilinear5x5(self, cfa, pattern): return self.o.bilinear5x5(cfa, pattern)
  def ppg(self, cfa, pattern, thr): return self.o.ppg(cfa, pattern, thr)
  def rcd(self, cfa, pattern): return self.o.rcd(cfa, pattern)

  def rcd_sequence(self, cfas, pattern):
    ws = self.o.RCDWorkspace(cfas[0].shape[1], cfas[0].shape[0], pattern)
    return [ws.process(c) for c in cfas]

  def postprocess(self, rgb, pattern, **kw): return self.o.postprocess(rgb, pattern, **kw)
  def color_convert(self, x, op, params): return self.o.color_convert(x, op, params)
  def compute_luminance(self, x): return self.o.compute_luminance(x)
  def compute_log_luminance(self, x, eps): return self.o.compute_log_luminance(x, eps)
  def modify_luminance(self, x, lum): return self.o.modify_luminance(x, lum)
  def modify_log_luminance(self, x, lum, eps): return self.o.modify_log_luminance(x, lum, eps)
  def compute_image_bounds(self, imgs, stride): return self.o.compute_image_bounds(imgs, stride)
  def compute_image_metrics(self, imgs, stride, mg, rs): return self.o.compute_image_metrics(imgs, stride, mg, rs)

  def tonemap(self, img, op, metrics, gamma, intensity, la, vib):
    return self.o.tonemap(img, op, metrics, gamma, intensity, la, vib)

  def wiener(self, x, sigmas, k, ov): return self.o.wiener(x, sigmas, k, ov)
  def wiener_log_luminance(self, x, noise, eps): return self.o.wiener_log_luminance(x, noise, eps)
  def bilateral(self, lum, ss, sr, d): return self.o.bilateral(lum, ss, sr, d)
  def bilateral_rgb(self, x, ss, sr, d): return self.o.bilateral_rgb(x, ss, sr, d)
  def laplacian(self, lum, s, sh, hi, cl): return self.o.laplacian(lum, s, sh, hi, cl)

  def pipeline(self, p, set0, set1):
    pl = self.o.Pipeline(p['width'], p['height'], white_balance=p['white_balance'], debayer=p['debayer'],
                         tone_mapping=p['tone_mapping'], moving_average=p['moving_average'], transform=p['transform'])
    r0 = pl.process_image_set(set0)
    out = {'set0_a': r0[0], 'set0_b': r0[1], 'bounds0': pl.bounds.copy(), 'metrics0': pl.metrics.copy()}
    r1 = pl.process_image_set(set1)
    out.update({'set1_a': r1[0], 'bounds1': pl.bounds.copy(), 'metrics1': pl.metrics.copy()})
    return out


PIPELINE_STATE_TOL = 2e-6  # bounds/metrics vectors inside the pipeline cases

# Comparisons that involve the CPU ORACLE on frames of more than a few tiles.  RCD picks between two interpolation
# directions with `|0.5 - c0| < |0.5 - nb| ? nb : c0` (reference rcd.cu:117-121) where c0/nb come out of approximate
# (--use_fast_math) divisions.  The oracle is IEEE C, so on ~2e-4 of the pixels the select flips and those pixels differ by
# up to ~2e-2 (measured against the reference itself: tools/three_way.py, profiles/r01_three_way_mid.log) while the CUDA
# path, compiled like the reference, stays within 2.4e-7 of it everywhere.  Against the reference's golden outputs the CUDA
# path is held to the strict TOLERANCE table; only oracle comparisons get the outlier allowance.
ORACLE_TOLERANCE = {'rcd': ('outliers', 5e-6, 5e-4, 0.05), 'pipeline': ('u8', 1e-3, 2e-4)}


def check_outputs(op: str, got: dict, ref: dict, oracle: bool = False) -> list[str]:
  problems = []
  for key, r in ref.items():
    if key not in got:
      problems.append(f'{key}: missing')
      continue
    if op == 'pipeline' and key.startswith(('bounds', 'metrics')):
      msg = compare(op, got[key], r, PIPELINE_STATE_TOL)
    elif oracle and op in ORACLE_TOLERANCE:
      msg = compare(op, got[key], r, ORACLE_TOLERANCE[op])
    else:
      msg = compare(op, got[key], r)
    if msg:
      problems.append(f'{key}: {msg}')
  return problems
