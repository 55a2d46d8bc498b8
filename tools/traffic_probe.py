"""DRAM bytes per pixel, fused (ours) vs unfused (reference), measured with ncu -- the north-star's "total HBM bytes per pixel against
the unfused reference" for the frame pipeline (BASELINE.json configs[2]) and for the 50 MP local-contrast stages (configs[3]).

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
      --csv --log-file gpurun_out/traffic_<impl>.csv python tools/traffic_probe.py --impl <ours|reference>
  python tools/traffic_probe.py --summarise gpurun_out/traffic_ours.csv gpurun_out/traffic_reference.csv > profiles/rNN_traffic_fused_vs_unfused.md

The probe runs three sections between cudaProfilerStart/Stop, separated by a marker kernel (torch.arange): one 3840x2160 frame through
ImageProcessor.process (after two warm-up frames), Bilateral.process_rgb (sigma_s 8, sigma_r 0.2, detail 0.4) on 8192x6144, Laplacian
(default parameters) on 8192x6144.  Under ncu every kernel is serialised and replayed with cold caches, so the byte counts are upper
bounds of what the back-to-back pipeline moves (L2 carry-over between launches is lost); both packages are measured the same way."""
import argparse, collections, csv, json, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SECTIONS = [('frame pipeline 3840x2160', 3840 * 2160), ('Bilateral.process_rgb 8192x6144', 8192 * 6144), ('Laplacian 8192x6144', 8192 * 6144)]


def probe(impl):
  sys.path.insert(0, str(ROOT / ('torch-darktable_b200' if impl == 'ours' else 'baseline/_ref')))
  sys.path.insert(0, str(ROOT))
  import torch
  import torch_darktable as td
  from torch_darktable.pipeline.config import Debayer, ImageProcessingSettings, ToneMapper
  from torch_darktable.pipeline.image_processor import ImageProcessor
  from torch_darktable.pipeline.transform import ImageTransform
  import bench
  dev = torch.device('cuda:0')
  frames = [torch.from_numpy(f).to(dev) for f in bench.make_frames(3, 0)]
  settings = ImageProcessingSettings(debayer=Debayer.rcd, tone_mapping=ToneMapper.adaptive_aces, **bench.settings_kwargs())
  proc = ImageProcessor((bench.WIDTH, bench.HEIGHT), td.BayerPattern.RGGB, td.PackedFormat.Packed12, settings, dev, None, ImageTransform.rotate_270)
  gen = torch.Generator(device=dev).manual_seed(1234)
  w, h = 8192, 6144
  rgb = [torch.rand((h, w, 3), device=dev, generator=gen) for _ in range(2)]
  lum = [torch.rand((h, w), device=dev, generator=gen) for _ in range(2)]
  bil = td.Bilateral(dev, (w, h), sigma_s=8.0, sigma_r=0.2)
  lap = td.Laplacian(dev, (w, h), td.LaplacianParams())
  for f in frames[:2]:
    proc.process(f, 'cam')
  bil.process_rgb(rgb[0], 0.4)
  lap.process(lum[0])
  torch.cuda.synchronize()
  marker = lambda: torch.arange(12345, device=dev)
  torch.cuda.profiler.start()
  proc.process(frames[2], 'cam')
  marker()
  bil.process_rgb(rgb[1], 0.4)
  marker()
  lap.process(lum[1])
  torch.cuda.synchronize()
  torch.cuda.profiler.stop()


def read(path):
  rows = list(csv.reader(open(path, errors='replace')))
  hdr = next(r for r in rows if 'Kernel Name' in r)
  body = rows[rows.index(hdr) + 1:]
  k, m, u, v, i = (hdr.index(c) for c in ('Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value', 'ID'))
  scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}
  launches = collections.OrderedDict()
  for r in body:
    if len(r) <= v:
      continue
    d = launches.setdefault(r[i], {'name': r[k]})
    d[r[m]] = float(r[v].replace(',', '')) * scale.get(r[u], 1)
  sections, cur = [], []
  for d in launches.values():
    if 'arange' in d['name'] or 'elementwise_kernel_with_index' in d['name']:
      sections.append(cur)
      cur = []
    else:
      cur.append(d)
  sections.append(cur)
  return sections


def summarise(paths):
  data = {('ours' if 'ours' in p else 'reference'): read(p) for p in paths}
  print('# Measured DRAM traffic, fused (ours) vs unfused (reference): ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, summed\n')
  print('Method: tools/traffic_probe.py under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none`;')
  print('one call of each workload after warm-up, every launch replayed alone with cold caches (an upper bound for the back-to-back pipeline).\n')
  print('| workload | package | launches | DRAM read MB | DRAM write MB | B/px | sum of kernel times ms |')
  print('|---|---|---|---|---|---|---|')
  for s, (name, px) in enumerate(SECTIONS):
    for impl in ('reference', 'ours'):
      if impl not in data or s >= len(data[impl]):
        continue
      sec = data[impl][s]
      rd = sum(d.get('dram__bytes_read.sum', 0) for d in sec)
      wr = sum(d.get('dram__bytes_write.sum', 0) for d in sec)
      t = sum(d.get('gpu__time_duration.sum', 0) for d in sec)
      print(f'| {name} | {impl} | {len(sec)} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / px:.1f} | {t / 1e3:.3f} |')
  print()
  for impl in data:
    for s, (name, px) in enumerate(SECTIONS):
      if s >= len(data[impl]):
        continue
      print(f'\n## {impl}: {name}\n\n| kernel | launches | read MB | write MB | us |\n|---|---|---|---|---|')
      agg = collections.OrderedDict()
      for d in data[impl][s]:
        a = agg.setdefault(d['name'].split('(')[0][-70:], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d.get('dram__bytes_read.sum', 0); a[2] += d.get('dram__bytes_write.sum', 0); a[3] += d.get('gpu__time_duration.sum', 0)
      for n, a in agg.items():
        print(f'| `{n}` | {a[0]} | {a[1] / 1e6:.1f} | {a[2] / 1e6:.1f} | {a[3]:.1f} |')


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--summarise', nargs='+')
  a = ap.parse_args()
  summarise(a.summarise) if a.summarise else probe(a.impl)
