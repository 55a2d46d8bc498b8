"""BASELINE.json configs[0] on the host cores: one 12 MP (4096 x 3000) synthetic RGGB frame -> white balance + bilinear 5x5 demosaic
+ colour conversion (rgb_to_lab) through the CPU oracle (oracle/, C + OpenMP) -- the reference has no CPU path of its own
(SURVEY.md 8c), so this restatement is the CPU baseline of that chain.  Wall clock, median of 5, thread count stated.
  python tools/bench_config1_cpu.py   ->   one JSON line
The same chain through the CUDA package takes 0.143 ms on a B200 (profiles/r01_stage_table.md, config1)."""
import json, os, statistics, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np
import oracle, synth

w, h = 4096, 3000
cfa = synth.mosaic(synth.scene_rgb(h, w, 1234), 'RGGB')
gains = (1.8, 1.0, 2.1)


def chain():
  return oracle.color_convert(oracle.bilinear5x5(oracle.white_balance(cfa, gains, 'RGGB'), 'RGGB'), 'rgb_to_lab')


chain()
times = []
for _ in range(5):
  t = time.perf_counter(); out = chain(); times.append(time.perf_counter() - t)
ms = statistics.median(times) * 1e3
print(json.dumps({'config': 'configs[0]: 12 MP RGGB -> white_balance + bilinear5x5 + rgb_to_lab', 'impl': 'CPU oracle (C + OpenMP)',
                  'ms': round(ms, 2), 'mp_per_s': round(w * h / 1e6 / (ms / 1e3), 1), 'threads': os.cpu_count(), 'runs_ms': [round(t * 1e3, 2) for t in times],
                  'checksum': float(np.asarray(out, dtype=np.float64).sum())}))
