#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "bilateral or local_contrast or ppg or bilinear or demosaic or pipeline or golden or tiled or bands" 2>&1 | tail -8 > gpurun_out/pytest_c.log; cat gpurun_out/pytest_c.log
python tools/bench_stages.py --configs 4,2 --kernels > gpurun_out/r02_stages_c.jsonl 2> gpurun_out/r02_stages_c.err
grep -E '"op"' gpurun_out/r02_stages_c.jsonl | grep -E "Bilateral|PPG|ppg|bilinear" | python -c "
import sys, json
for l in sys.stdin:
  d = json.loads(l); print(d['op'], d['ms'], d['frac_of_measured_hbm'])"
grep -A1 "sigma_s 2" gpurun_out/r02_stages_c.jsonl | tail -1 | cut -c1-300
grep sat gpurun_out/ref_live_report.jsonl | tail -2
