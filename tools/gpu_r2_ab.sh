#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "wiener or pipeline or tiled or batch or laplacian or local_contrast or repeatable or bands" 2>&1 | tail -8 > gpurun_out/pytest_ab.log; cat gpurun_out/pytest_ab.log
for c in 2 3; do
  TDB_WIENER_CTAS=$c python bench.py --frames 16 --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_wiener_ctas$c.json 2> gpurun_out/bench_wiener_ctas$c.err
  python - <<P
import json
d=json.loads(open('gpurun_out/bench_wiener_ctas$c.json').read().strip().splitlines()[-1])
print('ctas', $c, 'value', d['value'], [ (s['kernel'], s['ms_per_launch']) for s in d['stages'][:3] ])
P
done
python tools/bench_stages.py --configs 4 --kernels 2>/dev/null | grep -i "laplacian" | cut -c1-400
python tools/bench_tiled.py --breakdown 2>/dev/null | tail -1 | cut -c1-900
