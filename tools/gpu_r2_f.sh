#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "batch or wiener or host or submit or runner" 2>&1 | tail -2
for i in 1 2; do
python bench.py --steps 8 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], [(s['kernel'][:8], s['ms_per_launch']) for s in d['stages'][:2]], d['clocks'])"
done
