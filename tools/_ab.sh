python -m pytest tests -m gpu -x -q -k "rcd or RCD or demosaic or pipeline or fused or golden" 2>&1 | tail -3
python bench.py --frames 8 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v3', d['value'], [ (s['kernel'], s['ms_per_launch']) for s in d['stages'] if 'rcd' in s['kernel']])"
