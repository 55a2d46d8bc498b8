timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --frames 8 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], [ (s['kernel'], s['ms_per_launch']) for s in d['stages'] if 'smooth' in s['kernel'] or 'rcd' in s['kernel']])"
timeout 300 python tools/bench_stages.py --impl ours 2>&1 | grep -i "bilinear\|ppg\|PostProcess" | cut -c40-200
