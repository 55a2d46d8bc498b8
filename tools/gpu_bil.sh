mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "bilateral or pipeline or fused or tiled or band" 2>&1 | tail -6
python bench.py --frames 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 300 gpurun_out/bench_quick.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
for s in d['stages']: print(f"   {s['kernel']:28s} {s['ms_per_launch']:.4f} ms")
P
