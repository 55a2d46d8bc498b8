# quick GPU check: parity tests + short bench (per-kernel table), results under gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --frames 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 400 gpurun_out/bench_quick.err
python - <<'E'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
for s in d['stages']: print(f"{s['kernel']:28s} n={s['launches_per_step']:3d} {s['ms_per_launch']:.4f} ms  share {s['share']:.3f}  frac {s['frac']}")
E
