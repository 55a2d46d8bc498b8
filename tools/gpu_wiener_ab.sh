# Wiener A/B: shared-column kernel (default) against wiener32_kernel (TDB_WIENER_SHARED=0): parity tests, then the quick bench of both
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "wiener or pipeline or fused or tiled or band" 2>&1 | tail -15 > gpurun_out/pytest_wiener.log; cat gpurun_out/pytest_wiener.log
for mode in 1 0; do
  TDB_WIENER_SHARED=$mode python bench.py --frames 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wiener_$mode.json 2> gpurun_out/bench_wiener_$mode.err; tail -c 300 gpurun_out/bench_wiener_$mode.err
  python - <<P
import json
d=json.loads(open('gpurun_out/bench_wiener_$mode.json').read().strip().splitlines()[-1])
print('shared=$mode value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
for s in d['stages'][:3]: print(f"   {s['kernel']:28s} {s['ms_per_launch']:.4f} ms")
P
done
