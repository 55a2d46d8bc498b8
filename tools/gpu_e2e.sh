mkdir -p gpurun_out
python -m pytest tests/test_gpu_sizes.py -m gpu -x -q -k "runner or pipeline" 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; tail -c 300 gpurun_out/bench_e2e.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e'])
"
