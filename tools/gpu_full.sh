# full round-end style check: GPU parity tests, both bench arms, ncu launch list.  usage: bash tools/gpu_full.sh <tag>
set -x
TAG=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours_$TAG.json 2> gpurun_out/bench_ours_$TAG.err; tail -c 600 gpurun_out/bench_ours_$TAG.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 600 gpurun_out/bench_ref_$TAG.err
CMD="python bench.py --frames 2 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
python - <<P
import json
d=json.loads(open('gpurun_out/bench_ours_$TAG.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
for s in d['stages']: print(f"{s['kernel']:28s} n={s['launches_per_step']:3d} {s['ms_per_launch']:.4f} ms  share {s['share']:.3f}  frac {s['frac']}")
print(open('gpurun_out/bench_ref_$TAG.json').read()[:600])
P
