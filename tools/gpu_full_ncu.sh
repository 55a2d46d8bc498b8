# round-end style check + ncu --set full of one frame's nine kernels.  usage: bash tools/gpu_full_ncu.sh <tag>
TAG=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours_$TAG.json 2> gpurun_out/bench_ours_$TAG.err; tail -c 600 gpurun_out/bench_ours_$TAG.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 600 gpurun_out/bench_ref_$TAG.err
CMD="python bench.py --frames 2 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:rcd3_kernel|smooth_kernel|frame_stats_kernel|prepare_kernel|wiener32_kernel|wiener32_shared_kernel|wiener_normalize_kernel|grid_build_kernel|metrics_sliced_kernel|tonemap_kernel' -s 27 -c 9 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
python - <<P
import json
d=json.loads(open('gpurun_out/bench_ours_$TAG.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
for s in d['stages']: print(f"{s['kernel']:28s} n={s['launches_per_step']:3d} {s['ms_per_launch']:.4f} ms  share {s['share']:.3f}  frac {s['frac']}")
print(open('gpurun_out/bench_ref_$TAG.json').read()[:600])
P
