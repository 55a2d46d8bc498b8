#!/bin/bash
# round 2: full GPU test suite, both bench arms, ncu launch list + ncu --set full (with FLOP / XU counters) of one frame's kernels.
# usage: bash tools/gpu_r2_profile.sh <tag> [skip-tests]
TAG=${1:-r02}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -n 4 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
fi
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours_$TAG.json 2> gpurun_out/bench_ours_$TAG.err; tail -c 800 gpurun_out/bench_ours_$TAG.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 600 gpurun_out/bench_ref_$TAG.err
CMD="python bench.py --frames 2 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-graph"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,sm__inst_executed_pipe_xu.sum \
  --clock-control none --import-source on -k 'regex:rcd3_kernel|rcd_strip_kernel|smooth_kernel|frame_stats_kernel|prepare_kernel|wiener32_kernel|wiener32_shared_kernel|wiener_normalize_kernel|wiener_normalize_lum4_kernel|grid_build_kernel|metrics_sliced_kernel|tonemap_kernel' -s 27 -c 9 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
python tools/bench_batch.py > gpurun_out/batch_graph_$TAG.jsonl 2> gpurun_out/batch_graph_$TAG.err
python tools/bench_stages.py --kernels > gpurun_out/stages_ours_$TAG.jsonl 2> gpurun_out/stages_ours_$TAG.err
python - <<P
import json
d=json.loads(open('gpurun_out/bench_ours_$TAG.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']); print('peaks', d.get('pipe_peaks_measured')); print('parity', d.get('parity')); print('extra', d.get('extra')); print('cpu', d.get('cpu_baseline')); print('bind', d.get('host_binding'))
for s in d['stages']: print(f"{s['kernel']:28s} n={s['launches_per_step']:3d} {s['ms_per_launch']:.4f} ms  share {s['share']:.3f}  hbm frac {s['frac']}  bound {s.get('bound')} {s.get('frac_of_bound')}")
print(open('gpurun_out/bench_ref_$TAG.json').read()[:900])
P
