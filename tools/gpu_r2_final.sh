#!/bin/bash
# round 2, last call on the final tree: the bench line as the driver runs it, then the ncu launch list and the `ncu --set full`
# capture (FLOP / XU counters) of one frame's nine kernels of the same command.  usage: bash tools/gpu_r2_final.sh <tag>
TAG=${1:-final}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
python bench.py > gpurun_out/bench_ours_$TAG.json 2> gpurun_out/bench_ours_$TAG.err; tail -c 400 gpurun_out/bench_ours_$TAG.err
CMD="python bench.py --frames 2 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-graph"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 600 ncu --set full --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,sm__inst_executed_pipe_xu.sum \
  --clock-control none --import-source on -k 'regex:rcd3_kernel|rcd_strip_kernel|smooth_kernel|frame_stats_kernel|prepare_kernel|wiener32_kernel|wiener32_shared_kernel|wiener_normalize_kernel|wiener_normalize_lum4_kernel|grid_build_kernel|metrics_sliced_kernel|tonemap_kernel' -s 27 -c 9 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
python - <<P
import json
d=json.loads(open('gpurun_out/bench_ours_$TAG.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'steps', d['steps'], 'warmup', d['warmup'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'checksum', d.get('parity', {}).get('checksum_u8_sum'))
for s in d['stages']: print(f"{s['kernel']:28s} {s['ms_per_launch']:.4f} ms  share {s['share']:.3f}")
P
