# ncu --set full of selected kernels inside a 2-frame bench run.  usage: bash tools/gpu_ncu.sh '<kernel regex>' <skip> <count> <tag>
mkdir -p gpurun_out
CMD="python bench.py --frames 2 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c $3 -o gpurun_out/prof_$4 $CMD > gpurun_out/ncu_$4.log 2>&1
tail -2 gpurun_out/ncu_$4.log
