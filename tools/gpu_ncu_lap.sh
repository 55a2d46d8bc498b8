# ncu --set full of one Laplacian call at 8192 x 6144 (22 launches).  usage: bash tools/gpu_ncu_lap.sh <tag>
mkdir -p gpurun_out
CMD="python tools/run_laplacian.py"
$CMD > gpurun_out/plain_lap_$1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:reduce1_kernel|reduce_kernel|assemble_kernel" -s 44 -c 22 -o gpurun_out/prof_lap_$1 $CMD > gpurun_out/ncu_lap_$1.log 2>&1
tail -2 gpurun_out/ncu_lap_$1.log
