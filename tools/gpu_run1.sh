set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
./tools/red_bench > gpurun_out/red_bench.log 2>&1; cat gpurun_out/red_bench.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 600 gpurun_out/bench_ours.err
CMD="python bench.py --frames 2 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:wiener_tile|rcd_kernel|smooth_kernel|tonemap_kernel|splat_kernel|slice_kernel|blur_kernel|wiener_normalize|green_eq_kernel|loglum' -s 20 -c 10 -o gpurun_out/prof_r01b $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out
