#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_sizes.py tests/test_gpu_north_star.py tests/test_gpu_fullsize.py tests/test_gpu_reference_live.py -m gpu -x -q -p no:cacheprovider -k "laplacian or local_contrast or ppg or bilinear or demosaic or rcd" 2>&1 | tail -25 > gpurun_out/pytest_lap.log; cat gpurun_out/pytest_lap.log
python tools/bench_stages.py --configs 4,2 --kernels > gpurun_out/r02_stages_b.jsonl 2> gpurun_out/r02_stages_b.err
grep -E "Laplacian|laplacian|from packed|RCD|PPG|bilinear" gpurun_out/r02_stages_b.jsonl | cut -c1-330
cat gpurun_out/ref_live_report.jsonl | grep lap
