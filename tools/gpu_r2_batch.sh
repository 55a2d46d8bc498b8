#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_noise.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/pytest_batch.log; cat gpurun_out/pytest_batch.log
python tools/bench_rcd.py default | tail -1
