#!/bin/bash
# RCD strip kernel: parity (every test that touches a demosaic) + A/B timing against the tile kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_sizes.py tests/test_gpu_north_star.py tests/test_gpu_fullsize.py tests/test_gpu_reference_live.py -m gpu -x -q -p no:cacheprovider -k "rcd or demosaic or pipeline or RCD" 2>&1 | tail -25 > gpurun_out/pytest_rcd.log; cat gpurun_out/pytest_rcd.log
for cfg in "TDB_RCD_STRIPS=0" "TDB_RCD_STRIPS=1" "TDB_RCD_STRIPS=1 TDB_RCD_SEGMENTS=6" "TDB_RCD_STRIPS=1 TDB_RCD_SEGMENTS=14"; do
  env $cfg python tools/bench_rcd.py "$cfg" 2>&1 | tail -1 | tee -a gpurun_out/bench_rcd.jsonl
done
