#!/bin/bash
# round 2, call 1: the new parity tests (north_star additions, live reference at the BASELINE shapes, 200 MP bands) + config 4 stage timings
cd "$GRAFT_REPO_ROOT" || exit 1
rm -f gpurun_out/ref_live_report.jsonl
timeout 1300 python -m pytest tests/test_gpu_north_star.py tests/test_gpu_reference_live.py tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_new.log
tail -60 gpurun_out/r02_pytest_new.log
python tools/bench_stages.py --configs 4 > gpurun_out/r02_stages4_ours.jsonl 2> gpurun_out/r02_stages4_ours.err
python tools/bench_stages.py --configs 4 --impl reference > gpurun_out/r02_stages4_ref.jsonl 2> gpurun_out/r02_stages4_ref.err
cat gpurun_out/r02_stages4_ours.jsonl gpurun_out/r02_stages4_ref.jsonl | cut -c1-220
