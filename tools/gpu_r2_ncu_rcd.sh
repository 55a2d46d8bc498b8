#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/run_rcd_once.py > gpurun_out/plain_rcd.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:rcd_strip' -s 2 -c 1 -o gpurun_out/prof_rcd_strip python tools/run_rcd_once.py > gpurun_out/ncu_rcd.log 2>&1
tail -3 gpurun_out/ncu_rcd.log
