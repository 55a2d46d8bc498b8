#!/bin/bash
# round 2: full GPU test suite (pytest-xdist workers share the one GPU; failures are re-run serially) + A/B bench runs of
# environment toggles.  usage: bash tools/gpu_r2_ab.sh <tag> "<VAR=val ...>" ["<VAR=val ...>" ...]   (one bench run per toggle set)
TAG=${1:-ab}; shift
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
t0=$(date +%s)
# PYTEST_ARGS: a test selection instead of the whole suite (e.g. "tests/test_gpu_sizes.py -k white_balance"); "none" skips the tests
if [ "$PYTEST_ARGS" = none ]; then echo skipped > gpurun_out/pytest_gpu_$TAG.log; else
timeout 900 python -m pytest ${PYTEST_ARGS:-tests} -m gpu -q -n 4 > gpurun_out/pytest_gpu_$TAG.log 2>&1
fi
rc=$?
echo "xdist rc=$rc in $(( $(date +%s) - t0 )) s" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
if [ $rc -ne 0 ]; then
  grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_$TAG.log | head -20
  timeout 900 python -m pytest tests -m gpu -q --lf -x 2>&1 | tail -40 > gpurun_out/pytest_gpu_${TAG}_lf.log; cat gpurun_out/pytest_gpu_${TAG}_lf.log
fi
summary() {
python - "$1" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d.get('parity', {}).get('ok'), d.get('parity', {}).get('checksum_u8_sum'))
print('  ' + '  '.join(f"{s['kernel']}={s['ms_per_launch']:.4f}" for s in d['stages']))
P
}
python bench.py --steps 5 --warmup 3 --no-extra > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err || tail -5 gpurun_out/bench_${TAG}_default.err
summary gpurun_out/bench_${TAG}_default.json
i=0
for toggles in "$@"; do
  i=$((i + 1))
  env $toggles python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err || tail -5 gpurun_out/bench_${TAG}_$i.err
  echo "[$toggles]"; summary gpurun_out/bench_${TAG}_$i.json
done
