#!/bin/bash
# the driver's scaling step for one N: both bench arms under torch.distributed.run.  usage: bash tools/gpu_r2_scale.sh N
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --impl reference --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_ref_n${N}.json 2> gpurun_out/bench_ref_n${N}.err
$TR --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_ours_n${N}.json 2> gpurun_out/bench_ours_n${N}.err; tail -c 300 gpurun_out/bench_ours_n${N}.err
python - <<P
import json
for f in ('bench_ours_n${N}.json', 'bench_ref_n${N}.json'):
  try:
    d = json.loads(open('gpurun_out/' + f).read().strip().splitlines()[-1])
    print(f, 'value', d['value'], 'n_gpus', d['n_gpus'], 'e2e', d['e2e']); print(' extra', d.get('extra'))
  except Exception as e:
    print(f, 'unreadable', e)
P
