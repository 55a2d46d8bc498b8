#!/bin/bash
# round 2, N-GPU call: host-copy ceiling with all ranks copying at once (unbound / NUMA-bound), the bench at N ranks with its configs[4]
# legs, the A/B of the host binding, the 200 MP frame with its per-phase breakdown, and the NCCL tests.  usage: bash tools/gpu_r2_multi.sh N
N=${1:-8}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" > gpurun_out/lscpu_n$N.txt; nproc >> gpurun_out/lscpu_n$N.txt
$TR --master-port 29511 tools/pcie_probe.py 2> gpurun_out/pcie_n${N}.err | tail -1 > gpurun_out/pcie_n${N}_unbound.json; cut -c1-600 gpurun_out/pcie_n${N}_unbound.json
$TR --master-port 29512 tools/pcie_probe.py --bind 2>> gpurun_out/pcie_n${N}.err | tail -1 > gpurun_out/pcie_n${N}_bound.json; cut -c1-600 gpurun_out/pcie_n${N}_bound.json
$TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_ours_n${N}.json 2> gpurun_out/bench_ours_n${N}.err; tail -c 300 gpurun_out/bench_ours_n${N}.err
$TR --master-port 29515 tools/bench_tiled.py --breakdown > gpurun_out/tiled_n${N}.json 2> gpurun_out/tiled_n${N}.err; cat gpurun_out/tiled_n${N}.json | cut -c1-900
if [ "$N" -ge 2 ]; then timeout 300 python -m pytest tests/test_tiled.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/pytest_nccl_n${N}.log; cat gpurun_out/pytest_nccl_n${N}.log; fi
python - <<P
import json
for f in ('bench_ours_n${N}.json',):
  try:
    d = json.loads(open('gpurun_out/' + f).read().strip().splitlines()[-1])
    print(f, 'value', d['value'], 'e2e', d['e2e'], 'bind', d.get('host_binding')); print(' extra', d.get('extra'))
  except Exception as e:
    print(f, 'unreadable', e)
P
