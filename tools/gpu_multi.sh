# multi-GPU check: NCCL row-tile test + sharded bench under torchrun.  usage: bash tools/gpu_multi.sh <ngpus> <tag>
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_tiled.py -m gpu -x -q -k nccl 2>&1 | tail -5 > gpurun_out/pytest_nccl_$TAG.log; cat gpurun_out/pytest_nccl_$TAG.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
tail -c 300 gpurun_out/bench_n${N}_$TAG.err; cut -c1-700 gpurun_out/bench_n${N}_$TAG.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_ref_n${N}_$TAG.json 2> gpurun_out/bench_ref_n${N}_$TAG.err
tail -c 300 gpurun_out/bench_ref_n${N}_$TAG.err; cut -c1-400 gpurun_out/bench_ref_n${N}_$TAG.json
