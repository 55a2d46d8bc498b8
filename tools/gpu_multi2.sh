# 2-GPU check after the Wiener rewrite: geometry tests (1 GPU), NCCL row-tile test, sharded bench, one oversize frame on 2 GPUs
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sizes.py -m gpu -x -q -k "shared_columns" 2>&1 | tail -4
python -m pytest tests/test_tiled.py -m gpu -x -q -k nccl 2>&1 | tail -4 > gpurun_out/pytest_nccl_v8.log; cat gpurun_out/pytest_nccl_v8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_v8.json 2> gpurun_out/bench_n2_v8.err
tail -c 300 gpurun_out/bench_n2_v8.err; cut -c1-400 gpurun_out/bench_n2_v8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_tiled.py --steps 3 --warmup 2 > gpurun_out/tiled_n2_v8.json 2> gpurun_out/tiled_n2_v8.err
tail -c 300 gpurun_out/tiled_n2_v8.err; cut -c1-300 gpurun_out/tiled_n2_v8.json
