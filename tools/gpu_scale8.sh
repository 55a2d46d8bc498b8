# 8-GPU box: sharded batch at N = 8 and 4, and the oversize frame split 8 / 4 / 2 ways.  usage: bash tools/gpu_scale8.sh <tag>
TAG=${1:-scale}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
  tail -c 200 gpurun_out/bench_n${N}_$TAG.err; cut -c1-160 gpurun_out/bench_n${N}_$TAG.json
done
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/bench_tiled.py --steps 3 --warmup 2 > gpurun_out/tiled_n${N}_$TAG.json 2> gpurun_out/tiled_n${N}_$TAG.err
  tail -c 200 gpurun_out/tiled_n${N}_$TAG.err; cat gpurun_out/tiled_n${N}_$TAG.json
done
python tools/bench_tiled.py --steps 3 --warmup 2 > gpurun_out/tiled_n1_$TAG.json 2> gpurun_out/tiled_n1_$TAG.err; cat gpurun_out/tiled_n1_$TAG.json
