# closing check: GPU parity suite, PCIe copy bound, headline bench (both arms)
TAG=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
python tools/pcie_probe.py > gpurun_out/pcie_$TAG.json 2> gpurun_out/pcie_$TAG.err; cat gpurun_out/pcie_$TAG.json; tail -c 300 gpurun_out/pcie_$TAG.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours_$TAG.json 2> gpurun_out/bench_ours_$TAG.err; tail -c 300 gpurun_out/bench_ours_$TAG.err
python - <<P
import json
d=json.loads(open('gpurun_out/bench_ours_$TAG.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])
for s in d['stages']: print(f"{s['kernel']:28s} n={s['launches_per_step']:3d} {s['ms_per_launch']:.4f} ms  share {s['share']:.3f}  frac {s['frac']}")
P
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
