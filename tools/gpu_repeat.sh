mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_v11.log; cat gpurun_out/pytest_gpu_v11.log
for i in 1 2 3 4 5 6; do python -m pytest tests/test_gpu_sizes.py -m gpu -q -k "runner" 2>&1 | tail -3 | grep -E "passed|failed|frame"; done
