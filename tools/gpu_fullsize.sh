mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py -m gpu -q --durations=12 2>&1 | grep -v "^  \|^         \|^\.\.\.\|^$\|tensor(\[" | tail -60 > gpurun_out/pytest_fullsize.log; cat gpurun_out/pytest_fullsize.log
