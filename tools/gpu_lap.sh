mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "laplacian or local_contrast" 2>&1 | tail -4
python tools/run_laplacian.py 2>&1 | tail -3
