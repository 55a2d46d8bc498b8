# JPEG: streams of the reference (golden fixture) + our GPU tests against them
mkdir -p gpurun_out/golden
python tests/golden/make_golden.py gpurun_out/golden jpeg 2>&1 | tail -5
cp gpurun_out/golden/jpeg.npz tests/golden/jpeg.npz
python -m pytest tests/test_jpeg.py -q -m gpu 2>&1 | tail -15
