mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_photometric_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/time_photo.py 2>&1 | grep -E "photometric|ms_"
timeout 200 python tools/time_photo.py 8 320 1024 3 1 2>&1 | grep -E "photometric|ms_"
