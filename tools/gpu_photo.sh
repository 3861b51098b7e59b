mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python tools/time_photo.py 2>&1 | tail -9
timeout 200 python tools/time_photo.py 8 320 1024 3 1 2>&1 | tail -9
