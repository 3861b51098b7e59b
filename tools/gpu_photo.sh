mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_photometric_gpu.py -m gpu -x -q 2>&1 | tail -3
for f in 0 1 2 3; do echo "fwd cfg $f"; SQLX_FWD_CFG=$f timeout 200 python tools/time_photo.py 2>&1 | grep -E "photometric|photo_fwd" ; done
for f in 0 1; do echo "c3 cfg $f"; SQLX_FWD_CFG=$f timeout 200 python tools/time_photo.py 8 320 1024 3 1 2>&1 | grep -E "photometric|photo_fwd|photo_bwd" ; done
