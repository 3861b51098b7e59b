mkdir -p gpurun_out
for f in 4 5; do SQLX_FWD_CFG=$f timeout 600 python -m pytest tests/test_photometric_gpu.py -m gpu -x -q 2>&1 | tail -1; done
for f in 0 4 5; do echo "fwd cfg $f"; SQLX_FWD_CFG=$f timeout 200 python tools/time_photo.py 2>&1 | grep -E "photometric|photo_fwd" ; done
for f in 0 4 5; do echo "c3 fwd cfg $f"; SQLX_FWD_CFG=$f timeout 200 python tools/time_photo.py 8 320 1024 3 1 2>&1 | grep -E "photo_fwd" ; done
