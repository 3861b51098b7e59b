timeout 600 python -m pytest tests/test_photometric_gpu.py -m gpu -x -q 2>&1 | tail -2
for f in 0 3; do echo "bwd cfg $f"; SQLX_BWD_CFG=$f timeout 200 python tools/time_photo.py 2>&1 | grep -E "photometric|photo_bwd" ; done
SQLX_BWD_CFG=0 timeout 200 python tools/time_photo.py 8 320 1024 3 1 2>&1 | grep -E "photo_bwd"
SQLX_BWD_CFG=3 timeout 200 python tools/time_photo.py 8 320 1024 3 1 2>&1 | grep -E "photo_bwd"
