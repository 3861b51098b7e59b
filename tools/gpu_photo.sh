set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_photometric_gpu.py tests/test_layers_gpu.py tests/test_abi.py -m gpu -x -q > gpurun_out/pytest_photo.log 2>&1; tail -15 gpurun_out/pytest_photo.log
for f in 0 1 2 3; do SQLX_FWD_CFG=$f timeout 200 python tools/time_photo.py 2>&1 | grep -E "photometric|photo_fwd" ; done
for f in 0 1 2; do SQLX_BWD_CFG=$f timeout 200 python tools/time_photo.py 2>&1 | grep -E "photo_bwd" ; done
for f in 0 1 3; do SQLX_FWD_CFG=$f timeout 200 python tools/time_photo.py 8 320 1024 3 1 2>&1 | grep -E "photometric|photo_fwd|photo_bwd" ; done
