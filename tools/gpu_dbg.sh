mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'photo_fwd3|photo_bwd3' -s 6 -c 2 -o gpurun_out/full_v3c python tools/time_photo.py 12 192 640 2 1 > gpurun_out/ncu_v3.log 2>&1; tail -3 gpurun_out/ncu_v3.log
