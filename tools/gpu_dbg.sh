mkdir -p gpurun_out
timeout 200 python tools/debug_photo_grad.py > gpurun_out/debug_grad.log 2>&1; cat gpurun_out/debug_grad.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'photo_fwd3|photo_bwd3' -s 8 -c 2 -o gpurun_out/full_v3 python tools/time_photo.py 12 192 640 2 1 > gpurun_out/ncu_v3.log 2>&1; tail -3 gpurun_out/ncu_v3.log
