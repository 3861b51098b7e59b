mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ms_' -s 24 -c 5 -o gpurun_out/full_ms python tools/time_photo.py 12 192 640 2 4 > gpurun_out/ncu_ms.log 2>&1; tail -3 gpurun_out/ncu_ms.log
