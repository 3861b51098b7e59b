set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json
timeout 300 python bench.py --config 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 2500 gpurun_out/bench_c3.json
timeout 120 python tools/time_sql.py > gpurun_out/time_sql_c2.log 2>&1; cat gpurun_out/time_sql_c2.log
timeout 120 python tools/time_sql.py 8 160 512 128 128 > gpurun_out/time_sql_c3.log 2>&1; cat gpurun_out/time_sql_c3.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python profiles/run_step.py 2 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'photo_fwd_kernel|photo_bwd2_kernel|sql_tc' -c 14 -o gpurun_out/full_r01c python profiles/run_step.py 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
