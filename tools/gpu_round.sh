# Full round artefacts: GPU parity tests, bench lines (config 2 = the metric's configuration, config 3 informational),
# per-op timings, the ncu launch list and one ncu --set full capture of the main kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1500 gpurun_out/bench_c2.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
timeout 300 python bench.py --config 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.json
timeout 120 python tools/time_sql.py > gpurun_out/time_sql_c2.log 2>&1
timeout 120 python tools/time_sql.py 8 160 512 128 128 > gpurun_out/time_sql_c3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python profiles/run_step.py 2 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'photo_fwd3|photo_bwd3|identity3|sql_tc|ms_|head_' -c 28 -o gpurun_out/full_round python profiles/run_step.py 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
