# Full round artefacts on one B200: GPU parity tests, smoke, the bench line (config 2 headline + configs 3, 4 inside it),
# the reference arm, per-op timings, the ncu launch list and one ncu --set full capture of the main kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG 2>&1 | tail -40'
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 600 gpurun_out/${TAG}_bench_ref.json
for c in 2 3 4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c$c.csv python profiles/run_step.py 2 $c > gpurun_out/${TAG}_ncu_launch_c$c.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'photo_|identity3|sql_|ms_|head_|pack_|sum_partials' -c 40 -o gpurun_out/${TAG}_full_c2 python profiles/run_step.py 1 2 > gpurun_out/${TAG}_ncu_full_c2.log 2>&1
ls -la gpurun_out | tail -20
