mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sql_gpu.py tests/test_sql_tc_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches/step", d["gpu_launches_per_step"])
for k,v in d["kernels"].items(): print("   %-28s %5.1f x %8.1f us"%(k, v["launches_per_step"], v["ms_per_step"]/v["launches_per_step"]*1e3))
print(" kernel ms/step", d["kernel_ms_per_step"])
PY
