mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 400 gpurun_out/bench_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
for k,v in d["kernels"].items(): print("   %-28s %5.1f x %8.1f us"%(k, v["launches_per_step"], v["ms_per_step"]/v["launches_per_step"]*1e3))
print(" kernel ms/step", d["kernel_ms_per_step"])
PY
