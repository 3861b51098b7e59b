mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches/step", d["gpu_launches_per_step"], "loss", d["loss"])
print(" kernel ms/step", d["kernel_ms_per_step"])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | tail -c 400
