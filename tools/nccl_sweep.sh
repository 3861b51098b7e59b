# A/B of the in-step gradient exchange at N GPUs (one gpurun --gpus N call): SMs left to the NCCL kernel by the
# summary-path backward, and NCCL channel caps.
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/nccl_sweep.sh 8 "0 8 16 32"'
N=${1:-8}
SWEEP=${2:-"0 8 16 32"}
FORK=${3:-auto}
mkdir -p gpurun_out
run() {
  tag=$1; sms=$2; shift; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 100 --warmup 10 --no-extras --no-parity --no-cpu-baseline --workloads "" --exchange-sms $sms --prepare-fork $FORK \
    > gpurun_out/nccl_${tag}.json 2> gpurun_out/nccl_${tag}.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/nccl_%s.json" % tag) if l.startswith("{")][-1])
    print("%-28s %9.0f frames/s  %.4f ms/step  exposed comm %.1f us  e2e %.0f" % (tag, d["value"], d["ms_per_step"], 1e3 * (d.get("exposed_comm_ms") or 0), d["e2e"]["value"]))
except Exception as e:
    print(tag, "failed", e)
PY
}
for s in $SWEEP; do run n${N}_reserve${s}_${FORK} $s NCCL_DEBUG=WARN; done
