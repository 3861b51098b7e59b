# A/B of the compiled-in candidate kernels (see tools/check_candidates.py) on one B200, every step under a timeout:
#   gpurun --timeout 1500 -- 'bash tools/candidates_round.sh 2>&1 | tail -60'
# 1. bit-identity + per-kernel times at the bench shape, 2. for the variants worth it, the GPU parity suite and a bench
# line with the variant switched on.  Nothing here changes the defaults.
set -x
mkdir -p gpurun_out
timeout 900 python tools/check_candidates.py > gpurun_out/candidates.log 2>&1; echo "check_candidates rc=$?"; tail -20 gpurun_out/candidates.log
for v in "SQLX_SQL_PIPE=1" "SQLX_FWD_MS_STAGE=1"; do
  tag=$(echo $v | tr '=' '_')
  env $v timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "$v pytest rc=$?"; tail -2 gpurun_out/pytest_$tag.log
  env $v timeout 200 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("$v", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]))
except Exception as e:
    print("$v: no bench line", e)
PY
done
