"""Control flow of bench.py around the measurements, with the GPU-facing pieces mocked (tests/helpers/dry_bench.py):
one JSON line from rank 0 only, the fine-tune leg on one GPU only, and under torchrun a stalled extra workload does not
take the headline line with it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRY = os.path.join(ROOT, "tests", "helpers", "dry_bench.py")


def _run(env):
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, DRY], capture_output=True, text=True, timeout=240, env=e)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    return out.returncode, lines


def test_single_gpu_line_carries_all_workloads():
    rc, lines = _run({})
    assert rc == 0 and len(lines) == 1
    d = json.loads(lines[0])
    assert sorted(d["workloads"]) == ["config3", "config4", "config5"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["gpu_launches"] == 30 * d["steps"]


def test_torchrun_rank0_prints_other_ranks_stay_silent():
    rc0, l0 = _run({"WORLD_SIZE": "2", "RANK": "0"})
    rc1, l1 = _run({"WORLD_SIZE": "2", "RANK": "1"})
    assert rc0 == 0 and rc1 == 0 and len(l0) == 1 and l1 == []
    d = json.loads(l0[0])
    assert d["n_gpus"] == 2 and sorted(d["workloads"]) == ["config3", "config4"] and d["cpu_baseline"] is None


def test_stalled_extra_workload_keeps_the_headline():
    env = {"WORLD_SIZE": "2", "STALL": "1", "SQLX_BENCH_BAIL_AFTER": "2"}
    rc0, l0 = _run(dict(env, RANK="0"))
    rc1, l1 = _run(dict(env, RANK="1"))
    assert rc0 == 0 and rc1 == 0 and len(l0) == 1 and l1 == []
    d = json.loads(l0[0])
    assert d["value"] > 0 and "aborted" in d["workloads"]
