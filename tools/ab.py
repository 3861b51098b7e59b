"""A/B of library builds on one B200: every variant (`make variant NAME=.. DEFS=..` -> lib/libsqlx_NAME.so) runs the same
workloads in its own process (SQLX_LIB_PATH) and reports per-kernel times plus checksums of the results, so a variant is
only adopted when it is faster AND its results are the ones the parity suite pinned.

  python tools/ab.py [name ...]          default: every lib/libsqlx*.so;   run under gpurun
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "sfmnext-impl_b200", "lib")
CHILD = r'''
import json, os, sys
ROOT = %r
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from sqlx import _lib
from sqlx.hotpath import HotPath
from _workload import baseline_config, head_state, make_host_batch
res = {}
for tag, n, kw in (("c2", 2, {}), ("c2_noauto", 2, {"automask": False}), ("c3", 3, {}), ("c4", 4, {})):
    if os.environ.get("AB_ONLY") and tag not in os.environ["AB_ONLY"].split(","):
        continue
    cfg = baseline_config(n, **kw)
    hp = HotPath(cfg, use_graph=False)
    hp.load_state_dict(head_state(cfg), strict=True)
    hp.load(make_host_batch(cfg, 1234), non_blocking=False)
    for _ in range(3):
        hp.step_eager()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(10):
        hp.step_eager()
    torch.cuda.synchronize()
    prof = {k: v[1] / v[0] * 1e3 for k, v in _lib.profile_report().items()}
    _lib.profile_enable(False)
    g = torch.cuda.CUDAGraph
    hp2 = HotPath(cfg, use_graph=True)
    hp2.load_state_dict(head_state(cfg), strict=True)
    hp2.load(make_host_batch(cfg, 1234), non_blocking=False)
    for _ in range(3):
        hp2.step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(50):
        hp2.step()
    b.record(); torch.cuda.synchronize()
    res[tag] = {"us": prof, "step_ms": a.elapsed_time(b) / 50, "loss": float(hp.loss),
                "gsum": [float(p.grad.double().abs().sum()) for p in hp.param_list] +
                        [float(hp.slots[0][k].grad.double().abs().sum()) for k in hp.grad_inputs],
                "argmin": [int(v.long().sum()) for v in hp.argmins.values()]}
    del hp, hp2
    torch.cuda.empty_cache()
print("RESULT " + json.dumps(res))
''' % ROOT

KEYS = ["photo_fwd_ms_kernel", "photo_fwd_kernel", "photo_bwd_ms_kernel", "photo_bwd_kernel", "identity_loss_kernel",
        "sql_tc_summary_kernel", "sql_tc_pred_kernel", "sql_tc_bwd_pred_kernel", "sql_tc_bwd_sum_kernel"]


def run(lib):
    env = dict(os.environ, SQLX_LIB_PATH=lib)
    out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, timeout=600)
    for line in out.stdout.splitlines():
        if line.startswith("RESULT "):
            return json.loads(line[7:])
    raise RuntimeError(out.stderr[-3000:])


if __name__ == "__main__":
    names = sys.argv[1:]
    libs = [os.path.join(LIBDIR, "libsqlx%s.so" % ("_" + n if n != "base" else "")) for n in names] if names else \
        sorted(glob.glob(os.path.join(LIBDIR, "libsqlx*.so")))
    base = None
    for lib in libs:
        tag = os.path.basename(lib)[len("libsqlx"):-3].lstrip("_") or "base"
        try:
            r = run(lib)
        except Exception as e:
            print("%-12s FAILED: %s" % (tag, str(e)[-400:]))
            continue
        if base is None:
            base = r
        for wl, d in r.items():
            same = base is r or (abs(d["loss"] - base[wl]["loss"]) < 1e-7 and d["argmin"] == base[wl]["argmin"] and
                                 all(abs(a - b) <= 1e-5 * abs(b) + 1e-12 for a, b in zip(d["gsum"], base[wl]["gsum"])))
            print("%-12s %-10s step %.4f ms  loss %.9f  %s | %s" % (
                tag, wl, d["step_ms"], d["loss"], "same results" if same else "RESULTS DIFFER",
                "  ".join("%s %.1f" % (k.replace("_kernel", "").replace("sql_tc_", "").replace("photo_", "p_"), d["us"][k])
                          for k in KEYS if k in d["us"])))
