"""A/B of candidate kernel variants that are compiled in but NOT the default (selected through environment variables
read once per process): runs the multi-scale photometric step at the bench shape in one subprocess per setting and
compares loss, gradients and per-kernel times against the default.

  python tools/check_candidates.py                       # on a B200 (gpurun)

Candidates:  SQLX_FWD_MS_STAGE=1   photo_fwd3_kernel<MS, STG=1>: depth/target staging in one trip per thread
             SQLX_FWD_MS_CFG=1|5   multi-scale forward on 16x32 tiles (256 threads, 3 CTAs/SM | 128 threads, 4 CTAs/SM)
             SQLX_BWD_MS_CFG=2|3   multi-scale backward on 32x32 tiles, 2 CTAs/SM | 16x32 tiles, 2 CTAs/SM (no register cap)
             SQLX_SQL_PIPE=1       sql_tc_bwd_pred_kernel<DP, PIPE>, sql_tc_bwd_sum_kernel<QP, PIPE>: software-pipelined tile loop (run under `timeout`:
                                   a wrong mbarrier phase would hang)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys
ROOT = %r
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, sqlx
from sqlx import _lib
from _cases import synth_photo_case
kw = synth_photo_case(seed=5, B=12, H=192, W=640, S=2, scales=(0, 1, 2, 3))
dev = "cuda"
g = dict(kw)
g["disps"] = {s: v.to(dev).requires_grad_(True) for s, v in kw["disps"].items()}
g["target_pyr"] = {s: v.to(dev) for s, v in kw["target_pyr"].items()}
g["sources"] = [v.to(dev) for v in kw["sources"]]
g["K"], g["inv_K"] = kw["K"].to(dev), kw["inv_K"].to(dev)
g["poses"] = [{"axisangle": p["axisangle"].to(dev).requires_grad_(True), "translation": p["translation"].to(dev).requires_grad_(True),
               "invert": p["invert"]} for p in kw["poses"]]
g["noises"] = {s: v.to(dev) for s, v in kw["noises"].items()}
def step():
    for v in g["disps"].values():
        v.grad = None
    out = sqlx.photometric_losses(**g)
    out["loss"].backward()
    return out
for _ in range(3):
    out = step()
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(10):
    step()
torch.cuda.synchronize()
prof = {k: v[1] / v[0] * 1e3 for k, v in _lib.profile_report().items()}
res = {"loss": float(out["loss"]), "grads": [float(v.grad.double().abs().sum()) for v in g["disps"].values()],
       "argmin": [int(out[("argmin", s)].long().sum()) for s in kw["scales"]], "us": prof}
print("RESULT " + json.dumps(res))
''' % ROOT


SQL_CHILD = r'''
import json, os, sys
ROOT = %r
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200")):
    sys.path.insert(0, p)
import torch
from sqlx import sql as S
res = {}
for (B, h, w, Q, D) in ((12, 96, 320, 64, 64), (8, 160, 512, 128, 128), (2, 24, 40, 16, 24)):
    torch.manual_seed(0)
    x = torch.randn(B, 32, h, w, device="cuda"); q = 0.4 * torch.randn(B, Q, 32, device="cuda")
    Wp = 0.3 * torch.randn(D, Q, device="cuda"); bp = 0.1 * torch.randn(D, device="cuda")
    cen = torch.sort(torch.rand(B, D, device="cuda") * 80, dim=1).values.contiguous()
    g = torch.randn(B, 1, h, w, device="cuda")
    Mx = torch.matmul(Wp, q)
    outs = list(S.bwd_pred_mix(x, Mx, bp, cen, g))
    summ, m, l, _ = S.summary_fwd(x, q)
    ds = torch.randn_like(summ)
    outs += list(S.bwd_summary(x, q, summ, m, l, ds))                       # write mode
    acc0 = torch.randn_like(x)
    outs += list(S.bwd_summary(x, q, summ, m, l, ds, d_x=acc0.clone()))     # accumulate mode
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        S.bwd_pred_mix(x, Mx, bp, cen, g)
        S.bwd_summary(x, q, summ, m, l, ds)
    b.record(); torch.cuda.synchronize()
    res["%%dx%%dx%%d Q%%d D%%d" %% (B, h, w, Q, D)] = {"sums": [float(t.double().sum()) for t in outs],
                                                   "abs": [float(t.double().abs().sum()) for t in outs],
                                                   "us": a.elapsed_time(b) / 20 * 1e3}
print("RESULT " + json.dumps(res))
''' % ROOT


def run(env_extra, child=None):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, "-c", child or CHILD], capture_output=True, text=True, env=env, timeout=120)
    for line in out.stdout.splitlines():
        if line.startswith("RESULT "):
            return json.loads(line[7:])
    raise RuntimeError(out.stderr[-2000:])


if __name__ == "__main__":
    base = run({})
    print("default               loss %.9f  fwd %.1f us  bwd %.1f us" % (base["loss"], base["us"].get("photo_fwd_ms_kernel", 0),
                                                                       base["us"].get("photo_bwd_ms_kernel", 0)))
    for name, env in (("SQLX_FWD_MS_STAGE=1", {"SQLX_FWD_MS_STAGE": "1"}), ("SQLX_FWD_MS_CFG=1", {"SQLX_FWD_MS_CFG": "1"}),
                      ("SQLX_FWD_MS_CFG=5", {"SQLX_FWD_MS_CFG": "5"}), ("SQLX_BWD_MS_CFG=2", {"SQLX_BWD_MS_CFG": "2"}),
                      ("SQLX_BWD_MS_CFG=3", {"SQLX_BWD_MS_CFG": "3"})):
        r = run(env)
        same = (r["loss"] == base["loss"] and r["argmin"] == base["argmin"] and
                all(abs(a - b) <= 1e-6 * abs(b) for a, b in zip(r["grads"], base["grads"])))
        print("%-21s loss %.9f  fwd %.1f us  bwd %.1f us  %s" % (name, r["loss"], r["us"].get("photo_fwd_ms_kernel", 0),
                                                                 r["us"].get("photo_bwd_ms_kernel", 0),
                                                                 "bit-identical results" if same else "RESULTS DIFFER"))
    base = run({}, SQL_CHILD)
    pipe = run({"SQLX_SQL_PIPE": "1"}, SQL_CHILD)
    for k in base:
        same = base[k]["sums"] == pipe[k]["sums"] and base[k]["abs"] == pipe[k]["abs"]
        print("bwd_pred_mix + bwd_summary %-22s default %.1f us  SQLX_SQL_PIPE=1 %.1f us  %s" %
              (k, base[k]["us"], pipe[k]["us"], "bit-identical results" if same else "RESULTS DIFFER"))
