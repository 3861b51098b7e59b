"""A/B of candidate kernel variants that are compiled in but NOT the default (selected through environment variables
read once per process): runs the multi-scale photometric step at the bench shape in one subprocess per setting and
compares loss, gradients and per-kernel times against the default.

  python tools/check_candidates.py                       # on a B200 (gpurun)

Candidates:  SQLX_FWD_MS_STAGE=1   photo_fwd3_kernel<MS, STG=1>: depth/target staging in one trip per thread
             SQLX_FWD_MS_CFG=1|5   multi-scale forward on 16x32 tiles (256 threads, 3 CTAs/SM | 128 threads, 4 CTAs/SM)
             SQLX_BWD_MS_CFG=2|3   multi-scale backward on 32x32 tiles, 2 CTAs/SM | 16x32 tiles, 2 CTAs/SM (no register cap)
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


def run(env_extra):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, timeout=300)
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
