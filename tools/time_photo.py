"""Per-kernel device timing of the photometric block at a bench shape (CUDA events + libsqlx profile hooks):
  python tools/time_photo.py [B H W S scales]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import sqlx
from sqlx import _lib
from _cases import synth_photo_case

B, H, W, S, NS = [int(a) for a in sys.argv[1:6]] if len(sys.argv) >= 6 else (12, 192, 640, 2, 4)
kw = synth_photo_case(seed=5, B=B, H=H, W=W, S=S, scales=tuple(range(NS)))
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
    return out["loss"]

for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    step()
b.record(); torch.cuda.synchronize()
print("B=%d %dx%d S=%d scales=%d: photometric fwd+bwd %.1f us/step (eager)" % (B, H, W, S, NS, a.elapsed_time(b) / 20 * 1e3))
_lib.profile_enable(True)
for _ in range(10):
    step()
torch.cuda.synchronize()
for k, v in sorted(_lib.profile_report().items()):
    print("  %-28s %5.1f launches/step  %8.1f us/launch" % (k, v[0] / 10, v[1] / v[0] * 1e3))
_lib.profile_enable(False)
