"""Locate the largest gradient deviation from a golden fixture: python tools/debug_photo_grad.py [fixture]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import sqlx
import test_photometric_gpu as T
name = sys.argv[1] if len(sys.argv) > 1 else "photo_mono_s0"
kw, leaves, z, fids = T.photo_case(name)
g = T._to_dev(kw)
out = sqlx.photometric_losses(**g, materialize=True)
print("loss", float(out["loss"]), float(z["out_loss"]))
gl = T._grad_leaves(g)
grads = torch.autograd.grad(out["loss"], gl, allow_unused=True)
ref = torch.from_numpy(z["grad_disp0"]); gr = grads[0].cpu()
d = (gr - ref).abs()
print("shape", tuple(ref.shape), "max|ref|", float(ref.abs().max()), "max diff", float(d.max()))
flat = torch.topk(d.flatten(), 12)
for v, i in zip(flat.values, flat.indices):
    idx = np.unravel_index(int(i), ref.shape)
    print("  diff %.3e at %s ours %.4e ref %.4e" % (float(v), idx, float(gr[idx]), float(ref[idx])))
sel = out["identity_selection/0"].cpu().numpy().astype(np.uint8)
bad = np.argwhere(sel != z["out_idsel_s0"])
print("selection mismatches", len(bad), bad[:10].tolist())
