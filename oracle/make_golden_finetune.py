"""tests/golden/finetune_step.npz: the supervised fine-tuning step after the model forward
(finetune/train_ft_SQLdepth.py:232-274) on seeded inputs.  The loss is the UNMODIFIED reference class
finetune/loss.py:SILogLoss imported from /root/reference; the resize is the reference's own call (:235); the median loop
(:236-266) is inline code of train() with no callable to import, so its NumPy arithmetic is executed through the
line-by-line restatement in oracle/sqldepth_oracle.py:median_scale_ratios (itself pinned by tests/golden/median_scale.npz).
Nothing under sfmnext-impl_b200/ is imported here.  TEST INFRASTRUCTURE ONLY.
    python oracle/make_golden_finetune.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import sqldepth_oracle as O  # noqa: E402

REF = os.environ.get("SQLX_REFERENCE_SRC", "/root/reference")
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def reference_silog():
    spec = importlib.util.spec_from_file_location("ref_finetune_loss", os.path.join(REF, "finetune", "loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.SILogLoss()


if __name__ == "__main__":
    crit = reference_silog()
    rec = {}
    for name, (B, h, w, H, W), kw in [("kitti_garg", (4, 24, 80, 47, 156), dict(garg_crop=True)),
                                      ("kitti_eigen", (6, 20, 64, 40, 128), dict(eigen_crop=True, dataset="kitti"))]:
        g = torch.Generator().manual_seed(len(name) + B)
        pred = (0.5 + 20 * torch.rand(B, 1, h, w, generator=g)).requires_grad_(True)
        depth = 80 * torch.rand(B, 1, H, W, generator=g)
        depth[torch.rand(B, 1, H, W, generator=g) < 0.6] = 0.0                    # sparse ground truth
        min_depth, min_eval, max_eval = 1e-3, 1e-3, 80.0
        up = torch.nn.functional.interpolate(pred, depth.shape[-2:], mode="bilinear", align_corners=True)    # :235
        ratios = O.median_scale_ratios(up, depth, min_eval, max_eval, **kw)                                  # :236-266
        up = up * torch.as_tensor(ratios).view(-1, 1, 1, 1)                                                  # pred[i] *= ratio
        mask = depth > min_depth                                                                             # :271
        loss = crit(up, depth, mask=mask.to(torch.bool), interpolate=False)                                  # :274
        grad, = torch.autograd.grad(loss, pred)
        rec.update({name + "/pred": pred.detach().numpy(), name + "/depth": depth.numpy(), name + "/ratios": ratios,
                    name + "/loss": np.float32(loss.item()), name + "/grad": grad.numpy(),
                    name + "/garg": np.int32(bool(kw.get("garg_crop"))), name + "/eigen": np.int32(bool(kw.get("eigen_crop")))})
    np.savez_compressed(os.path.join(GOLD, "finetune_step.npz"), **rec)
    print("wrote", os.path.join(GOLD, "finetune_step.npz"), {k: v.shape for k, v in rec.items() if hasattr(v, "shape")})
