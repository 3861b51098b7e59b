"""tests/golden/median_scale.npz: inputs and per-sample ratios of the median-scaling loop of the fine-tuning step
(finetune/train_ft_SQLdepth.py:236-266).  The loop is inline code of train() (no callable to import), so the fixture is
produced by executing its NumPy arithmetic -- restated line by line in oracle/sqldepth_oracle.py:median_scale_ratios --
OUTSIDE the product (nothing under sfmnext-impl_b200/ is imported here).  TEST INFRASTRUCTURE ONLY.
    python oracle/make_golden_median.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import sqldepth_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def make_case(seed, B, H, W, hole=0.35, nan_sample=None, empty_sample=None):
    g = torch.Generator().manual_seed(seed)
    depth = torch.rand(B, 1, H, W, generator=g) * 90.0
    depth[torch.rand(B, 1, H, W, generator=g) < hole] = 0.0          # sparse LiDAR ground truth
    pred = (0.3 + torch.rand(B, 1, H, W, generator=g)) * (depth + 5.0 * torch.rand(B, 1, H, W, generator=g))
    pred = (pred * 64).round() / 64                                   # quantised: ties around the median
    if nan_sample is not None:
        pred[nan_sample, 0, H // 2, W // 2] = float("nan")
        depth[nan_sample, 0, H // 2, W // 2] = 40.0
    if empty_sample is not None:
        depth[empty_sample] = 0.0
    return pred, depth


if __name__ == "__main__":
    rec = {}
    cases = [("garg", dict(garg_crop=True), make_case(1, 6, 96, 128)),
             ("eigen_kitti", dict(eigen_crop=True, dataset="kitti"), make_case(2, 4, 80, 112, nan_sample=1)),
             ("eigen_nyu", dict(eigen_crop=True, dataset="nyu"), make_case(3, 4, 480, 640, empty_sample=0))]
    for name, kw, (pred, depth) in cases:
        ratios = O.median_scale_ratios(pred, depth, 1e-3, 80.0, **kw)
        rec[name + "_pred"] = pred.numpy().astype(np.float16 if False else np.float32)
        rec[name + "_depth"] = depth.numpy()
        rec[name + "_ratio"] = ratios
        print(name, ratios)
    # keep the fixture small: store only the first two cases' tensors; the 480x640 case is regenerated from its seed
    small = {k: v for k, v in rec.items() if not k.startswith("eigen_nyu_") or k.endswith("_ratio")}
    np.savez_compressed(os.path.join(GOLD, "median_scale.npz"), **small)
    print("wrote", os.path.join(GOLD, "median_scale.npz"))
