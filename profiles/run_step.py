"""Minimal driver for ncu: N eager hot-path steps at the bench workload (no graph, no timing).
  ncu ... python profiles/run_step.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bench  # noqa: E402
from sqlx.hotpath import HotPath, HotPathConfig  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = HotPathConfig()
torch.manual_seed(0)
hp = HotPath(cfg, use_graph=False)
hp.load(bench.make_host_batch(cfg, 1234, pin=False), non_blocking=False)
for _ in range(steps):
    hp.step_eager()
torch.cuda.synchronize()
print("loss", float(hp.loss))
