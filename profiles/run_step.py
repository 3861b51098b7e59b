"""Minimal driver for ncu: N eager hot-path steps at a BASELINE workload (no graph, no timing).
  ncu ... python profiles/run_step.py [steps] [config = 2|3|4] [batch override]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from _workload import baseline_config, head_state, make_host_batch  # noqa: E402
from sqlx.hotpath import HotPath  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = baseline_config(n, B=int(sys.argv[3]) if len(sys.argv) > 3 else None)
hp = HotPath(cfg, use_graph=False)
hp.load_state_dict(head_state(cfg), strict=True)
hp.load(make_host_batch(cfg, 1234, pin=False), non_blocking=False)
for _ in range(steps):
    hp.step_eager()
torch.cuda.synchronize()
print("config", n, "loss", float(hp.loss))
