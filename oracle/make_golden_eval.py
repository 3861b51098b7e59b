"""Golden vector for the flip-TTA disparity blend (N3): runs the reference's own
evaluate_depth_config.batch_post_process_disparity (evaluate_depth_config.py:51-59) through oracle/ref_shim.py.
  python oracle/make_golden_eval.py   ->  tests/golden/eval_postprocess.npz"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.load()
ev = importlib.import_module("evaluate_depth_config")
rng = np.random.RandomState(11)
rec = {}
for i, (n, h, w) in enumerate([(2, 24, 40), (1, 17, 33), (3, 96, 320)]):
    l = (0.01 + rng.rand(n, h, w)).astype(np.float32)
    r = (0.01 + rng.rand(n, h, w)).astype(np.float32)
    rec["l%d" % i], rec["r%d" % i] = l, r
    rec["out%d" % i] = ev.batch_post_process_disparity(l, r)
np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "eval_postprocess.npz"), **rec)
print({k: v.shape for k, v in rec.items()})
