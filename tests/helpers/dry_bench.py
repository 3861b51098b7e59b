"""bench.main() with every GPU-facing piece mocked: exercises the control flow around the measurements (which workloads a
rank runs, which rank prints, the N > 1 watchdog) on a CPU box.  Driven by tests/test_bench_flow.py."""
import sys, os, json, time, types, threading
sys.argv = ["bench.py", "--steps", "7", "--no-extras"] + sys.argv[1:]
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, _p)
import torch
import bench
import sqlx
from sqlx import affinity

torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *_: None
torch.cuda.synchronize = lambda *_: None
affinity.bind_to_gpu = lambda *a, **k: "mock binding"
class L:  # fake lib
    def sqlx_device_ok(self, i): return 1
sqlx.lib = lambda: L()
def fake_run(cx, n, cfg, steps, warmup, full):
    if os.environ.get("STALL") and not full:
        time.sleep(10)
    if cx.rank != 0:
        return None
    return {"config": {"workload": "fake %d" % n}, "value": 100.0 * n, "ms_per_step": 1.0, "gpu_launches_per_step": 30,
            "e2e": {"value": 1.0, "unit": "frames/s", "h2d_bytes_per_step": 1, "d2h_bytes_per_step": 4}, "loss": 0.1}
bench.run_workload = fake_run
bench.run_finetune_workload = lambda cx, s, w: {"value": 5.0}
bench.time_reference = lambda *a, **k: {"value": 9.0, "unit": "frames/s", "cores": 16, "kind": "reference", "sample": "fake",
                                        "ms_per_step": 1.0, "loss": 0.0}
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    import torch.distributed as dist
    dist.init_process_group = lambda *a, **k: None
    dist.barrier = lambda *a, **k: None
    dist.destroy_process_group = lambda *a, **k: None
bench.main()
