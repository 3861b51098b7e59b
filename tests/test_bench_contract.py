"""bench.py contract checks that need no GPU: the reference arm (the oracle port timed on the host cores) prints ONE
JSON line with the keys the driver reads, and the algorithmic-byte table follows SURVEY 8d's per-scale formulas."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-sample-batch", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_frames_per_sec_hot_path" and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["batch_per_gpu"] == 12 and d["config"]["loss_scales"] == 4


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                          "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_follow_survey_8d():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "sfmnext-impl_b200"))
    import bench
    from sqlx.hotpath import HotPathConfig
    c = HotPathConfig()                                   # BASELINE config 2
    ab = bench.algorithmic_bytes(c)
    N, S, B = c.H * c.W, c.S, c.B
    ns = [c.scale_hw(s)[0] * c.scale_hw(s)[1] for s in c.scales]
    # SURVEY 8d, photometric forward per scale: 4n + 12N + 12NS + 4NS (noise) [+ 4NS identity planes, + N arg-min]
    per_scale = [4 * n + 12 * N + 12 * N * S + 4 * N * S + 4 * N * S + N for n in ns]
    assert ab["photo_fwd_ms_kernel"] == B * sum(per_scale)
    assert abs(ab["photo_fwd_kernel"] - B * sum(per_scale) / len(ns)) < 1e-6 * ab["photo_fwd_kernel"]
    assert ab["sql_tc_summary_kernel"] == B * 4 * c.h * c.w * c.E
    assert ab["sql_tc_bwd_pred_kernel"] == B * (8 * c.h * c.w * c.E + 4 * c.h * c.w)
