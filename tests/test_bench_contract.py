"""bench.py contract checks that need no GPU: the reference arm (the unmodified reference from /root/reference or
oracle/_ref, timed on the host cores) prints ONE JSON line with the keys the driver reads and reports the batch it
actually ran; the algorithmic-byte table counts each input once per launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--batch", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_frames_per_sec_hot_path" and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    # the unmodified reference where its tree is available (build container: /root/reference; GPU box: oracle/_ref)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the line reports the batch it ran and honours --steps / --warmup
    assert "workload" in d["config"] and d["config"]["batch_per_gpu"] == 2 and d["config"]["loss_scales"] == 4
    assert d["steps"] == 1 and d["warmup"] == 1
    assert "batch of the workload (2 x 192x640" in d["cpu_baseline"]["sample"]


def test_reference_arm_matches_oracle_loss():
    """The reference arm's step (unmodified Trainer methods fed by bench.py's batch builder) and the oracle restatement
    compute the same photometric loss on the same batch: the two CPU legs are the same workload."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
    import ref_shim
    if not ref_shim.available():
        import pytest
        pytest.skip("reference tree not available")
    import torch
    import bench
    from _workload import baseline_config, make_host_batch
    cfg = baseline_config(2, B=2)
    hb = make_host_batch(cfg, seed=1234)
    step, _ = bench.reference_step_factory(cfg, 2, hb, "cpu")
    torch.manual_seed(7)
    loss = float(step())
    assert 0.0 < loss < 1.0 and loss == loss


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                          "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_count_each_input_once():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "sfmnext-impl_b200"))
    import bench
    from sqlx.hotpath import HotPathConfig
    c = HotPathConfig()                                   # BASELINE config 2
    ab = bench.algorithmic_bytes(c)
    N, S, B = c.H * c.W, c.S, c.B
    ns = [c.scale_hw(s)[0] * c.scale_hw(s)[1] for s in c.scales]
    # one launch for all scales: target + S sources + S identity planes ONCE; depth, noise, arg-min per scale
    once = 12 * N + 12 * N * S + 4 * N * S
    assert ab["photo_fwd_ms_kernel"] == B * (once + sum(4 * n + 4 * N * S + N for n in ns))
    assert ab["photo_bwd_ms_kernel"] == B * (12 * N * (1 + S) + sum(8 * n + N for n in ns))
    # the padded round-1 unit (SURVEY 8d per scale x scales) is kept beside it and is larger
    assert bench.survey_8d_bytes(c)["photo_fwd_ms_kernel"] > 2 * ab["photo_fwd_ms_kernel"]
    assert ab["sql_tc_summary_kernel"] == B * 4 * c.h * c.w * c.E
    assert ab["sql_tc_bwd_pred_kernel"] == B * (8 * c.h * c.w * c.E + 4 * c.h * c.w)
