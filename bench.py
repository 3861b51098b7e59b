#!/usr/bin/env python
"""Benchmark of the SQLdepth training hot path (BASELINE.json metric: train frames/s at 192x640 and 320x1024).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (libsqlx)
  python bench.py --impl reference [...]                         the UNMODIFIED reference's CPU path (oracle/_ref)

One "step" = SQL decoder tail forward -> photometric losses forward -> backward of both, on one batch of synthetic
KITTI-shape inputs.  frames/s = target frames processed per second = batch / step time (trainer.py:584).
The headline workload is BASELINE config 2 (batch 12 per GPU, 192x640, decoder features 32x96x320, Q = D = 64, 4 loss
scales); the same JSON line carries configs 3 and 4 (320x1024, Q = D = 128) under "workloads".  Prints ONE JSON line
(rank 0).  See DESIGN.md "Measurement" for the byte accounting behind `roofline`.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "train_frames_per_sec_hot_path"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="sqlx", choices=["sqlx", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4],
                    help="headline workload = BASELINE config (default 2: the configuration the metric is quoted on "
                         "that fits one GPU step for step with the reference's CPU-runnable case)")
    ap.add_argument("--workloads", default="3,4,5",
                    help="further BASELINE configs measured into the same JSON line under 'workloads' ('' = none)")
    ap.add_argument("--batch", type=int, default=0, help="override the batch per GPU of the headline workload")
    ap.add_argument("--no-graph", action="store_true", help="submit the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--f32-frames", action="store_true",
                    help="ship the frames as float32 in the end-to-end loop (default: uint8, converted on the device)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the fp64 oracle evaluation of the timed batch")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the informational legs (reprojection-dominant batch, h2d-only, stock-PyTorch-on-GPU)")
    ap.add_argument("--allreduce-after-step", action="store_true",
                    help="N > 1: issue the gradient all-reduce after the step instead of inside it (A/B switch)")
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: where the unmodified reference runs (cpu = the contract's arm)")
    ap.add_argument("--exchange-sms", type=int, default=-1,
                    help="N > 1: SMs the summary-path backward leaves to the NCCL kernel of the in-step exchange "
                         "(-1 = default 32)")
    ap.add_argument("--prepare-fork", default="auto", choices=["auto", "start", "after_pred_fwd", "after_bwd_pred", "off"],
                    help="where the step forks the next input set's frame-only work (HotPath.prepare_fork; off = computed "
                         "at the start of the step it belongs to, as in round 1)")
    ap.add_argument("--prefetch", type=int, default=2, help="batches in flight ahead of the step in the e2e loop")
    return ap.parse_args()


def workload_text(n, cfg):
    return ("BASELINE config %d hot path: SQL decoder tail (x0 32x%dx%d, Q=%d, D=%d) + photometric loss (%dx%d, %d source "
            "frames%s, %d loss scale%s: scale 0 = decoder output, coarser scales synthetic since the reference decoder "
            "emits scale 0 only), forward+backward"
            % (n, cfg.h, cfg.w, cfg.Q, cfg.D, cfg.H, cfg.W, cfg.S, " incl. the stereo frame" if cfg.stereo else "",
               len(cfg.scales), "s" if len(cfg.scales) > 1 else ""))


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        # the median UNDER LOAD: samples taken while the SM clock was above idle (the sampler also sees setup phases)
        busy = sorted(x for x in sm if x > 0.5 * max(mx or [0]))
        allv = sorted(sm)
        med = (busy or allv)
        return {"sm_mhz": med[len(med) // 2] if med else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_under_load": len(busy), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- reference arm
def reference_step_factory(cfg, cfg_id, hb, device="cpu"):
    """The UNMODIFIED reference on this workload, through its own public API (oracle/ref_shim.py imports it from
    /root/reference, or from oracle/_ref -- the copy oracle/build_ref.py makes -- on the GPU box):
        networks.{Lite_,}Depth_Decoder_QueryTr.forward   (networks/depth_decoder_QTR.py:36-74)
        Trainer.generate_images_pred + compute_losses     (trainer.py:386-549), then loss.backward().
    The decoder's forward includes its patch-embedding conv, 4-layer transformer and conv3x3 (lines 37-45), which are not
    on our path: the reference arm does slightly MORE work than ours per step (reported as `front_ms`)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    ref = ref_shim.load(force_cpu=(device == "cpu"))
    c = cfg
    dev = torch.device(device)
    T = ref_shim.make_trainer(c.B, c.H, c.W, scales=c.scales, use_stereo=c.stereo, device=device)
    patch = {2: 16, 3: 20, 4: 32}[cfg_id]
    cls = ref.networks.Depth_Decoder_QueryTr if cfg_id == 4 else ref.networks.Lite_Depth_Decoder_QueryTr
    torch.manual_seed(0)
    dec = cls(in_channels=c.E, patch_size=patch, dim_out=c.D, embedding_dim=c.E, query_nums=c.Q, num_heads=4,
              min_val=c.min_depth, max_val=c.max_depth).to(dev)
    dec.train()
    t = {k: (v.float() / 255.0 if v.dtype == torch.uint8 else v).to(dev) for k, v in hb.items()}
    fids = [-1, 1] + (["s"] if c.stereo else [])
    inputs = {("K", 0): t["K"], ("inv_K", 0): t["inv_K"], ("color", 0, 0): t["target"]}
    for i, f in enumerate(fids):
        inputs[("color", f, 0)] = t["source%d" % i]
    for s in c.scales:
        if s > 0:
            inputs[("color", 0, s)] = t["target%d" % s]
    if c.stereo:
        inputs["stereo_T"] = t["stereo_T"]
    x0 = t["x"].clone().requires_grad_(True)
    leaves = {("disp", s): t["disp%d" % s].clone().requires_grad_(True) for s in c.scales if s > 0}
    poses = {}
    for i in c.pose_sources:
        poses[fids[i]] = (t["axisangle%d" % i].clone().requires_grad_(True),
                          t["translation%d" % i].clone().requires_grad_(True))
    params = list(dec.parameters())

    def step():
        for p_ in params + [x0] + list(leaves.values()) + [q for pr in poses.values() for q in pr]:
            p_.grad = None
        outputs = dec(x0)
        outputs.update(leaves)
        for f, (aa, tr) in poses.items():           # what Trainer.predict_poses leaves in `outputs` (trainer.py:330-337)
            outputs[("axisangle", 0, f)] = aa
            outputs[("translation", 0, f)] = tr
            outputs[("cam_T_cam", 0, f)] = ref.layers.transformation_from_parameters(aa[:, 0], tr[:, 0], invert=(f < 0))
        T.generate_images_pred(inputs, outputs)
        losses = T.compute_losses(inputs, outputs)
        losses["loss"].backward()
        return losses["loss"]

    def front():
        """forward + backward of depth_decoder_QTR.py:37-45 alone (the part of decoder.forward outside our path)"""
        xx = x0.detach().clone().requires_grad_(True)
        emb = dec.embedding_convPxP(xx).flatten(2)
        emb = emb + dec.positional_encodings[:emb.shape[2], :].T.unsqueeze(0)
        tok = dec.transformer_encoder(emb.permute(2, 0, 1))
        y = dec.conv3x3(xx)
        (tok.sum() + y.sum()).backward()
    return step, front


def port_step_factory(cfg, hb, state):
    """Fallback when the reference tree is not available (oracle/_ref missing): the oracle port, same ATen primitives."""
    from _workload import oracle_step
    return lambda: oracle_step(cfg, hb, state, dtype=torch.float32)["loss"]


def time_reference(cfg, cfg_id, steps, warmup, device="cpu"):
    from _workload import make_host_batch, head_state
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)
    hb = make_host_batch(cfg, seed=1234, pin=False)
    front_ms = None
    if ref_shim.available():
        kind = "reference"
        step, front = reference_step_factory(cfg, cfg_id, hb, device)
    else:
        kind = "port"
        step, front = port_step_factory(cfg, hb, head_state(cfg)), None
    sync = torch.cuda.synchronize if device == "cuda" else (lambda: None)
    for _ in range(warmup):
        step()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    sync()
    dt = (time.perf_counter() - t0) / steps
    if front is not None:
        front(); sync()
        t1 = time.perf_counter()
        front(); sync()
        front_ms = (time.perf_counter() - t1) * 1e3
    what = ("unmodified reference (oracle/_ref): Lite_/Depth_Decoder_QueryTr.forward + Trainer.generate_images_pred + "
            "compute_losses + backward" if kind == "reference" else "oracle port (reference tree not available)")
    return {"value": cfg.B / dt, "unit": UNIT, "cores": cores, "kind": kind, "ms_per_step": dt * 1e3,
            "front_ms": front_ms, "loss": float(loss),
            "sample": "%s; the FULL batch of the workload (%d x %dx%d, %d loss scales), %d timed steps after %d warm-up, "
                      "torch %s fp32 threads=%d%s"
                      % (what, cfg.B, cfg.H, cfg.W, len(cfg.scales), steps, warmup, device, cores,
                         ("; of which %.0f ms/step is the decoder front (patch embedding, transformer, conv3x3) that is "
                          "not on our path" % front_ms) if front_ms is not None else "")}


# --------------------------------------------------------------------------------------------- roofline
def algorithmic_bytes(c):
    """Compulsory fp32 bytes of ONE LAUNCH of every main kernel, counting each input ONCE per launch (DESIGN.md,
    "Measurement"): frames, identity losses and intrinsics are scale-invariant; depth, noise and arg-min are per scale."""
    N, S, B, E = c.H * c.W, c.S, c.B, c.E
    n0 = c.h * c.w
    ns = [c.scale_hw(s)[0] * c.scale_hw(s)[1] for s in c.scales]
    L = len(ns)
    frames = 12 * N * (1 + S)                      # target + S sources, read once
    fwd = B * (frames + 4 * N * S + sum(4 * n + 4 * N * S + N for n in ns))       # + identity; depth, noise, arg-min
    bwd = B * (frames + sum(4 * n + N + 4 * n for n in ns))                       # depth + arg-min in, d_depth out
    return {
        "photo_fwd_ms_kernel": fwd, "photo_fwd_kernel": fwd / L if L else fwd,
        "photo_bwd_ms_kernel": bwd, "photo_bwd_kernel": bwd / L if L else bwd,
        # identity losses of all S sources in one launch: target once, every source once, S loss maps out
        "identity_loss_kernel": B * (12 * N + 12 * N * S + 4 * N * S),
        "sql_tc_summary_kernel": B * 4 * n0 * E,
        "sql_tc_pred_kernel": B * (4 * n0 * E + 4 * n0),
        # regression backward reads x + g_pred and writes d_x; the summary-path backward reads x and accumulates into d_x
        "sql_tc_bwd_pred_kernel": B * (4 * n0 * E + 4 * n0 + 4 * n0 * E),
        "sql_tc_bwd_sum_kernel": B * (4 * n0 * E + 2 * 4 * n0 * E),
        "sql_summary_kernel": B * 4 * n0 * E, "sql_pred_kernel": B * (4 * n0 * E + 4 * n0),
        "sql_bwd_reduce_kernel": B * (4 * n0 * E + 4 * n0), "sql_bwd_dx_kernel": B * (8 * n0 * E + 4 * n0),
    }


def survey_8d_bytes(c):
    """SURVEY 8d's per-scale unit (every scale re-reads the frames) x the scales one launch processes: what round 1
    reported; kept beside the unique-byte figure for comparison."""
    N, S, B = c.H * c.W, c.S, c.B
    ns = [c.scale_hw(s)[0] * c.scale_hw(s)[1] for s in c.scales]
    return {"photo_fwd_ms_kernel": B * sum(4 * n + 12 * N + 12 * N * S + 4 * N * S for n in ns),
            "photo_bwd_ms_kernel": B * sum(4 * n + 12 * N + 12 * N * S + 4 * N * S + 4 * n for n in ns)}


def algorithmic_flops(c):
    """Algorithmic (1x, not 3xTF32) tensor FLOPs per launch of the SQL kernels."""
    n0, E, Q, D, B = c.h * c.w, c.E, c.Q, c.D, c.B
    return {"sql_tc_summary_kernel": B * 4.0 * n0 * E * Q,          # y = x^T K and S = P x^T
            "sql_tc_pred_kernel": B * 2.0 * n0 * E * D,             # z = M x (mixed weights: contraction over E)
            "sql_tc_bwd_pred_kernel": B * 6.0 * n0 * E * D,         # z recompute, dM = dz^T x, d_x = dz M
            "sql_tc_bwd_sum_kernel": B * 10.0 * n0 * E * Q}         # y, t = x^T ds^T, d_x = dy K + a ds, dK = dy^T x


def selection_stats(hp, cfg):
    """Arg-min histogram of the timed batch per scale (identity / per-source shares) and the share of (tile, source)
    pairs the backward kernel skips because no pixel of the tile's halo selected that source (photo_v3.cu)."""
    import torch.nn.functional as F
    S = cfg.S
    n_ident = S if cfg.automask else 0
    out = {"per_scale": {}, "backward_tile_skip_rate": None}
    skipped, total = 0, 0
    ident_tot, n_tot = 0.0, 0
    for s, am in hp.argmins.items():
        hist = torch.bincount(am.flatten().long(), minlength=n_ident + S).float()
        hist = (hist / hist.sum()).tolist()
        out["per_scale"][str(s)] = {"identity": sum(hist[:n_ident]), "reprojection": hist[n_ident:n_ident + S]}
        ident_tot += sum(hist[:n_ident]); n_tot += 1
        for k in range(S):
            sel = (am == n_ident + k).float()[:, None]
            tiles = F.max_pool2d(sel, kernel_size=(22, 38), stride=(16, 32), padding=(3, 3), ceil_mode=True)
            skipped += int((tiles == 0).sum()); total += tiles.numel()
    out["identity_share"] = ident_tot / max(1, n_tot)
    out["backward_tile_skip_rate"] = skipped / max(1, total)
    return out


def load_traffic():
    """{configN: {profile-hook kernel name: DRAM bytes per launch}} from the committed ncu --set full summary
    (profiles/traffic_latest.json, written by tools/ncu_summary.py)."""
    out = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
    except Exception:
        return out
    by_cfg = tj.get("by_config") or {"config2": tj.get("kernels", {})}
    for ck, kern in by_cfg.items():
        d = {}
        for k, v in kern.items():              # ncu names -> the names of the library's profile hooks
            full = k.replace("void ", "")
            base = full.split("<")[0].split("::")[-1]
            ms = full.rstrip().endswith(", 1>")      # last template argument of the photometric kernels: all scales per launch
            base = {"photo_fwd3_kernel": "photo_fwd_ms_kernel" if ms else "photo_fwd_kernel",
                    "photo_bwd3_kernel": "photo_bwd_ms_kernel" if ms else "photo_bwd_kernel",
                    "identity3_kernel": "identity_loss_kernel",
                    "sql_tc_pred2_kernel": "sql_tc_pred_kernel", "sql_ws_pred_kernel": "sql_tc_pred_kernel",
                    "sql_ws_summary_kernel": "sql_tc_summary_kernel", "sql_ws_bwd_pred_kernel": "sql_tc_bwd_pred_kernel",
                    "sql_ws_bwd_sum_kernel": "sql_tc_bwd_sum_kernel"}.get(base, base)
            d[base] = v["dram_bytes_per_launch"]
        out[ck] = d
    return out


# --------------------------------------------------------------------------------------------- the fine-tune workload
def run_finetune_workload(cx, steps, warmup):
    """BASELINE config 5 (ConvNeXt-L 320x1024 metric-depth fine-tune, SURVEY 8 table: B = 8, x0 32x160x512, Q = D = 64):
    the supervised step of finetune/train_ft_SQLdepth.py:232-278 behind the model's backbone -- SQL decoder tail ->
    align_corners=True resize to the ground truth -> per-sample median scaling -> SILog -> backward -- as one CUDA graph.
    value: inputs resident; e2e: decoder features, queries and the ground-truth depth map copied from pinned host memory
    every step (copy of batch i+1 under the step on batch i) + loss read-back."""
    import sqlx
    from sqlx import sql as S
    from sqlx.hotpath import HotPath, HotPathConfig
    from _workload import head_state
    dev, world, rank, dist = cx.dev, cx.world, cx.rank, cx.dist
    B, H, W, h, w, Q, D, E = 8, 320, 1024, 160, 512, 64, 64, 32
    min_depth, max_depth = 1e-3, 80.0
    cfg = HotPathConfig(B=B, H=H, W=W, h=h, w=w, E=E, Q=Q, D=D, S=2, scales=(0,), min_depth=min_depth, max_depth=max_depth)
    hp = HotPath(cfg, device=dev, use_graph=False, num_slots=2, prepare_next=False)     # owns the head and the flat gradient bucket
    hp.load_state_dict(head_state(cfg), strict=True)
    conv = hp.convert_to_prob[0]
    g = torch.Generator().manual_seed(4321 + rank)
    host = {"x": torch.randn(B, E, h, w, generator=g).pin_memory(),
            "queries": (0.4 * torch.randn(B, Q, E, generator=g)).pin_memory()}
    gt = 1.0 + 79.0 * torch.rand(B, 1, H, W, generator=g)
    gt[torch.rand(B, 1, H, W, generator=g) < 0.8] = 0.0               # sparse LiDAR ground truth: ~20 % of the pixels valid
    host["depth"] = gt.pin_memory()
    sets = [{k: v.to(dev) for k, v in host.items()} for _ in range(2)]
    for st in sets:
        st["x"].requires_grad_(True)
        st["queries"].requires_grad_(True)
    one = torch.ones((), device=dev)
    losses = [None, None]

    def step_fn(slot):
        I, gv = sets[slot], hp.grad_views[slot]
        I["x"].grad = None
        I["queries"].grad = None
        pred = S.sql_tail(I["x"], I["queries"], conv.weight.view(D, Q), conv.bias, hp._centers_fn(slot), (),
                          head_grad_out=(gv[0].view(D, Q), gv[1]))
        loss = sqlx.finetune_loss(pred, I["depth"], min_depth, min_depth, max_depth, garg_crop=True)
        torch.autograd.backward(loss, grad_tensors=one)
        return loss.detach(), pred.detach()

    graphs = []
    for slot in range(2):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step_fn(slot)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            losses[slot], pred0 = step_fn(slot)
        graphs.append(gr)
    from sqlx import _lib
    _lib.profile_enable(True)
    for _ in range(5):
        step_fn(0)
    torch.cuda.synchronize()
    prof = _lib.profile_report()
    _lib.profile_enable(False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / k

    for _ in range(max(warmup, 3)):
        graphs[0].replay()
    ms_dev = timed(lambda: graphs[0].replay(), steps)
    # end to end: two input sets, the copy of batch i+1 on a copy stream under the step on batch i
    copy_stream, main = torch.cuda.Stream(), torch.cuda.current_stream()
    ev_loaded = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    loss_ring = [torch.zeros(1).pin_memory() for _ in range(2)]
    idx = [0]
    e2e_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def enqueue(slot):
        copy_stream.wait_event(ev_done[slot])
        with torch.cuda.stream(copy_stream), torch.no_grad():
            for k, v in host.items():
                sets[slot][k].copy_(v, non_blocking=True)
            ev_loaded[slot].record(copy_stream)

    def e2e_step():
        slot = idx[0] % 2
        main.wait_event(ev_loaded[slot])
        graphs[slot].replay()
        ev_done[slot].record(main)
        loss_ring[slot].copy_(losses[slot].reshape(1), non_blocking=True)
        enqueue((slot + 1) % 2)
        idx[0] += 1

    for ev in ev_done:
        ev.record(main)
    enqueue(0)
    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, steps)
    copy_stream.synchronize()
    res = {"config": {"workload": "BASELINE config 5 hot path: SQL decoder tail (x0 32x%dx%d, Q=%d, D=%d) + supervised "
                                  "fine-tune loss (resize to the %dx%d ground truth, per-sample median scaling of the first "
                                  "B/2 samples, SILog), forward+backward" % (h, w, Q, D, H, W),
                      "batch_per_gpu": B, "global_batch": B * world, "height": H, "width": W, "parallelism": "dp%d" % world,
                      "submission": "cuda_graph"},
           "value": B * world / (ms_dev * 1e-3), "unit": UNIT, "ms_per_step": ms_dev,
           "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": int(e2e_bytes), "d2h_bytes_per_step": 4},
           "loss": float(losses[0]),
           "kernels": {k: {"launches_per_step": v[0] / 5.0, "us_per_launch": 1e3 * v[1] / max(1, v[0])}
                       for k, v in sorted(prof.items())}}
    if world == 1 and not cx.args.no_parity:
        # parity of the timed batch against the float64 oracle on the host (forward: depth, ratios, loss)
        from oracle import sqldepth_oracle as O
        t0 = time.time()
        state = {k: v.detach().double().cpu() for k, v in hp.state_dict().items()}
        mlp = [state["bins_regressor.%d.%s" % (i, k)] for i in (0, 2, 4) for k in ("weight", "bias")]
        with torch.no_grad():
            graphs[0].replay()
            torch.cuda.synchronize()
            # (set 0 holds the host batch after the e2e loop: the same values the oracle sees)
            tail = O.sql_tail(host["x"].double(), host["queries"].double(), mlp,
                              state["convert_to_prob.0.weight"].view(D, Q), state["convert_to_prob.0.bias"], min_depth, max_depth)
            ref_loss, ref_ratio = O.finetune_loss(tail["pred"], host["depth"].double(), min_depth, min_depth, max_depth,
                                                  garg_crop=True)
        res["parity"] = {"loss_gpu": float(losses[0]), "loss_oracle_fp64": float(ref_loss),
                         "abs_diff": abs(float(losses[0]) - float(ref_loss)), "bar": 1e-5,
                         "depth_max_rel": float(((pred0.double().cpu() - tail["pred"]) / tail["pred"]).abs().max()),
                         "oracle": "oracle/sqldepth_oracle.py in float64 on the host CPU, forward only, %.1f s" % (time.time() - t0)}
    graphs.clear()
    return res


# --------------------------------------------------------------------------------------------- one workload on the GPU
class Ctx:
    pass


def run_workload(cx, cfg_id, cfg, steps, warmup, full):
    """Times one workload on this rank's GPU.  full: the headline workload (CPU legs, extras, per-N exchange A/B);
    otherwise a shorter measurement (value, e2e, per-kernel table, roofline, parity)."""
    import sqlx
    from sqlx import _lib
    from sqlx.hotpath import HotPath
    from _workload import make_host_batch, head_state, oracle_step
    args, dev, world, rank, dist = cx.args, cx.dev, cx.world, cx.rank, cx.dist
    res = {"config": {"workload": workload_text(cfg_id, cfg), "batch_per_gpu": cfg.B, "global_batch": cfg.B * world,
                      "height": cfg.H, "width": cfg.W, "source_frames": cfg.S, "loss_scales": len(cfg.scales),
                      "stereo": cfg.stereo, "parallelism": "dp%d" % world}}
    torch.manual_seed(0)
    exchange_in_step = world > 1 and not args.allreduce_after_step
    sm_reserve = 32 if args.exchange_sms < 0 else args.exchange_sms

    def exchange(flat):
        """the path's one exchange step: ONE NCCL all-reduce (average) of the flat gradient bucket, in place"""
        sqlx.dist.allreduce_flat_(flat)

    nslots = max(2, args.prefetch + 1)
    hp = HotPath(cfg, device=dev, use_graph=not args.no_graph, num_slots=nslots,
                 grad_exchange=exchange if exchange_in_step else None, exchange_sm_reserve=sm_reserve,
                 prepare_next=args.prepare_fork != "off",
                 prepare_fork=args.prepare_fork if args.prepare_fork != "off" else "auto")
    state = head_state(cfg)                       # identical initial weights on every rank (seeded)
    hp.load_state_dict(state, strict=True)
    hb = make_host_batch(cfg, seed=1234 + rank, pin=True, u8_frames=not args.f32_frames)
    for sl in range(nslots):
        hp.load(hb, non_blocking=False, slot=sl)
    torch.cuda.synchronize()

    # launches of OUR kernels in one step (counted eagerly; a graph replay re-issues the same nodes)
    n0 = sqlx.lib().sqlx_launch_count()
    hp.step_eager()
    torch.cuda.synchronize()
    launches_per_step = int(sqlx.lib().sqlx_launch_count() - n0)
    if hp.use_graph:
        for sl in range(nslots):
            hp.capture(slot=sl)

    exchange_check = None
    if exchange_in_step and full:
        # pre-flight: the gradients of a step with the in-step (overlapped, graph-captured) exchange equal those of an
        # eager step followed by a plain all-reduce of the bucket
        hp.step(0)
        torch.cuda.synchronize()
        fused = hp.grad_flat[0].clone()
        hp.grad_exchange = None
        hp.step_eager(0)
        ref = hp.grad_flat[0].clone()
        exchange(ref)
        torch.cuda.synchronize()
        exchange_check = float((fused - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        hp.grad_exchange = exchange
        assert exchange_check < 1e-4, "in-step gradient exchange disagrees with the plain all-reduce: %g" % exchange_check

    def allreduce_after():
        if world > 1 and not exchange_in_step:
            exchange(hp.grad_flat[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / k

    def dev_step():
        hp.step()
        allreduce_after()

    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    ev_loaded = [torch.cuda.Event() for _ in range(nslots)]
    ev_done = [torch.cuda.Event() for _ in range(nslots)]
    e2e_i = [0]
    # The auto-mask tie-break noise is NOT shipped from the host in the end-to-end loop: the reference draws it with
    # the CPU generator and copies it every step (trainer.py:516-517), our API draws it on the device when no noise
    # tensor is supplied.  Everything else the step consumes (frames, decoder features, queries, intrinsics, poses,
    # coarser-scale depth maps) crosses PCIe every step.
    hb_e2e = {k: (torch.zeros_like(v) if k.startswith("noise") else v) for k, v in hb.items()}
    host_frames, host_other = hp.pack_host(hb_e2e, pin=True)     # what a collate_fn would fill: two pinned buffers
    other_ship = hp.other_numel_without(("noise",))
    e2e_bytes = host_frames.numel() * host_frames.element_size() + other_ship * 4

    def enqueue_load(slot):
        """host -> device copy of one batch into input set `slot` on the copy stream (+ fresh device noise)"""
        copy_stream.wait_event(ev_done[slot])          # the previous step on this set has finished reading it
        with torch.cuda.stream(copy_stream):
            hp.load_flat(host_frames, host_other[:other_ship], non_blocking=True, slot=slot)
            hp.noise_region(slot).normal_()             # all scales' noise: one contiguous region, one kernel
            ev_loaded[slot].record(copy_stream)

    loss_ring = [torch.zeros(1).pin_memory() for _ in range(nslots)]
    ev_loss = [torch.cuda.Event() for _ in range(nslots)]
    loss_seen = [0.0]
    ahead = nslots - 1

    def e2e_step():
        """What a training loop with a pinned-memory prefetching loader does: the H2D copies of the next `ahead`
        batches overlap the step on batch i.  The loss of EVERY step is copied to pinned host memory and read by the
        host one step later, so the host never stalls the device (the reference reads the loss only on logging steps,
        trainer.py:242-262)."""
        i = e2e_i[0]
        slot = i % nslots
        main.wait_event(ev_loaded[slot])
        if ahead >= 2:                                  # the step also prepares the NEXT set's frame-only work
            main.wait_event(ev_loaded[(slot + 1) % nslots])     # (HotPath.prepare_next): its frames must have landed
        hp.step(slot)
        allreduce_after()
        ev_done[slot].record(main)
        loss_ring[slot].copy_(hp.loss.reshape(1), non_blocking=True)
        ev_loss[slot].record(main)
        enqueue_load((i + ahead) % nslots)             # keep `ahead` batches in flight
        if i > 0:                                      # host-side read of the previous step's loss
            pslot = (i - 1) % nslots
            ev_loss[pslot].synchronize()
            loss_seen[0] = float(loss_ring[pslot])
        e2e_i[0] = i + 1

    for _ in range(max(warmup, 3)):
        dev_step()
    sampler = None
    if rank == 0 and full:
        sampler = ClockSampler(cx.local_rank)
        sampler.start()
    ms_dev = timed(dev_step, steps)
    for ev in ev_done:
        ev.record(main)
    for sl in range(ahead):                            # the first batches; afterwards every step prefetches one more
        enqueue_load(sl)
    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, steps)                    # (timed() ends with a device synchronize: the last loss is in)
    copy_stream.synchronize()
    if sampler is not None:
        res["clocks"] = sampler.stop()
    res.update(value=cfg.B * world / (ms_dev * 1e-3), ms_per_step=ms_dev,
               e2e={"value": cfg.B * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(e2e_bytes), "d2h_bytes_per_step": 4},
               gpu_launches_per_step=launches_per_step, loss=float(hp.loss))

    if not args.no_extras:
        # the copy ceiling of the end-to-end loop: the same host -> device traffic with no compute behind it
        h2d_i = [0]

        def h2d_only():
            with torch.cuda.stream(copy_stream):
                hp.load_flat(host_frames, host_other[:other_ship], non_blocking=True, slot=h2d_i[0] % nslots)
            main.wait_stream(copy_stream)
            h2d_i[0] += 1
        h2d_only(); torch.cuda.synchronize()
        ms_h2d = timed(h2d_only, min(steps, 50))
        res["e2e"]["h2d_only_ms_per_step"] = ms_h2d
        res["e2e"]["h2d_only_GBps_per_gpu"] = e2e_bytes / (ms_h2d * 1e-3) / 1e9

    if world > 1 and full:
        # exposed communication: the same step replayed with and without the in-step exchange
        if exchange_in_step:
            hp.grad_exchange = None
            hp.drop_graphs()
            if hp.use_graph:
                hp.capture(slot=0)
            for _ in range(3):
                hp.step(0)
            ms_nocomm = timed(lambda: hp.step(0), steps)
            res["exposed_comm_ms"] = ms_dev - ms_nocomm
            res["ms_per_step_without_exchange"] = ms_nocomm
            hp.grad_exchange = exchange
            hp.drop_graphs()

    # per-kernel device times: the same steps submitted eagerly with CUDA events around each main kernel
    ge, hp.grad_exchange = hp.grad_exchange, None
    _lib.profile_enable(True)
    nprof = min(steps, 20)
    for _ in range(nprof):
        hp.step_eager()
    torch.cuda.synchronize()
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    hp.grad_exchange = ge
    if rank != 0:
        hp.drop_graphs()
        del hp
        gc.collect()
        return None

    ab, fl = algorithmic_bytes(cfg), algorithmic_flops(cfg)
    s8d = survey_8d_bytes(cfg)
    peak, peak_src, tpeak = cx.peak, cx.peak_src, cx.tf32_peak
    ktot = sum(v[1] for v in prof.values())
    kernels = {}
    for k, v in sorted(prof.items()):
        us = v[1] / v[0] * 1e3
        row = {"launches_per_step": v[0] / max(1, nprof), "us_per_launch": us, "share_of_kernel_time": v[1] / ktot,
               "hbm_frac": (ab[k] / (us * 1e-6) / 1e9 / peak) if k in ab else None}
        if k in fl:
            row["tensor_frac_1xtf32"] = fl[k] / (us * 1e-6) / 1e12 / tpeak
        kernels[k] = row
    res["kernels"] = kernels
    res["kernel_ms_per_step"] = ktot / max(1, nprof)
    cand = {k: v for k, v in prof.items() if k in ab}
    dom = max(cand.items(), key=lambda kv: kv[1][1])[0] if cand else None
    traffic = cx.traffic.get("config%d" % cfg_id, {})
    if dom is not None:
        cnt, tot_ms = prof[dom]
        sec = tot_ms / cnt * 1e-3
        achieved = ab[dom] / sec / 1e9
        tr = traffic.get(dom)
        res["roofline"] = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": tr, "peak_source": peak_src,
                           "avg_launch_us": sec * 1e6, "algorithmic_bytes_per_launch": ab[dom],
                           "algorithmic_bytes": "unique bytes per launch: frames / identity losses once, depth / noise / "
                                                "arg-min per scale (fp32)",
                           "traffic_over_algorithmic": (tr / ab[dom]) if tr else None,
                           "survey_8d_bytes_per_launch": s8d.get(dom),
                           "survey_8d_frac": (s8d[dom] / sec / 1e9 / peak) if dom in s8d else None,
                           "share_of_kernel_time": tot_ms / ktot,
                           "timing": "CUDA events around each launch on the launching stream (eager submission of the "
                                     "same steps after the timed region); traffic = dram bytes/launch from the "
                                     "committed ncu capture profiles/traffic_latest.json"}
    # step-level roofline: SURVEY 8d's compulsory bytes of the whole step over the device-timed step
    N, n0 = cfg.H * cfg.W, cfg.h * cfg.w
    step_bytes = cfg.B * (len(cfg.scales) * ((24 + 32 * cfg.S) * N + 12 * n0) + 4 * n0 * (3 * cfg.E + 2))
    res["step_roofline"] = {"survey_8d_bytes_per_step": step_bytes, "frac": step_bytes / (ms_dev * 1e-3) / 1e9 / peak}
    res["selection"] = selection_stats(hp, cfg)

    if not args.no_parity and world == 1:
        # the fp64 oracle on the timed batch (forward only, CPU, two samples at a time: every loss term is a per-sample
        # mean followed by a batch mean, trainer.py:532,535,546, so the batch loss is the mean of the chunk losses)
        t0 = time.perf_counter()
        tot = 0.0
        hb_cpu = {k: v for k, v in hb.items()}
        for b0 in range(0, cfg.B, 2):
            sub = {k: v[b0:b0 + 2] for k, v in hb_cpu.items()}
            ccfg = type(cfg)(**{**cfg_kwargs(cfg), "B": sub["x"].shape[0]})
            tot += float(oracle_step(ccfg, sub, state, backward=False)["loss"]) * sub["x"].shape[0]
        loss_ref = tot / cfg.B
        # the GPU loss on the SAME noise (the timed loops draw fresh device noise per step)
        hp.load(hb, non_blocking=False, slot=0)
        hp.step_eager(0)
        torch.cuda.synchronize()
        res["parity"] = {"loss_gpu": float(hp.loss), "loss_oracle_fp64": loss_ref,
                         "abs_diff": abs(float(hp.loss) - loss_ref), "bar": 1e-5,
                         "oracle": "oracle/sqldepth_oracle.py in float64 on the host CPU, forward only, %.1f s"
                                   % (time.perf_counter() - t0)}

    hp.drop_graphs()
    del hp
    gc.collect()
    torch.cuda.empty_cache()
    if full and not args.no_extras and world == 1:
        # The backward's coefficient reads and tile skipping depend on which candidate wins the per-pixel minimum, and the
        # synthetic batch is auto-masked on most pixels (see `selection`).  Upper bound of the selection-dependent cost:
        # the same batch with --disable_automasking (trainer.py:480,514,520): EVERY pixel selects a reprojection, as a
        # well-trained model on real driving data does on most pixels -- no tile is skipped, every coefficient is read.
        cfg2 = type(cfg)(**{**cfg_kwargs(cfg), "automask": False})
        hp2 = HotPath(cfg2, device=dev, use_graph=not args.no_graph, num_slots=1)
        hp2.load_state_dict(state, strict=True)
        hp2.load(hb, non_blocking=False)
        for _ in range(3):
            hp2.step()
        ms2 = timed(lambda: hp2.step(), steps)
        _lib.profile_enable(True)
        for _ in range(10):
            hp2.step_eager()
        torch.cuda.synchronize()
        prof2 = _lib.profile_report()
        _lib.profile_enable(False)
        res["reprojection_dominant"] = {
            "value": cfg.B / (ms2 * 1e-3), "ms_per_step": ms2, "selection": selection_stats(hp2, cfg2),
            "kernels_us": {k: v[1] / v[0] * 1e3 for k, v in prof2.items() if k.startswith("photo_")},
            "batch": "the timed batch with --disable_automasking: 100 % of the pixels select a reprojection"}
        hp2.drop_graphs()
        del hp2
        gc.collect()
        torch.cuda.empty_cache()
    return res


def cfg_kwargs(c):
    return dict(B=c.B, H=c.H, W=c.W, h=c.h, w=c.w, E=c.E, Q=c.Q, D=c.D, S=c.S, scales=c.scales, min_depth=c.min_depth,
                max_depth=c.max_depth, disparity_smoothness=c.disparity_smoothness, stereo=c.stereo, automask=c.automask)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from _workload import baseline_config
    cfg = baseline_config(args.config, B=args.batch or None)

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, args.steps), max(0, args.warmup)
        cb = time_reference(cfg, args.config, steps=steps, warmup=warmup, device=args.device)
        config = {"workload": workload_text(args.config, cfg), "batch_per_gpu": cfg.B, "global_batch": cfg.B,
                  "height": cfg.H, "width": cfg.W, "source_frames": cfg.S, "loss_scales": len(cfg.scales),
                  "stereo": cfg.stereo, "parallelism": "dp1"}
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
                "n_gpus": 1 if args.device == "cuda" else 0, "steps": steps, "warmup": warmup,
                "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config, "gpu_launches": 0, "device": args.device,
                "loss": cb["loss"],
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import sqlx
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (use --impl reference for the CPU arm)"
    from sqlx.affinity import bind_to_gpu
    binding = bind_to_gpu(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        # keep stdout to the single JSON line: NCCL's version banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    assert sqlx.lib().sqlx_device_ok(local_rank) == 1, "libsqlx targets sm_100a (B200) only"

    cx = Ctx()
    cx.args, cx.dev, cx.world, cx.rank, cx.local_rank, cx.dist = args, dev, world, rank, local_rank, dist
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    cx.peak = float(peaks.get("hbm_gbs", 6650.0))
    cx.peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # 1xTF32 tensor peak = half the measured sustained bf16 rate
    cx.tf32_peak = 0.5 * float(peaks.get("bf16_tflops_sustained", 1422.0))
    cx.traffic = load_traffic()

    def build_line(res, extra):
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        config = res.pop("config")
        config["l2"] = ("no explicit flush: every step streams its inputs (frames, decoder features, noise: > 126 MB at every "
                        "workload) plus the saved planes of the loss scales through the 126 MB L2")
        config["submission"] = "eager" if args.no_graph else "cuda_graph"
        config["frame_only_work"] = ("identity reprojection losses + pixel-interleaved source copies: every step computes them "
                                     "for the NEXT input set inside its own graph, at low stream priority under its backward, "
                                     "and consumes what the previous step prepared (HotPath.prepare_next; once per step, as "
                                     "in the reference)")
        if world > 1:
            config["grad_exchange"] = (
                "(summary-path backward leaves %d SMs to the NCCL kernel) " % (32 if args.exchange_sms < 0 else args.exchange_sms) +
                "one NCCL all-reduce (average) of the flat gradient bucket per step (the kernels write the parameter gradients "
                "into views of it: no pack / unpack), issued inside the step on a communication stream when the last parameter "
                "gradient exists (overlaps the summary-path backward kernel; part of the CUDA graph)"
                if not args.allreduce_after_step else
                "one NCCL all-reduce (average) of the flat gradient bucket after the step")
        config["e2e_pipeline"] = ("H2D of the next %d batches on a copy stream overlaps the step on batch i (%d device input "
                                  "sets); tie-break noise drawn on the device instead of copied from the host; the loss of "
                                  "every step is copied to pinned host memory and read by the host one step later; frames "
                                  "shipped as %s" % (max(1, args.prefetch), max(2, args.prefetch + 1),
                                                     "float32" if args.f32_frames else "uint8 and scaled to [0,1] on the device"))
        config["cpu_binding"] = binding
        line["config"] = config
        line["e2e"] = res.pop("e2e")
        line["gpu_launches"] = res["gpu_launches_per_step"] * args.steps
        for k in ("gpu_launches_per_step", "clocks", "roofline", "step_roofline", "kernels", "kernel_ms_per_step", "loss",
                  "parity", "selection", "reprojection_dominant", "exposed_comm_ms", "ms_per_step_without_exchange"):
            if k in res:
                line[k] = res[k]
        line["workloads"] = extra
        return line

    res = run_workload(cx, args.config, cfg, args.steps, args.warmup, full=True)
    extra = {}
    # N > 1 insurance: the headline workload is measured; should a later workload of this launch stall (a collective one
    # rank never enters cannot be caught as an exception), every rank leaves after `bail_after` seconds and rank 0 prints
    # the line it has, with the reason under "workloads"
    emitted = threading.Event()
    bail_after = float(os.environ.get("SQLX_BENCH_BAIL_AFTER", "240"))
    watchdog = None
    if world > 1 and args.workloads.strip():
        def bail():
            if rank == 0 and not emitted.is_set():
                emitted.set()
                try:
                    partial = build_line(dict(res), {"aborted": "a further workload did not finish within %.0f s; "
                                                                "headline workload only" % bail_after})
                    partial["cpu_baseline"] = None
                    print(json.dumps(partial))
                    sys.stdout.flush()
                except Exception:
                    pass
            os._exit(0)
        watchdog = threading.Timer(bail_after, bail)
        watchdog.daemon = True
        watchdog.start()
    for tok in [t for t in args.workloads.split(",") if t.strip()]:
        n = int(tok)
        if n == args.config:
            continue
        if n == 5:
            # single-GPU leg only: under torchrun the line carries the self-supervised workloads (whose in-step gradient
            # exchange is the multi-GPU path that is measured); nothing rank-dependent may run between collectives
            if world == 1:
                extra["config5"] = run_finetune_workload(cx, max(20, min(args.steps, 100)), max(args.warmup, 3))
            continue
        r = run_workload(cx, n, baseline_config(n), max(20, min(args.steps, 100)), max(args.warmup, 3), full=False)
        if r is not None:
            extra["config%d" % n] = r
    if watchdog is not None:
        watchdog.cancel()

    def finish():
        """N > 1 teardown: the graphs holding captured collectives are gone (run_workload drops them), drain the device
        and destroy the process group; a watchdog ends the process if the communicator teardown does not return."""
        if world == 1:
            return
        torch.cuda.synchronize()
        gc.collect()
        sys.stdout.flush()
        sys.stderr.flush()
        t = threading.Timer(20.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        dist.barrier()
        dist.destroy_process_group()
        t.cancel()

    if rank != 0:
        finish()
        return

    if emitted.is_set():          # the watchdog already printed (cannot happen after cancel(); belt and braces)
        finish()
        return
    emitted.set()
    line = build_line(res, extra)

    cpu = None
    if not args.no_cpu_baseline and world == 1:          # reported beside the N = 1 line only
        cb = time_reference(cfg, args.config, steps=4, warmup=1, device="cpu")
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line["cpu_baseline"] = cpu
    if not args.no_extras and world == 1:
        # informational: the same unmodified reference code on THIS GPU (stock PyTorch / cuDNN / cuBLAS op chain)
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--device", "cuda",
                                  "--config", str(args.config), "--steps", "20", "--warmup", "5"],
                                 capture_output=True, text=True, timeout=300)
            g = json.loads(out.stdout.strip().splitlines()[-1])
            line["torch_gpu_baseline"] = {"value": g["value"], "unit": UNIT, "ms_per_step": g["ms_per_step"],
                                          "what": "unmodified reference modules (oracle/_ref) on the same B200, fp32, "
                                                  "eager PyTorch; includes the decoder front (see cpu_baseline.sample)"}
        except Exception as e:       # never fail the bench line on the informational leg
            line["torch_gpu_baseline"] = {"unavailable": repr(e)[:200]}
    print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
