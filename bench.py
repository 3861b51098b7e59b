#!/usr/bin/env python
"""Benchmark of the SQLdepth training hot path (BASELINE.json metric: train frames/s).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (libsqlx)
  python bench.py --impl reference [...]                         the reference's CPU path (oracle port)

One "step" = SQL decoder tail forward -> photometric losses (4 loss scales) forward -> backward of both, on one
batch of synthetic KITTI-shape 3-frame inputs (BASELINE config 2: batch 12 per GPU, 192x640, decoder features
32x96x320, Q = D = 64).  frames/s = target frames processed per second = batch / step time (trainer.py:584).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the byte accounting behind `roofline`.
"""
import argparse
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
import torch.nn.functional as F  # noqa: E402

METRIC = "train_frames_per_sec_hot_path"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="sqlx", choices=["sqlx", "reference"])
    ap.add_argument("--batch", type=int, default=12, help="batch per GPU (BASELINE config 2: 12)")
    ap.add_argument("--height", type=int, default=192)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--queries", type=int, default=64)
    ap.add_argument("--bins", type=int, default=64)
    ap.add_argument("--scales", type=int, default=4, help="number of loss scales (BASELINE config 2: 4)")
    ap.add_argument("--sources", type=int, default=2, help="source frames (2 = [-1,+1]; 3 with --use_stereo)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3],
                    help="2 = BASELINE config 2 (default, the bench line); 3 = config 3 shapes: 320x1024, batch 8/GPU, "
                         "3 sources, Q = D = 128, single loss scale (informational)")
    ap.add_argument("--no-graph", action="store_true", help="submit the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--f32-frames", action="store_true",
                    help="ship the frames as float32 in the end-to-end loop (default: uint8, converted on the device)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--allreduce-after-step", action="store_true",
                    help="N > 1: issue the gradient all-reduce after the step instead of inside it (overlapping the last "
                         "backward kernel); A/B switch")
    ap.add_argument("--cpu-sample-batch", type=int, default=2)
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- synthetic data
def make_host_batch(cfg, seed, pin, u8_frames=False):
    """Seeded KITTI-shape synthetic batch on the host (SURVEY 8d recipe): smooth frames, KITTI intrinsics,
    PoseCNN-scale poses, decoder-feature-like x and queries.  Frames are quantised to 8 bits (as decoded images
    are); with u8_frames they stay uint8 on the host and are scaled to [0,1] on the device by HotPath.load."""
    from _cases import smooth_images, kitti_K, depth_like
    g = torch.Generator().manual_seed(seed)
    c = cfg
    frames = smooth_images(g, c.B, c.H, c.W, c.S + 1)
    mid = (c.S + 1) // 2
    hb = {"target": frames[mid]}
    for i, fr in enumerate([f for j, f in enumerate(frames) if j != mid]):
        hb["source%d" % i] = fr
    hb["K"], hb["inv_K"] = kitti_K(c.B, c.H, c.W)
    hb["x"] = torch.randn(c.B, c.E, c.h, c.w, generator=g)
    hb["queries"] = 0.4 * torch.randn(c.B, c.Q, c.E, generator=g)
    for i in range(c.S):
        hb["axisangle%d" % i] = 0.01 * torch.randn(c.B, 1, 1, 3, generator=g)
        hb["translation%d" % i] = 0.01 * torch.randn(c.B, 1, 1, 3, generator=g)
    for s in c.scales:
        hb["noise%d" % s] = torch.randn(c.B, c.S, c.H, c.W, generator=g)
        if s > 0:
            hs, ws = c.scale_hw(s)
            hb["disp%d" % s] = depth_like(g, c.B, hs, ws)
            hb["target%d" % s] = F.interpolate(hb["target"], [c.H // 2 ** s, c.W // 2 ** s], mode="bilinear",
                                               align_corners=False)
    hb = {k: v.contiguous().float() for k, v in hb.items()}
    for k in list(hb):
        if k.startswith("target") or k.startswith("source"):
            q = (hb[k].clamp(0, 1) * 255.0).round()
            hb[k] = q.to(torch.uint8) if u8_frames else q / 255.0
    if pin:
        hb = {k: v.pin_memory() for k, v in hb.items()}
    return hb


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
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_factory(cfg, hb, mlp_state):
    """The reference's CPU path for the same step, restated by oracle/sqldepth_oracle.py (same ATen primitives as
    the reference: matmul, softmax, F.interpolate, F.grid_sample, avg_pool2d), autograd backward included."""
    from oracle import sqldepth_oracle as O
    c = cfg
    leaves = {k: hb[k].clone().requires_grad_(True) for k in ["x", "queries"] +
              ["disp%d" % s for s in c.scales if s > 0] +
              ["axisangle%d" % i for i in range(c.S)] + ["translation%d" % i for i in range(c.S)]}
    params = {k: v.clone().requires_grad_(True) for k, v in mlp_state.items()}

    def step():
        for t in list(leaves.values()) + list(params.values()):
            t.grad = None
        mlp = [params["bins_regressor.%d.%s" % (i, k)] for i in (0, 2, 4) for k in ("weight", "bias")]
        Wp = params["convert_to_prob.0.weight"].view(c.D, c.Q)
        tail = O.sql_tail(leaves["x"], leaves["queries"], mlp, Wp, params["convert_to_prob.0.bias"], c.min_depth,
                          c.max_depth)
        disps = {s: (tail["pred"] if s == 0 else leaves["disp%d" % s]) for s in c.scales}
        target_pyr = {s: (hb["target"] if s == 0 else hb["target%d" % s]) for s in c.scales}
        poses = [{"axisangle": leaves["axisangle%d" % i], "translation": leaves["translation%d" % i], "invert": i == 0}
                 for i in range(c.S)]
        out = O.photometric_losses(disps, target_pyr, [hb["source%d" % i] for i in range(c.S)], hb["K"], hb["inv_K"],
                                   poses, {s: hb["noise%d" % s] for s in c.scales}, height=c.H, width=c.W,
                                   scales=c.scales, disparity_smoothness=c.disparity_smoothness)
        out["loss"].backward()
        return float(out["loss"].detach())
    return step


def time_cpu(cfg_full, args, steps, warmup):
    from sqlx.hotpath import HotPathConfig
    c = cfg_full
    Bs = min(args.cpu_sample_batch, c.B)
    cs = HotPathConfig(B=Bs, H=c.H, W=c.W, h=c.h, w=c.w, E=c.E, Q=c.Q, D=c.D, S=c.S, scales=c.scales,
                       min_depth=c.min_depth, max_depth=c.max_depth)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hb = make_host_batch(cs, seed=1234, pin=False)
    torch.manual_seed(0)
    nn = torch.nn
    conv = nn.Conv2d(c.Q, c.D, 1)
    mlp = nn.Sequential(nn.Linear(c.E * c.Q, 16 * c.Q), nn.LeakyReLU(), nn.Linear(16 * c.Q, 256), nn.LeakyReLU(),
                        nn.Linear(256, c.D))
    state = {"convert_to_prob.0.weight": conv.weight.detach(), "convert_to_prob.0.bias": conv.bias.detach()}
    for k, v in mlp.state_dict().items():
        state["bins_regressor." + k] = v
    step = cpu_reference_step_factory(cs, hb, state)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": Bs / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "batch %d of the same workload (%dx%d, %d loss scales, fwd+bwd), %d timed steps, torch CPU fp32 "
                      "threads=%d" % (Bs, c.H, c.W, len(c.scales), steps, cores),
            "ms_per_step": dt * 1e3}


# --------------------------------------------------------------------------------------------- roofline
def algorithmic_bytes(c):
    """Compulsory fp32 bytes per launch of every main kernel ("Measurement"), averaged over the
    loss scales for the per-scale kernels."""
    N, S, B, E = c.H * c.W, c.S, c.B, c.E
    n0 = c.h * c.w
    ns = [c.scale_hw(s)[0] * c.scale_hw(s)[1] for s in c.scales]
    n_avg = sum(ns) / len(ns)
    return {
        # depth_lr + target + S sources + S identity + S noise (reads); argmin u8 (write)
        "photo_fwd_kernel": B * (4 * n_avg + 12 * N + 12 * N * S + 4 * N * S + 4 * N * S + N),
        # all loss scales in one launch: the per-scale figure (SURVEY 8d's unit) times the scales one launch processes
        "photo_fwd_ms_kernel": B * sum(4 * n + 12 * N + 12 * N * S + 4 * N * S + 4 * N * S + N for n in ns),
        # depth_lr + target + S sources + argmin (reads); d_depth_lr (write)
        "photo_bwd_kernel": B * (4 * n_avg + 12 * N + 12 * N * S + N + 4 * n_avg),
        "photo_bwd_ms_kernel": B * sum(4 * n + 12 * N + 12 * N * S + N + 4 * n for n in ns),
        "reproj_loss_kernel": B * (24 * N + 4 * N),
        # identity losses of all S sources in one launch: target once, every source once, S loss maps out
        "identity_loss_kernel": B * (12 * N + 12 * N * S + 4 * N * S),
        "sql_summary_kernel": B * 4 * n0 * E,
        "sql_pred_kernel": B * (4 * n0 * E + 4 * n0),
        "sql_bwd_reduce_kernel": B * (4 * n0 * E + 4 * n0),
        "sql_bwd_dx_kernel": B * (4 * n0 * E + 4 * n0 + 4 * n0 * E),
        # tensor-core versions: same compulsory traffic
        "sql_tc_summary_kernel": B * 4 * n0 * E,
        "sql_tc_pred_kernel": B * (4 * n0 * E + 4 * n0),
        "sql_tc_bwd_reduce_kernel": B * (4 * n0 * E + 4 * n0),
        "sql_tc_bwd_dx_kernel": B * (4 * n0 * E + 4 * n0 + 4 * n0 * E),
        # mixed-weight decomposition (sql_tc.cu): regression backward reads x + g_pred and writes d_x; the summary-path
        # backward reads x and accumulates into d_x (read + write)
        "sql_tc_bwd_pred_kernel": B * (4 * n0 * E + 4 * n0 + 4 * n0 * E),
        "sql_tc_bwd_sum_kernel": B * (4 * n0 * E + 2 * 4 * n0 * E),
    }


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from sqlx.hotpath import HotPath, HotPathConfig
    if args.config == 3:
        args.height, args.width, args.batch, args.queries, args.bins, args.sources, args.scales = 320, 1024, 8, 128, 128, 3, 1
    H, W = args.height, args.width
    cfg = HotPathConfig(B=args.batch, H=H, W=W, h=H // 2, w=W // 2, E=32, Q=args.queries, D=args.bins, S=args.sources,
                        scales=tuple(range(args.scales)), min_depth=0.01 if args.config == 3 else 0.001)
    config = {"workload": "BASELINE config %d hot path: SQL decoder tail (x0 32x%dx%d, Q=%d, D=%d) + photometric loss "
                          "(%dx%d, %d loss scales: scale 0 = decoder output, coarser scales synthetic "
                          "since the reference decoder emits scale 0 only), forward+backward"
                          % (args.config, H // 2, W // 2, args.queries, args.bins, H, W, args.scales),
              "batch_per_gpu": args.batch, "global_batch": args.batch * world, "height": H, "width": W,
              "source_frames": args.sources, "loss_scales": args.scales, "parallelism": "dp%d" % world}

    if args.impl == "reference":
        if rank != 0:
            return
        cb = time_cpu(cfg, args, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": 0,
                "steps": max(1, args.steps), "warmup": max(1, min(args.warmup, 2)), "ms_per_step": cb["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "gpu_launches": 0,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import sqlx
    from sqlx import _lib
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        # keep stdout to the single JSON line: NCCL's version banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    assert sqlx.lib().sqlx_device_ok(local_rank) == 1, "libsqlx targets sm_100a (B200) only"

    torch.manual_seed(0)
    # N > 1: the path's one exchange step -- a single bucketed NCCL all-reduce (average) of the parameter gradients --
    # is issued inside the step on a communication stream as soon as the last parameter gradient exists, so it
    # overlaps the summary-path backward kernel (captured into the step's CUDA graph with everything else)
    exchange_in_step = world > 1 and not args.allreduce_after_step
    bucket_box = [None]

    def exchange(grads):
        if bucket_box[0] is None:
            from sqlx.dist import GradBucket
            bucket_box[0] = GradBucket(grads)
        bucket_box[0].allreduce_(grads)

    hp = HotPath(cfg, device=dev, use_graph=not args.no_graph, num_slots=2,
                 grad_exchange=exchange if exchange_in_step else None)
    if world > 1:   # identical initial weights on every rank
        for p in hp.parameters():
            dist.broadcast(p.data, 0)
    hb = make_host_batch(cfg, seed=1234 + rank, pin=True, u8_frames=not args.f32_frames)
    h2d_bytes = hp.load(hb, non_blocking=False, slot=0)
    hp.load(hb, non_blocking=False, slot=1)
    torch.cuda.synchronize()

    # launches of OUR kernels in one step (counted eagerly; a graph replay re-issues the same nodes)
    n0 = sqlx.lib().sqlx_launch_count()
    hp.step_eager()
    torch.cuda.synchronize()
    launches_per_step = int(sqlx.lib().sqlx_launch_count() - n0)
    if hp.use_graph:
        hp.capture(slot=0)
        hp.capture(slot=1)

    exchange_check = None
    if exchange_in_step:
        # pre-flight: the gradients of a step with the in-step (overlapped, graph-captured) exchange equal those of an
        # eager step followed by a plain bucketed all-reduce
        hp.step(0)
        torch.cuda.synchronize()
        fused = [g.clone() for g in hp.param_grads()]
        hp.grad_exchange = None
        hp.step_eager(0)
        from sqlx.dist import GradBucket
        ref = hp.param_grads()
        GradBucket(ref).allreduce_(ref)
        torch.cuda.synchronize()
        exchange_check = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(fused, ref))
        hp.grad_exchange = exchange
        assert exchange_check < 1e-4, "in-step gradient exchange disagrees with the plain all-reduce: %g" % exchange_check

    bucket = None

    def allreduce_grads():
        """the path's one exchange step: a single bucketed NCCL all-reduce (average) of the parameter gradients"""
        nonlocal bucket
        if world == 1 or exchange_in_step:
            return
        grads = hp.param_grads()
        if bucket is None:
            from sqlx.dist import GradBucket
            bucket = GradBucket(grads)
        bucket.allreduce_(grads)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    def dev_step():
        hp.step()
        allreduce_grads()

    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    ev_loaded = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_i = [0]

    # The auto-mask tie-break noise is NOT shipped from the host in the end-to-end loop: the reference draws it with
    # the CPU generator and copies it every step (trainer.py:516-517), our API draws it on the device when no noise
    # tensor is supplied.  Everything else the step consumes (frames, decoder features, queries, intrinsics, poses,
    # coarser-scale depth maps) crosses PCIe every step.
    hb_e2e = {k: (torch.zeros_like(v) if k.startswith("noise") else v) for k, v in hb.items()}
    host_frames, host_other = hp.pack_host(hb_e2e, pin=True)     # what a collate_fn would fill: two pinned buffers
    noise_floats = sum(hb[k].numel() for k in hb if k.startswith("noise"))
    # the noise region is part of the flat buffer but is NOT shipped: only the leading part (everything else) is
    other_ship = hp.other_numel_without(("noise",))
    e2e_bytes = host_frames.numel() * host_frames.element_size() + other_ship * 4

    def enqueue_load(slot):
        """host -> device copy of one batch into input set `slot` on the copy stream (+ fresh device noise)"""
        copy_stream.wait_event(ev_done[slot])          # the previous step on this set has finished reading it
        with torch.cuda.stream(copy_stream):
            hp.load_flat(host_frames, host_other[:other_ship], non_blocking=True, slot=slot)
            hp.noise_region(slot).normal_()             # all scales' noise: one contiguous region, one kernel
            ev_loaded[slot].record(copy_stream)

    loss_ring = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
    ev_loss = [torch.cuda.Event(), torch.cuda.Event()]
    loss_seen = [0.0]

    def e2e_step():
        """What a training loop with a pinned-memory prefetching loader does: the H2D copy of batch i+1 overlaps the
        step on batch i (two device input sets).  The loss of EVERY step is copied to pinned host memory and read by
        the host; the read of step i happens after step i+1 has been submitted, so the host never stalls the device
        (the reference reads the loss only on logging steps, trainer.py:242-262)."""
        i = e2e_i[0]
        slot = i & 1
        main.wait_event(ev_loaded[slot])
        hp.step(slot)
        allreduce_grads()
        ev_done[slot].record(main)
        loss_ring[slot].copy_(hp.loss.reshape(1), non_blocking=True)
        ev_loss[slot].record(main)
        enqueue_load(slot ^ 1)                         # prefetch the next batch while this one computes
        if i > 0:                                      # host-side read of the previous step's loss
            ev_loss[slot ^ 1].synchronize()
            loss_seen[0] = float(loss_ring[slot ^ 1])
        e2e_i[0] = i + 1

    for _ in range(max(args.warmup, 3)):
        dev_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(dev_step, args.steps)
    for ev in ev_done:
        ev.record(main)
    enqueue_load(0)                                    # the very first batch; afterwards every step prefetches the next
    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)               # (timed() ends with a device synchronize: the last loss is in)
    copy_stream.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel device times: the same steps submitted eagerly with CUDA events around each main kernel
    _lib.profile_enable(True)
    nprof = min(args.steps, 20)
    for _ in range(nprof):
        hp.step_eager()
    torch.cuda.synchronize()
    prof = _lib.profile_report()
    _lib.profile_enable(False)

    def finish():
        """N > 1 teardown.  With the all-reduce captured inside the step's CUDA graphs, destroying the NCCL communicator
        while those graphs are alive blocks forever (observed on 2 x B200, torch 2.11 / NCCL 2.28): drop the graphs, drain
        the device, and leave without the communicator teardown (the process is exiting anyway)."""
        if world == 1:
            return
        torch.cuda.synchronize()
        if exchange_in_step:
            hp.graphs = [None] * len(hp.graphs)
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()

    if rank != 0:
        finish()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    ab = algorithmic_bytes(cfg)
    traffic = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
        for k, v in tj["kernels"].items():     # ncu names -> the names of the library's profile hooks
            full = k.replace("void ", "")
            k = full.split("<")[0].split("::")[-1]
            ms = full.rstrip().endswith(", 1>")       # last template argument of the photometric kernels: all scales per launch
            k = {"photo_fwd3_kernel": "photo_fwd_ms_kernel" if ms else "photo_fwd_kernel",
                 "photo_bwd3_kernel": "photo_bwd_ms_kernel" if ms else "photo_bwd_kernel",
                 "sql_tc_pred2_kernel": "sql_tc_pred_kernel"}.get(k, k)
            traffic[k] = v["dram_bytes_per_launch"]
    except Exception:
        pass
    cand = {k: v for k, v in prof.items() if k in ab}
    dom = max(cand.items(), key=lambda kv: kv[1][1])[0] if cand else None
    roofline = None
    if dom is not None:
        cnt, tot_ms = prof[dom]
        achieved = ab[dom] / (tot_ms / cnt * 1e-3) / 1e9
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic.get(dom), "peak_source": peak_src,
                    "avg_launch_us": tot_ms / cnt * 1e3, "algorithmic_bytes_per_launch": ab[dom],
                    "share_of_kernel_time": tot_ms / sum(v[1] for v in prof.values()),
                    "timing": "CUDA events around each launch on the launching stream (eager submission of the same "
                              "steps after the timed region); traffic = dram bytes/launch from the committed ncu "
                              "capture profiles/traffic_latest.json"}
    step_ms_kernels = sum(v[1] for v in prof.values()) / max(1, nprof)
    kernels = {k: {"launches_per_step": v[0] / max(1, nprof),
                   "ms_per_step": v[1] / max(1, nprof),
                   "hbm_frac": (ab[k] / (v[1] / v[0] * 1e-3) / 1e9 / peak) if k in ab else None}
               for k, v in sorted(prof.items())}

    cpu = None
    if not args.no_cpu_baseline and world == 1:          # reported beside the N = 1 line only
        cpu = time_cpu(cfg, args, steps=3, warmup=1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    total_bytes = sum(v.numel() * 4 for v in hb.values())   # device-resident footprint (frames are fp32 on the device)
    config["l2"] = ("no explicit flush: every step streams %.0f MB of inputs plus %.0f MB of gradients through a 126 MB L2"
                    % (total_bytes / 1e6, (hb["x"].numel() * 4) / 1e6))
    config["submission"] = "eager" if args.no_graph else "cuda_graph"
    if world > 1:
        config["grad_exchange"] = ("one bucketed NCCL all-reduce (average) of the parameter gradients per step, issued inside "
                                   "the step on a communication stream when the last parameter gradient exists (overlaps the "
                                   "summary-path backward kernel; part of the CUDA graph); max rel. difference to a plain "
                                   "all-reduce after the step: %.2g" % exchange_check) if exchange_in_step else \
            "one bucketed NCCL all-reduce (average) of the parameter gradients after the step"
    config["e2e_pipeline"] = ("H2D of batch i+1 on a copy stream overlaps the step on batch i (2 device input sets); "
                              "tie-break noise drawn on the device instead of copied from the host; the loss of every step is "
                              "copied to pinned host memory and read by the host one step later; frames shipped as %s"
                              % ("float32" if args.f32_frames else "uint8 and scaled to [0,1] on the device"))
    line = {"metric": METRIC, "value": cfg.B * world / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": cfg.B * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(e2e_bytes), "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "kernels": kernels, "kernel_ms_per_step": step_ms_kernels, "loss": float(hp.loss)}
    print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
