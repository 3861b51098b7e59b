"""The BASELINE.json workloads as seeded synthetic batches, shared by bench.py and the full-size parity tests.

  baseline_config(n)     HotPathConfig of BASELINE config n (SURVEY 8 table "Config -> shapes": arg files
                         args_files/hisfog/kitti/resnet_192x640.txt, resnet_320x1024.txt, effb5_320x1024.txt)
  make_host_batch        seeded KITTI-shape host batch (SURVEY 8d recipe)
  head_state             default-init weights of the decoder's 1x1 conv and bins MLP under torch.manual_seed(0)
  oracle_step            the same step through oracle/sqldepth_oracle.py (TEST INFRASTRUCTURE: the checker), any dtype,
                         returning loss, depth and every gradient
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from _cases import smooth_images, kitti_K, depth_like  # noqa: E402

# BASELINE.json configs -> (B per GPU, H, W, decoder map h x w, Q, D, S, loss scales, min_depth, stereo)
#   2: ResNet-50 192x640, batch 12, 4 loss scales (3-frame sequence: S = 2)
#   3: ResNet-50 320x1024, batch 8/GPU, --use_stereo (S = 3, last source = stereo frame), patch 20 -> Q = D = 128
#   4: EfficientNet-B5 320x1024, batch 8/GPU, decoder map at full resolution (h = H, w = W), Q = D = 128,
#      multi-scale warp / SSIM (4 loss scales), --use_stereo
#   5: ConvNeXt-L 320x1024 metric fine-tune (SILog): handled by the fine-tune workload, not a HotPathConfig
BASELINE_CONFIGS = {
    2: dict(B=12, H=192, W=640, h=96, w=320, Q=64, D=64, S=2, scales=(0, 1, 2, 3), min_depth=0.001, stereo=False),
    3: dict(B=8, H=320, W=1024, h=160, w=512, Q=128, D=128, S=3, scales=(0,), min_depth=0.01, stereo=True),
    4: dict(B=8, H=320, W=1024, h=320, w=1024, Q=128, D=128, S=3, scales=(0, 1, 2, 3), min_depth=0.001, stereo=True),
}


def baseline_config(n, B=None, **over):
    from sqlx.hotpath import HotPathConfig
    kw = dict(BASELINE_CONFIGS[n])
    if B is not None:
        kw["B"] = B
    kw.update(over)
    return HotPathConfig(E=32, max_depth=80.0, **kw)


def make_host_batch(cfg, seed, pin=False, u8_frames=False):
    """Seeded KITTI-shape synthetic batch on the host (SURVEY 8d recipe): smooth frames, KITTI intrinsics,
    PoseCNN-scale poses, decoder-feature-like x and queries.  Frames are quantised to 8 bits (as decoded images
    are); with u8_frames they stay uint8 on the host and are scaled to [0,1] on the device by HotPath.load."""
    g = torch.Generator().manual_seed(seed)
    c = cfg
    frames = smooth_images(g, c.B, c.H, c.W, c.S + 1)
    mid = (c.S + 1) // 2 if not c.stereo else c.S // 2
    if c.stereo:
        # temporal frames [-1, 0, +1] and the stereo frame last
        order = [j for j in range(c.S) if j != mid] + [c.S]
    else:
        order = [j for j in range(c.S + 1) if j != mid]
    hb = {"target": frames[mid]}
    for i, j in enumerate(order):
        hb["source%d" % i] = frames[j]
    hb["K"], hb["inv_K"] = kitti_K(c.B, c.H, c.W)
    hb["x"] = torch.randn(c.B, c.E, c.h, c.w, generator=g)
    hb["queries"] = 0.4 * torch.randn(c.B, c.Q, c.E, generator=g)
    for i in c.pose_sources:
        hb["axisangle%d" % i] = 0.01 * torch.randn(c.B, 1, 1, 3, generator=g)
        hb["translation%d" % i] = 0.01 * torch.randn(c.B, 1, 1, 3, generator=g)
    if c.stereo:
        st = torch.eye(4).unsqueeze(0).repeat(c.B, 1, 1)
        st[:, 0, 3] = 0.1                                      # datasets/mono_dataset.py:193-199
        hb["stereo_T"] = st
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


def head_state(cfg, seed=0):
    """{state_dict key: tensor} of convert_to_prob.0 and bins_regressor with the reference decoder's shapes
    (networks/depth_decoder_QTR.py:22-28) under the default init and torch.manual_seed(seed)."""
    c = cfg
    nn = torch.nn
    torch.manual_seed(seed)
    conv = nn.Conv2d(c.Q, c.D, 1)
    mlp = nn.Sequential(nn.Linear(c.E * c.Q, 16 * c.Q), nn.LeakyReLU(), nn.Linear(16 * c.Q, 256), nn.LeakyReLU(),
                        nn.Linear(256, c.D))
    state = {"convert_to_prob.0.weight": conv.weight.detach().clone(), "convert_to_prob.0.bias": conv.bias.detach().clone()}
    for k, v in mlp.state_dict().items():
        state["bins_regressor." + k] = v.detach().clone()
    return state


GRAD_INPUT_PREFIXES = ("x", "queries", "disp", "axisangle", "translation")


def oracle_step(cfg, hb, state, dtype=torch.float64, backward=True, device="cpu"):
    """One step of the workload through the oracle (torch restatement of the reference, oracle/sqldepth_oracle.py):
    SQL tail -> photometric losses of every scale -> autograd backward.  Returns
    {"loss", "pred", "losses": {s: .}, "argmin": {s: [B,H,W]}, "grads": {name: tensor}} with grads for every
    differentiable input (x, queries, disp<s>, axisangle<i>, translation<i>) and every parameter (state_dict key)."""
    from oracle import sqldepth_oracle as O
    c = cfg
    cast = lambda t: t.to(device=device, dtype=dtype)  # noqa: E731
    inp = {k: cast(v.float() / 255.0 if v.dtype == torch.uint8 else v) for k, v in hb.items()}
    leaves = {k: inp[k].clone().requires_grad_(backward) for k in inp if k.rstrip("0123456789") in GRAD_INPUT_PREFIXES}
    params = {k: cast(v).clone().requires_grad_(backward) for k, v in state.items()}
    mlp = [params["bins_regressor.%d.%s" % (i, k)] for i in (0, 2, 4) for k in ("weight", "bias")]
    Wp = params["convert_to_prob.0.weight"].view(c.D, c.Q)
    ctx = torch.enable_grad() if backward else torch.no_grad()
    with ctx:
        tail = O.sql_tail(leaves["x"], leaves["queries"], mlp, Wp, params["convert_to_prob.0.bias"], c.min_depth,
                          c.max_depth)
        disps = {s: (tail["pred"] if s == 0 else leaves["disp%d" % s]) for s in c.scales}
        target_pyr = {s: (inp["target"] if s == 0 else inp["target%d" % s]) for s in c.scales}
        poses = [{"axisangle": leaves["axisangle%d" % i], "translation": leaves["translation%d" % i], "invert": i == 0}
                 for i in c.pose_sources]
        if c.stereo:
            poses.append({"T": inp["stereo_T"]})
        out = O.photometric_losses(disps, target_pyr, [inp["source%d" % i] for i in range(c.S)], inp["K"], inp["inv_K"],
                                   poses, {s: inp["noise%d" % s] for s in c.scales}, height=c.H, width=c.W,
                                   scales=c.scales, disparity_smoothness=c.disparity_smoothness,
                                   rescale_translation=not c.stereo, disable_automasking=not c.automask)
        res = {"loss": out["loss"].detach(), "pred": tail["pred"].detach(),
               "losses": {s: out["loss/%d" % s].detach() for s in c.scales},
               "identity_selection": {s: out["identity_selection/%d" % s] for s in c.scales
                                      if ("identity_selection/%d" % s) in out}}
        if backward:
            names = list(leaves) + list(params)
            grads = torch.autograd.grad(out["loss"], [leaves[k] for k in leaves] + [params[k] for k in params],
                                        allow_unused=True)
            res["grads"] = {k: g for k, g in zip(names, grads)}
    return res


def oracle_step_chunked(cfg, hb, state, chunk=2, dtype=torch.float64, backward=True):
    """oracle_step evaluated `chunk` samples at a time (bounded memory at the 320x1024 / Q = D = 128 shapes) and
    recombined: every loss term is a per-sample mean followed by a batch mean (trainer.py:532,535,546; layers.py:280),
    so with equal chunks  loss = mean_k loss_k,  d loss / d input_k = (d loss_k / d input_k) / n_chunks  and the
    parameter gradients are the chunk gradients averaged."""
    from sqlx.hotpath import HotPathConfig
    assert cfg.B % chunk == 0
    nch = cfg.B // chunk
    kw = dict(B=chunk, H=cfg.H, W=cfg.W, h=cfg.h, w=cfg.w, E=cfg.E, Q=cfg.Q, D=cfg.D, S=cfg.S, scales=cfg.scales,
              min_depth=cfg.min_depth, max_depth=cfg.max_depth, disparity_smoothness=cfg.disparity_smoothness,
              stereo=cfg.stereo, automask=cfg.automask)
    ccfg = HotPathConfig(**kw)
    parts = []
    for k in range(nch):
        sub = {key: v[k * chunk:(k + 1) * chunk] for key, v in hb.items()}
        parts.append(oracle_step(ccfg, sub, state, dtype=dtype, backward=backward))
    res = {"loss": sum(p["loss"] for p in parts) / nch,
           "pred": torch.cat([p["pred"] for p in parts], 0),
           "losses": {s: sum(p["losses"][s] for p in parts) / nch for s in cfg.scales},
           "identity_selection": {s: torch.cat([p["identity_selection"][s] for p in parts], 0)
                                  for s in parts[0]["identity_selection"]}}
    if backward:
        g = {}
        for name in parts[0]["grads"]:
            gs = [p["grads"][name] for p in parts]
            if any(x is None for x in gs):
                g[name] = None
            elif name in state:
                g[name] = sum(gs) / nch
            else:
                g[name] = torch.cat(gs, 0) / nch
        res["grads"] = g
    return res
