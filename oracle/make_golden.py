"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE's own code on CPU.

TEST INFRASTRUCTURE ONLY; runs only where /root/reference is mounted (the build
container).  Re-run with:   python oracle/make_golden.py
The fixtures are small seeded synthetic cases fed through
  - trainer.Trainer.generate_images_pred + compute_losses  (trainer.py:386-549)
  - networks.Depth_Decoder_QueryTr / Lite_Depth_Decoder_QueryTr forward (depth_decoder_QTR.py:36-74)
  - layers.SSIM, BackprojectDepth, Project3D, get_smooth_loss, transformation_from_parameters
  - finetune/loss.py SILogLoss
  - trainer_indoor.Trainer.generate_images_pred + compute_losses_with_occ  (trainer_indoor.py:512-719)
and store inputs, outputs and autograd gradients.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def smooth_images(g, B, H, W, n_frames, shift=2.5, noise=0.02):
    """KITTI-like smooth frames: one bicubic-upsampled random base, shifted per frame + a little noise."""
    base = torch.rand(B, 3, H // 8 + 2, W // 8 + 2, generator=g)
    big = F.interpolate(base, size=(H + 16, W + 16), mode="bicubic", align_corners=False).clamp(0, 1)
    frames = []
    for i in range(n_frames):
        dx = int(round((i - (n_frames - 1) / 2) * shift))
        fr = big[:, :, 8:8 + H, 8 + dx:8 + dx + W]
        fr = (fr + noise * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1)
        frames.append(fr.contiguous())
    return frames


def kitti_K(B, H, W):
    K = np.array([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    K[0, :] *= W
    K[1, :] *= H
    inv_K = np.linalg.pinv(K)
    return (torch.from_numpy(K).unsqueeze(0).repeat(B, 1, 1),
            torch.from_numpy(inv_K).unsqueeze(0).repeat(B, 1, 1))


def depth_like(g, B, h, w, lo=2.0, hi=30.0):
    d = torch.rand(B, 1, max(h // 6, 2), max(w // 6, 2), generator=g)
    d = F.interpolate(d, size=(h, w), mode="bicubic", align_corners=False).clamp(0, 1)
    return (lo + (hi - lo) * d).contiguous()


def photometric_case(name, seed, B, H, W, scales, use_stereo=False, extra_args=(), half_res_scale0=True):
    ref = ref_shim.load()
    g = torch.Generator().manual_seed(seed)
    frame_ids = [0, -1, 1]
    T = ref_shim.make_trainer(B, H, W, scales=scales, frame_ids=frame_ids, use_stereo=use_stereo,
                              extra_args=extra_args)
    fids = T.opt.frame_ids                         # [0,-1,1] (+ "s")
    frames = smooth_images(g, B, H, W, len(fids))
    K, inv_K = kitti_K(B, H, W)
    inputs = {("K", 0): K, ("inv_K", 0): inv_K}
    order = [0] + [f for f in fids if f != 0]
    # frame 0 is the middle one so that -1 / +1 are shifted either way
    mid = len(frames) // 2
    perm = [mid] + [i for i in range(len(frames)) if i != mid]
    for f, idx in zip(order, perm):
        inputs[("color", f, 0)] = frames[idx]
    for s in scales:
        if s > 0:
            inputs[("color", 0, s)] = F.interpolate(inputs[("color", 0, 0)], [H // 2 ** s, W // 2 ** s],
                                                    mode="bilinear", align_corners=False)
    if use_stereo:
        st = torch.eye(4).unsqueeze(0).repeat(B, 1, 1)
        st[:, 0, 3] = torch.tensor([0.1 if b % 2 == 0 else -0.1 for b in range(B)])
        inputs["stereo_T"] = st
    outputs = {}
    leaves = {}
    for s in scales:
        if s == 0 and half_res_scale0:
            hs, ws = H // 2, W // 2
        else:
            hs, ws = H // 2 ** s, W // 2 ** s
        d = depth_like(g, B, hs, ws).requires_grad_(True)
        outputs[("disp", s)] = d
        leaves["disp%d" % s] = d
    for f in fids[1:]:
        if f == "s":
            continue
        aa = (0.01 * torch.randn(B, 1, 1, 3, generator=g)).requires_grad_(True)
        tr = (0.3 * torch.randn(B, 1, 1, 3, generator=g)).requires_grad_(True)
        outputs[("axisangle", 0, f)] = aa
        outputs[("translation", 0, f)] = tr
        outputs[("cam_T_cam", 0, f)] = ref.layers.transformation_from_parameters(aa[:, 0], tr[:, 0], invert=(f < 0))
        leaves["axisangle_%d" % f] = aa
        leaves["translation_%d" % f] = tr
    S = len(fids) - 1
    noise_seed = seed + 1000
    # the reference draws torch.randn(identity.shape) once per scale from the global CPU generator
    torch.manual_seed(noise_seed)
    n_noise_ch = 1 if T.opt.avg_reprojection else S
    noises = [torch.randn(B, n_noise_ch, H, W) for _ in scales] if not T.opt.disable_automasking else []
    T.generate_images_pred(inputs, outputs)
    torch.manual_seed(noise_seed)
    losses = T.compute_losses(inputs, outputs)
    names = list(leaves.keys())
    grads = torch.autograd.grad(losses["loss"], [leaves[n] for n in names], allow_unused=True)
    rec = {"B": B, "H": H, "W": W, "scales": np.array(scales), "use_stereo": int(use_stereo),
           "frame_ids": np.array([(99 if f == "s" else f) for f in fids]),
           "avg_reprojection": int(T.opt.avg_reprojection), "disable_automasking": int(T.opt.disable_automasking),
           "no_ssim": int(T.opt.no_ssim), "K": K.numpy(), "inv_K": inv_K.numpy()}
    for f in fids:
        rec["color_%s" % f] = inputs[("color", f, 0)].numpy()
    for s in scales:
        if s > 0:
            rec["color_0_s%d" % s] = inputs[("color", 0, s)].numpy()
    if use_stereo:
        rec["stereo_T"] = inputs["stereo_T"].numpy()
    for i, s in enumerate(scales):
        if noises:
            rec["noise_s%d" % s] = noises[i].numpy()
        rec["out_loss_s%d" % s] = losses["loss/%d" % s].detach().numpy()
        rec["out_depth_s%d" % s] = outputs[("depth", 0, s)].detach().numpy()
        if not T.opt.disable_automasking:
            rec["out_idsel_s%d" % s] = outputs["identity_selection/%d" % s].numpy().astype(np.uint8)
        for f in fids[1:]:
            if s == scales[0]:
                rec["out_sample_%s_s%d" % (f, s)] = outputs[("sample", f, s)].detach().numpy()
                rec["out_color_%s_s%d" % (f, s)] = outputs[("color", f, s)].detach().numpy()
    rec["out_loss"] = losses["loss"].detach().numpy()
    for n, gr in zip(names, grads):
        rec["in_" + n] = leaves[n].detach().numpy()
        rec["grad_" + n] = (gr if gr is not None else torch.zeros_like(leaves[n])).numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(name, "loss", float(losses["loss"]))


def decoder_case(name, seed, lite, B, E, h, w, P, Q, D, min_val, max_val):
    ref = ref_shim.load()
    torch.manual_seed(seed)
    cls = ref.networks.Lite_Depth_Decoder_QueryTr if lite else ref.networks.Depth_Decoder_QueryTr
    dec = cls(in_channels=E, embedding_dim=E, patch_size=P, num_heads=4, query_nums=Q, dim_out=D,
              min_val=min_val, max_val=max_val)
    dec.eval()
    g = torch.Generator().manual_seed(seed + 1)
    x0 = torch.randn(B, E, h, w, generator=g)
    cap = {}
    h1 = dec.conv3x3.register_forward_hook(lambda m, i, o: cap.__setitem__("x", o))
    h2 = dec.transformer_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("tokens", o))
    h3 = dec.full_query_layer.register_forward_hook(lambda m, i, o: cap.__setitem__("fq", o))
    h4 = dec.bins_regressor.register_forward_hook(lambda m, i, o: cap.__setitem__("raw", o))
    out = dec(x0)
    for hh in (h1, h2, h3, h4):
        hh.remove()
    pred = out[("disp", 0)]
    x = cap["x"]
    queries = cap["tokens"][:Q].permute(1, 0, 2)
    energy, summary = cap["fq"]
    gout = torch.randn(pred.shape, generator=g)
    Wp = dec.convert_to_prob[0].weight
    bp = dec.convert_to_prob[0].bias
    x.retain_grad()
    cap["tokens"].retain_grad()
    summary.retain_grad()
    cap["raw"].retain_grad()
    (pred * gout).sum().backward()
    dq = cap["tokens"].grad[:Q].permute(1, 0, 2)
    rec = dict(B=B, E=E, h=h, w=w, P=P, Q=Q, D=D, min_val=min_val, max_val=max_val, lite=int(lite),
               x0=x0.numpy(), x=x.detach().numpy(), queries=queries.detach().numpy(),
               # weights live in <name>_state.npz (bins_regressor.{0,2,4}.*, convert_to_prob.0.*)
               out_summary=summary.detach().numpy(), out_raw=cap["raw"].detach().numpy(),
               out_pred=pred.detach().numpy(), out_energy_sample=energy.detach().numpy()[:, :, ::4, ::4],
               gout=gout.numpy(), grad_x=x.grad.numpy(), grad_queries=dq.numpy(),
               grad_Wp=Wp.grad.numpy().reshape(D, Q), grad_bp=bp.grad.numpy(),
               grad_summary=summary.grad.numpy(), grad_raw=cap["raw"].grad.numpy())
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    sd = {k: v.numpy() for k, v in dec.state_dict().items()}
    np.savez_compressed(os.path.join(GOLD, name + "_state.npz"), **sd)
    print(name, "pred range", float(pred.min()), float(pred.max()))


def module_case(name, seed, B, H, W):
    """Module-level drop-ins: SSIM, BackprojectDepth, Project3D, get_smooth_loss, pose matrix, SILog."""
    ref = ref_shim.load()
    g = torch.Generator().manual_seed(seed)
    a, b = smooth_images(g, B, H, W, 2)
    a.requires_grad_(True)
    ss = ref.layers.SSIM()(a, b)
    gs = torch.randn(ss.shape, generator=g)
    (ga,) = torch.autograd.grad((ss * gs).sum(), a)
    K, inv_K = kitti_K(B, H, W)
    depth = depth_like(g, B, H, W).requires_grad_(True)
    aa = (0.02 * torch.randn(B, 1, 3, generator=g)).requires_grad_(True)
    tr = (0.2 * torch.randn(B, 1, 3, generator=g)).requires_grad_(True)
    Tm = ref.layers.transformation_from_parameters(aa, tr, invert=False)
    Tinv = ref.layers.transformation_from_parameters(aa, tr, invert=True)
    pts = ref.layers.BackprojectDepth(B, H, W)(depth, inv_K)
    grid = ref.layers.Project3D(B, H, W)(pts, K, Tm)
    gg = torch.randn(grid.shape, generator=g)
    gd, gaa, gtr = torch.autograd.grad((grid * gg).sum(), [depth, aa, tr])
    disp = depth_like(g, B, H, W, 0.5, 1.5).requires_grad_(True)
    sm = ref.layers.get_smooth_loss(disp, b)
    (gdisp,) = torch.autograd.grad(sm, disp)
    # SILog (finetune/loss.py:24-42), loaded by path because finetune/ is not a package on sys.path
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_ft_loss", os.path.join(ref_shim.REFERENCE_ROOT, "finetune", "loss.py"))
    ft = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ft)
    pred_lr = depth_like(g, B, H // 2, W // 2).requires_grad_(True)
    gt = depth_like(g, B, H, W) * (torch.rand(B, 1, H, W, generator=g) > 0.4)
    mask = gt > 1e-3
    sil = ft.SILogLoss()(pred_lr, gt, mask=mask, interpolate=True)
    (gsil,) = torch.autograd.grad(sil, pred_lr)
    rec = dict(B=B, H=H, W=W, a=a.detach().numpy(), b=b.numpy(), out_ssim=ss.detach().numpy(), g_ssim=gs.numpy(),
               grad_a=ga.numpy(), K=K.numpy(), inv_K=inv_K.numpy(), depth=depth.detach().numpy(),
               axisangle=aa.detach().numpy(), translation=tr.detach().numpy(), out_T=Tm.detach().numpy(),
               out_T_inv=Tinv.detach().numpy(), out_points=pts.detach().numpy(), out_grid=grid.detach().numpy(),
               g_grid=gg.numpy(), grad_depth=gd.numpy(), grad_axisangle=gaa.numpy(), grad_translation=gtr.numpy(),
               disp=disp.detach().numpy(), out_smooth=sm.detach().numpy(), grad_disp=gdisp.numpy(),
               silog_pred=pred_lr.detach().numpy(), silog_gt=gt.numpy(), out_silog=sil.detach().numpy(),
               grad_silog=gsil.numpy())
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(name, "ssim mean", float(ss.mean()), "silog", float(sil))


def indoor_case(name, seed, B, H, W, extra_args=()):
    """trainer_indoor.Trainer.generate_images_pred + compute_losses_with_occ (trainer_indoor.py:512-599, 615-719)
    with --use_improved_mini_reproj_loss, scales = [0]."""
    ref = ref_shim.load()
    import trainer_indoor
    g = torch.Generator().manual_seed(seed)
    base = ref_shim.make_trainer(B, H, W, scales=[0], frame_ids=[0, -1, 1],
                                 extra_args=["--use_improved_mini_reproj_loss"] + list(extra_args))
    T = trainer_indoor.Trainer.__new__(trainer_indoor.Trainer)
    T.__dict__.update(base.__dict__)
    fids = T.opt.frame_ids
    frames = smooth_images(g, B, H, W, 3)
    # a dark band in one source so that valid_mask (mean |pred| > 1e-3) is exercised
    frames[0][:, :, H // 3:H // 3 + 5, W // 4:W // 4 + 12] = 0.0
    K, inv_K = kitti_K(B, H, W)
    inputs = {("K", 0): K, ("inv_K", 0): inv_K, ("color", 0, 0): frames[1], ("color", -1, 0): frames[0],
              ("color", 1, 0): frames[2]}
    outputs, leaves = {}, {}
    d = depth_like(g, B, H // 2, W // 2).requires_grad_(True)
    outputs[("disp", 0)] = d
    leaves["disp0"] = d
    for f in fids[1:]:
        # depth of the reference frames: close to the target depth (the consistency term compares them)
        dr = (F.interpolate(d.detach(), [H, W], mode="bilinear", align_corners=False) *
              (1.0 + 0.2 * (torch.rand(B, 1, H, W, generator=g) - 0.5))).requires_grad_(True)
        outputs[("depth_ref", f, 0)] = dr
        leaves["depth_ref_%d" % f] = dr
        aa = (0.01 * torch.randn(B, 1, 1, 3, generator=g)).requires_grad_(True)
        tr = (0.3 * torch.randn(B, 1, 1, 3, generator=g)).requires_grad_(True)
        outputs[("axisangle", 0, f)] = aa
        outputs[("translation", 0, f)] = tr
        outputs[("cam_T_cam", 0, f)] = ref.layers.transformation_from_parameters(aa[:, 0], tr[:, 0], invert=(f < 0))
        leaves["axisangle_%d" % f] = aa
        leaves["translation_%d" % f] = tr
    S = 2
    noise_seed = seed + 1000
    torch.manual_seed(noise_seed)
    noise = torch.randn(B, 1 if T.opt.avg_reprojection else S, H, W)
    T.generate_images_pred(inputs, outputs)
    torch.manual_seed(noise_seed)
    total, losses = T.compute_losses_with_occ(inputs, outputs)
    total = total / T.num_scales                                   # trainer_indoor.py:413
    names = list(leaves.keys())
    grads = torch.autograd.grad(total, [leaves[n] for n in names], allow_unused=True)
    rec = {"B": B, "H": H, "W": W, "K": K.numpy(), "inv_K": inv_K.numpy(), "noise": noise.numpy(),
           "reg_wt": np.float32(T.opt.reg_wt), "avg_reprojection": int(T.opt.avg_reprojection),
           "disable_automasking": int(T.opt.disable_automasking), "no_ssim": int(T.opt.no_ssim),
           "out_loss": total.detach().numpy(), "out_loss_s0": losses["loss/0"].detach().numpy(),
           "out_depth_s0": outputs[("depth", 0, 0)].detach().numpy()}
    for f in fids:
        rec["color_%s" % f] = inputs[("color", f, 0)].numpy()
    for f in fids[1:]:
        rec["out_color_%s" % f] = outputs[("color", f, 0)].detach().numpy()
        rec["out_pred_dep_%s" % f] = outputs[("pred_dep", f, 0)].detach().numpy()
    for n, gr in zip(names, grads):
        rec["in_" + n] = leaves[n].detach().numpy()
        rec["grad_" + n] = (gr if gr is not None else torch.zeros_like(leaves[n])).numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(name, "loss", float(total))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "indoor":      # only the fixtures added for SURVEY 8f row N4
        indoor_case("indoor_occ", 41, B=2, H=48, W=80)
        indoor_case("indoor_occ_avg", 42, B=1, H=48, W=80, extra_args=["--avg_reprojection"])
        return
    photometric_case("photo_mono_s0", 11, B=2, H=48, W=80, scales=[0])
    photometric_case("photo_mono_ms4", 12, B=1, H=64, W=96, scales=[0, 1, 2, 3])
    photometric_case("photo_stereo_s0", 13, B=2, H=48, W=80, scales=[0], use_stereo=True)
    photometric_case("photo_mono_fullres", 14, B=1, H=48, W=80, scales=[0], half_res_scale0=False)
    photometric_case("photo_avg", 15, B=1, H=48, W=80, scales=[0], extra_args=["--avg_reprojection"])
    photometric_case("photo_noauto", 16, B=1, H=48, W=80, scales=[0], extra_args=["--disable_automasking"])
    photometric_case("photo_nossim", 17, B=1, H=48, W=80, scales=[0], extra_args=["--no_ssim"])
    decoder_case("decoder_full", 21, lite=False, B=2, E=32, h=24, w=40, P=8, Q=12, D=16, min_val=0.001, max_val=80.0)
    decoder_case("decoder_lite", 22, lite=True, B=2, E=32, h=32, w=32, P=8, Q=16, D=24, min_val=0.01, max_val=80.0)
    module_case("modules", 31, B=2, H=40, W=72)
    indoor_case("indoor_occ", 41, B=2, H=48, W=80)
    indoor_case("indoor_occ_avg", 42, B=1, H=48, W=80, extra_args=["--avg_reprojection"])


if __name__ == "__main__":
    main()
