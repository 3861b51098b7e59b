"""Shared helpers: load tests/golden/*.npz (reference outputs) and build seeded synthetic cases."""
import os

import numpy as np
import torch
import torch.nn.functional as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PHOTO_CASES = ["photo_mono_s0", "photo_mono_ms4", "photo_stereo_s0", "photo_mono_fullres",
               "photo_avg", "photo_noauto", "photo_nossim"]


def load_npz(name):
    with np.load(os.path.join(GOLD, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def photo_case(name, dtype=torch.float32, device="cpu"):
    """Golden photometric case -> kwargs for oracle.photometric_losses / sqlx.photometric_loss + expected."""
    z = load_npz(name)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(device=device, dtype=dtype)  # noqa: E731
    fids = [("s" if f == 99 else int(f)) for f in z["frame_ids"]]
    scales = [int(s) for s in z["scales"]]
    H, W = int(z["H"]), int(z["W"])
    disps = {s: t(z["in_disp%d" % s]).requires_grad_(True) for s in scales}
    target_pyr = {0: t(z["color_0"])}
    for s in scales:
        if s > 0:
            target_pyr[s] = t(z["color_0_s%d" % s])
    sources, poses, leaves = [], [], {}
    use_stereo = bool(z["use_stereo"])
    for f in fids[1:]:
        sources.append(t(z["color_%s" % f]))
        if f == "s":
            poses.append({"T": t(z["stereo_T"])})
        else:
            aa = t(z["in_axisangle_%d" % f]).requires_grad_(True)
            tr = t(z["in_translation_%d" % f]).requires_grad_(True)
            leaves["axisangle_%d" % f] = aa
            leaves["translation_%d" % f] = tr
            poses.append({"axisangle": aa, "translation": tr, "invert": f < 0})
    noises = {s: t(z["noise_s%d" % s]) for s in scales if ("noise_s%d" % s) in z}
    for s in scales:
        leaves["disp%d" % s] = disps[s]
    kw = dict(disps=disps, target_pyr=target_pyr, sources=sources, K=t(z["K"]), inv_K=t(z["inv_K"]),
              poses=poses, noises=noises, height=H, width=W, scales=tuple(scales),
              rescale_translation=not use_stereo, no_ssim=bool(z["no_ssim"]),
              avg_reprojection=bool(z["avg_reprojection"]), disable_automasking=bool(z["disable_automasking"]))
    return kw, leaves, z, fids


INDOOR_CASES = ["indoor_occ", "indoor_occ_avg"]


def indoor_case(name, dtype=torch.float32, device="cpu"):
    """Golden indoor case (trainer_indoor.py compute_losses_with_occ) -> kwargs for oracle.indoor_losses /
    sqlx.indoor_losses + the leaves whose gradients the fixture holds."""
    z = load_npz(name)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(device=device, dtype=dtype)  # noqa: E731
    leaves = {"disp0": t(z["in_disp0"]).requires_grad_(True)}
    sources, ref_depths, poses = [], [], []
    for f in (-1, 1):
        sources.append(t(z["color_%d" % f]))
        dr = t(z["in_depth_ref_%d" % f]).requires_grad_(True)
        aa = t(z["in_axisangle_%d" % f]).requires_grad_(True)
        tr = t(z["in_translation_%d" % f]).requires_grad_(True)
        leaves["depth_ref_%d" % f] = dr
        leaves["axisangle_%d" % f] = aa
        leaves["translation_%d" % f] = tr
        ref_depths.append(dr)
        poses.append({"axisangle": aa, "translation": tr, "invert": f < 0})
    kw = dict(disp=leaves["disp0"], target=t(z["color_0"]), sources=sources, ref_depths=ref_depths, K=t(z["K"]),
              inv_K=t(z["inv_K"]), poses=poses, noise=t(z["noise"]), height=int(z["H"]), width=int(z["W"]),
              reg_wt=float(z["reg_wt"]), no_ssim=bool(z["no_ssim"]), avg_reprojection=bool(z["avg_reprojection"]),
              disable_automasking=bool(z["disable_automasking"]))
    return kw, leaves, z


def smooth_images(g, B, H, W, n_frames, shift=2.5, noise=0.02):
    base = torch.rand(B, 3, H // 8 + 2, W // 8 + 2, generator=g)
    big = F.interpolate(base, size=(H + 16, W + 16), mode="bicubic", align_corners=False).clamp(0, 1)
    frames = []
    for i in range(n_frames):
        dx = int(round((i - (n_frames - 1) / 2) * shift))
        fr = big[:, :, 8:8 + H, 8 + dx:8 + dx + W]
        frames.append((fr + noise * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1).contiguous())
    return frames


def kitti_K(B, H, W):
    K = np.array([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    K[0, :] *= W
    K[1, :] *= H
    inv_K = np.linalg.pinv(K)
    return (torch.from_numpy(K).unsqueeze(0).repeat(B, 1, 1).contiguous(),
            torch.from_numpy(inv_K).unsqueeze(0).repeat(B, 1, 1).contiguous())


def depth_like(g, B, h, w, lo=2.0, hi=30.0):
    d = torch.rand(B, 1, max(h // 6, 2), max(w // 6, 2), generator=g)
    d = F.interpolate(d, size=(h, w), mode="bicubic", align_corners=False).clamp(0, 1)
    return (lo + (hi - lo) * d).contiguous()


def synth_photo_case(seed, B, H, W, S=2, scales=(0,), stereo=False, white_noise=False, full_res_disp=False):
    """Seeded synthetic photometric case (same recipe as oracle/make_golden.py) for oracle-vs-CUDA tests."""
    g = torch.Generator().manual_seed(seed)
    n_frames = S + 1
    if white_noise:
        frames = [torch.rand(B, 3, H, W, generator=g) for _ in range(n_frames)]
    else:
        frames = smooth_images(g, B, H, W, n_frames)
    mid = n_frames // 2
    target = frames[mid]
    sources = [frames[i] for i in range(n_frames) if i != mid]
    K, inv_K = kitti_K(B, H, W)
    target_pyr = {0: target}
    disps = {}
    for s in scales:
        if s > 0:
            target_pyr[s] = F.interpolate(target, [H // 2 ** s, W // 2 ** s], mode="bilinear", align_corners=False)
        hs, ws = ((H, W) if full_res_disp else (H // 2, W // 2)) if s == 0 else (H // 2 ** s, W // 2 ** s)
        disps[s] = depth_like(g, B, hs, ws)
    poses = []
    for i in range(S):
        if stereo and i == S - 1:
            st = torch.eye(4).unsqueeze(0).repeat(B, 1, 1)
            st[:, 0, 3] = 0.1
            poses.append({"T": st})
        else:
            poses.append({"axisangle": 0.01 * torch.randn(B, 1, 1, 3, generator=g),
                          "translation": 0.3 * torch.randn(B, 1, 1, 3, generator=g), "invert": i == 0})
    noises = {s: torch.randn(B, S, H, W, generator=g) for s in scales}
    return dict(disps=disps, target_pyr=target_pyr, sources=sources, K=K, inv_K=inv_K, poses=poses,
                noises=noises, height=H, width=W, scales=tuple(scales), rescale_translation=not stereo)
