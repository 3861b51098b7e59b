"""Indoor loss variant (trainer_indoor.py:512-599 + 615-719, --use_improved_mini_reproj_loss):
the CUDA path through the C ABI (sqlx_photo_occ_fwd/bwd) against the reference-generated fixtures and against the
float64 oracle on seeded inputs.  Loss 1e-5 absolute; gradients relative to max|grad| + cosine (argmin ties excluded
exactly as in test_photometric_gpu.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _cases import INDOOR_CASES, depth_like, indoor_case, kitti_K, smooth_images
from test_photometric_gpu import _cos, _mask_flips, _rel

pytestmark = pytest.mark.gpu


def _to_dev(kw, dtype=torch.float32, dev="cuda"):
    def leaf(t):
        return t.detach().to(device=dev, dtype=dtype).requires_grad_(True)

    def plain(t):
        return t.detach().to(device=dev, dtype=dtype)
    out = dict(kw)
    out["disp"] = leaf(kw["disp"])
    out["target"] = plain(kw["target"])
    out["sources"] = [plain(s) for s in kw["sources"]]
    out["ref_depths"] = [leaf(r) for r in kw["ref_depths"]]
    out["K"], out["inv_K"], out["noise"] = plain(kw["K"]), plain(kw["inv_K"]), plain(kw["noise"])
    out["poses"] = [{"axisangle": leaf(p["axisangle"]), "translation": leaf(p["translation"]), "invert": p["invert"]}
                    for p in kw["poses"]]
    return out


def _leaves(kw):
    names, leaves = ["disp0"], [kw["disp"]]
    for i, p in enumerate(kw["poses"]):
        names += ["depth_ref_%d" % i, "axisangle_%d" % i, "translation_%d" % i]
        leaves += [kw["ref_depths"][i], p["axisangle"], p["translation"]]
    return names, leaves


def _compare_grads(names, got, ref, flips, tol):
    for n, a, b in zip(names, got, ref):
        a, b = a.detach().cpu().double(), b.detach().cpu().double()
        if n == "disp0":
            a, b = _mask_flips(a, b, flips)
        elif n.startswith("depth_ref"):
            # full-resolution map: cells within the SSIM window / bilinear tap of a flipped pixel are excluded
            if flips is not None and bool(flips.any()):
                m = F.max_pool2d(flips[:, None].float(), 9, 1, 4) > 0
                a, b = a.masked_fill(m, 0.0), b.masked_fill(m, 0.0)
        assert _rel(a, b) < tol, n
        assert _cos(a, b) > 0.999, n


@pytest.mark.parametrize("name", INDOOR_CASES)
def test_indoor_golden(name):
    import sqlx
    from oracle import sqldepth_oracle as O
    kw, leaves, z = indoor_case(name)
    g = _to_dev(kw)
    out = sqlx.indoor_losses(**g)
    assert abs(float(out["loss"]) - float(z["out_loss"])) < 1e-5
    assert abs(float(out["loss/0"]) - float(z["out_loss_s0"])) < 1e-5
    ref = O.indoor_losses(**kw)                      # (CPU, fp32) only for the arg-min map the fixture does not hold
    flips = out[("argmin", 0)].cpu().long() != ref[("argmin", 0)]
    assert float(flips.float().mean()) < 2e-3
    names, gl = _leaves(g)
    got = torch.autograd.grad(out["loss"], gl)
    fid = {0: -1, 1: 1}
    want = []
    for n in names:
        if n == "disp0":
            want.append(torch.from_numpy(z["grad_disp0"]))
        else:
            base, i = n.rsplit("_", 1)
            want.append(torch.from_numpy(z["grad_%s_%d" % (base, fid[int(i)])]))
    _compare_grads(names, got, want, flips, 2e-2)


def _synth(seed, B, H, W, S=2, full_res_disp=False):
    g = torch.Generator().manual_seed(seed)
    frames = smooth_images(g, B, H, W, S + 1)
    mid = (S + 1) // 2
    target = frames[mid]
    sources = [f for i, f in enumerate(frames) if i != mid]
    sources[0][:, :, H // 4:H // 4 + 6, W // 3:W // 3 + 9] = 0.0           # exercises valid_mask
    K, inv_K = kitti_K(B, H, W)
    h, w = (H, W) if full_res_disp else (H // 2, W // 2)
    disp = depth_like(g, B, h, w)
    up = F.interpolate(disp, [H, W], mode="bilinear", align_corners=False)
    ref_depths = [up * (1.0 + 0.3 * (torch.rand(B, 1, H, W, generator=g) - 0.5)) for _ in range(S)]
    poses = [{"axisangle": 0.01 * torch.randn(B, 1, 1, 3, generator=g),
              "translation": 0.3 * torch.randn(B, 1, 1, 3, generator=g), "invert": i == 0} for i in range(S)]
    noise = torch.randn(B, S, H, W, generator=g)
    return dict(disp=disp, target=target, sources=sources, ref_depths=ref_depths, K=K, inv_K=inv_K, poses=poses,
                noise=noise, height=H, width=W)


@pytest.mark.parametrize("cfg", [
    dict(seed=51, B=2, H=64, W=96),
    dict(seed=52, B=1, H=50, W=70),                       # ragged: not a multiple of the tile
    dict(seed=53, B=1, H=64, W=96, S=3),
    dict(seed=54, B=1, H=48, W=80, full_res_disp=True),
])
@pytest.mark.parametrize("variant", [{}, {"disable_automasking": True}, {"no_ssim": True}])
def test_indoor_oracle_fp64(cfg, variant):
    import sqlx
    from oracle import sqldepth_oracle as O
    kw = _synth(**cfg)
    kw.update(variant)
    g = _to_dev(kw)
    out = sqlx.indoor_losses(**g)
    kd = _to_dev(kw, dtype=torch.float64, dev="cpu")
    ref = O.indoor_losses(**kd)
    assert abs(float(out["loss"]) - float(ref["loss"])) < 1e-5
    flips = out[("argmin", 0)].cpu().long() != ref[("argmin", 0)]
    assert float(flips.float().mean()) < 2e-3
    names, gl = _leaves(g)
    _, rl = _leaves(kd)
    got = torch.autograd.grad(out["loss"], gl)
    want = torch.autograd.grad(ref["loss"], rl)
    _compare_grads(names, got, want, flips, 3e-2)


def test_indoor_abi_validation():
    """bad arguments come back as SqlxError through the C ABI, not as a crash"""
    import sqlx
    kw = _to_dev(_synth(seed=55, B=1, H=48, W=64))
    kw["ref_depths"] = [r[:, :, ::2, ::2].contiguous() for r in kw["ref_depths"]]    # wrong shape
    with pytest.raises((AssertionError, sqlx.SqlxError)):
        sqlx.indoor_losses(**kw)


def test_indoor_trainer_mixin_dict_contract():
    """sqlx.IndoorFusedLossMixin: generate_images_pred / compute_losses_with_occ with trainer_indoor.py's signatures,
    dict keys and (total, losses) return value, against the reference's own results."""
    import types
    import sqlx
    kw, leaves, z = indoor_case("indoor_occ")
    g = _to_dev(kw)

    class T(sqlx.IndoorFusedLossMixin):
        pass
    tr = T()
    fids = [0, -1, 1]
    tr.num_scales = 1
    tr.opt = types.SimpleNamespace(scales=[0], frame_ids=fids, height=kw["height"], width=kw["width"],
                                   pose_model_type="posecnn", use_stereo=False, no_ssim=False, avg_reprojection=False,
                                   disable_automasking=False, disparity_smoothness=1e-3, v1_multiscale=False,
                                   predictive_mask=False, use_improved_mini_reproj_loss=True, use_rectify_net=False,
                                   reg_wt=kw["reg_wt"])
    tr.sqlx_noises = {0: g["noise"]}
    inputs = {("K", 0): g["K"], ("inv_K", 0): g["inv_K"], ("color", 0, 0): g["target"]}
    outputs = {("disp", 0): g["disp"]}
    for f, src, ref, pose in zip(fids[1:], g["sources"], g["ref_depths"], g["poses"]):
        inputs[("color", f, 0)] = src
        outputs[("depth_ref", f, 0)] = ref
        outputs[("axisangle", 0, f)], outputs[("translation", 0, f)] = pose["axisangle"], pose["translation"]
    tr.generate_images_pred(inputs, outputs)
    total, losses = tr.compute_losses_with_occ(inputs, outputs)
    assert abs(float(total) - float(z["out_loss"])) < 1e-5
    assert abs(float(losses["loss/0"]) - float(z["out_loss_s0"])) < 1e-5
    np.testing.assert_allclose(outputs[("depth", 0, 0)].cpu().numpy(), z["out_depth_s0"], rtol=1e-5, atol=1e-5)
    for f in fids[1:]:
        np.testing.assert_allclose(outputs[("color", f, 0)].cpu().numpy(), z["out_color_%d" % f], atol=1e-4)
        assert outputs[("color_identity", f, 0)] is inputs[("color", f, 0)]
    total.backward()
    for t in [g["disp"]] + g["ref_depths"]:
        assert t.grad is not None and bool(torch.isfinite(t.grad).all())


@pytest.mark.parametrize("B,H,W", [(2, 48, 64), (1, 37, 53)])
def test_inverse_rotation_warp_vs_oracle(B, H, W):
    """layers.inverse_rotation_warp (layers.py:460-479) on libsqlx vs the float64 restatement: warped frames and the
    gradient wrt the Euler angles (what trainer_indoor.rectify_imgs back-propagates, trainer_indoor.py:877-920)."""
    import sqlx
    from oracle import sqldepth_oracle as O
    from _cases import smooth_images
    g = torch.Generator().manual_seed(B + H)
    img = smooth_images(g, B, H + (8 - H % 8) % 8, W + (8 - W % 8) % 8, 1)[0][:, :, :H, :W].contiguous()
    rot = 0.05 * torch.randn(B, 3, generator=g)
    K = torch.tensor([[0.58 * W, 0, 0.5 * W], [0, 1.92 * H / 2, 0.5 * H], [0, 0, 1]]).repeat(B, 1, 1)
    gout = torch.randn(B, 3, H, W, generator=g)
    rd = rot.double().requires_grad_(True)
    want = O.inverse_rotation_warp(img.double(), rd, K.double())
    (g_want,) = torch.autograd.grad((want * gout.double()).sum(), [rd])
    rc = rot.cuda().requires_grad_(True)
    got = sqlx.inverse_rotation_warp(img.cuda(), rc, K.cuda())
    assert float((got.cpu().double() - want).abs().max()) < 2e-4
    (g_got,) = torch.autograd.grad((got * gout.cuda()).sum(), [rc])
    rel = float((g_got.cpu().double() - g_want).abs().max() / g_want.abs().max())
    assert rel < 5e-3, rel
