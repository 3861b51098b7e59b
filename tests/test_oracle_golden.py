"""Pin the CPU oracle (oracle/sqldepth_oracle.py) against outputs of the reference itself.

The golden vectors under tests/golden/ were produced by oracle/make_golden.py, which executes the
unmodified reference (trainer.py:386-549, networks/depth_decoder_QTR.py:36-74, layers.py) on CPU.
"""
import numpy as np
import pytest
import torch

from oracle import sqldepth_oracle as O
from _cases import INDOOR_CASES, PHOTO_CASES, indoor_case, load_npz, photo_case


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", PHOTO_CASES)
def test_photometric_matches_reference(name):
    kw, leaves, z, fids = photo_case(name)
    out = O.photometric_losses(**kw)
    assert abs(float(out["loss"]) - float(z["out_loss"])) < 2e-7
    for s in kw["scales"]:
        assert abs(float(out["loss/%d" % s]) - float(z["out_loss_s%d" % s])) < 2e-7
        np.testing.assert_allclose(out[("depth", 0, s)].detach().numpy(), z["out_depth_s%d" % s], rtol=1e-6, atol=1e-6)
        if not kw["disable_automasking"]:
            sel = out["identity_selection/%d" % s].numpy().astype(np.uint8)
            assert (sel != z["out_idsel_s%d" % s]).mean() < 1e-4
    s0 = kw["scales"][0]
    for i, f in enumerate(fids[1:]):
        np.testing.assert_allclose(out[("sample", i, s0)].detach().numpy(), z["out_sample_%s_s%d" % (f, s0)], atol=2e-6)
        np.testing.assert_allclose(out[("color", i, s0)].detach().numpy(), z["out_color_%s_s%d" % (f, s0)], atol=2e-5)
    names = list(leaves)
    grads = torch.autograd.grad(out["loss"], [leaves[n] for n in names], allow_unused=True)
    for n, g in zip(names, grads):
        ref = z["grad_" + n]
        g = np.zeros_like(ref) if g is None else g.numpy()
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(g - ref).max() / scale < 2e-3, n


@pytest.mark.parametrize("name", INDOOR_CASES)
def test_indoor_losses_match_reference(name):
    """oracle.indoor_losses vs trainer_indoor.py generate_images_pred + compute_losses_with_occ."""
    kw, leaves, z = indoor_case(name)
    out = O.indoor_losses(**kw)
    assert abs(float(out["loss"]) - float(z["out_loss"])) < 2e-7
    assert abs(float(out["loss/0"]) - float(z["out_loss_s0"])) < 2e-7
    np.testing.assert_allclose(out[("depth", 0, 0)].detach().numpy(), z["out_depth_s0"], rtol=1e-6, atol=1e-6)
    for i, f in enumerate((-1, 1)):
        np.testing.assert_allclose(out[("color", i, 0)].detach().numpy(), z["out_color_%d" % f], atol=2e-5)
        np.testing.assert_allclose(out[("pred_dep", i, 0)].detach().numpy(), z["out_pred_dep_%d" % f], rtol=2e-5, atol=2e-4)
    names = list(leaves)
    grads = torch.autograd.grad(out["loss"], [leaves[n] for n in names], allow_unused=True)
    for n, g in zip(names, grads):
        ref = z["grad_" + n]
        g = np.zeros_like(ref) if g is None else g.numpy()
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(g - ref).max() / scale < 2e-3, n


@pytest.mark.parametrize("name", ["decoder_full", "decoder_lite"])
def test_sql_tail_matches_reference(name):
    z = load_npz(name)
    st = load_npz(name + "_state")
    x = _t(z["x"]).requires_grad_(True)
    q = _t(z["queries"]).requires_grad_(True)
    mlp = [_t(st["bins_regressor.%d.%s" % (i, k)]) for i in (0, 2, 4) for k in ("weight", "bias")]
    D, Q = int(z["D"]), int(z["Q"])
    Wp = _t(st["convert_to_prob.0.weight"]).reshape(D, Q).clone().requires_grad_(True)
    bp = _t(st["convert_to_prob.0.bias"]).clone().requires_grad_(True)
    out = O.sql_tail(x, q, mlp, Wp, bp, float(z["min_val"]), float(z["max_val"]))
    np.testing.assert_allclose(out["summary"].detach().numpy(), z["out_summary"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["energy"].detach().numpy()[:, :, ::4, ::4], z["out_energy_sample"], rtol=1e-5, atol=1e-5)
    rel = (out["pred"].detach().numpy() - z["out_pred"]) / z["out_pred"]
    assert np.abs(rel).max() < 2e-5
    (out["pred"] * _t(z["gout"])).sum().backward()
    for got, ref in ((x.grad, z["grad_x"]), (q.grad, z["grad_queries"]), (Wp.grad, z["grad_Wp"]), (bp.grad, z["grad_bp"])):
        assert np.abs(got.numpy() - ref).max() / np.abs(ref).max() < 1e-3


def test_modules_match_reference():
    z = load_npz("modules")
    a = _t(z["a"]).requires_grad_(True)
    ss = O.ssim(a, _t(z["b"]))
    np.testing.assert_allclose(ss.detach().numpy(), z["out_ssim"], atol=1e-6)
    (ga,) = torch.autograd.grad((ss * _t(z["g_ssim"])).sum(), a)
    assert np.abs(ga.numpy() - z["grad_a"]).max() / np.abs(z["grad_a"]).max() < 1e-3
    aa, tr = _t(z["axisangle"]), _t(z["translation"])
    np.testing.assert_allclose(O.pose_matrix(aa, tr, False).numpy(), z["out_T"], atol=1e-7)
    np.testing.assert_allclose(O.pose_matrix(aa, tr, True).numpy(), z["out_T_inv"], atol=1e-7)
    B, H, W = int(z["B"]), int(z["H"]), int(z["W"])
    pts = O.backproject(_t(z["depth"]), _t(z["inv_K"]))
    np.testing.assert_allclose(pts.numpy(), z["out_points"], rtol=1e-6, atol=1e-6)
    grid = O.project(pts, _t(z["K"]), _t(z["out_T"]), H, W)
    np.testing.assert_allclose(grid.numpy(), z["out_grid"], atol=2e-6)
    sm = O.smooth_loss(_t(z["disp"]), _t(z["b"]))
    assert abs(float(sm) - float(z["out_smooth"])) < 1e-7
    sil = O.silog_loss(_t(z["silog_pred"]), _t(z["silog_gt"]), mask=_t(z["silog_gt"]) > 1e-3)
    assert abs(float(sil) - float(z["out_silog"])) < 1e-5


def test_postprocess_disparity_golden():
    """Flip-TTA blend (evaluate_depth_config.py:51-59): oracle vs the reference's own numpy output."""
    z = load_npz("eval_postprocess")
    for i in range(3):
        out = O.batch_post_process_disparity(_t(z["l%d" % i]), _t(z["r%d" % i]))
        assert float((out - torch.from_numpy(z["out%d" % i])).abs().max()) < 1e-12


def test_indoor_reduces_to_outdoor_when_depths_agree():
    """Property of compute_losses_with_occ (trainer_indoor.py:636-651): with a fixed identity camera transform the source
    depth is sampled at the pixel itself; if it equals the target depth, diff = 0, the weight is exactly 1, the
    regularisation term vanishes and the photometric part equals compute_losses' (trainer.py:474-532)."""
    import torch.nn.functional as F
    from _cases import synth_photo_case
    kw = synth_photo_case(seed=9, B=1, H=48, W=80, S=2)
    for k in ("K", "inv_K"):
        kw[k] = kw[k].double()
    kw["disps"] = {0: kw["disps"][0].double()}
    kw["target_pyr"] = {0: kw["target_pyr"][0].double()}
    kw["sources"] = [t.double() for t in kw["sources"]]
    kw["noises"] = {0: kw["noises"][0].double()}
    B, H, W = 1, 48, 80
    eye = torch.eye(4, dtype=torch.float64).unsqueeze(0).repeat(B, 1, 1)
    poses = [{"T": eye}, {"T": eye}]
    disp = kw["disps"][0]
    up = F.interpolate(disp, [H, W], mode="bilinear", align_corners=False)
    out = O.indoor_losses(disp, kw["target_pyr"][0], kw["sources"], [up, up], kw["K"], kw["inv_K"], poses,
                          kw["noises"][0], height=H, width=W)
    ref = O.photometric_losses({0: disp}, kw["target_pyr"], kw["sources"], kw["K"], kw["inv_K"], poses, kw["noises"],
                               height=H, width=W, scales=(0,), rescale_translation=False, disparity_smoothness=0.0)
    # (not exactly 0: inv_K is a float32 pseudo-inverse and Project3D adds eps = 1e-7 to z, layers.py:252)
    assert float(out["reg"]) < 1e-6
    # weight = 1 - sqrt(1 - (diff - 1)^2) = 1 - sqrt(2 diff - diff^2): a residual diff of 1e-7 leaves 1 - 4.5e-4
    assert abs(float(out["photo"]) - float(ref["loss"])) < 1e-3 * float(ref["loss"])
    # (the arg-min itself is not comparable here: with an identity transform every reprojection loss ties with its
    # identity loss up to the 1e-5 noise)


@pytest.mark.parametrize("variant", [{"no_ssim": True}, {"disable_automasking": True}, {"avg_reprojection": True}])
def test_indoor_variants_are_finite_and_differentiable(variant):
    kw, leaves, z = indoor_case("indoor_occ")
    kw.update(variant)
    if variant.get("avg_reprojection"):
        kw["noise"] = kw["noise"][:, :1]
    out = O.indoor_losses(**kw)
    grads = torch.autograd.grad(out["loss"], list(leaves.values()), allow_unused=True)
    assert bool(torch.isfinite(out["loss"]))
    assert all(g is None or bool(torch.isfinite(g).all()) for g in grads)
    assert grads[0] is not None and float(grads[0].abs().max()) > 0


# ---- the explicit-index rules (oracle/explicit.py) against the library primitives the reference calls
def test_explicit_upsample_and_grid_sample_rules():
    import torch.nn.functional as F
    from oracle import explicit as X
    g = torch.Generator().manual_seed(0)
    for (h, w, H, W) in ((6, 10, 12, 20), (5, 7, 13, 17), (8, 8, 8, 8)):
        x = torch.rand(1, 1, h, w, generator=g, dtype=torch.float64)
        ref = F.interpolate(x, [H, W], mode="bilinear", align_corners=False)[0, 0].numpy()
        assert np.abs(X.upsample_bilinear(x[0, 0].numpy(), H, W) - ref).max() < 1e-13
    src = torch.rand(1, 3, 9, 11, generator=g, dtype=torch.float64)
    grid = torch.rand(1, 6, 7, 2, generator=g, dtype=torch.float64) * 2.6 - 1.3      # some coordinates outside the frame
    grid[0, 0, 0] = torch.tensor([1.0, 1.0])                                          # exactly the last pixel
    grid[0, 0, 1] = torch.tensor([-1.0, -1.0])
    ref = F.grid_sample(src, grid, padding_mode="border", align_corners=True)[0].numpy()
    assert np.abs(X.grid_sample_border(src[0].numpy(), grid[0].numpy()) - ref).max() < 1e-13


def test_explicit_ssim_softmax_centers_rules():
    from oracle import explicit as X
    g = torch.Generator().manual_seed(1)
    a = torch.rand(1, 1, 12, 15, generator=g, dtype=torch.float64)
    b = (a + 0.1 * torch.rand(1, 1, 12, 15, generator=g, dtype=torch.float64)).clamp(0, 1)
    for radius in (3, 1):
        ref = O.ssim(a, b, radius)[0, 0].numpy()
        assert np.abs(X.ssim_reflect(a[0, 0].numpy(), b[0, 0].numpy(), radius) - ref).max() < 1e-12
    x = torch.randn(1, 8, 5, 6, generator=g, dtype=torch.float64)
    K = torch.randn(1, 4, 8, generator=g, dtype=torch.float64)
    energy, summary = O.full_query(x, K)
    e2, s2 = X.pixel_softmax_summary(x[0].reshape(8, 30).numpy(), K[0].numpy())
    assert np.abs(e2 - energy[0].reshape(4, 30).numpy()).max() < 1e-12
    assert np.abs(s2 - summary[0].numpy()).max() < 1e-12
    r = torch.randn(2, 10, generator=g, dtype=torch.float64)
    ref = O.bin_centers(r, 0.001, 80.0).numpy()
    for i in range(2):
        assert np.abs(X.bin_centers(r[i].numpy(), 0.001, 80.0) - ref[i]).max() < 1e-12
