"""Per-sample median scaling of the supervised fine-tuning step (finetune/train_ft_SQLdepth.py:236-266).

The reference loop is inline NumPy code of train(); oracle/sqldepth_oracle.py:median_scale_ratios restates it line by line
and oracle/make_golden_median.py executed it into tests/golden/median_scale.npz (outside the product).  CPU: the oracle
against the fixture.  GPU: the radix-select kernel (csrc/median.cu, through the C ABI) against the fixture and against
the oracle on seeded cases with ties, NaNs, empty selections and negative predictions -- ratios are order statistics, so
the comparison is EXACT (same two float32 values averaged and divided)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from _cases import load_npz  # noqa: E402

CASES = {"garg": dict(garg_crop=True), "eigen_kitti": dict(eigen_crop=True, dataset="kitti")}


def test_oracle_matches_fixture():
    from oracle import sqldepth_oracle as O
    z = load_npz("median_scale")
    for name, kw in CASES.items():
        got = O.median_scale_ratios(torch.from_numpy(z[name + "_pred"]), torch.from_numpy(z[name + "_depth"]), 1e-3, 80.0, **kw)
        np.testing.assert_array_equal(got, z[name + "_ratio"])
    # the 480x640 case is regenerated from its seed (make_golden_median.py keeps only its ratios)
    import make_golden_median as G
    pred, depth = G.make_case(3, 4, 480, 640, empty_sample=0)
    got = O.median_scale_ratios(pred, depth, 1e-3, 80.0, eigen_crop=True, dataset="nyu")
    np.testing.assert_array_equal(got, z["eigen_nyu_ratio"])


@pytest.mark.gpu
def test_kernel_matches_fixture():
    import sqlx
    z = load_npz("median_scale")
    for name, kw in CASES.items():
        pred, depth = torch.from_numpy(z[name + "_pred"]).cuda(), torch.from_numpy(z[name + "_depth"]).cuda()
        got = sqlx.median_scale_ratios(pred, depth, 1e-3, 80.0, **kw).cpu().numpy()
        np.testing.assert_array_equal(got, z[name + "_ratio"])
    import make_golden_median as G
    pred, depth = G.make_case(3, 4, 480, 640, empty_sample=0)
    got = sqlx.median_scale_ratios(pred.cuda(), depth.cuda(), 1e-3, 80.0, eigen_crop=True, dataset="nyu").cpu().numpy()
    np.testing.assert_array_equal(got, z["eigen_nyu_ratio"])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,B,H,W,kw", [
    (11, 8, 352, 1216, dict(garg_crop=True)),                      # KITTI ground-truth resolution
    (12, 2, 37, 53, dict(eigen_crop=True, dataset="kitti")),       # odd sizes, tiny selections
    (13, 6, 480, 640, dict(eigen_crop=True, dataset="nyu")),
])
def test_kernel_matches_oracle(seed, B, H, W, kw):
    import sqlx
    from oracle import sqldepth_oracle as O
    import make_golden_median as G
    pred, depth = G.make_case(seed, B, H, W, nan_sample=1 if B > 2 else None, empty_sample=2 if B > 5 else None)
    pred[0] = -pred[0]                                              # negative predictions order correctly
    want = O.median_scale_ratios(pred, depth, 1e-3, 80.0, **kw)
    got = sqlx.median_scale_ratios(pred.cuda(), depth.cuda(), 1e-3, 80.0, **kw).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    # the drop-in: pred scaled per sample, differentiable wrt pred with constant ratios (as the reference's in-place *=)
    p = pred.cuda().requires_grad_(True)
    out = sqlx.median_scale(p, depth.cuda(), 1e-3, 80.0, **kw)
    np.testing.assert_array_equal(out.detach().cpu().numpy(), (pred * torch.from_numpy(want).view(-1, 1, 1, 1)).numpy())
    out.nan_to_num().sum().backward()
    assert torch.equal(p.grad[B - 1], torch.ones_like(p.grad[B - 1]))           # second half of the batch: ratio 1


def test_needs_a_crop_flag_and_cuda():
    import sqlx
    with pytest.raises((ValueError, sqlx.SqlxError)):
        sqlx.median_scale_ratios(torch.rand(2, 1, 8, 8), torch.rand(2, 1, 8, 8), 1e-3, 80.0)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 24, 80, 47, 156), (6, 40, 128, 93, 310), (2, 16, 16, 16, 16)])
def test_finetune_loss_matches_oracle(shape):
    """sqlx.finetune_loss (median ratios and SILog both read the low-resolution prediction through the align_corners=True
    resize; nothing of ground-truth size is materialised) against the fp64 oracle of train_ft_SQLdepth.py:232-274: same
    ratios (float32 resize arithmetic: 1e-5), loss 1e-5, gradient 1e-4 of its largest entry."""
    import sqlx
    from oracle import sqldepth_oracle as O
    B, h, w, H, W = shape
    g = torch.Generator().manual_seed(B + h)
    pred = (0.5 + 20 * torch.rand(B, 1, h, w, generator=g)).requires_grad_(True)
    depth = 80 * torch.rand(B, 1, H, W, generator=g)
    depth[torch.rand(B, 1, H, W, generator=g) < 0.6] = 0.0          # sparse ground truth
    ref_loss, ref_ratio = O.finetune_loss(pred.double(), depth.double(), 1e-3, 1e-3, 80.0, garg_crop=True)
    ref_grad, = torch.autograd.grad(ref_loss, pred)
    pc = pred.detach().cuda().requires_grad_(True)
    dc = depth.cuda()
    ratio = sqlx.median_scale_ratios(pc.detach(), dc, 1e-3, 80.0, garg_crop=True)
    assert float((ratio.cpu().double() - ref_ratio.double()).abs().max()) < 1e-5 * float(ref_ratio.abs().max())
    loss = sqlx.finetune_loss(pc, dc, 1e-3, 1e-3, 80.0, garg_crop=True)
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 1e-5 * max(1.0, abs(float(ref_loss)))
    gd = (pc.grad.cpu().double() - ref_grad.double()).abs().max()
    assert float(gd) < 1e-4 * float(ref_grad.abs().max())


@pytest.mark.parametrize("case", ["kitti_garg", "kitti_eigen"])
def test_oracle_finetune_step_matches_reference_fixture(case):
    """oracle.finetune_loss against tests/golden/finetune_step.npz, produced with the reference's own SILogLoss class and
    resize call (oracle/make_golden_finetune.py): ratios exact, loss and gradient at float32 round-off."""
    import numpy as np
    from oracle import sqldepth_oracle as O
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "finetune_step.npz"))
    pred = torch.from_numpy(z[case + "/pred"]).requires_grad_(True)
    depth = torch.from_numpy(z[case + "/depth"])
    kw = dict(garg_crop=bool(z[case + "/garg"]), eigen_crop=bool(z[case + "/eigen"]), dataset="kitti")
    loss, ratio = O.finetune_loss(pred, depth, 1e-3, 1e-3, 80.0, **kw)
    grad, = torch.autograd.grad(loss, pred)
    assert np.array_equal(ratio.numpy(), z[case + "/ratios"])
    assert abs(float(loss) - float(z[case + "/loss"])) < 1e-6 * max(1.0, abs(float(z[case + "/loss"])))
    assert float((grad - torch.from_numpy(z[case + "/grad"])).abs().max()) < 1e-6 * float(np.abs(z[case + "/grad"]).max())
