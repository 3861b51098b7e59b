"""Per-sample median scaling of the supervised fine-tuning step (finetune/train_ft_SQLdepth.py:236-266):
the device-side ratios (sort-based, no host round trip) against a NumPy restatement of the reference loop.  The maths
is device-agnostic torch code, so it is checked here on CPU tensors; the public entry point takes CUDA tensors only."""
import numpy as np
import pytest
import torch


def _reference_ratios(pred, depth, lo, hi, garg_crop, eigen_crop, dataset):
    """restates train_ft_SQLdepth.py:236-266 (boolean-mask gather + np.median per sample, first half of the batch)"""
    B = pred.shape[0]
    out = np.ones(B, dtype=np.float32)
    for i in range(B // 2):
        p = pred[i, 0].numpy()
        d = depth[i, 0].numpy()
        valid = np.logical_and(d > lo, d < hi)
        H, W = d.shape
        ev = np.zeros(valid.shape)
        if garg_crop:
            ev[int(0.40810811 * H):int(0.99189189 * H), int(0.03594771 * W):int(0.96405229 * W)] = 1
        elif eigen_crop:
            if dataset == "kitti":
                ev[int(0.3324324 * H):int(0.91351351 * H), int(0.0359477 * W):int(0.96405229 * W)] = 1
            else:
                ev[45:471, 41:601] = 1
        valid = np.logical_and(valid, ev)
        with np.errstate(all="ignore"):
            mp = np.median(p[valid]) if valid.any() else np.nan
            md = np.median(d[valid]) if valid.any() else np.nan
        out[i] = 1.0 if (np.isnan(md) or np.isnan(mp)) else md / mp
    return out


@pytest.mark.parametrize("crop", [dict(garg_crop=True), dict(eigen_crop=True), dict(eigen_crop=True, dataset="nyu")])
def test_median_ratios_match_numpy_loop(crop):
    from sqlx.layers import median_scale_ratios
    g = torch.Generator().manual_seed(4)
    B, H, W = 6, 480, 640
    depth = torch.rand(B, 1, H, W, generator=g) * 90.0
    depth[depth < 20.0] = 0.0                      # sparse ground truth: most pixels invalid
    pred = torch.rand(B, 1, H, W, generator=g) * 40.0 + 0.5
    pred[1, 0, 300, 300] = float("nan")            # a NaN inside the crop -> ratio 1 for that sample (:261-262)
    depth[2] = 0.0                                 # no valid pixel at all -> ratio 1
    kw = dict(garg_crop=False, eigen_crop=False, dataset="kitti")
    kw.update(crop)
    got = median_scale_ratios(pred, depth, 1e-3, 80.0, **kw).numpy()
    want = _reference_ratios(pred, depth, 1e-3, 80.0, kw["garg_crop"], kw["eigen_crop"], kw["dataset"])
    assert got[1] == 1.0 and got[2] == 1.0 and (got[3:] == 1.0).all()
    np.testing.assert_allclose(got, want, rtol=1e-6)


def test_even_and_odd_counts():
    from sqlx.layers import _masked_median
    v = torch.tensor([[5.0, 1.0, 9.0, 3.0, 7.0], [5.0, 1.0, 9.0, 3.0, 7.0]])
    m = torch.tensor([[True, True, True, True, True], [True, True, False, True, True]])
    med = _masked_median(v, m)
    assert med.tolist() == [5.0, 4.0]              # odd: middle element; even: mean of the two middle ones


def test_crop_flag_is_required_like_the_reference():
    from sqlx.layers import median_scale_ratios
    with pytest.raises(ValueError):
        median_scale_ratios(torch.ones(2, 1, 8, 8), torch.ones(2, 1, 8, 8), 1e-3, 80.0)


def test_public_entry_point_takes_cuda_tensors_only():
    import sqlx
    from sqlx.layers import median_scale
    with pytest.raises(sqlx.SqlxError):
        median_scale(torch.ones(2, 1, 8, 8), torch.ones(2, 1, 8, 8), 1e-3, 80.0, garg_crop=True)
