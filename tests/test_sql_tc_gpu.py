"""Tensor-core (tcgen05 / TMEM / TMA) SQL kernels against the float64 oracle and the exact-fp32 CUDA kernels."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [
    dict(B=2, h=24, w=40, Q=64),
    dict(B=1, h=17, w=20, Q=120),       # ragged last tile, Q padded to 128
    dict(B=3, h=8, w=12, Q=16),         # fewer pixels than one tile
    dict(B=2, h=96, w=320, Q=128),
])
def test_energy_tc(cfg):
    from sqlx import sql as S
    B, h, w, Q = cfg["B"], cfg["h"], cfg["w"], cfg["Q"]
    g = torch.Generator().manual_seed(7 + Q)
    x = torch.randn(B, 32, h, w, generator=g)
    q = 0.5 * torch.randn(B, Q, 32, generator=g)
    assert S.tc_supported(32, Q, 0, h * w)
    en = S.energy_tc(x.cuda(), q.cuda())
    ref = torch.einsum("bep,bqe->bqp", x.double().reshape(B, 32, -1), q.double()).reshape(B, Q, h, w)
    err = float((en.cpu().double() - ref).abs().max())
    # 3xTF32 reproduces fp32: |y| <= ~12 here, fp32 dot-product error is ~1e-6; single-pass TF32 would be ~5e-3
    assert err < 2e-6 * float(ref.abs().max()) + 1e-5, err


def _set_tc(on):
    import sqlx
    return sqlx.lib().sqlx_sql_set_tensor_cores(int(on))


@pytest.mark.parametrize("cfg", [
    dict(B=2, h=24, w=40, Q=64, D=64),
    dict(B=1, h=17, w=20, Q=120, D=128),      # cfg-1 style: Q padded to 128
    dict(B=3, h=8, w=12, Q=12, D=16),         # golden-fixture sized
    dict(B=2, h=96, w=320, Q=128, D=128),     # cfg-3 sized Q, D (one CTA per SM)
    dict(B=12, h=96, w=320, Q=64, D=64),      # BASELINE config 2 full size
])
def test_pred_tc(cfg):
    """tcgen05 depth-regression kernel (mixed weights: logits = (Wp K) x + b) vs the float64 oracle (1e-4 relative, the
    north-star bar) at EVERY size, and vs the exact-fp32 CUDA-core kernel of the un-mixed formulation."""
    from sqlx import sql as S
    from oracle import sqldepth_oracle as O
    B, h, w, Q, D = (cfg[k] for k in ("B", "h", "w", "Q", "D"))
    g = torch.Generator().manual_seed(11 + Q + D)
    x = torch.randn(B, 32, h, w, generator=g)
    q = 0.4 * torch.randn(B, Q, 32, generator=g)
    Wp = 0.3 * torch.randn(D, Q, generator=g)
    bp = 0.1 * torch.randn(D, generator=g)
    centers = torch.sort(0.1 + 80 * torch.rand(B, D, generator=g), dim=1).values
    xc, qc, Wc, bc, cc = (t.cuda() for t in (x, q, Wp, bp, centers))
    assert S.tc_supported(32, Q, D, h * w)
    pred_tc, stats = S.pred_mix_fwd(xc, S.mix_weights(Wc, qc), bc, cc)
    assert torch.isfinite(stats).all() and bool((stats[1] > 0).all()) and bool((stats[1] <= 1.0 + 1e-6).all())
    pred_fp = S.pred_fwd(xc, qc, Wc, bc, cc)
    assert float(((pred_tc - pred_fp) / pred_fp).abs().max()) < 5e-5
    ref = []
    for b in range(B):       # one sample at a time: the oracle's [n x Q] float64 energy maps are 31 MB per sample here
        energy, _ = O.full_query(x[b:b + 1].double(), q[b:b + 1].double())
        ref.append(O.bins_expectation(energy, Wp.double(), bp.double(), centers[b:b + 1].double()))
    ref = torch.cat(ref)
    assert float(((pred_tc.cpu().double() - ref) / ref).abs().max()) < 1e-4


@pytest.mark.parametrize("cfg", [
    dict(B=2, h=24, w=40, Q=64),
    dict(B=1, h=17, w=20, Q=120),        # 340 pixels: ragged last 64-pixel tile
    dict(B=3, h=2, w=6, Q=12),           # 12 pixels: a single partial tile
    dict(B=12, h=96, w=320, Q=64),       # BASELINE config 2 full size
    dict(B=2, h=160, w=512, Q=128),      # cfg-3 sized
    dict(B=1, h=320, w=1024, Q=128),     # cfg-4 sized: 327,680 pixels per sample
])
def test_summary_tc(cfg):
    """tcgen05 flash-style pixel-softmax summaries vs the float64 oracle and the fp32 CUDA kernel."""
    from sqlx import sql as S
    from oracle import sqldepth_oracle as O
    B, h, w, Q = (cfg[k] for k in ("B", "h", "w", "Q"))
    g = torch.Generator().manual_seed(23 + Q)
    x = torch.randn(B, 32, h, w, generator=g)
    q = 0.5 * torch.randn(B, Q, 32, generator=g)
    xc, qc = x.cuda(), q.cuda()
    prev = _set_tc(1)
    try:
        s_tc, m_tc, l_tc, _ = S.summary_fwd(xc, qc)
        _set_tc(0)
        s_fp, m_fp, l_fp, _ = S.summary_fwd(xc, qc)
    finally:
        _set_tc(prev)
    scale = float(s_fp.abs().max())
    assert float((s_tc - s_fp).abs().max()) < 2e-5 * max(scale, 1.0)
    # the (max, sum) statistics may use different reference points; log-sum-exp must agree
    lse_tc = m_tc + torch.log(l_tc)
    lse_fp = m_fp + torch.log(l_fp)
    assert float((lse_tc - lse_fp).abs().max()) < 1e-4
    ref = torch.cat([O.full_query(x[b:b + 1].double(), q[b:b + 1].double())[1] for b in range(B)])
    assert float((s_tc.cpu().double() - ref).abs().max()) < 2e-5 * max(scale, 1.0)


@pytest.mark.parametrize("cfg", [
    dict(B=2, h=24, w=40, Q=64, D=64),
    dict(B=1, h=17, w=20, Q=120, D=128),
    dict(B=2, h=16, w=24, Q=128, D=128),
    dict(B=3, h=8, w=12, Q=12, D=16),
])
def test_tail_paths_agree(cfg):
    """sql_tail through the mixed-weight tensor-core decomposition vs the exact-fp32 CUDA-core kernels:
    forward depth and every gradient (x, queries, Wp, bp, MLP weights)."""
    import sqlx
    B, h, w, Q, D = (cfg[k] for k in ("B", "h", "w", "Q", "D"))
    g = torch.Generator().manual_seed(5 + Q)
    x = torch.randn(B, 32, h, w, generator=g)
    q = 0.4 * torch.randn(B, Q, 32, generator=g)
    Wp = 0.3 * torch.randn(D, Q, generator=g)
    bp = 0.1 * torch.randn(D, generator=g)
    W1 = torch.randn(D, Q * 32, generator=g) / (Q * 32) ** 0.5
    gout = torch.randn(B, 1, h, w, generator=g)
    F = torch.nn.functional
    res = {}
    for on in (1, 0):
        prev = _set_tc(on)
        try:
            leaves = [t.clone().cuda().requires_grad_(True) for t in (x, q, Wp, bp, W1)]
            xc, qc, Wc, bc, W1c = leaves
            pred = sqlx.sql_tail(xc, qc, Wc, bc, lambda s: sqlx.bin_centers(F.linear(s.reshape(B, -1), W1c), 0.01, 80.0), (W1c,))
            grads = torch.autograd.grad((pred * gout.cuda()).sum(), leaves)
            res[on] = (pred.detach(), grads)
        finally:
            _set_tc(prev)
    assert float(((res[1][0] - res[0][0]) / res[0][0]).abs().max()) < 5e-5
    for a, b, nm in zip(res[1][1], res[0][1], ("x", "queries", "Wp", "bp", "W1")):
        rel = float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
        assert rel < 2e-3, (nm, rel)


@pytest.mark.parametrize("cfg", [
    dict(B=2, h=24, w=40, Q=64, D=64),
    dict(B=1, h=17, w=20, Q=120, D=128),      # ragged last tile, two bin halves
    dict(B=3, h=8, w=12, Q=12, D=16),         # fewer pixels than a tile, three column groups all padding
    dict(B=5, h=3, w=4, Q=8, D=100),          # more CTAs than tiles: empty chunks
    dict(B=12, h=96, w=320, Q=64, D=64),      # BASELINE config 2
    dict(B=8, h=160, w=512, Q=128, D=128),    # BASELINE config 3
    dict(B=2, h=320, w=1024, Q=128, D=128),   # BASELINE config 4 (two of its eight samples)
])
def test_ws_kernels_match_v1(cfg):
    """Warp-specialised depth-regression forward / backward (csrc/sql_ws.cu: 16 epilogue warps, double-buffered TMEM, saved
    softmax statistics) against the round-1 single-warpgroup kernels on the same inputs: same arithmetic up to summation
    order (2e-6 of the depth; gradients 1e-4 of max |grad|)."""
    from sqlx import sql as S
    B, h, w, Q, D = (cfg[k] for k in ("B", "h", "w", "Q", "D"))
    g = torch.Generator().manual_seed(3 + Q + D)
    x = torch.randn(B, 32, h, w, generator=g).cuda()
    q = (0.4 * torch.randn(B, Q, 32, generator=g)).cuda()
    Wp = (0.3 * torch.randn(D, Q, generator=g)).cuda()
    bp = (0.1 * torch.randn(D, generator=g)).cuda()
    cen = torch.sort(0.1 + 80 * torch.rand(B, D, generator=g), dim=1).values.cuda()
    gp = torch.randn(B, 1, h, w, generator=g).cuda()
    Mx = S.mix_weights(Wp, q)
    p1, _ = S.pred_mix_fwd(x, Mx, bp, cen, version=1)
    p2, stats = S.pred_mix_fwd(x, Mx, bp, cen)
    assert float(((p2 - p1) / p1).abs().max()) < 2e-6
    r1 = S.bwd_pred_mix(x, Mx, bp, cen, gp)
    r2 = S.bwd_pred_mix(x, Mx, bp, cen, gp, p2, stats)
    for a, b, nm in zip(r2, r1, ("d_M", "d_bp", "d_centers", "d_x")):
        assert torch.isfinite(a).all(), nm
        rel = float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
        assert rel < 1e-3, (nm, rel)     # single-pass TF32 gradient contractions: operand-truncation noise ~2e-4
    # summary-path backward, write and accumulate modes
    summ, mx, sm, _ = S.summary_fwd(x, q)
    ds = torch.randn(summ.shape, generator=g).cuda()
    acc0 = torch.randn(x.shape, generator=g).cuda()
    for mode in ("write", "accumulate"):
        a1 = S.bwd_summary(x, q, summ, mx, sm, ds, d_x=acc0.clone() if mode == "accumulate" else None, version=1)
        a2 = S.bwd_summary(x, q, summ, mx, sm, ds, d_x=acc0.clone() if mode == "accumulate" else None)
        for a, b, nm in zip(a2, a1, ("d_x", "d_queries")):
            assert torch.isfinite(a).all(), (mode, nm)
            rel = float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
            assert rel < 1e-3, (mode, nm, rel)


@pytest.mark.parametrize("cfg", [
    dict(B=2, h=24, w=40, Q=64),
    dict(B=1, h=17, w=20, Q=120),              # ragged last step (340 = 10 * 32 + 20), padded query rows
    dict(B=3, h=8, w=12, Q=12),                # three steps: one group of epilogue warps stays idle
    dict(B=5, h=3, w=4, Q=8),                  # one partial step (12 pixels: the upper lane half owns none of them)
    dict(B=2, h=9, w=4, Q=40),                 # 36 pixels: the second step holds 4 (upper lane half empty)
    dict(B=2, h=64, w=96, Q=64, qscale=6.0),   # energies spread over +-60: the running maximum keeps moving (rescale path)
    dict(B=2, h=64, w=96, Q=128, qscale=6.0),
    dict(B=150, h=8, w=16, Q=64),              # more samples than SMs: one chunk per sample
    dict(B=12, h=96, w=320, Q=64),             # BASELINE config 2
    dict(B=8, h=160, w=512, Q=128),            # BASELINE config 3
    dict(B=2, h=320, w=1024, Q=128),           # BASELINE config 4 (two of its eight samples)
])
def test_ws_summary_matches_v1_and_fp64(cfg):
    """Warp-specialised summary kernel (csrc/sql_ws.cu: lane = query, four groups of epilogue warps, accumulators resident in
    TMEM, lazily rescaled) against the round-1 kernel and against an fp64 restatement of networks/layers.py:17-20
    (softmax over the pixels of y = K x, summary = a x^T): 3xTF32 on both contractions, so 1e-5 of the largest summary."""
    from sqlx import sql as S
    B, h, w, Q = (cfg[k] for k in ("B", "h", "w", "Q"))
    g = torch.Generator().manual_seed(11 + Q + h)
    x = torch.randn(B, 32, h, w, generator=g).cuda()
    q = (cfg.get("qscale", 1.0) * 0.4 * torch.randn(B, Q, 32, generator=g)).cuda()
    s2, m2, l2, _ = S.summary_fwd(x, q)
    s1, m1, l1, _ = S.summary_fwd(x, q, version=1)
    xd, qd = x.double().flatten(2), q.double()
    y = torch.bmm(qd, xd)                                         # [B, Q, n]
    ref = torch.bmm(torch.softmax(y, dim=2), xd.transpose(1, 2))  # [B, Q, E]
    scale = float(ref.abs().max())
    assert torch.isfinite(s2).all()
    assert float((s2.double() - ref).abs().max()) < 1e-5 * max(scale, 1.0), float((s2.double() - ref).abs().max())
    assert float((s2 - s1).abs().max()) < 1e-5 * max(scale, 1.0)
    # the saved row statistics describe the same softmax: log-sum-exp agrees with fp64
    lse = torch.logsumexp(y, dim=2)
    assert float(((m2.double() + l2.double().log()) - lse).abs().max()) < 1e-4
    assert float(((m1.double() + l1.double().log()) - lse).abs().max()) < 1e-4


def test_sm_budget_changes_grid_not_result():
    """sqlx_sql_set_sm_budget (SMs left to the communication kernel beside the summary-path backward)
    only changes how the tiles are dealt to CTAs: same gradients up to the summation order of the per-CTA partials."""
    from sqlx import sql as S
    from sqlx._lib import lib
    g = torch.Generator().manual_seed(5)
    B, h, w, Q = 12, 48, 160, 64
    x = torch.randn(B, 32, h, w, generator=g).cuda()
    q = (0.4 * torch.randn(B, Q, 32, generator=g)).cuda()
    summ, mx, sm, _ = S.summary_fwd(x, q)
    ds = torch.randn(summ.shape, generator=g).cuda()
    ref = S.bwd_summary(x, q, summ, mx, sm, ds)
    prev = lib().sqlx_sql_set_sm_budget(116)
    try:
        assert prev == 148
        got = S.bwd_summary(x, q, summ, mx, sm, ds)
        assert lib().sqlx_sql_set_sm_budget(3) == 116          # clamped to [8, 148]
        assert lib().sqlx_sql_set_sm_budget(1000) == 8
        assert lib().sqlx_sql_set_sm_budget(148) == 148
    finally:
        lib().sqlx_sql_set_sm_budget(148)
    for a, b, nm in zip(got, ref, ("d_x", "d_queries")):
        rel = float((a - b).abs().max() / b.abs().max())
        assert rel < 1e-5, (nm, rel)
