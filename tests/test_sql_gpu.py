"""Parity of the SQL block kernels (through the C ABI) against the reference golden vectors
(tests/golden/decoder_*.npz, produced by executing networks/depth_decoder_QTR.py) and against the CPU oracle.
Tolerance: depth 1e-4 relative (BASELINE.json north_star); gradients relative to max|grad|."""
import numpy as np
import pytest
import torch

from _cases import load_npz

pytestmark = pytest.mark.gpu


def _t(a, dev="cuda"):
    return torch.from_numpy(np.asarray(a)).to(dev)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


@pytest.mark.parametrize("name", ["decoder_full", "decoder_lite"])
def test_decoder_golden(name):
    """Drop-in Depth_Decoder_QueryTr with the reference's own state_dict: forward depth and all gradients."""
    import sqlx
    z = load_npz(name)
    st = load_npz(name + "_state")
    cls = sqlx.Lite_Depth_Decoder_QueryTr if int(z["lite"]) else sqlx.Depth_Decoder_QueryTr
    E, P, Q, D = int(z["E"]), int(z["P"]), int(z["Q"]), int(z["D"])
    dec = cls(in_channels=E, embedding_dim=E, patch_size=P, num_heads=4, query_nums=Q, dim_out=D,
              min_val=float(z["min_val"]), max_val=float(z["max_val"]))
    missing = dec.load_state_dict({k: torch.from_numpy(v) for k, v in st.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    dec = dec.cuda().eval()
    x0 = _t(z["x0"]).requires_grad_(True)
    # the golden run is true fp32 (CPU); cuDNN / cuBLAS default to TF32 convolutions on the GPU, which alone
    # moves depth by ~1e-3 -- switch it off for the PyTorch-side layers so the comparison isolates our kernels
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out = dec(x0)
    pred = out[("disp", 0)]
    ref = _t(z["out_pred"])
    assert pred.shape == ref.shape
    assert float(((pred - ref) / ref).abs().max()) < 1e-4
    (pred * _t(z["gout"])).sum().backward()
    conv = dec.convert_to_prob[0]
    assert _rel(conv.weight.grad.view(D, Q), _t(z["grad_Wp"])) < 2e-3
    assert _rel(conv.bias.grad, _t(z["grad_bp"])) < 2e-3
    assert torch.isfinite(x0.grad).all()


@pytest.mark.parametrize("name", ["decoder_full", "decoder_lite"])
def test_tail_golden(name):
    """Kernel-level: x (conv3x3 output) and queries captured from the reference run."""
    import sqlx
    z = load_npz(name)
    st = load_npz(name + "_state")
    D, Q = int(z["D"]), int(z["Q"])
    x = _t(z["x"]).requires_grad_(True)
    q = _t(z["queries"]).contiguous().requires_grad_(True)
    mlp = [_t(st["bins_regressor.%d.%s" % (i, k)]) for i in (0, 2, 4) for k in ("weight", "bias")]
    Wp = _t(st["convert_to_prob.0.weight"]).reshape(D, Q).clone().requires_grad_(True)
    bp = _t(st["convert_to_prob.0.bias"]).clone().requires_grad_(True)
    B = x.shape[0]
    F = torch.nn.functional

    def centers_fn(s):
        r = F.linear(s.reshape(B, -1), mlp[0], mlp[1])
        r = F.linear(F.leaky_relu(r, 0.01), mlp[2], mlp[3])
        r = F.linear(F.leaky_relu(r, 0.01), mlp[4], mlp[5])
        return sqlx.bin_centers(r, float(z["min_val"]), float(z["max_val"]))
    pred = sqlx.sql_tail(x, q, Wp, bp, centers_fn)
    ref = _t(z["out_pred"])
    assert float(((pred - ref) / ref).abs().max()) < 1e-4
    (pred * _t(z["gout"])).sum().backward()
    assert _rel(x.grad, _t(z["grad_x"])) < 2e-3
    assert _rel(q.grad, _t(z["grad_queries"])) < 2e-3
    assert _rel(Wp.grad, _t(z["grad_Wp"])) < 2e-3
    assert _rel(bp.grad, _t(z["grad_bp"])) < 2e-3
    # module-level FullQueryLayer: energy maps + summaries
    energy, summary = sqlx.FullQueryLayer()(x.detach(), q.detach())
    np.testing.assert_allclose(summary.cpu().numpy(), z["out_summary"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(energy.cpu().numpy()[:, :, ::4, ::4], z["out_energy_sample"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("cfg", [
    dict(B=2, E=32, h=24, w=40, Q=64, D=64),
    dict(B=1, E=32, h=17, w=23, Q=120, D=128),       # ragged pixel count, Q not a multiple of 8
    dict(B=3, E=16, h=16, w=16, Q=5, D=7),           # tiny, odd Q and D
    dict(B=1, E=32, h=20, w=32, Q=128, D=128),       # maximum Q, D of the fp32 path
    dict(B=1, E=64, h=12, w=20, Q=64, D=64),         # widest embedding
    dict(B=2, E=32, h=3, w=5, Q=8, D=8),             # fewer pixels than one tile
])
def test_oracle_fp64(cfg):
    """sql_tail + FullQueryLayer vs the oracle evaluated in float64, forward and every gradient."""
    import sqlx
    from oracle import sqldepth_oracle as O
    B, E, h, w, Q, D = (cfg[k] for k in "BEhwQD")
    g = torch.Generator().manual_seed(100 + Q)
    x = torch.randn(B, E, h, w, generator=g)
    q = 0.4 * torch.randn(B, Q, E, generator=g)
    Wp = 0.3 * torch.randn(D, Q, generator=g)
    bp = 0.1 * torch.randn(D, generator=g)
    W1 = torch.randn(D, Q * E, generator=g) / (Q * E) ** 0.5
    b1 = 0.1 * torch.randn(D, generator=g)
    gout = torch.randn(B, 1, h, w, generator=g)
    F = torch.nn.functional
    # oracle, float64
    leaves = [t.double().requires_grad_(True) for t in (x, q, Wp, bp, W1, b1)]
    xd, qd, Wd, bd, W1d, b1d = leaves
    energy, summ = O.full_query(xd, qd)
    centers = O.bin_centers(F.linear(summ.reshape(B, -1), W1d, b1d), 0.001, 80.0)
    pred_ref = O.bins_expectation(energy, Wd, bd, centers)
    gref = torch.autograd.grad((pred_ref * gout.double()).sum(), leaves)
    # CUDA
    cl = [t.cuda().requires_grad_(True) for t in (x, q, Wp, bp, W1, b1)]
    xc, qc, Wc, bc, W1c, b1c = cl
    pred = sqlx.sql_tail(xc, qc, Wc, bc, lambda s: sqlx.bin_centers(F.linear(s.reshape(B, -1), W1c, b1c), 0.001, 80.0),
                         (W1c, b1c))
    assert float(((pred.cpu().double() - pred_ref) / pred_ref).abs().max()) < 1e-4
    gg = torch.autograd.grad((pred * gout.cuda()).sum(), cl)
    for a, b_, nm in zip(gg, gref, ("x", "queries", "Wp", "bp", "W1", "b1")):
        assert _rel(a.cpu().double(), b_) < 2e-3, nm
    # FullQueryLayer backward with upstream gradients on both outputs
    xe = x.cuda().requires_grad_(True)
    qe = q.cuda().requires_grad_(True)
    en, su = sqlx.FullQueryLayer()(xe, qe)
    ge = torch.randn(en.shape, generator=g)
    gs = torch.randn(su.shape, generator=g)
    ga = torch.autograd.grad((en * ge.cuda()).sum() + (su * gs.cuda()).sum(), [xe, qe])
    xr, qr = x.double().requires_grad_(True), q.double().requires_grad_(True)
    en_r, su_r = O.full_query(xr, qr)
    gr = torch.autograd.grad((en_r * ge.double()).sum() + (su_r * gs.double()).sum(), [xr, qr])
    assert float((en.cpu().double() - en_r).abs().max()) < 1e-4
    assert float((su.cpu().double() - su_r).abs().max()) < 1e-4
    assert _rel(ga[0].cpu().double(), gr[0]) < 2e-3
    assert _rel(ga[1].cpu().double(), gr[1]) < 2e-3


def test_full_size_properties():
    """BASELINE config-2 size (B=12, x0 32x96x320, Q=D=64): size-independent properties.
    (1) pred lies strictly inside (min centre, max centre); (2) permuting pixels permutes pred and leaves the
    summaries unchanged; (3) a constant shift of all logits (bp + c) leaves pred unchanged."""
    import sqlx
    torch.manual_seed(5)
    B, E, h, w, Q, D = 12, 32, 96, 320, 64, 64
    x = torch.randn(B, E, h, w, device="cuda")
    q = 0.4 * torch.randn(B, Q, E, device="cuda")
    Wp = 0.3 * torch.randn(D, Q, device="cuda")
    bp = 0.1 * torch.randn(D, device="cuda")
    centers = torch.sort(torch.rand(B, D, device="cuda") * 80, dim=1).values
    from sqlx import sql as S
    pred = S.pred_fwd(x, q, Wp, bp, centers)
    pv = pred.view(B, -1)
    assert bool((pv >= centers[:, :1] * (1 - 1e-6)).all()) and bool((pv <= centers[:, -1:] * (1 + 1e-6)).all())
    pred2 = S.pred_fwd(x, q, Wp, bp + 3.0, centers)
    assert float(((pred - pred2) / pred).abs().max()) < 1e-5
    perm = torch.randperm(h * w, device="cuda")
    xp = x.view(B, E, -1)[:, :, perm].view(B, E, h, w).contiguous()
    predp = S.pred_fwd(xp, q, Wp, bp, centers)
    assert float((predp.view(B, -1) - pred.view(B, -1)[:, perm]).abs().max()) == 0.0
    s1 = S.summary_fwd(x, q)[0]
    s2 = S.summary_fwd(xp, q)[0]
    assert float((s1 - s2).abs().max()) < 1e-5


def test_config4_size_properties():
    """BASELINE config-4 size (EffB5: x0 32x320x1024 = 327,680 pixels per sample, Q = D = 128), through the public
    sql_tail (tensor-core mixed-weight path), forward and backward: (1) pred inside the centre range; (2) a constant
    shift of the logits leaves pred and the gradient wrt x unchanged; (3) permuting the pixels permutes pred and
    d_x and leaves d_queries / d_Wp unchanged (every reduction over pixels is order-independent up to rounding)."""
    import sqlx
    torch.manual_seed(6)
    B, E, h, w, Q, D = 2, 32, 320, 1024, 128, 128
    x = torch.randn(B, E, h, w, device="cuda")
    q = 0.4 * torch.randn(B, Q, E, device="cuda")
    Wp = 0.3 * torch.randn(D, Q, device="cuda")
    bp = 0.1 * torch.randn(D, device="cuda")
    lin = torch.nn.Linear(Q * E, D).cuda()
    gout = torch.randn(B, 1, h, w, device="cuda")

    def run(xin, bias, g):
        xl, ql, Wl = xin.clone().requires_grad_(True), q.clone().requires_grad_(True), Wp.clone().requires_grad_(True)
        cen = []

        def centers_fn(s):
            c = sqlx.bin_centers(lin(s.reshape(B, -1)), 0.01, 80.0)
            cen.append(c.detach())
            return c
        pred = sqlx.sql_tail(xl, ql, Wl, bias, centers_fn, tuple(lin.parameters()))
        gx, gq, gW = torch.autograd.grad((pred * g).sum(), [xl, ql, Wl])
        return pred.detach(), gx, gq, gW, cen[0]

    pred, gx, gq, gW, cen = run(x, bp, gout)
    pv = pred.view(B, -1)
    assert bool(torch.isfinite(pv).all()) and bool(torch.isfinite(gx).all())
    assert bool((pv >= cen[:, :1] * (1 - 1e-5)).all()) and bool((pv <= cen[:, -1:] * (1 + 1e-5)).all())
    pred2, gx2, _, _, _ = run(x, bp + 3.0, gout)
    assert float(((pred - pred2) / pred).abs().max()) < 1e-4
    assert float((gx - gx2).abs().max() / gx.abs().max()) < 2e-3
    perm = torch.randperm(h * w, device="cuda")
    xp = x.view(B, E, -1)[:, :, perm].view(B, E, h, w).contiguous()
    gp = gout.view(B, 1, -1)[:, :, perm].view(B, 1, h, w).contiguous()
    predp, gxp, gqp, gWp, _ = run(xp, bp, gp)
    assert float(((predp.view(B, -1) - pv[:, perm]) / pv[:, perm]).abs().max()) < 1e-4
    assert float((gxp.view(B, E, -1) - gx.view(B, E, -1)[:, :, perm]).abs().max() / gx.abs().max()) < 2e-3
    assert float((gqp - gq).abs().max() / gq.abs().max()) < 2e-3
    assert float((gWp - gW).abs().max() / gW.abs().max()) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("B,Q,E,D", [(12, 64, 32, 64), (2, 120, 32, 128), (16, 16, 32, 24)])
def test_bins_head_matches_torch(B, Q, E, D):
    """csrc/bins_head.cu (three Linear layers + centre arithmetic, depth_decoder_QTR.py:48-66) against the same
    arithmetic in float64 PyTorch: values and every gradient."""
    from sqlx import sql as S
    torch.manual_seed(B + Q)
    nn = torch.nn
    reg = nn.Sequential(nn.Linear(Q * E, 16 * Q), nn.LeakyReLU(), nn.Linear(16 * Q, 256), nn.LeakyReLU(),
                        nn.Linear(256, D)).cuda()
    s = torch.randn(B, Q * E, device="cuda", requires_grad=True)
    g = torch.randn(B, D, device="cuda")
    c = S.bins_head(s, reg, 0.001, 80.0)
    grads = torch.autograd.grad(c, [s] + list(reg.parameters()), g)
    reg64 = nn.Sequential(nn.Linear(Q * E, 16 * Q), nn.LeakyReLU(), nn.Linear(16 * Q, 256), nn.LeakyReLU(),
                          nn.Linear(256, D)).double().cuda()
    reg64.load_state_dict({k: v.double() for k, v in reg.state_dict().items()})
    s64 = s.detach().double().requires_grad_(True)
    c64 = S.bin_centers(reg64(s64), 0.001, 80.0)
    grads64 = torch.autograd.grad(c64, [s64] + list(reg64.parameters()), g.double())
    assert float((c.double() - c64).abs().max() / c64.abs().max()) < 1e-5
    for a, b in zip(grads, grads64):
        assert float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)) < 1e-4


def test_param_grad_hook_overlap_semantics():
    """sql_tail(on_param_grads=...): the hook sees every parameter gradient of the tail before the summary-path kernel
    runs; an in-place exchange it starts on a side stream (here: halving, standing in for the all-reduce average of two
    identical ranks' doubled gradients) is what ends up in .grad, and the gradients of x / queries are untouched."""
    import sqlx
    torch.manual_seed(3)
    B, E, h, w, Q, D = 4, 32, 48, 64, 64, 64
    x = torch.randn(B, E, h, w, device="cuda")
    q = 0.4 * torch.randn(B, Q, E, device="cuda")
    conv = torch.nn.Conv2d(Q, D, 1).cuda()
    mlp = torch.nn.Sequential(torch.nn.Linear(Q * E, 16 * Q), torch.nn.LeakyReLU(), torch.nn.Linear(16 * Q, 256),
                              torch.nn.LeakyReLU(), torch.nn.Linear(256, D)).cuda()
    params = [conv.weight, conv.bias] + list(mlp.parameters())
    gout = torch.randn(B, 1, h, w, device="cuda")
    side = torch.cuda.Stream()
    seen = []

    def hook(grads):
        seen.append(len(grads))
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for g in grads:
                g.mul_(0.5)
        return lambda: torch.cuda.current_stream().wait_stream(side)

    def run(cb):
        xs, qs = x.clone().requires_grad_(True), q.clone().requires_grad_(True)
        for p in params:
            p.grad = None
        pred = sqlx.sql_tail(xs, qs, conv.weight.view(D, Q), conv.bias,
                             lambda s: sqlx.sql.bins_head(s.reshape(B, Q * E), mlp, 0.001, 80.0), tuple(mlp.parameters()),
                             on_param_grads=cb)
        (pred * gout).sum().backward()
        torch.cuda.synchronize()
        return [xs.grad.clone(), qs.grad.clone()] + [p.grad.clone() for p in params]

    ref = run(None)
    got = run(hook)
    assert seen == [len(params)]
    assert _rel(got[0], ref[0]) < 1e-6 and _rel(got[1], ref[1]) < 1e-6
    for a, b in zip(got[2:], ref[2:]):
        assert _rel(a, 0.5 * b) < 1e-6


@pytest.mark.parametrize("B,Q,D,E", [(12, 64, 64, 32), (8, 128, 128, 32), (3, 12, 16, 32), (2, 120, 128, 32)])
def test_mix_weight_glue_kernels(B, Q, D, E):
    """csrc/sql_glue.cu: M = Wp K, dWp = sum_b dM K^T, dK += Wp^T dM against float64 matmuls (these replaced three cuBLAS
    calls inside the step); both halves of the backward launch, together and separately."""
    from sqlx import sql as S
    g = torch.Generator().manual_seed(B + Q)
    Wp = torch.randn(D, Q, generator=g).cuda()
    K = torch.randn(B, Q, E, generator=g).cuda()
    dM = torch.randn(B, D, E, generator=g).cuda()
    dK0 = torch.randn(B, Q, E, generator=g).cuda()
    M = S.mix_weights(Wp, K)
    assert _rel(M.double(), torch.matmul(Wp.double(), K.double())) < 1e-6
    dWp_ref = torch.einsum("bde,bqe->dq", dM.double(), K.double())
    dK_ref = dK0.double() + torch.matmul(Wp.double().t(), dM.double())
    dK = dK0.clone()
    dWp = S.mix_weights_bwd(dM, K, Wp, dK)
    assert _rel(dWp.double(), dWp_ref) < 1e-6 and _rel(dK.double(), dK_ref) < 1e-6
    buf = torch.full((D, Q), float("nan"), device="cuda")
    S.mix_weights_bwd(dM, K, Wp, None, d_Wp=buf)                  # weight half only, into a caller-owned buffer
    assert torch.equal(buf, dWp)
    dK2 = dK0.clone()
    S.mix_weights_bwd(dM, K, Wp, dK2, want_d_Wp=False)            # query half only
    assert torch.equal(dK2, dK)
