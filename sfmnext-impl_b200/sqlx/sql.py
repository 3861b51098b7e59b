"""Self Query Layer tail of the SQLdepth depth decoder on libsqlx kernels.

Mirrors FullQueryLayer.forward (networks/layers.py:7-21) and the tail of Depth_Decoder_QueryTr.forward
(networks/depth_decoder_QTR.py:47-74).  The [pixels x queries] self-cost volume and the [D x pixels]
logits / probabilities never reach HBM: each kernel recomputes them tile by tile on chip.  The tiny bins
MLP (nn.Linear x3) and the centres arithmetic stay in PyTorch (cuBLAS GEMV-ish work);
autograd carries d_centers -> d_summary through them between the two backward kernels.
"""
import torch
import torch.nn.functional as F

from ._lib import check, lib, ptr, require_cuda, stream_ptr


def _f32c(t):
    return t.detach().contiguous().float() if t is not None else None


def _workspace(B, E, Q, D, n, device):
    nbytes = lib().sqlx_sql_workspace_bytes(B, E, Q, D, n)
    return torch.empty(nbytes, device=device, dtype=torch.uint8), nbytes


def summary_fwd(x, queries, want_energy=False, version=2):
    """x [B,E,h,w], queries [B,Q,E] -> summary [B,Q,E], row_max [B,Q], row_sum [B,Q], energy [B,Q,h,w] | None.
    version=1: the round-1 tensor-core kernel (cross-check of the warp-specialised one; no energy output)."""
    require_cuda(x, queries)
    B, E, h, w = x.shape
    Q = queries.shape[1]
    if queries.shape[0] != B or queries.shape[2] != E:
        # same guard as networks/layers.py:16
        raise AssertionError("Number of channels in x and Embedding dimension (at dim 2) of K matrix must match")
    n = h * w
    dev = x.device
    summary = torch.empty(B, Q, E, device=dev, dtype=torch.float32)
    row_max = torch.empty(B, Q, device=dev, dtype=torch.float32)
    row_sum = torch.empty(B, Q, device=dev, dtype=torch.float32)
    energy = torch.empty(B, Q, h, w, device=dev, dtype=torch.float32) if want_energy else None
    ws, nbytes = _workspace(B, E, Q, 0, n, dev)
    if version == 1:
        check(lib().sqlx_sql_summary_fwd_v1(ptr(x), ptr(queries), B, E, Q, n, ptr(summary), ptr(row_max), ptr(row_sum),
                                            ptr(ws), nbytes, stream_ptr()), "sqlx_sql_summary_fwd_v1")
        return summary, row_max, row_sum, None
    check(lib().sqlx_sql_summary_fwd(ptr(x), ptr(queries), B, E, Q, n, ptr(summary), ptr(row_max), ptr(row_sum),
                                     ptr(energy), ptr(ws), nbytes, stream_ptr()), "sqlx_sql_summary_fwd")
    return summary, row_max, row_sum, energy


def pred_fwd(x, queries, Wp, bp, centers):
    require_cuda(x, queries, Wp, bp, centers)
    B, E, h, w = x.shape
    Q, D = queries.shape[1], Wp.shape[0]
    pred = torch.empty(B, 1, h, w, device=x.device, dtype=torch.float32)
    check(lib().sqlx_sql_pred_fwd(ptr(x), ptr(queries), ptr(Wp), ptr(bp), ptr(centers), B, E, Q, D, h * w, ptr(pred),
                                  stream_ptr()), "sqlx_sql_pred_fwd")
    return pred


def bwd_reduce(x, queries, Wp, bp, centers, g_pred):
    B, E, h, w = x.shape
    Q, D = queries.shape[1], Wp.shape[0]
    dev = x.device
    d_centers = torch.empty(B, D, device=dev, dtype=torch.float32)
    d_Wp = torch.empty(D, Q, device=dev, dtype=torch.float32)
    d_bp = torch.empty(D, device=dev, dtype=torch.float32)
    ws, nbytes = _workspace(B, E, Q, D, h * w, dev)
    check(lib().sqlx_sql_bwd_reduce(ptr(x), ptr(queries), ptr(Wp), ptr(bp), ptr(centers), None, ptr(g_pred), B, E, Q, D,
                                    h * w, ptr(d_centers), ptr(d_Wp), ptr(d_bp), ptr(ws), nbytes, stream_ptr()),
          "sqlx_sql_bwd_reduce")
    return d_centers, d_Wp, d_bp


def bwd_dx(x, queries, Wp=None, bp=None, centers=None, g_pred=None, summary=None, row_max=None, row_sum=None,
           d_summary=None, g_energy=None):
    B, E, h, w = x.shape
    Q = queries.shape[1]
    D = Wp.shape[0] if Wp is not None else 0
    dev = x.device
    d_x = torch.empty_like(x)
    d_q = torch.empty_like(queries)
    ws, nbytes = _workspace(B, E, Q, D, h * w, dev)
    check(lib().sqlx_sql_bwd_dx(ptr(x), ptr(queries), ptr(Wp), ptr(bp), ptr(centers), None, ptr(g_pred), ptr(summary),
                                ptr(row_max), ptr(row_sum), ptr(d_summary), ptr(g_energy), B, E, Q, D, h * w,
                                ptr(d_x), ptr(d_q), ptr(ws), nbytes, stream_ptr()), "sqlx_sql_bwd_dx")
    return d_x, d_q


# ----------------------------------------------------------------------------- mixed-weight decomposition (tensor cores)
def use_mix(E, Q, D, n):
    """True when the tensor-core kernels take the shape: the decoder tail then runs as
    logits = (Wp K) x + b (regression contracts over E = 32, no Wp tiles on chip; csrc/sql_tc.cu)."""
    L = lib()
    return bool(L.sqlx_sql_get_tensor_cores()) and bool(L.sqlx_sql_tc_supported(E, Q, D, n))


def mix_weights(Wp, queries):
    """Mx [B,D,E] = Wp [D,Q] . queries [B,Q,E]: the 1x1 conv weight folded into the queries (one small kernel)."""
    B, Q, E = queries.shape
    D = Wp.shape[0]
    Mx = torch.empty(B, D, E, device=queries.device, dtype=torch.float32)
    check(lib().sqlx_sql_mix_weights(ptr(Wp), ptr(queries), B, Q, D, E, ptr(Mx), stream_ptr()), "sqlx_sql_mix_weights")
    return Mx


def mix_weights_bwd(d_M, queries, Wp, d_queries=None, d_Wp=None, want_d_Wp=True):
    """d_Wp [D,Q] = sum_b d_M[b] queries[b]^T (written into `d_Wp` when given) and / or d_queries += Wp^T d_M, in one
    launch; d_queries=None or want_d_Wp=False skips that half."""
    B, Q, E = queries.shape
    D = Wp.shape[0]
    if want_d_Wp and d_Wp is None:
        d_Wp = torch.empty(D, Q, device=queries.device, dtype=torch.float32)
    check(lib().sqlx_sql_mix_weights_bwd(ptr(d_M), ptr(queries), ptr(Wp), B, Q, D, E, 1,
                                         ptr(d_Wp) if want_d_Wp else None, ptr(d_queries), stream_ptr()),
          "sqlx_sql_mix_weights_bwd")
    return d_Wp


def _mix_workspace(B, Q, D, n, device):
    nbytes = lib().sqlx_sql_mix_workspace_bytes(B, Q, D, n)
    return torch.empty(nbytes, device=device, dtype=torch.uint8), nbytes


def pred_mix_fwd(x, Mx, bp, centers, version=2):
    """pred [B,1,h,w] and the per-pixel softmax statistics [2,B,n] the backward reads (version=1: the round-1 kernel, no
    statistics -- A/B and cross-check only)."""
    B, E, h, w = x.shape
    D = Mx.shape[1]
    pred = torch.empty(B, 1, h, w, device=x.device, dtype=torch.float32)
    if version == 1:
        check(lib().sqlx_sql_pred_mix_fwd_v1(ptr(x), ptr(Mx), ptr(bp), ptr(centers), B, E, D, h * w, ptr(pred),
                                             stream_ptr()), "sqlx_sql_pred_mix_fwd_v1")
        return pred, None
    stats = torch.empty(2, B, h * w, device=x.device, dtype=torch.float32)
    check(lib().sqlx_sql_pred_mix_fwd(ptr(x), ptr(Mx), ptr(bp), ptr(centers), B, E, D, h * w, ptr(pred), ptr(stats),
                                      stream_ptr()), "sqlx_sql_pred_mix_fwd")
    return pred, stats


def bwd_pred_mix(x, Mx, bp, centers, g_pred, pred=None, stats=None, d_bp=None):
    """Regression-path backward.  pred / stats from pred_mix_fwd; stats=None runs the round-1 kernel (cross-check)."""
    B, E, h, w = x.shape
    D = Mx.shape[1]
    dev = x.device
    d_M = torch.empty_like(Mx)
    if d_bp is None:
        d_bp = torch.empty(D, device=dev, dtype=torch.float32)
    d_centers = torch.empty(B, D, device=dev, dtype=torch.float32)
    d_x = torch.empty_like(x)
    ws, nbytes = _mix_workspace(B, 1, D, h * w, dev)
    if stats is None:
        check(lib().sqlx_sql_bwd_pred_mix_v1(ptr(x), ptr(Mx), ptr(bp), ptr(centers), ptr(g_pred), B, E, D, h * w, ptr(d_M),
                                             ptr(d_bp), ptr(d_centers), ptr(d_x), ptr(ws), nbytes, stream_ptr()),
              "sqlx_sql_bwd_pred_mix_v1")
    else:
        check(lib().sqlx_sql_bwd_pred_mix(ptr(x), ptr(Mx), ptr(bp), ptr(centers), ptr(g_pred), ptr(pred), ptr(stats), B, E,
                                          D, h * w, ptr(d_M), ptr(d_bp), ptr(d_centers), ptr(d_x), ptr(ws), nbytes,
                                          stream_ptr()), "sqlx_sql_bwd_pred_mix")
    return d_M, d_bp, d_centers, d_x


def bwd_summary(x, queries, summary, row_max, row_sum, d_summary, d_x=None, version=2):
    """Summary-path backward; accumulates into d_x when given (else writes a fresh tensor).  version=1: round-1 kernel."""
    B, E, h, w = x.shape
    Q = queries.shape[1]
    accumulate = d_x is not None
    if d_x is None:
        d_x = torch.empty_like(x)
    d_q = torch.empty_like(queries)
    ws, nbytes = _mix_workspace(B, Q, 0, h * w, x.device)
    fn = lib().sqlx_sql_bwd_summary if version == 2 else lib().sqlx_sql_bwd_summary_v1
    check(fn(ptr(x), ptr(queries), ptr(summary), ptr(row_max), ptr(row_sum), ptr(d_summary), B, E, Q, h * w, int(accumulate),
             ptr(d_x), ptr(d_q), ptr(ws), nbytes, stream_ptr()), "sqlx_sql_bwd_summary")
    return d_x, d_q


# ----------------------------------------------------------------------------- module-level FullQueryLayer
class _FullQuery(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, queries):
        xc, qc = _f32c(x), _f32c(queries)
        summary, row_max, row_sum, energy = summary_fwd(xc, qc, want_energy=True)
        ctx.save_for_backward(xc, qc, summary, row_max, row_sum)
        return energy, summary

    @staticmethod
    def backward(ctx, g_energy, g_summary):
        xc, qc, summary, row_max, row_sum = ctx.saved_tensors
        d_x, d_q = bwd_dx(xc, qc, summary=summary, row_max=row_max, row_sum=row_sum,
                          d_summary=_f32c(g_summary), g_energy=_f32c(g_energy))
        return d_x, d_q


class FullQueryLayer(torch.nn.Module):
    """Drop-in for networks.layers.FullQueryLayer (networks/layers.py:4-21): same signature, same outputs.
    Materialises the energy maps (the reference API returns them); the fused decoder below does not."""

    def forward(self, x, K):
        return _FullQuery.apply(x, K)


# ----------------------------------------------------------------------------- fused decoder tail
def bin_centers(raw, min_val, max_val, norm="linear"):
    """depth_decoder_QTR.py:51-66: regressor output [B,D] -> bin centres [B,D]."""
    if norm == "linear":
        y = torch.relu(raw) + 0.1
    else:
        y = torch.sigmoid(raw)
    y = y / y.sum(dim=1, keepdim=True)
    widths = (max_val - min_val) * y
    widths = F.pad(widths, (1, 0), mode="constant", value=min_val)
    edges = torch.cumsum(widths, dim=1)
    return 0.5 * (edges[:, :-1] + edges[:, 1:])


class _BinsHead(torch.autograd.Function):
    """summary [B, Q*E] -> bin centres [B, D] through the three Linear layers of bins_regressor and the centre
    arithmetic, on the weight-streaming kernels of csrc/bins_head.cu (4 launches forward, 4 backward)."""

    @staticmethod
    def forward(ctx, s, W1, b1, W2, b2, W3, b3, min_val, max_val, grad_buffers=None):
        L = lib()
        ctx.grad_buffers = grad_buffers
        sc = _f32c(s)
        Ws = [_f32c(W1), _f32c(W2), _f32c(W3)]
        bs = [_f32c(b1), _f32c(b2), _f32c(b3)]
        B = sc.shape[0]
        acts = [sc]
        for i in range(3):
            N, K = Ws[i].shape
            y = torch.empty(B, N, device=sc.device, dtype=torch.float32)
            check(L.sqlx_head_linear_fwd(ptr(Ws[i]), ptr(bs[i]), ptr(acts[-1]), B, N, K, int(i < 2), ptr(y), stream_ptr()),
                  "sqlx_head_linear_fwd")
            acts.append(y)
        raw = acts[3]
        D = raw.shape[1]
        centers = torch.empty_like(raw)
        check(L.sqlx_head_centers_fwd(ptr(raw), B, D, float(min_val), float(max_val), ptr(centers), stream_ptr()),
              "sqlx_head_centers_fwd")
        ctx.save_for_backward(*acts, centers, *Ws)
        ctx.lims = (float(min_val), float(max_val))
        return centers

    @staticmethod
    def backward(ctx, g_centers):
        L = lib()
        s, h1, h2, raw, centers, W1, W2, W3 = ctx.saved_tensors
        B, D = raw.shape
        g = _f32c(g_centers)
        d_raw = torch.empty_like(raw)
        check(L.sqlx_head_centers_bwd(ptr(raw), ptr(centers), ptr(g), B, D, ctx.lims[0], ctx.lims[1], ptr(d_raw),
                                      stream_ptr()), "sqlx_head_centers_bwd")
        acts, Ws = [s, h1, h2, raw], [W1, W2, W3]
        dy = d_raw
        dWs, dbs = [None] * 3, [None] * 3
        gb = ctx.grad_buffers          # caller-owned [dW1, db1, dW2, db2, dW3, db3] (views of a flat gradient bucket)
        for i in (2, 1, 0):
            N, K = Ws[i].shape
            dz = torch.empty(B, N, device=g.device, dtype=torch.float32)
            dWs[i] = gb[2 * i] if gb is not None else torch.empty_like(Ws[i])
            dbs[i] = gb[2 * i + 1] if gb is not None else torch.empty(N, device=g.device, dtype=torch.float32)
            need_dx = i > 0 or ctx.needs_input_grad[0]
            dx = torch.empty(B, K, device=g.device, dtype=torch.float32) if need_dx else None
            check(L.sqlx_head_linear_bwd(ptr(Ws[i]), ptr(acts[i]), ptr(acts[i + 1]), ptr(dy), B, N, K, int(i < 2), ptr(dz),
                                         ptr(dWs[i]), ptr(dbs[i]), ptr(dx), stream_ptr()), "sqlx_head_linear_bwd")
            dy = dx
        if gb is not None:              # the gradients live in the caller's buffers: nothing for autograd to accumulate
            return dy, None, None, None, None, None, None, None, None, None
        return dy, dWs[0], dbs[0], dWs[1], dbs[1], dWs[2], dbs[2], None, None, None


def bins_head(summary_flat, regressor, min_val, max_val, norm="linear", grad_buffers=None):
    """bin centres [B,D] = centres(bins_regressor(summary)) (depth_decoder_QTR.py:48-66).  `regressor` is the
    nn.Sequential(Linear, LeakyReLU, Linear, LeakyReLU, Linear) of the reference decoder.  The weight-streaming
    kernels take batch <= 16, norm == 'linear', in_features % 4 == 0 and dim_out <= 256 (every reference config);
    other cases run the same arithmetic through cuBLAS + elementwise PyTorch ops on the GPU.
    grad_buffers: optional [dW1, db1, dW2, db2, dW3, db3] the backward WRITES the parameter gradients into (e.g. views
    of one flat all-reduce bucket) instead of handing them to autograd."""
    lins = [m for m in regressor if isinstance(m, torch.nn.Linear)]
    acts = [m for m in regressor if isinstance(m, torch.nn.LeakyReLU)]
    ok = (norm == "linear" and len(lins) == 3 and len(acts) == 2 and all(a.negative_slope == 0.01 for a in acts) and
          summary_flat.is_cuda and summary_flat.shape[0] <= 16 and lins[2].out_features <= 256 and
          all(l.in_features % 4 == 0 and l.bias is not None for l in lins))
    if not ok:
        if grad_buffers is not None:
            raise ValueError("grad_buffers needs the bins-head kernels (batch <= 16, norm='linear', dim_out <= 256)")
        return bin_centers(regressor(summary_flat), min_val, max_val, norm)
    return _BinsHead.apply(summary_flat, lins[0].weight, lins[0].bias, lins[1].weight, lins[1].bias, lins[2].weight,
                           lins[2].bias, min_val, max_val, grad_buffers)


class _SqlTail(torch.autograd.Function):
    """x, queries, Wp, bp -> pred, with the bins MLP evaluated by a caller-supplied closure.

    forward : summary kernel -> centers = centers_fn(summary) (PyTorch) -> pred kernel
    backward: reduce kernel (d_centers, d_Wp, d_bp) -> autograd through centers_fn (d_summary, MLP grads)
              -> dx kernel (d_x, d_queries): two passes over x instead of the reference's ~10.
    """

    @staticmethod
    def forward(ctx, x, queries, Wp, bp, centers_fn, n_params, on_param_grads, head_grad_out, on_stage, *params):
        ctx.set_materialize_grads(False)
        ctx.head_grad_out = head_grad_out
        ctx.on_stage = on_stage
        xc, qc, Wc, bc = _f32c(x), _f32c(queries), _f32c(Wp), _f32c(bp)
        summary, row_max, row_sum, _ = summary_fwd(xc, qc)
        with torch.enable_grad():
            s_leaf = summary.detach().requires_grad_(True)
            centers = centers_fn(s_leaf)
        cc = _f32c(centers)
        B, E, h, w = xc.shape
        ctx.mix = use_mix(E, qc.shape[1], Wc.shape[0], h * w)
        if ctx.mix:
            Mx = mix_weights(Wc, qc)                     # [B,D,E] = Wp . K
            pred, stats = pred_mix_fwd(xc, Mx, bc, cc)
            ctx.save_for_backward(xc, qc, Wc, bc, cc, summary, row_max, row_sum, Mx, pred, stats)
        else:
            pred = pred_fwd(xc, qc, Wc, bc, cc)
            ctx.save_for_backward(xc, qc, Wc, bc, cc, summary, row_max, row_sum)
        ctx.graph = (s_leaf, centers)
        ctx.params = params
        ctx.on_param_grads = on_param_grads
        if on_stage is not None:
            on_stage("after_pred_fwd")         # the last one-CTA-per-SM kernel of the forward is enqueued
        return pred

    @staticmethod
    def backward(ctx, g_pred):
        s_leaf, centers = ctx.graph
        none = (None,) * (9 + len(ctx.params))
        if g_pred is None:
            return none
        g = _f32c(g_pred)
        # head_grad_out = (d_Wp, d_bp) buffers: the 1x1 conv gradients are WRITTEN there (views of the caller's flat
        # all-reduce bucket, like bins_head(grad_buffers=...)) and autograd gets nothing to accumulate for Wp / bp
        out_Wp, out_bp = ctx.head_grad_out if ctx.head_grad_out is not None else (None, None)
        need = [p for p in ctx.params if p.requires_grad]
        if ctx.mix:
            xc, qc, Wc, bc, cc, summary, row_max, row_sum, Mx, pred, stats = ctx.saved_tensors
            d_M, d_bp, d_centers, d_x = bwd_pred_mix(xc, Mx, bc, cc, g, pred, stats, d_bp=out_bp)
            if ctx.on_stage is not None:
                ctx.on_stage("after_bwd_pred")     # what follows is ~70 us of small grids (partial sums, bins-head backward)
        else:
            xc, qc, Wc, bc, cc, summary, row_max, row_sum = ctx.saved_tensors
            d_centers, d_Wp, d_bp = bwd_reduce(xc, qc, Wc, bc, cc, g)
            if ctx.on_stage is not None:
                ctx.on_stage("after_bwd_pred")
            if out_Wp is not None:
                out_Wp.copy_(d_Wp.view_as(out_Wp)); out_bp.copy_(d_bp)
                d_Wp, d_bp = out_Wp, out_bp
        grads = torch.autograd.grad(centers, [s_leaf] + need, d_centers, allow_unused=True)
        d_summary = grads[0] if grads[0] is not None else torch.zeros_like(summary)
        early = ctx.mix and ctx.on_param_grads is not None
        if early:     # d_Wp = sum_b dM K^T now (the exchange needs it); the d_K half runs after the summary-path kernel
            d_Wp = mix_weights_bwd(d_M, qc, Wc, None, d_Wp=out_Wp)
        join = None
        if ctx.on_param_grads is not None:
            # every parameter gradient of the tail exists now; the summary-path backward (the longest kernel of the
            # tail) is still to run: the caller's gradient exchange overlaps it
            join = ctx.on_param_grads([d_Wp, d_bp] + [g_ for g_ in grads[1:] if g_ is not None])
        if ctx.mix:
            # beside a running exchange the summary-path kernel leaves `exchange_sm_reserve` SMs to the communication
            # kernel (it would otherwise queue behind this kernel's one-CTA-per-SM grid and run exposed afterwards)
            prev = lib().sqlx_sql_set_sm_budget(148 - exchange_sm_reserve) if (join is not None and exchange_sm_reserve) else None
            try:
                d_x, d_q = bwd_summary(xc, qc, summary, row_max, row_sum, _f32c(d_summary), d_x=d_x)
            finally:
                if prev is not None:
                    lib().sqlx_sql_set_sm_budget(prev)
            if early:
                mix_weights_bwd(d_M, qc, Wc, d_q, want_d_Wp=False)         # d_q += Wp^T dM
            else:
                d_Wp = mix_weights_bwd(d_M, qc, Wc, d_q, d_Wp=out_Wp)      # d_Wp and d_q += Wp^T dM in one launch
        else:
            d_x, d_q = bwd_dx(xc, qc, Wc, bc, cc, g, summary, row_max, row_sum, _f32c(d_summary))
        if join is not None:
            join()
        it = iter(grads[1:])
        d_params = tuple((next(it) if p.requires_grad else None) for p in ctx.params)
        ctx.graph = None
        if out_Wp is not None:
            return (d_x, d_q, None, None, None, None, None, None, None) + d_params
        return (d_x, d_q, d_Wp.view_as(Wc), d_bp, None, None, None, None, None) + d_params


# SMs left free for the communication kernel while the summary-path backward runs beside an in-step gradient exchange
# (0 = none).  Set by the caller that owns the exchange (sqlx.hotpath.HotPath, bench.py --exchange-sms).
exchange_sm_reserve = 0


def sql_tail(x, queries, Wp, bp, centers_fn, params=(), on_param_grads=None, head_grad_out=None, on_stage=None):
    """pred [B,1,h,w] = sum_d softmax_d(Wp (x^T K) + bp) * centers_fn(summary(x, K)).

    on_param_grads(list of gradient tensors) -> join callable or None: called in the backward as soon as the gradients
    of Wp, bp and `params` exist (before the summary-path kernel runs), e.g. to start their all-reduce on a side
    stream; the returned callable is invoked once the remaining backward kernels are enqueued.
    head_grad_out = (d_Wp [D,Q], d_bp [D]) buffers: the gradients of Wp / bp are written there instead of being
    returned to autograd (views of a flat gradient bucket: no pack / unpack around the all-reduce).
    on_stage(name): optional callback at points of the backward where a caller may fork independent work onto another
    stream; "after_pred_fwd" = the forward's last kernel is enqueued; "after_bwd_pred" = the regression-path backward
    kernel is enqueued and a stretch of small grids follows."""
    params = tuple(params)
    return _SqlTail.apply(x, queries, Wp, bp, centers_fn, len(params), on_param_grads, head_grad_out, on_stage, *params)


class Depth_Decoder_QueryTr(torch.nn.Module):
    """Drop-in for networks.Depth_Decoder_QueryTr (networks/depth_decoder_QTR.py:6-74): same constructor,
    same parameter names / shapes (depth.pth strict-loads), same forward contract.  Patch embedding, the
    4-layer transformer encoder and conv3x3 stay PyTorch (cuDNN/cuBLAS); lines 47-70 run on libsqlx."""

    dim_feedforward = 1024

    def __init__(self, in_channels, embedding_dim=128, patch_size=16, num_heads=4, query_nums=100, dim_out=256,
                 norm="linear", min_val=0.001, max_val=10):
        super().__init__()
        nn = torch.nn
        self.norm = norm
        self.embedding_convPxP = nn.Conv2d(in_channels, embedding_dim, kernel_size=patch_size, stride=patch_size,
                                           padding=0)
        self.positional_encodings = nn.Parameter(torch.rand(500, embedding_dim), requires_grad=True)
        layer = nn.TransformerEncoderLayer(embedding_dim, num_heads, dim_feedforward=self.dim_feedforward)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=4)
        self.conv3x3 = nn.Conv2d(in_channels, embedding_dim, kernel_size=3, stride=1, padding=1)
        self.full_query_layer = FullQueryLayer()
        self.bins_regressor = nn.Sequential(nn.Linear(embedding_dim * query_nums, 16 * query_nums), nn.LeakyReLU(),
                                            nn.Linear(16 * query_nums, 16 * 16), nn.LeakyReLU(),
                                            nn.Linear(16 * 16, dim_out))
        self.convert_to_prob = nn.Sequential(nn.Conv2d(query_nums, dim_out, kernel_size=1, stride=1, padding=0),
                                             nn.Softmax(dim=1))
        self.query_nums = query_nums
        self.min_val = min_val
        self.max_val = max_val

    def queries_and_features(self, x0):
        """depth_decoder_QTR.py:37-45 (PyTorch): tokens -> first Q as queries [B,Q,E]; x = conv3x3(x0)."""
        emb = self.embedding_convPxP(x0).flatten(2)
        emb = emb + self.positional_encodings[:emb.shape[2], :].T.unsqueeze(0)
        tokens = self.transformer_encoder(emb.permute(2, 0, 1))
        x = self.conv3x3(x0)
        queries = tokens[:self.query_nums, ...].permute(1, 0, 2)
        return x, queries

    def forward(self, x0):
        x, queries = self.queries_and_features(x0)
        if self.norm == "softmax":          # depth_decoder_QTR.py:55-56 returns (softmax(y), energy_maps)
            energy, summary = self.full_query_layer(x, queries)
            B, Q, E = summary.shape
            return torch.softmax(self.bins_regressor(summary.view(B, Q * E)), dim=1), energy
        conv = self.convert_to_prob[0]
        Wp = conv.weight.view(conv.out_channels, conv.in_channels)

        def centers_fn(summary):
            B, Q, E = summary.shape
            return bins_head(summary.view(B, Q * E), self.bins_regressor, self.min_val, self.max_val, self.norm)

        pred = sql_tail(x, queries.contiguous(), Wp, conv.bias, centers_fn, tuple(self.bins_regressor.parameters()))
        return {("disp", 0): pred}


class Lite_Depth_Decoder_QueryTr(Depth_Decoder_QueryTr):
    """networks/lite_depth_decoder_QTR.py: identical but dim_feedforward=512."""
    dim_feedforward = 512


def convert_depth_decoder(ref_decoder):
    """A sqlx decoder with the hyper-parameters and weights of a reference decoder instance -- networks.Depth_Decoder_QueryTr,
    networks.Lite_Depth_Decoder_QueryTr (networks/{,lite_}depth_decoder_QTR.py) or the duplicate class inside SQLdepth.py
    (:154-221): same parameter names, so its state_dict strict-loads."""
    conv = ref_decoder.embedding_convPxP
    layer = ref_decoder.transformer_encoder.layers[0]
    ff = layer.linear1.out_features
    cls = Lite_Depth_Decoder_QueryTr if ff == Lite_Depth_Decoder_QueryTr.dim_feedforward else Depth_Decoder_QueryTr
    if ff not in (512, 1024):
        raise ValueError("unexpected dim_feedforward %d" % ff)
    out = cls(in_channels=conv.in_channels, embedding_dim=conv.out_channels, patch_size=conv.kernel_size[0],
              num_heads=layer.self_attn.num_heads, query_nums=ref_decoder.query_nums,
              dim_out=ref_decoder.convert_to_prob[0].out_channels, norm=ref_decoder.norm, min_val=ref_decoder.min_val,
              max_val=ref_decoder.max_val)
    out.load_state_dict(ref_decoder.state_dict(), strict=True)
    p0 = next(ref_decoder.parameters())
    out = out.to(p0.device)
    out.train(ref_decoder.training)
    return out


def fuse_depth_decoder(model):
    """Replace every reference SQL decoder inside `model` (e.g. SQLdepth(opt).depth_decoder, SQLdepth.py:22-27, or
    Trainer.models["depth"]) by its sqlx drop-in, in place; returns the number of decoders replaced.  The inference
    wrapper's forward -- self.depth_decoder(self.encoder(x))["disp", 0], SQLdepth.py:48-50 -- then runs lines 47-70 of the
    decoder on libsqlx with the checkpoint it was loaded with."""
    names = ("Depth_Decoder_QueryTr", "Lite_Depth_Decoder_QueryTr")
    count = 0
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if type(child).__name__ in names and not isinstance(child, Depth_Decoder_QueryTr):
                setattr(parent, name, convert_depth_decoder(child))
                count += 1
    return count


# ----------------------------------------------------------------------------- tensor-core entry points (tests / profiling)
def tc_supported(E, Q, D, n):
    return bool(lib().sqlx_sql_tc_supported(E, Q, D, n))


def energy_tc(x, queries):
    """energy [B,Q,h,w] = x^T K on tcgen05 (3xTF32).  Raises when the shape is not supported."""
    require_cuda(x, queries)
    B, E, h, w = x.shape
    Q = queries.shape[1]
    energy = torch.empty(B, Q, h, w, device=x.device, dtype=torch.float32)
    check(lib().sqlx_sql_energy_tc(ptr(x), ptr(queries), B, E, Q, h * w, ptr(energy), stream_ptr()), "sqlx_sql_energy_tc")
    return energy
