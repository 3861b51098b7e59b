"""Module-level drop-ins for the reference's layers.py (same class names, constructor arguments, forward
signatures and error behaviour), each backed by libsqlx kernels with hand-written backward passes.

  SSIM                             layers.py:13-46
  transformation_from_parameters   layers.py:75-92 (+ rot_from_axisangle :111-150, get_translation_matrix :95-108)
  BackprojectDepth                 layers.py:186-215
  Project3D                        layers.py:236-258
  get_smooth_loss                  layers.py:267-280
"""
import torch

from ._lib import check, lib, ptr, require_cuda, stream_ptr
from .photometric import _Smooth, _f32c, pose_matrix


class _SSIMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, radius):
        require_cuda(x, y)
        xc, yc = _f32c(x), _f32c(y)
        if xc.shape != yc.shape or xc.dim() != 4:
            raise RuntimeError("SSIM expects two [B,C,H,W] tensors of the same shape")
        B, C, H, W = xc.shape
        out = torch.empty_like(xc)
        check(lib().sqlx_ssim_fwd(ptr(xc), ptr(yc), B, C, H, W, radius, ptr(out), stream_ptr()), "sqlx_ssim_fwd")
        ctx.save_for_backward(xc, yc)
        ctx.radius = radius
        return out

    @staticmethod
    def backward(ctx, g):
        xc, yc = ctx.saved_tensors
        B, C, H, W = xc.shape
        g = g.contiguous().float()
        gx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(yc) if ctx.needs_input_grad[1] else None
        if gx is not None or gy is not None:
            check(lib().sqlx_ssim_bwd(ptr(xc), ptr(yc), ptr(g), B, C, H, W, ctx.radius, ptr(gx), ptr(gy), stream_ptr()),
                  "sqlx_ssim_bwd")
        return gx, gy, None


class SSIM(torch.nn.Module):
    """SSIM()(x, y) -> clamp((1 - SSIM) / 2, 0, 1), 7x7 window with reflection padding (layers.py:19-26).
    `radius=1` gives the 3x3 monodepth2 variant of calc_layers.py:223-229."""

    def __init__(self, radius=3):
        super().__init__()
        self.radius = radius

    def forward(self, x, y):
        return _SSIMFn.apply(x, y, self.radius)


def transformation_from_parameters(axisangle, translation, invert=False):
    """axisangle, translation [B,1,3] -> [B,4,4] (one kernel instead of ~60 tiny ATen launches)."""
    return pose_matrix(axisangle, translation, None, invert)


class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, inv_K, H, W):
        require_cuda(depth, inv_K)
        d, iK = _f32c(depth), _f32c(inv_K)
        B = iK.shape[0]
        pts = torch.empty(B, 4, H * W, device=d.device, dtype=torch.float32)
        check(lib().sqlx_backproject_fwd(ptr(d), ptr(iK), B, H, W, ptr(pts), stream_ptr()), "sqlx_backproject_fwd")
        ctx.save_for_backward(iK)
        ctx.shape = (depth.shape, H, W)
        return pts

    @staticmethod
    def backward(ctx, g):
        (iK,) = ctx.saved_tensors
        shape, H, W = ctx.shape
        B = iK.shape[0]
        g = g.contiguous().float()
        dd = torch.empty(B, 1, H, W, device=g.device, dtype=torch.float32)
        check(lib().sqlx_backproject_bwd(ptr(g), ptr(iK), B, H, W, ptr(dd), stream_ptr()), "sqlx_backproject_bwd")
        return dd.view(shape), None, None, None


class BackprojectDepth(torch.nn.Module):
    """BackprojectDepth(batch_size, height, width)(depth [B,1,H,W], inv_K [B,4,4]) -> [B,4,H*W].
    Like the reference (layers.py:212) the batch size is fixed at construction; a mismatching depth raises."""

    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width

    def forward(self, depth, inv_K):
        if depth.numel() != self.batch_size * self.height * self.width:
            raise RuntimeError("shape '[%d, 1, -1]' is invalid for input of size %d" % (self.batch_size, depth.numel()))
        return _Backproject.apply(depth, inv_K, self.height, self.width)


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, K, T, H, W, eps):
        require_cuda(points, K, T)
        p, Kc, Tc = _f32c(points), _f32c(K), _f32c(T)
        B = Kc.shape[0]
        grid = torch.empty(B, H, W, 2, device=p.device, dtype=torch.float32)
        check(lib().sqlx_project_fwd(ptr(p), ptr(Kc), ptr(Tc), B, H, W, eps, ptr(grid), stream_ptr()), "sqlx_project_fwd")
        ctx.save_for_backward(p, Kc, Tc)
        ctx.cfg = (H, W, eps)
        return grid

    @staticmethod
    def backward(ctx, g):
        p, Kc, Tc = ctx.saved_tensors
        H, W, eps = ctx.cfg
        B = Kc.shape[0]
        g = g.contiguous().float()
        dp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        dT = torch.empty_like(Tc)
        ws = torch.empty(48 * B, device=p.device, dtype=torch.uint8)
        check(lib().sqlx_project_bwd(ptr(p), ptr(Kc), ptr(Tc), ptr(g), B, H, W, eps, ptr(dp), ptr(dT), ptr(ws), 48 * B,
                                     stream_ptr()), "sqlx_project_bwd")
        return dp, None, dT, None, None, None


class Project3D(torch.nn.Module):
    """Project3D(batch_size, height, width, eps=1e-7)(points [B,4,N], K, T [B,4,4]) -> grid [B,H,W,2]."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super().__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points, K, T):
        if points.shape[0] != self.batch_size or points.shape[-1] != self.height * self.width:
            raise RuntimeError("shape '[%d, 2, %d, %d]' is invalid for input of size %d"
                               % (self.batch_size, self.height, self.width, points.numel() // 2))
        return _Project.apply(points, K, T, self.height, self.width, self.eps)


def get_smooth_loss(disp, img):
    """Edge-aware smoothness of `disp` [B,1,H,W] against `img` [B,3,H,W] -> scalar (layers.py:267-280)."""
    if disp.shape[-2:] != img.shape[-2:]:
        raise RuntimeError("The size of tensor a (%d) must match the size of tensor b (%d)" % (disp.shape[-1], img.shape[-1]))
    sums = _Smooth.apply(disp, img)
    B = disp.shape[0]
    H, W = img.shape[-2:]
    return sums[:, 0].sum() / float(B * H * (W - 1)) + sums[:, 1].sum() / float(B * (H - 1) * W)


class _SILogFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, variance_focus):
        require_cuda(pred, target)
        p, t = _f32c(pred), _f32c(target)
        if p.dim() != 4 or t.dim() != 4 or p.shape[:2] != t.shape[:2]:
            raise RuntimeError("SILogLoss expects input [B,C,h,w] and target [B,C,H,W]")
        B, C, h, w = p.shape
        H, W = t.shape[-2:]
        m = None
        if mask is not None:
            if mask.shape != t.shape:
                raise RuntimeError("SILogLoss mask must have the target's shape")
            m = mask.contiguous()
            m = m.view(torch.uint8) if m.dtype == torch.bool else (m != 0).view(torch.uint8)
        loss = torch.empty(1, device=p.device, dtype=torch.float32)
        saved = torch.empty(4, device=p.device, dtype=torch.float32)
        nws = lib().sqlx_silog_workspace_bytes()
        ws = torch.empty(nws, device=p.device, dtype=torch.uint8)
        check(lib().sqlx_silog_fwd(ptr(p), ptr(t), ptr(m), B * C, h, w, H, W, float(variance_focus), ptr(loss), ptr(saved),
                                   ptr(ws), nws, stream_ptr()), "sqlx_silog_fwd")
        ctx.save_for_backward(p, t, saved, *([m] if m is not None else []))
        ctx.vf = float(variance_focus)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        p, t, saved, *rest = ctx.saved_tensors
        m = rest[0] if rest else None
        B, C, h, w = p.shape
        H, W = t.shape[-2:]
        d = torch.empty_like(p)
        gl = g.contiguous().float().reshape(1)
        check(lib().sqlx_silog_bwd(ptr(p), ptr(t), ptr(m), B * C, h, w, H, W, ctx.vf, ptr(saved), ptr(gl), ptr(d),
                                   stream_ptr()), "sqlx_silog_bwd")
        return d, None, None, None


class SILogLoss(torch.nn.Module):
    """Drop-in for finetune/loss.py:24-42 (the cfg-5 metric-depth fine-tuning loss): same constructor, attribute
    `name` and forward(input, target, mask=None, interpolate=True) -> scalar.  The align_corners=True resize, the
    mask gather and both reductions run in one kernel."""

    def __init__(self, variance_focus=0.15):
        super().__init__()
        self.name = "SILog"
        self.variance_focus = variance_focus

    def forward(self, input, target, mask=None, interpolate=True):
        if not interpolate and input.shape[-2:] != target.shape[-2:]:
            raise RuntimeError("The size of tensor a must match the size of tensor b")   # what the reference raises
        return _SILogFn.apply(input, target, mask, self.variance_focus)


def batch_post_process_disparity(l_disp, r_disp, r_is_flipped=False):
    """Drop-in for evaluate_depth_config.batch_post_process_disparity (evaluate_depth_config.py:51-59) on CUDA tensors:
    l_disp, r_disp [N,h,w] -> [N,h,w].  With r_is_flipped=True `r_disp` is the raw prediction for the mirrored frames
    (the reference un-flips it on the host first, :157); the kernel reads it mirrored instead."""
    require_cuda(l_disp, r_disp)
    l, r = _f32c(l_disp), _f32c(r_disp)
    if l.dim() != 3 or l.shape != r.shape:
        raise RuntimeError("batch_post_process_disparity expects two [N,h,w] tensors of the same shape")
    N, h, w = l.shape
    out = torch.empty_like(l)
    check(lib().sqlx_postprocess_disparity(ptr(l), ptr(r), N, h, w, int(bool(r_is_flipped)), ptr(out), stream_ptr()),
          "sqlx_postprocess_disparity")
    return out


def predict_disparity(encoder, depth_decoder, input_color, post_process=False):
    """The evaluation forward of evaluate_depth_config.py:126-158 kept on the device: optional flip test-time
    augmentation (one batched pass over [frames; mirrored frames]) and the Monodepth-v1 blend.  `encoder` is the
    reference's backbone, `depth_decoder` a sqlx.Depth_Decoder_QueryTr / Lite_Depth_Decoder_QueryTr.
    Returns pred_disp [N,h,w]."""
    with torch.no_grad():
        if post_process:
            input_color = torch.cat((input_color, torch.flip(input_color, [3])), 0)
        pred = depth_decoder(encoder(input_color))[("disp", 0)][:, 0]
        if post_process:
            N = pred.shape[0] // 2
            pred = batch_post_process_disparity(pred[:N], pred[N:], r_is_flipped=True)
    return pred


# ----------------------------------------------------------------------------- fine-tuning: per-sample median scaling
def _crop_mask(H, W, garg_crop, eigen_crop, dataset, device):
    """eval_mask of finetune/train_ft_SQLdepth.py:240-253 as a bool [H, W] tensor."""
    if not (garg_crop or eigen_crop):
        # the reference's loop reads eval_mask unconditionally (:254): without a crop flag it raises NameError
        raise ValueError("median scaling needs garg_crop or eigen_crop (finetune/train_ft_SQLdepth.py:240-254)")
    m = torch.zeros(H, W, dtype=torch.bool, device=device)
    if garg_crop:
        m[int(0.40810811 * H):int(0.99189189 * H), int(0.03594771 * W):int(0.96405229 * W)] = True
    elif dataset == "kitti":
        m[int(0.3324324 * H):int(0.91351351 * H), int(0.0359477 * W):int(0.96405229 * W)] = True
    else:
        m[45:471, 41:601] = True
    return m


def _masked_median(v, mask):
    """numpy.median of v[mask] per row (mean of the two middle order statistics; NaN when the selection is empty or
    holds a NaN) for v, mask [R, n] -- one sort, no boolean indexing, no device->host synchronisation."""
    k = mask.sum(dim=1)
    key = torch.where(mask, v, torch.full_like(v, float("inf")))
    key = torch.where(torch.isnan(key), torch.full_like(v, float("inf")), key)      # NaNs are flagged separately
    srt = torch.sort(key, dim=1).values
    lo = ((k - 1).clamp_min(0) // 2).unsqueeze(1)
    hi = (k // 2).clamp_max(v.shape[1] - 1).unsqueeze(1)
    med = (srt.gather(1, lo) + srt.gather(1, hi))[:, 0] * 0.5
    bad = (k == 0) | (torch.isnan(v) & mask).any(dim=1)
    return torch.where(bad, torch.full_like(med, float("nan")), med)


def median_scale_ratios(pred, depth, min_depth_eval, max_depth_eval, garg_crop=False, eigen_crop=False,
                        dataset="kitti", count=None):
    """The per-sample factors of finetune/train_ft_SQLdepth.py:236-266, computed on the tensors' device:
    ratio_i = median(depth_i[valid]) / median(pred_i[valid]) for the first `count` (default B // 2, :236) samples,
    1 where either median is NaN (:261-264), 1 for the remaining samples.  pred, depth: [B,1,H,W] -> [B] (detached)."""
    B, _, H, W = pred.shape
    count = B // 2 if count is None else count
    with torch.no_grad():
        p = pred.detach().reshape(B, H * W).float()
        d = depth.detach().reshape(B, H * W).float()
        valid = (d > min_depth_eval) & (d < max_depth_eval)
        valid = valid & _crop_mask(H, W, garg_crop, eigen_crop, dataset, pred.device).reshape(1, H * W)
        md, mp = _masked_median(d, valid), _masked_median(p, valid)
        ratio = md / mp
        ratio = torch.where(torch.isnan(md) | torch.isnan(mp), torch.ones_like(ratio), ratio)
        ratio = torch.where(torch.arange(B, device=pred.device) < count, ratio, torch.ones_like(ratio))
    return ratio


def median_scale(pred, depth, min_depth_eval, max_depth_eval, garg_crop=False, eigen_crop=False, dataset="kitti",
                 count=None):
    """Drop-in for the NumPy loop of finetune/train_ft_SQLdepth.py:236-266 (`pred[i] *= ratio`, one device->host->device
    round trip per sample in the reference): returns pred scaled per sample, differentiable wrt pred (the ratios are
    constants, as in the reference).  Sort-based torch ops on the device; the fused loss that follows is
    sqlx.SILogLoss (sqlx_silog_fwd/bwd)."""
    require_cuda(pred, depth)
    ratio = median_scale_ratios(pred, depth, min_depth_eval, max_depth_eval, garg_crop, eigen_crop, dataset, count)
    return pred * ratio.view(-1, 1, 1, 1)
