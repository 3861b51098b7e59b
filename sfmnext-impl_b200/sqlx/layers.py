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
def _crop_box(H, W, garg_crop, eigen_crop, dataset):
    """eval_mask of finetune/train_ft_SQLdepth.py:240-253 as (r0, r1, c0, c1)."""
    if not (garg_crop or eigen_crop):
        # the reference's loop reads eval_mask unconditionally (:254): without a crop flag it raises NameError
        raise ValueError("median scaling needs garg_crop or eigen_crop (finetune/train_ft_SQLdepth.py:240-254)")
    if garg_crop:
        return int(0.40810811 * H), int(0.99189189 * H), int(0.03594771 * W), int(0.96405229 * W)
    if dataset == "kitti":
        return int(0.3324324 * H), int(0.91351351 * H), int(0.0359477 * W), int(0.96405229 * W)
    return 45, 471, 41, 601


_median_ws = {}


def median_scale_ratios(pred, depth, min_depth_eval, max_depth_eval, garg_crop=False, eigen_crop=False,
                        dataset="kitti", count=None):
    """The per-sample factors of finetune/train_ft_SQLdepth.py:236-266 in ONE kernel launch on the tensors' device
    (csrc/median.cu: exact radix select of numpy.median's two middle order statistics over the valid pixels):
    ratio_i = median(depth_i[valid]) / median(pred_i[valid]) for the first `count` (default B // 2, :236) samples,
    1 where either median is NaN (:261-264), 1 for the remaining samples.  depth: [B,1,H,W]; pred: [B,1,H,W], or
    [B,1,h,w] at the network's resolution -- the kernel then reads it through the align_corners=True bilinear resize of
    :235 without materialising the resized map.  Returns [B] (detached)."""
    require_cuda(pred, depth)
    B, _, H, W = depth.shape
    h, w = pred.shape[-2:]
    count = B // 2 if count is None else count
    r0, r1, c0, c1 = _crop_box(H, W, garg_crop, eigen_crop, dataset)
    p, d = _f32c(pred), _f32c(depth)
    ratio = torch.empty(B, device=p.device, dtype=torch.float32)
    key = (p.device, B)
    ws = _median_ws.get(key)
    if ws is None:                                  # arrival counters: zeroed once, the kernel leaves them zero
        ws = _median_ws[key] = torch.zeros(lib().sqlx_median_ratio_workspace_bytes(B), device=p.device, dtype=torch.uint8)
    check(lib().sqlx_median_ratio_resized(ptr(p), h, w, ptr(d), B, H, W, int(count), float(min_depth_eval),
                                          float(max_depth_eval), r0, r1, c0, c1, ptr(ratio), ptr(ws), ws.numel(),
                                          stream_ptr()), "sqlx_median_ratio_resized")
    return ratio


def median_scale(pred, depth, min_depth_eval, max_depth_eval, garg_crop=False, eigen_crop=False, dataset="kitti",
                 count=None):
    """Drop-in for the NumPy loop of finetune/train_ft_SQLdepth.py:236-266 (`pred[i] *= ratio`, one device->host->device
    round trip per sample in the reference): returns pred scaled per sample, differentiable wrt pred (the ratios are
    constants, as in the reference).  The loss that follows is sqlx.SILogLoss (sqlx_silog_fwd/bwd)."""
    require_cuda(pred, depth)
    ratio = median_scale_ratios(pred, depth, min_depth_eval, max_depth_eval, garg_crop, eigen_crop, dataset, count)
    return pred * ratio.view(-1, 1, 1, 1)


def finetune_loss(pred, depth, min_depth, min_depth_eval, max_depth_eval, garg_crop=False, eigen_crop=False,
                  dataset="kitti", variance_focus=0.15, mask=None):
    """The supervised fine-tuning step of finetune/train_ft_SQLdepth.py:232-274 after the model forward, on the device:
    pred [B,1,h,w] (the decoder output) is NOT resized: the median-ratio kernel and the SILog kernel both read it through
    the align_corners=True bilinear resize to the ground truth's shape (:235), and the per-sample scaling commutes with
    that (linear) resize.  mask defaults to depth > min_depth (:271).  Returns the SILog loss (:274), differentiable wrt
    pred with the ratios as constants, exactly like `pred[i] *= ratio` in the reference."""
    require_cuda(pred, depth)
    ratio = median_scale_ratios(pred.detach(), depth, min_depth_eval, max_depth_eval, garg_crop, eigen_crop, dataset)
    if mask is None:
        mask = depth > min_depth
    return _SILogFn.apply(pred * ratio.view(-1, 1, 1, 1), depth, mask, variance_focus)


# ----------------------------------------------------------------------------- indoor rectification warp (N4)
def euler2mat(angle):
    """layers.py:422-457: rotation matrix [B,3,3] = X(x) . Y(y) . Z(z) of Euler angles [B,3] (radians)."""
    x, y, z = angle[:, 0], angle[:, 1], angle[:, 2]
    zeros, ones = torch.zeros_like(z), torch.ones_like(z)
    cz, sz, cy, sy, cx, sx = torch.cos(z), torch.sin(z), torch.cos(y), torch.sin(y), torch.cos(x), torch.sin(x)
    B = angle.shape[0]
    zmat = torch.stack([cz, -sz, zeros, sz, cz, zeros, zeros, zeros, ones], dim=1).reshape(B, 3, 3)
    ymat = torch.stack([cy, zeros, sy, zeros, ones, zeros, -sy, zeros, cy], dim=1).reshape(B, 3, 3)
    xmat = torch.stack([ones, zeros, zeros, zeros, cx, -sx, zeros, sx, cx], dim=1).reshape(B, 3, 3)
    return xmat @ ymat @ zmat


_rw_ws = {}


class _RotationWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, P, K3):
        require_cuda(img, P, K3)
        im, Pc, Kc = _f32c(img), _f32c(P), _f32c(K3)
        B, C, H, W = im.shape
        if C != 3:
            raise RuntimeError("inverse_rotation_warp expects [B,3,H,W] images")
        out = torch.empty_like(im)
        check(lib().sqlx_rotation_warp_fwd(ptr(im), ptr(Pc), ptr(Kc), B, H, W, ptr(out), stream_ptr()), "sqlx_rotation_warp_fwd")
        ctx.save_for_backward(im, Pc, Kc)
        return out

    @staticmethod
    def backward(ctx, g):
        im, Pc, Kc = ctx.saved_tensors
        B, _, H, W = im.shape
        key = (im.device, B)
        ws = _rw_ws.get(key)
        if ws is None:
            ws = _rw_ws[key] = torch.zeros(lib().sqlx_rotation_warp_workspace_bytes(B), device=im.device, dtype=torch.uint8)
        dP = torch.empty(B, 3, 3, device=im.device, dtype=torch.float32)
        check(lib().sqlx_rotation_warp_bwd(ptr(im), ptr(Pc), ptr(Kc), ptr(_f32c(g)), B, H, W, ptr(dP), ptr(ws), ws.numel(),
                                           stream_ptr()), "sqlx_rotation_warp_bwd")
        return None, dP, None


def inverse_rotation_warp(img, rot, intrinsics, padding_mode="zeros"):
    """Drop-in for layers.inverse_rotation_warp (layers.py:460-479): img [B,3,H,W], rot [B,3] Euler angles, intrinsics
    [B,3,3] -> img resampled under the pure rotation, differentiable wrt `rot` (what trainer_indoor.rectify_imgs needs,
    trainer_indoor.py:877-920; the frames themselves carry no gradient there).  The per-pixel work -- back-projection of
    the unit-depth plane, projection, zero-padded bilinear gather, and the reduction of the gradient to dL/d(K R) -- is one
    kernel each way; K . euler2mat(rot) is three tiny torch ops."""
    if padding_mode != "zeros":
        raise NotImplementedError("inverse_rotation_warp: the reference only ever passes padding_mode='zeros'")
    P = torch.matmul(intrinsics.float(), euler2mat(rot.float()))
    return _RotationWarp.apply(img, P, intrinsics)
