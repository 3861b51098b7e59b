"""Photometric reprojection loss on libsqlx kernels (autograd Functions + the functional pipeline).

Mirrors Trainer.generate_images_pred + Trainer.compute_losses of the reference
(trainer.py:386-439, 455-549): same inputs, same loss, same `identity_selection` masks, gradients
to the per-scale network outputs and to the pose parameters.  Everything device-side is a call
into libsqlx.so through ctypes; there is no PyTorch fallback.
"""
import ctypes

import torch

from . import _lib
from ._lib import PhotoDesc, check, lib, ptr, require_cuda, stream_ptr


def _f32c(t):
    return t.detach().contiguous().float() if t is not None else None


def make_desc(B, H, W, h, w, S, *, ssim_radius=3, automask=True, avg=False, no_ssim=False,
              w_ssim=0.85, w_l1=0.15, noise_scale=1e-5, eps=1e-7):
    flags = (_lib.SQLX_AUTOMASK if automask else 0) | (_lib.SQLX_AVG_REPROJ if avg else 0) | \
            (_lib.SQLX_NO_SSIM if no_ssim else 0)
    return PhotoDesc(B, H, W, h, w, S, ssim_radius, flags, w_ssim, w_l1, noise_scale, eps)


def _src_array(sources):
    arr = (ctypes.c_void_p * _lib.MAX_SOURCES)()
    for i, s in enumerate(sources):
        arr[i] = s.data_ptr()
    return arr


def pack_rgba(image, out=None):
    """[B,3,H,W] planar frame -> [B,H,W,4] pixel-interleaved frame (r,g,b,0): the layout the fused kernels gather
    source frames from (one 128-bit load per bilinear tap).  Packed once per step, shared by all loss scales.
    out: optional preallocated [B,H,W,4] fp32 result buffer."""
    require_cuda(image)
    img = _f32c(image)
    B, C, H, W = img.shape
    assert C == 3
    if out is None:
        out = torch.empty(B, H, W, 4, device=img.device, dtype=torch.float32)
    assert out.shape == (B, H, W, 4) and out.dtype == torch.float32 and out.is_contiguous()
    check(lib().sqlx_pack_rgba(ptr(img), B, H, W, ptr(out), stream_ptr()), "sqlx_pack_rgba")
    return out


# ----------------------------------------------------------------------------- small ops
class _DepthStats(torch.autograd.Function):
    """[B,1,h,w] -> [B,2] = (mean, mean of reciprocal) of the map bilinearly upsampled to HxW."""

    @staticmethod
    def forward(ctx, depth_lr, H, W):
        require_cuda(depth_lr)
        d = _f32c(depth_lr)
        B, _, h, w = d.shape
        stats = torch.empty(B, 2, device=d.device, dtype=torch.float32)
        nbytes = lib().sqlx_depth_stats_workspace_bytes(B, H, W)
        ws = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
        check(lib().sqlx_depth_stats_fwd(ptr(d), B, h, w, H, W, ptr(stats), ptr(ws), nbytes, stream_ptr()),
              "sqlx_depth_stats_fwd")
        ctx.save_for_backward(d)
        ctx.hw = (H, W)
        return stats

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        B, _, h, w = d.shape
        H, W = ctx.hw
        out = torch.zeros_like(d)
        g = g.contiguous().float()
        check(lib().sqlx_depth_stats_bwd(ptr(d), B, h, w, H, W, ptr(g), ptr(out), stream_ptr()),
              "sqlx_depth_stats_bwd")
        return out, None, None


def depth_stats(depth_lr, H, W):
    return _DepthStats.apply(depth_lr, H, W)


class _PoseMatrix(torch.autograd.Function):
    """transformation_from_parameters(axisangle, translation*scale, invert) (layers.py:75-150)."""

    @staticmethod
    def forward(ctx, axisangle, translation, scale, invert):
        require_cuda(axisangle, translation, scale)
        aa = _f32c(axisangle).reshape(-1, 3)
        tr = _f32c(translation).reshape(-1, 3)
        sc = _f32c(scale).reshape(-1) if scale is not None else None
        B = aa.shape[0]
        T = torch.empty(B, 4, 4, device=aa.device, dtype=torch.float32)
        check(lib().sqlx_pose_fwd(ptr(aa), ptr(tr), ptr(sc), B, int(bool(invert)), ptr(T), stream_ptr()), "sqlx_pose_fwd")
        ctx.save_for_backward(aa, tr, sc)
        ctx.invert = int(bool(invert))
        ctx.shapes = (axisangle.shape, translation.shape, None if scale is None else scale.shape)
        return T

    @staticmethod
    def backward(ctx, dT):
        aa, tr, sc = ctx.saved_tensors
        B = aa.shape[0]
        dT = dT.contiguous().float()
        daa = torch.empty_like(aa)
        dtr = torch.empty_like(tr)
        dsc = torch.empty(B, device=aa.device, dtype=torch.float32) if sc is not None else None
        check(lib().sqlx_pose_bwd(ptr(aa), ptr(tr), ptr(sc), B, ctx.invert, ptr(dT), ptr(daa), ptr(dtr), ptr(dsc),
                                  stream_ptr()), "sqlx_pose_bwd")
        s_aa, s_tr, s_sc = ctx.shapes
        return daa.reshape(s_aa), dtr.reshape(s_tr), (dsc.reshape(s_sc) if dsc is not None else None), None


def pose_matrix(axisangle, translation, scale=None, invert=False):
    """axisangle, translation: [B,1,3] (or [B,3]); scale: [B] or None -> [B,4,4]."""
    return _PoseMatrix.apply(axisangle, translation, scale, invert)


def reprojection_loss(pred, target, *, no_ssim=False, ssim_radius=3, w_ssim=0.85, w_l1=0.15):
    """Forward-only compute_reprojection_loss (trainer.py:441-453) -> [B,1,H,W]; used for the identity losses."""
    require_cuda(pred, target)
    p, t = _f32c(pred), _f32c(target)
    B, C, H, W = p.shape
    assert C == 3 and t.shape == p.shape
    out = torch.empty(B, 1, H, W, device=p.device, dtype=torch.float32)
    check(lib().sqlx_reprojection_loss_fwd(ptr(p), ptr(t), B, H, W, ssim_radius, w_ssim, w_l1, int(no_ssim), ptr(out),
                                           stream_ptr()), "sqlx_reprojection_loss_fwd")
    return out


class _Smooth(torch.autograd.Function):
    """disp [B,1,h,w] (upsampled to the colour resolution when different), color [B,3,Hc,Wc] -> sums [B,3]."""

    @staticmethod
    def forward(ctx, disp, color):
        require_cuda(disp, color)
        d, c = _f32c(disp), _f32c(color)
        B, _, h, w = d.shape
        Hc, Wc = c.shape[-2:]
        sums = torch.empty(B, 3, device=d.device, dtype=torch.float32)
        nbytes = lib().sqlx_smooth_workspace_bytes(B, Hc, Wc)
        ws = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
        check(lib().sqlx_smooth_fwd(ptr(d), ptr(c), B, h, w, Hc, Wc, ptr(sums), ptr(ws), nbytes, stream_ptr()),
              "sqlx_smooth_fwd")
        ctx.save_for_backward(d, c)
        return sums

    @staticmethod
    def backward(ctx, g):
        d, c = ctx.saved_tensors
        B, _, h, w = d.shape
        Hc, Wc = c.shape[-2:]
        out = torch.zeros_like(d)
        g = g.contiguous().float()
        check(lib().sqlx_smooth_bwd(ptr(d), ptr(c), B, h, w, Hc, Wc, ptr(g), ptr(out), stream_ptr()), "sqlx_smooth_bwd")
        return out, None


def smooth_loss_normalised(disp, color):
    """get_smooth_loss(disp / (mean(disp)+1e-7), color) with the upsample of trainer.py:533-534 folded in."""
    sums = _Smooth.apply(disp, color)
    B = disp.shape[0]
    Hc, Wc = color.shape[-2:]
    mean = sums[:, 2] / float(Hc * Wc)
    inv = 1.0 / (mean + 1e-7)
    return (sums[:, 0] * inv).sum() / float(B * Hc * (Wc - 1)) + (sums[:, 1] * inv).sum() / float(B * (Hc - 1) * Wc)


# ----------------------------------------------------------------------------- fused photometric loss
class _PhotoLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth_lr, T, target, K, inv_K, identity, noise, cfg, *sources):
        require_cuda(depth_lr, T, target, K, inv_K, identity, noise, *sources)
        d = _f32c(depth_lr)
        Tm = _f32c(T)
        B, _, h, w = d.shape
        H, W = target.shape[-2:]
        S = len(sources)
        desc = make_desc(B, H, W, h, w, S, **cfg)
        srcs = [pack_rgba(s) for s in sources]
        tgt, Kc, iKc = _f32c(target), _f32c(K), _f32c(inv_K)
        ident, nz = _f32c(identity), _f32c(noise)
        nbytes = lib().sqlx_photo_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
        coef = torch.empty(max(1, lib().sqlx_photo_coef_bytes(ctypes.byref(desc))), device=d.device, dtype=torch.uint8)
        loss_sum = torch.empty(1, device=d.device, dtype=torch.float32)
        argmin = torch.empty(B, H, W, device=d.device, dtype=torch.uint8)
        check(lib().sqlx_photo_fwd(ctypes.byref(desc), ptr(d), ptr(tgt), _src_array(srcs), ptr(Kc), ptr(iKc), ptr(Tm),
                                   ptr(ident), ptr(nz), ptr(loss_sum), ptr(argmin), ptr(coef), ptr(ws), nbytes,
                                   stream_ptr()), "sqlx_photo_fwd")
        ctx.save_for_backward(d, Tm, tgt, Kc, iKc, argmin, coef, *srcs)
        ctx.desc = desc
        ctx.mark_non_differentiable(argmin)
        return loss_sum, argmin

    @staticmethod
    def backward(ctx, g_loss, _g_argmin):
        d, Tm, tgt, Kc, iKc, argmin, coef, *srcs = ctx.saved_tensors
        desc = ctx.desc
        nbytes = lib().sqlx_photo_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
        d_depth = torch.zeros_like(d)
        d_T = torch.empty_like(Tm)
        g = g_loss.contiguous().float()
        check(lib().sqlx_photo_bwd(ctypes.byref(desc), ptr(d), ptr(tgt), _src_array(srcs), ptr(Kc), ptr(iKc), ptr(Tm),
                                   ptr(argmin), ptr(coef), ptr(g), 1.0, ptr(d_depth), ptr(d_T), ptr(ws), nbytes, stream_ptr()),
              "sqlx_photo_bwd")
        return (d_depth, d_T, None, None, None, None, None, None) + (None,) * len(srcs)


class _PhotoOccLoss(torch.autograd.Function):
    """Indoor variant of _PhotoLoss (sqlx_photo_occ_fwd/bwd): depth-consistency weighted reprojection losses and
    the regularisation sum; differentiable wrt the depth map, the camera transforms and the source frames' depths."""

    @staticmethod
    def forward(ctx, depth_lr, T, target, K, inv_K, identity, noise, cfg, sources, *ref_depths):
        require_cuda(depth_lr, T, target, K, inv_K, identity, noise, *sources, *ref_depths)
        d = _f32c(depth_lr)
        Tm = _f32c(T)
        B, _, h, w = d.shape
        H, W = target.shape[-2:]
        S = len(sources)
        assert len(ref_depths) == S
        desc = make_desc(B, H, W, h, w, S, **cfg)
        srcs = [pack_rgba(s) for s in sources]
        refs = [_f32c(r) for r in ref_depths]
        for r in refs:
            assert tuple(r.shape) == (B, 1, H, W), "depth_ref must be [B,1,H,W] (trainer_indoor.py:377)"
        tgt, Kc, iKc = _f32c(target), _f32c(K), _f32c(inv_K)
        ident, nz = _f32c(identity), _f32c(noise)
        nbytes = lib().sqlx_photo_occ_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
        coef = torch.empty(max(1, lib().sqlx_photo_coef_bytes(ctypes.byref(desc))), device=d.device, dtype=torch.uint8)
        sums = torch.empty(2, device=d.device, dtype=torch.float32)
        argmin = torch.empty(B, H, W, device=d.device, dtype=torch.uint8)
        check(lib().sqlx_photo_occ_fwd(ctypes.byref(desc), ptr(d), ptr(tgt), _src_array(srcs), _src_array(refs), ptr(Kc),
                                       ptr(iKc), ptr(Tm), ptr(ident), ptr(nz), ptr(sums), ptr(argmin), ptr(coef), ptr(ws),
                                       nbytes, stream_ptr()), "sqlx_photo_occ_fwd")
        ctx.save_for_backward(d, Tm, tgt, Kc, iKc, argmin, coef, *srcs, *refs)
        ctx.desc = desc
        ctx.S = S
        ctx.ref_shapes = [r.shape for r in ref_depths]
        ctx.mark_non_differentiable(argmin)
        return sums, argmin

    @staticmethod
    def backward(ctx, g_sums, _g_argmin):
        d, Tm, tgt, Kc, iKc, argmin, coef, *rest = ctx.saved_tensors
        S = ctx.S
        srcs, refs = rest[:S], rest[S:]
        desc = ctx.desc
        nbytes = lib().sqlx_photo_occ_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
        d_depth = torch.zeros_like(d)
        d_T = torch.empty_like(Tm)
        d_refs = [torch.zeros_like(r) for r in refs]
        g = g_sums.contiguous().float()
        check(lib().sqlx_photo_occ_bwd(ctypes.byref(desc), ptr(d), ptr(tgt), _src_array(srcs), _src_array(refs), ptr(Kc),
                                       ptr(iKc), ptr(Tm), ptr(argmin), ptr(coef), ptr(g), 1.0, ptr(d_depth), ptr(d_T),
                                       _src_array(d_refs), ptr(ws), nbytes, stream_ptr()), "sqlx_photo_occ_bwd")
        return (d_depth, d_T, None, None, None, None, None, None, None) + \
            tuple(g.reshape(sh) for g, sh in zip(d_refs, ctx.ref_shapes))


def indoor_losses(disp, target, sources, ref_depths, K, inv_K, poses, noise=None, *, height, width,
                  disparity_smoothness=1e-3, reg_wt=0.01, rescale_translation=True, no_ssim=False,
                  avg_reprojection=False, disable_automasking=False, ssim_radius=3):
    """Indoor loss variant: trainer_indoor.py generate_images_pred (:512-599) +
    compute_losses_with_occ (:615-719) for --use_improved_mini_reproj_loss at the single loss scale the SQL decoder
    emits (scales = [0]; every indoor arg file of the reference uses it).

    disp [B,1,h,w] = outputs[("disp",0)]; ref_depths: list of S [B,1,H,W] = outputs[("depth_ref",f,0)] (network
    depth of every source frame, differentiable); the other arguments as for photometric_losses.  Returns
    {"loss", "loss/0", ("argmin",0)}; gradients reach disp, the pose parameters and ref_depths."""
    H, W = height, width
    S = len(sources)
    B = target.shape[0]
    require_cuda(disp, target, K, inv_K, *sources, *ref_depths)
    automask = not disable_automasking
    cfg = dict(ssim_radius=ssim_radius, automask=automask, avg=avg_reprojection, no_ssim=no_ssim)
    identity = identity_losses(target, sources, no_ssim=no_ssim, ssim_radius=ssim_radius) if automask else None
    if automask and noise is None:
        noise = torch.randn(B, 1 if avg_reprojection else S, H, W, device=target.device)
    rescale = bool(rescale_translation) and any("T" not in p for p in poses)
    stats = depth_stats(disp, H, W) if rescale else None          # mean(1/depth) of the upsampled map (:543-544)
    Ts = []
    for pose in poses:
        if "T" in pose:
            Ts.append(pose["T"].float())
        else:
            Ts.append(pose_matrix(pose["axisangle"][:, 0], pose["translation"][:, 0],
                                  stats[:, 1] if rescale else None, pose["invert"]))
    T = torch.stack(Ts, 1)                                        # [B,S,4,4]
    sums, argmin = _PhotoOccLoss.apply(disp, T, target, K, inv_K, identity, noise, cfg, tuple(sources), *ref_depths)
    n = float(B * H * W)
    loss = sums[0] / n + reg_wt * sums[1] / (n * S)               # :698-699 (mean over sources, then over pixels)
    h, w = disp.shape[-2:]
    # the colour is always resized to the depth map's shape (`shape[-2:] != [H, W]` is always True, :704-707)
    color = torch.nn.functional.interpolate(target, [h, w], mode="bilinear", align_corners=False)
    loss = loss + disparity_smoothness * smooth_loss_normalised(disp, color)     # :701-711
    return {"loss": loss, "loss/0": loss, ("argmin", 0): argmin}


def warp(depth_lr, source, K, inv_K, T, H, W, *, want_depth=True, want_sample=True, want_color=True, eps=1e-7):
    """Materialise outputs[("depth",0,s)], ("sample",f,s), ("color",f,s) for logging (no autograd)."""
    require_cuda(depth_lr, source, K, inv_K, T)
    d, src, Kc, iKc, Tm = _f32c(depth_lr), _f32c(source), _f32c(K), _f32c(inv_K), _f32c(T)
    B, _, h, w = d.shape
    dev = d.device
    depth_up = torch.empty(B, 1, H, W, device=dev) if want_depth else None
    sample = torch.empty(B, H, W, 2, device=dev) if want_sample else None
    color = torch.empty(B, 3, H, W, device=dev) if want_color else None
    check(lib().sqlx_warp_fwd(ptr(d), ptr(src), ptr(Kc), ptr(iKc), ptr(Tm), 16, B, h, w, H, W, eps,
                              ptr(depth_up), ptr(sample), ptr(color), stream_ptr()), "sqlx_warp_fwd")
    return depth_up, sample, color


class _ScaleLoss(torch.autograd.Function):
    """One loss scale as one library call forward and one backward (csrc/scale_loss.cu)."""

    @staticmethod
    def forward(ctx, depth_lr, meta, *pose_tensors):
        (target, color_s, K, inv_K, identity, noise, sources, pose_spec, cfg, smooth_weight, rescale) = meta
        d = _f32c(depth_lr)
        B, _, h, w = d.shape
        H, W = target.shape[-2:]
        S = len(sources)
        Hc, Wc = color_s.shape[-2:]
        desc = _lib.ScaleDesc(make_desc(B, H, W, h, w, S, **cfg), Hc, Wc, float(smooth_weight), int(bool(rescale)))
        keep = []                                  # contiguous fp32 views that must outlive the call
        pin = _lib.PoseInputs()
        mask = 0
        it = iter(pose_tensors)
        for i, spec in enumerate(pose_spec):
            if spec[0] == "fixed":
                Tm = _f32c(spec[1]); keep.append(Tm)
                pin.fixed_T[i] = Tm.data_ptr()
            else:
                aa = _f32c(next(it)).reshape(B, 3); tr = _f32c(next(it)).reshape(B, 3)
                keep += [aa, tr]
                pin.axisangle[i] = aa.data_ptr(); pin.translation[i] = tr.data_ptr()
                if spec[1]:
                    mask |= 1 << i
        pin.invert_mask = mask
        dev = d.device
        saved = torch.empty(lib().sqlx_scale_saved_bytes(ctypes.byref(desc)), device=dev, dtype=torch.uint8)
        nws = lib().sqlx_scale_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(nws, device=dev, dtype=torch.uint8)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        argmin = torch.empty(B, H, W, device=dev, dtype=torch.uint8)
        srcs = list(sources)                       # [B,H,W,4] pixel-interleaved copies (pack_rgba)
        tgt, col, Kc, iKc, ident, nz = (_f32c(t) for t in (target, color_s, K, inv_K, identity, noise))
        check(lib().sqlx_scale_loss_fwd(ctypes.byref(desc), ptr(d), ptr(tgt), _src_array(srcs), ptr(col), ptr(Kc), ptr(iKc),
                                        ctypes.byref(pin), ptr(ident), ptr(nz), ptr(loss), ptr(argmin), ptr(saved),
                                        saved.numel(), ptr(ws), nws, stream_ptr()), "sqlx_scale_loss_fwd")
        ctx.save_for_backward(d, tgt, col, Kc, iKc, argmin, saved, *srcs, *keep)
        ctx.state = (desc, pose_spec, mask, S, len(keep), [t.shape for t in pose_tensors])
        ctx.mark_non_differentiable(argmin)
        return loss.reshape(()), argmin

    @staticmethod
    def backward(ctx, g_loss, _g_argmin):
        desc, pose_spec, mask, S, nkeep, shapes = ctx.state
        d, tgt, col, Kc, iKc, argmin, saved, *rest = ctx.saved_tensors
        srcs, keep = rest[:S], rest[S:]
        B = d.shape[0]
        dev = d.device
        pin = _lib.PoseInputs()
        pin.invert_mask = mask
        d_aa = (ctypes.c_void_p * _lib.MAX_SOURCES)()
        d_tr = (ctypes.c_void_p * _lib.MAX_SOURCES)()
        grads = []
        it = iter(keep)
        for i, spec in enumerate(pose_spec):
            if spec[0] == "fixed":
                pin.fixed_T[i] = next(it).data_ptr()
            else:
                aa, tr = next(it), next(it)
                pin.axisangle[i] = aa.data_ptr(); pin.translation[i] = tr.data_ptr()
                ga = torch.empty(B, 3, device=dev, dtype=torch.float32)
                gt = torch.empty(B, 3, device=dev, dtype=torch.float32)
                d_aa[i] = ga.data_ptr(); d_tr[i] = gt.data_ptr()
                grads += [ga, gt]
        nws = lib().sqlx_scale_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(nws, device=dev, dtype=torch.uint8)
        d_depth = torch.empty_like(d)
        g = g_loss.contiguous().float().reshape(1)
        check(lib().sqlx_scale_loss_bwd(ctypes.byref(desc), ptr(d), ptr(tgt), _src_array(srcs), ptr(col), ptr(Kc), ptr(iKc),
                                        ctypes.byref(pin), ptr(argmin), ptr(g), ptr(saved), ptr(d_depth), d_aa, d_tr,
                                        ptr(ws), nws, stream_ptr()), "sqlx_scale_loss_bwd")
        return (d_depth, None) + tuple(gr.reshape(sh) for gr, sh in zip(grads, shapes))


def _ptr_array(tensors, n):
    arr = (ctypes.c_void_p * n)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


class _MultiScaleLoss(torch.autograd.Function):
    """Every loss scale in one library call forward and one backward (csrc/multiscale.cu): 3 + 4 kernel launches per
    step whatever the number of scales (the fused photometric forward / backward kernels take all scales in one launch),
    one autograd node, pose gradients summed over scales inside the library."""

    @staticmethod
    def forward(ctx, meta, *tensors):
        ctx.set_materialize_grads(False)       # no zero "gradients" for the per-scale losses / arg-min outputs
        (target, colors, K, inv_K, identity, noises, sources, pose_spec, cfg, smooth_weights, rescale) = meta
        ns = len(colors)
        disps = [_f32c(t) for t in tensors[:ns]]
        pose_tensors = tensors[ns:]
        B = disps[0].shape[0]
        H, W = target.shape[-2:]
        S = len(sources)
        desc = _lib.MsDesc()
        desc.photo = make_desc(B, H, W, disps[0].shape[-2], disps[0].shape[-1], S, **cfg)
        desc.num_scales = ns
        cols = [_f32c(c) for c in colors]
        for i in range(ns):
            desc.h[i], desc.w[i] = disps[i].shape[-2:]
            desc.Hc[i], desc.Wc[i] = cols[i].shape[-2:]
            desc.smooth_weight[i] = float(smooth_weights[i])
        desc.rescale_translation = int(bool(rescale))
        keep = []                                  # contiguous fp32 views that must outlive the call
        pin = _lib.PoseInputs()
        mask = 0
        it = iter(pose_tensors)
        for i, spec in enumerate(pose_spec):
            if spec[0] == "fixed":
                Tm = _f32c(spec[1]); keep.append(Tm)
                pin.fixed_T[i] = Tm.data_ptr()
            else:
                aa = _f32c(next(it)).reshape(B, 3); tr = _f32c(next(it)).reshape(B, 3)
                keep += [aa, tr]
                pin.axisangle[i] = aa.data_ptr(); pin.translation[i] = tr.data_ptr()
                if spec[1]:
                    mask |= 1 << i
        pin.invert_mask = mask
        dev = disps[0].device
        saved = torch.empty(lib().sqlx_ms_saved_bytes(ctypes.byref(desc)), device=dev, dtype=torch.uint8)
        nws = lib().sqlx_ms_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(nws, device=dev, dtype=torch.uint8)
        losses = torch.empty(1 + ns, device=dev, dtype=torch.float32)
        argmins = [torch.empty(B, H, W, device=dev, dtype=torch.uint8) for _ in range(ns)]
        srcs = list(sources)                       # [B,H,W,4] pixel-interleaved copies (pack_rgba)
        tgt, Kc, iKc, ident = (_f32c(t) for t in (target, K, inv_K, identity))
        nzs = [_f32c(n) for n in noises]
        M = _lib.MAX_SCALES
        check(lib().sqlx_ms_loss_fwd(ctypes.byref(desc), _ptr_array(disps, M), ptr(tgt), _src_array(srcs),
                                     _ptr_array(cols, M), ptr(Kc), ptr(iKc), ctypes.byref(pin), ptr(ident),
                                     _ptr_array(nzs, M), ptr(losses), _ptr_array(argmins, M), ptr(saved), saved.numel(),
                                     ptr(ws), nws, stream_ptr()), "sqlx_ms_loss_fwd")
        ctx.save_for_backward(tgt, Kc, iKc, saved, *disps, *cols, *argmins, *srcs, *keep)
        ctx.state = (desc, pose_spec, mask, S, ns, [t.shape for t in pose_tensors], [t.shape for t in tensors[:ns]])
        per_scale = losses[1:].detach()
        ctx.mark_non_differentiable(per_scale, *argmins)
        return (losses[0], per_scale) + tuple(argmins)

    @staticmethod
    def backward(ctx, g_loss, _g_per_scale, *_g_argmins):
        desc, pose_spec, mask, S, ns, pose_shapes, disp_shapes = ctx.state
        if g_loss is None:
            return (None,) * (1 + ns + len(pose_shapes))
        tgt, Kc, iKc, saved, *rest = ctx.saved_tensors
        disps, cols, argmins = rest[:ns], rest[ns:2 * ns], rest[2 * ns:3 * ns]
        srcs, keep = rest[3 * ns:3 * ns + S], rest[3 * ns + S:]
        B = disps[0].shape[0]
        dev = disps[0].device
        pin = _lib.PoseInputs()
        pin.invert_mask = mask
        d_aa = (ctypes.c_void_p * _lib.MAX_SOURCES)()
        d_tr = (ctypes.c_void_p * _lib.MAX_SOURCES)()
        grads = []
        it = iter(keep)
        for i, spec in enumerate(pose_spec):
            if spec[0] == "fixed":
                pin.fixed_T[i] = next(it).data_ptr()
            else:
                aa, tr = next(it), next(it)
                pin.axisangle[i] = aa.data_ptr(); pin.translation[i] = tr.data_ptr()
                ga = torch.empty(B, 3, device=dev, dtype=torch.float32)
                gt = torch.empty(B, 3, device=dev, dtype=torch.float32)
                d_aa[i] = ga.data_ptr(); d_tr[i] = gt.data_ptr()
                grads += [ga, gt]
        nws = lib().sqlx_ms_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(nws, device=dev, dtype=torch.uint8)
        d_disps = [torch.empty_like(d) for d in disps]
        g = g_loss.contiguous().float().reshape(1)
        M = _lib.MAX_SCALES
        check(lib().sqlx_ms_loss_bwd(ctypes.byref(desc), _ptr_array(disps, M), ptr(tgt), _src_array(srcs),
                                     _ptr_array(cols, M), ptr(Kc), ptr(iKc), ctypes.byref(pin), _ptr_array(argmins, M),
                                     ptr(g), ptr(saved), _ptr_array(d_disps, M), d_aa, d_tr, ptr(ws), nws, stream_ptr()),
              "sqlx_ms_loss_bwd")
        return (None,) + tuple(dd.reshape(sh) for dd, sh in zip(d_disps, disp_shapes)) + \
            tuple(gr.reshape(sh) for gr, sh in zip(grads, pose_shapes))


def identity_losses(target, sources, *, no_ssim=False, ssim_radius=3, w_ssim=0.85, w_l1=0.15, out=None):
    """[B,S,H,W]: compute_reprojection_loss(source_f, target) for every source (trainer.py:480-493), no torch.cat.
    out: optional preallocated [B,S,H,W] fp32 result buffer."""
    require_cuda(target, *sources)
    tgt = _f32c(target)
    srcs = [_f32c(x) for x in sources]
    B, _, H, W = tgt.shape
    S = len(srcs)
    if out is None:
        out = torch.empty(B, S, H, W, device=tgt.device, dtype=torch.float32)
    assert out.shape == (B, S, H, W) and out.dtype == torch.float32 and out.is_contiguous()
    check(lib().sqlx_identity_losses_fwd(ptr(tgt), _src_array(srcs), S, B, H, W, ssim_radius, w_ssim, w_l1, int(no_ssim),
                                         ptr(out), stream_ptr()), "sqlx_identity_losses_fwd")
    return out


def photometric_losses(disps, target_pyr, sources, K, inv_K, poses, noises, *, height, width, scales=(0,),
                       disparity_smoothness=1e-3, rescale_translation=True, no_ssim=False, avg_reprojection=False,
                       disable_automasking=False, ssim_radius=3, materialize=False, identity=None,
                       packed_sources=None, per_scale_calls=False):
    """generate_images_pred + compute_losses (trainer.py:386-549): ONE fused library call for all loss scales
    (`per_scale_calls=True`: one call per scale through sqlx_scale_loss_fwd/bwd instead, same results).

    disps {s: [B,1,h_s,w_s]} network outputs (these ARE depth, trainer.py:399-402); target_pyr {s: [B,3,H_s,W_s]};
    sources: list of S [B,3,H,W]; poses: per source {"T": [B,4,4]} or {"axisangle", "translation" [B,1,1,3], "invert"};
    noises {s: [B,S,H,W]} standard-normal tie-break noise, or None / missing: drawn on the device (the reference draws
    it on the CPU generator and copies it, trainer.py:516-517).  Returns loss, loss/<s>, identity_selection/<s>
    (only materialised on request or when automasking... see `materialize`), and with materialize=True the
    depth / sample / color tensors Trainer.log reads.
    """
    H, W = height, width
    S = len(sources)
    target = target_pyr[0]
    B = target.shape[0]
    automask = not disable_automasking
    cfg = dict(ssim_radius=ssim_radius, automask=automask, avg=avg_reprojection, no_ssim=no_ssim)
    out = {}
    require_cuda(target, K, inv_K, *sources)
    if automask and identity is None:
        identity = identity_losses(target, sources, no_ssim=no_ssim, ssim_radius=ssim_radius)
    pose_spec, pose_tensors = [], []
    for pose in poses:
        if "T" in pose:
            pose_spec.append(("fixed", pose["T"]))
        else:
            pose_spec.append(("net", bool(pose["invert"])))
            pose_tensors += [pose["axisangle"], pose["translation"]]
    rescale = bool(rescale_translation) and any(sp[0] == "net" for sp in pose_spec)
    n_ident = 0 if not automask else (1 if avg_reprojection else S)
    # once per step: every scale, forward and backward, gathers from these (`identity` and `packed_sources` depend on
    # the input frames only: a caller may compute them early, e.g. on a side stream while the decoder runs)
    packed = list(packed_sources) if packed_sources is not None else [pack_rgba(src) for src in sources]
    scales = tuple(scales)
    if len(scales) > _lib.MAX_SCALES:
        raise ValueError("at most %d loss scales" % _lib.MAX_SCALES)
    noise_list = []
    for s in scales:
        noise = None
        if automask:
            noise = noises.get(s) if noises is not None else None
            if noise is None:
                noise = torch.randn(B, 1 if avg_reprojection else S, H, W, device=target.device)
        noise_list.append(noise)
    if per_scale_calls:
        total = 0
        for s, noise in zip(scales, noise_list):
            meta = (target, target_pyr[s], K, inv_K, identity, noise, packed, pose_spec, cfg,
                    disparity_smoothness / (2 ** s), rescale)
            loss, argmin = _ScaleLoss.apply(disps[s], meta, *pose_tensors)
            out["loss/%d" % s] = loss
            out[("argmin", s)] = argmin
            total = total + loss
        out["loss"] = total / len(scales)
    else:
        meta = (target, [target_pyr[s] for s in scales], K, inv_K, identity, noise_list, packed, pose_spec, cfg,
                [disparity_smoothness / (2 ** s) for s in scales], rescale)
        res = _MultiScaleLoss.apply(meta, *[disps[s] for s in scales], *pose_tensors)
        out["loss"] = res[0]
        for i, s in enumerate(scales):
            out["loss/%d" % s] = res[1][i]
            out[("argmin", s)] = res[2 + i]
    for s in scales:
        disp = disps[s]
        argmin = out[("argmin", s)]
        if automask:
            out["identity_selection/%d" % s] = _LazyMask(argmin, n_ident)
        if materialize:
            stats = depth_stats(disp.detach(), H, W) if rescale else None
            for i, src in enumerate(sources):
                sp = pose_spec[i]
                if sp[0] == "fixed":
                    Ti = sp[1].float()
                else:
                    j = sum(1 for q in pose_spec[:i] if q[0] == "net")
                    Ti = pose_matrix(pose_tensors[2 * j].detach()[:, 0], pose_tensors[2 * j + 1].detach()[:, 0],
                                     stats[:, 1] if rescale else None, sp[1])
                depth_up, sample, color = warp(disp.detach(), src, K, inv_K, Ti, H, W)
                out[("depth", 0, s)] = depth_up
                out[("sample", i, s)] = sample
                out[("color", i, s)] = color
    if materialize and automask:
        for s in scales:
            out["identity_selection/%d" % s] = out["identity_selection/%d" % s].float()
    return out


class _LazyMask:
    """identity_selection mask (trainer.py:529-530) materialised on first use: the training step itself never reads
    it (only Trainer.log does, every `log_frequency` steps), so the cast kernels are skipped on ordinary steps."""

    def __init__(self, argmin, n_ident):
        self.argmin, self.n_ident = argmin, n_ident

    def float(self):
        return (self.argmin >= self.n_ident).float()

    def cpu(self):
        return self.float().cpu()

    @property
    def shape(self):
        return self.argmin.shape

    def __len__(self):
        return self.argmin.shape[0]

    def __getitem__(self, idx):             # Trainer.log indexes outputs["identity_selection/s"][j] (trainer.py:620-623)
        return (self.argmin[idx] >= self.n_ident).float()

    def __getattr__(self, name):            # behave like the float tensor for anything else
        return getattr(self.float(), name)
