"""Trainer-level drop-ins ("fused path"): the three Trainer methods on the photometric hot path with the
reference's signatures and dict contract, backed by the fused multi-scale library call.

    from trainer import Trainer            # the reference's trainer.py
    import sqlx

    class FusedTrainer(sqlx.FusedLossMixin, Trainer):
        pass

Replaces (same names, same arguments, same keys read and written):
  Trainer.generate_images_pred(inputs, outputs)       trainer.py:386-439
  Trainer.compute_reprojection_loss(pred, target)     trainer.py:441-453
  Trainer.compute_losses(inputs, outputs)             trainer.py:455-549
The whole computation happens in generate_images_pred (one sqlx_ms_loss_fwd call for all scales); compute_losses
returns the dict stashed there.  `outputs[("depth",0,s)]`, `("sample",f,s)`, `("color",f,s)`, `("color_identity",f,s)`
and `"identity_selection/s"` -- read only by Trainer.log -- are materialised when `self.sqlx_materialize` is true
(default); set it per step to `will_log` to skip those kernels on ordinary steps.
Unsupported exactly where the reference is: --predictive_mask (model never built, trainer.py:116-126) and
--v1_multiscale (latent bug, trainer.py:392-402) raise NotImplementedError.
"""
import torch

from .layers import SSIM
from .photometric import photometric_losses


class FusedLossMixin:
    sqlx_materialize = True
    sqlx_noises = None        # optional {scale: [B,S,H,W] standard-normal tensor}: bit parity with a seeded reference run

    def generate_images_pred(self, inputs, outputs):
        opt = self.opt
        if getattr(opt, "v1_multiscale", False) or getattr(opt, "predictive_mask", False):
            raise NotImplementedError("--v1_multiscale / --predictive_mask are not supported by the fused loss path")
        fids = list(opt.frame_ids[1:])
        posecnn_rescale = opt.pose_model_type == "posecnn" and not opt.use_stereo       # trainer.py:412-421
        poses = []
        for i, f in enumerate(fids):
            if f == "s":
                poses.append({"T": inputs["stereo_T"]})
                continue
            aa, tr = outputs[("axisangle", 0, f)], outputs[("translation", 0, f)]
            if aa.shape[1] == 1:                       # one pose per forward pass (num_pose_frames == 2, trainer.py:333-337)
                poses.append({"axisangle": aa, "translation": tr, "invert": f < 0})
            elif posecnn_rescale:                      # trainer.py:414-421 reads [:, 0] and inverts for f < 0
                poses.append({"axisangle": aa[:, 0:1], "translation": tr[:, 0:1], "invert": f < 0})
            else:                                      # all poses predicted together (trainer.py:365-368): [:, i], no invert
                poses.append({"axisangle": aa[:, i:i + 1], "translation": tr[:, i:i + 1], "invert": False})
        out = photometric_losses(
            {s: outputs[("disp", s)] for s in opt.scales},
            {s: inputs[("color", 0, s)] for s in opt.scales},
            [inputs[("color", f, 0)] for f in fids], inputs[("K", 0)], inputs[("inv_K", 0)], poses,
            noises=self.sqlx_noises, height=opt.height, width=opt.width, scales=tuple(opt.scales),
            disparity_smoothness=opt.disparity_smoothness, rescale_translation=posecnn_rescale,
            no_ssim=opt.no_ssim, avg_reprojection=opt.avg_reprojection,
            disable_automasking=opt.disable_automasking, materialize=bool(self.sqlx_materialize))
        for s in opt.scales:
            if ("depth", 0, s) in out:
                outputs[("depth", 0, s)] = out[("depth", 0, s)]
            for i, f in enumerate(fids):
                for key in ("sample", "color"):
                    if (key, i, s) in out:
                        outputs[(key, f, s)] = out[(key, i, s)]
                if not opt.disable_automasking:
                    outputs[("color_identity", f, s)] = inputs[("color", f, 0)]           # trainer.py:437-439
            k = "identity_selection/%d" % s
            if k in out:
                outputs[k] = out[k]
        self._sqlx_losses = {k: v for k, v in out.items() if isinstance(k, str) and k.startswith("loss")}

    def compute_reprojection_loss(self, pred, target):
        """trainer.py:441-453, differentiable wrt both arguments (sqlx.SSIM carries its own backward)."""
        l1_loss = torch.abs(target - pred).mean(1, True)
        if self.opt.no_ssim:
            return l1_loss
        if not isinstance(getattr(self, "ssim", None), SSIM):
            self.ssim = SSIM()
        return 0.85 * self.ssim(pred, target).mean(1, True) + 0.15 * l1_loss

    def compute_losses(self, inputs, outputs):
        losses = getattr(self, "_sqlx_losses", None)
        if losses is None:
            raise RuntimeError("compute_losses called before generate_images_pred (trainer.py:296-297 order)")
        self._sqlx_losses = None
        return losses
