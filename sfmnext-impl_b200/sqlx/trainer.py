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
from .photometric import indoor_losses, photometric_losses, warp


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
            disable_automasking=opt.disable_automasking,
            # Trainer.val() (trainer.py:363-384, under no_grad) always logs and reads outputs[("depth",0,0)] in
            # compute_depth_losses (:557): materialise there whatever the per-step switch says
            materialize=bool(self.sqlx_materialize) or not torch.is_grad_enabled())
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


class IndoorFusedLossMixin(FusedLossMixin):
    """Indoor trainer (trainer_indoor.py, --use_improved_mini_reproj_loss):

        class FusedIndoorTrainer(sqlx.IndoorFusedLossMixin, trainer_indoor.Trainer): ...

    Replaces generate_images_pred (trainer_indoor.py:512-599) and compute_losses_with_occ (:615-719, returns
    `(total_loss, losses)` as the reference does).  The single loss scale of the SQL decoder (opt.scales == [0]) and the
    un-rectified frames (not --use_rectify_net) are supported; `outputs[("depth_ref", f, 0)]` must hold the network
    depth of every source frame (:370-377).  `outputs[("depth",0,0)]` and `("color",f,0)` (read by log()) are
    materialised when `self.sqlx_materialize` is true."""

    def generate_images_pred(self, inputs, outputs):
        opt = self.opt
        if not getattr(opt, "use_improved_mini_reproj_loss", False):
            return FusedLossMixin.generate_images_pred(self, inputs, outputs)
        if list(opt.scales) != [0] or getattr(opt, "v1_multiscale", False) or getattr(opt, "use_rectify_net", False):
            raise NotImplementedError("the fused indoor loss supports scales == [0] without --v1_multiscale / "
                                      "--use_rectify_net")
        fids = list(opt.frame_ids[1:])
        rescale = opt.pose_model_type == "posecnn" and not opt.use_stereo               # trainer_indoor.py:537
        poses = []
        for f in fids:
            if f == "s":
                poses.append({"T": inputs["stereo_T"]})
            else:
                poses.append({"axisangle": outputs[("axisangle", 0, f)][:, 0:1],
                              "translation": outputs[("translation", 0, f)][:, 0:1], "invert": f < 0})
        if not rescale:                         # the camera transforms predict_poses produced (:533-535)
            poses = [{"T": inputs["stereo_T"]} if f == "s" else {"T": outputs[("cam_T_cam", 0, f)]} for f in fids]
        noise = self.sqlx_noises.get(0) if self.sqlx_noises else None
        out = indoor_losses(outputs[("disp", 0)], inputs[("color", 0, 0)], [inputs[("color", f, 0)] for f in fids],
                            [outputs[("depth_ref", f, 0)] for f in fids], inputs[("K", 0)], inputs[("inv_K", 0)], poses,
                            noise, height=opt.height, width=opt.width, disparity_smoothness=opt.disparity_smoothness,
                            reg_wt=opt.reg_wt, rescale_translation=rescale, no_ssim=opt.no_ssim,
                            avg_reprojection=opt.avg_reprojection, disable_automasking=opt.disable_automasking)
        if self.sqlx_materialize:
            from .photometric import depth_stats, pose_matrix
            disp = outputs[("disp", 0)].detach()
            stats = depth_stats(disp, opt.height, opt.width) if rescale else None
            for f, pose in zip(fids, poses):
                T = pose["T"].detach().float() if "T" in pose else pose_matrix(
                    pose["axisangle"].detach()[:, 0], pose["translation"].detach()[:, 0], stats[:, 1], pose["invert"])
                depth_up, _, color = warp(disp, inputs[("color", f, 0)], inputs[("K", 0)], inputs[("inv_K", 0)], T,
                                          opt.height, opt.width, want_sample=False)
                outputs[("depth", 0, 0)] = depth_up
                outputs[("color", f, 0)] = color
                if not opt.disable_automasking:
                    outputs[("color_identity", f, 0)] = inputs[("color", f, 0)]          # :596-598
        self._sqlx_losses = {"loss/0": out["loss/0"]}
        self._sqlx_total = out["loss"]

    def compute_losses_with_occ(self, inputs, outputs):
        losses = getattr(self, "_sqlx_losses", None)
        if losses is None:
            raise RuntimeError("compute_losses_with_occ called before generate_images_pred (trainer_indoor.py:378-407)")
        self._sqlx_losses = None
        total = self._sqlx_total / self.num_scales                                        # :714-715
        losses["loss"] = total
        return total, losses
