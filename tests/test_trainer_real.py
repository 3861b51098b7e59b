"""End-to-end drop-in through the reference's REAL Trainer class (trainer.py:266-299):

    class FusedTrainer(sqlx.FusedLossMixin, Trainer): pass

with `Trainer` imported unmodified from /root/reference (build container) or oracle/_ref (GPU box; oracle/build_ref.py).
CPU: the MRO / attribute contract of the composed class.  GPU: generate_images_pred + compute_losses of the composed class
against the same two methods of the reference class on the same inputs, noise and device -- loss, per-scale losses, the
`outputs` keys Trainer.log and compute_depth_losses read, and the gradients autograd hands back to the networks."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not available (oracle/build_ref.py)")


def _compose(device):
    import sqlx
    ref = ref_shim.load(force_cpu=(device == "cpu"))

    class FusedTrainer(sqlx.FusedLossMixin, ref.trainer.Trainer):
        pass
    return ref, FusedTrainer


@needs_ref
def test_composed_class_contract():
    import sqlx
    ref, FusedTrainer = _compose("cpu")
    T = ref.trainer.Trainer
    # the three hot-path methods resolve to the mixin, everything else to the reference class
    for name in ("generate_images_pred", "compute_reprojection_loss", "compute_losses"):
        assert getattr(FusedTrainer, name) is getattr(sqlx.FusedLossMixin, name)
    for name in ("process_batch", "predict_poses", "train", "run_epoch", "val", "log", "save_model", "compute_depth_losses",
                 "set_train", "set_eval"):
        assert getattr(FusedTrainer, name) is getattr(T, name)
    assert FusedTrainer.__mro__[1] is sqlx.FusedLossMixin and FusedTrainer.__mro__[2] is T
    # same call order contract as trainer.py:296-297
    ft = FusedTrainer.__new__(FusedTrainer)
    with pytest.raises(RuntimeError):
        ft.compute_losses({}, {})


def _batch(B, H, W, dev, stereo):
    from _cases import smooth_images, kitti_K, depth_like
    g = torch.Generator().manual_seed(17)
    n = 4 if stereo else 3
    fr = smooth_images(g, B, H, W, n)
    K, iK = kitti_K(B, H, W)
    fids = [0, -1, 1] + (["s"] if stereo else [])
    inputs = {("K", 0): K.to(dev), ("inv_K", 0): iK.to(dev)}
    for f, i in zip(fids, [1, 0, 2, 3][:n]):
        inputs[("color", f, 0)] = fr[i].to(dev)
    if stereo:
        st = torch.eye(4).repeat(B, 1, 1)
        st[:, 0, 3] = 0.1
        inputs["stereo_T"] = st.to(dev)
    leaves = {"disp": depth_like(g, B, H // 2, W // 2).to(dev).requires_grad_(True)}
    for f in (-1, 1):
        leaves["aa%d" % f] = (0.01 * torch.randn(B, 1, 1, 3, generator=g)).to(dev).requires_grad_(True)
        leaves["tr%d" % f] = (0.05 * torch.randn(B, 1, 1, 3, generator=g)).to(dev).requires_grad_(True)
    return inputs, leaves


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("stereo", [False, True])
def test_fused_trainer_matches_reference_trainer(stereo):
    ref, FusedTrainer = _compose("cuda")
    B, H, W = 2, 96, 160
    dev = "cuda"
    T_ref = ref_shim.make_trainer(B, H, W, scales=(0,), use_stereo=stereo, device=dev)
    T_fus = FusedTrainer.__new__(FusedTrainer)
    T_fus.__dict__.update(T_ref.__dict__)
    inputs, leaves = _batch(B, H, W, dev, stereo)

    def outputs_of():
        out = {("disp", 0): leaves["disp"]}
        for f in (-1, 1):
            out[("axisangle", 0, f)], out[("translation", 0, f)] = leaves["aa%d" % f], leaves["tr%d" % f]
            out[("cam_T_cam", 0, f)] = ref.layers.transformation_from_parameters(leaves["aa%d" % f][:, 0],
                                                                                 leaves["tr%d" % f][:, 0], invert=(f < 0))
        return out
    names = list(leaves)
    # reference: the tie-break noise is torch.randn(...) on the CPU generator (trainer.py:516) -> seed it, and hand the
    # same draw to the fused path
    S = 3 if stereo else 2
    torch.manual_seed(5)
    noise = torch.randn(B, S, H, W)
    out_ref = outputs_of()
    T_ref.generate_images_pred(inputs, out_ref)
    torch.manual_seed(5)
    loss_ref = T_ref.compute_losses(inputs, out_ref)
    g_ref = torch.autograd.grad(loss_ref["loss"], [leaves[k] for k in names])
    T_fus.sqlx_noises = {0: noise.to(dev)}
    T_fus.sqlx_materialize = True
    out_fus = outputs_of()
    T_fus.generate_images_pred(inputs, out_fus)
    loss_fus = T_fus.compute_losses(inputs, out_fus)
    g_fus = torch.autograd.grad(loss_fus["loss"], [leaves[k] for k in names])
    assert abs(float(loss_fus["loss"]) - float(loss_ref["loss"])) < 1e-5
    assert abs(float(loss_fus["loss/0"]) - float(loss_ref["loss/0"])) < 1e-5
    # what Trainer.log / compute_depth_losses read (trainer.py:557, 593-625)
    fids = [-1, 1] + (["s"] if stereo else [])
    assert float((out_fus[("depth", 0, 0)] - out_ref[("depth", 0, 0)]).abs().max()) < 1e-4
    for f in fids:
        assert float((out_fus[("color", f, 0)] - out_ref[("color", f, 0)]).abs().max()) < 2e-4
        assert float((out_fus[("sample", f, 0)] - out_ref[("sample", f, 0)]).abs().max()) < 1e-4
        assert out_fus[("color_identity", f, 0)] is inputs[("color", f, 0)]
    sel_f, sel_r = out_fus["identity_selection/0"], out_ref["identity_selection/0"]
    assert sel_f.shape == sel_r.shape and float((sel_f.float() != sel_r.float()).float().mean()) < 2e-3
    assert sel_f[0].shape == sel_r[0].shape                       # Trainer.log indexes it per sample
    for k, a, b in zip(names, g_fus, g_ref):
        rel = float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
        assert rel < 3e-2, (k, rel)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("lite", [True, False])
def test_inference_wrapper_and_flip_tta(lite):
    """the inference path of SQLdepth.py:9-50 / evaluate_depth_config.py:126-161 with the reference's own
    encoder and a reference decoder converted in place by sqlx.fuse_depth_decoder: depth 1e-4 relative against the
    unmodified modules on the same device, and predict_disparity(post_process=True) against the reference's
    batch_post_process_disparity (evaluate_depth_config.py:51-59) applied to the reference forward."""
    import numpy as np
    import sqlx
    ref = ref_shim.load(force_cpu=False)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    nn = torch.nn

    class Wrapper(nn.Module):                       # the two attributes and the forward of SQLdepth.py:9-50
        def __init__(self):
            super().__init__()
            self.encoder = ref.networks.LiteResnetEncoderDecoder(model_dim=32)
            cls = ref.networks.Lite_Depth_Decoder_QueryTr if lite else ref.networks.Depth_Decoder_QueryTr
            self.depth_decoder = cls(in_channels=32, patch_size=16, dim_out=64, embedding_dim=32, query_nums=64, num_heads=4,
                                     min_val=0.001, max_val=80.0)

        def forward(self, x):
            return self.depth_decoder(self.encoder(x))["disp", 0]
    model = Wrapper().cuda().eval()
    x = torch.rand(2, 3, 192, 640, device="cuda")
    with torch.no_grad():
        want = model(x)
        both = model(torch.cat((x, torch.flip(x, [3])), 0)).cpu()[:, 0].numpy()
    state_before = {k: v.clone() for k, v in model.depth_decoder.state_dict().items()}
    assert sqlx.fuse_depth_decoder(model) == 1
    assert isinstance(model.depth_decoder, sqlx.Depth_Decoder_QueryTr) and not model.depth_decoder.training
    for k, v in model.depth_decoder.state_dict().items():
        assert torch.equal(v, state_before[k]), k
    with torch.no_grad():
        got = model(x)
    assert float(((got - want) / want).abs().max()) < 1e-4
    # flip test-time augmentation: reference numpy blend of the reference forward vs the device-side path
    import evaluate_depth_config as E
    want_pp = E.batch_post_process_disparity(both[:2], both[2:, :, ::-1])
    got_pp = sqlx.predict_disparity(model.encoder, model.depth_decoder, x, post_process=True).cpu().numpy()
    assert got_pp.shape == want_pp.shape
    assert float(np.abs(got_pp - want_pp).max() / np.abs(want_pp).max()) < 1e-4
    got_plain = sqlx.predict_disparity(model.encoder, model.depth_decoder, x, post_process=False)
    assert float(((got_plain - want[:, 0]) / want[:, 0]).abs().max()) < 1e-4
