"""Import the UNMODIFIED reference (/root/reference) in-process, CPU only.

TEST INFRASTRUCTURE ONLY.  This module exists so that `oracle/make_golden.py`
can execute the reference's own code (trainer.Trainer.generate_images_pred /
compute_losses, networks.Depth_Decoder_QueryTr, layers.SSIM ...) to produce the
golden vectors under tests/golden/, and so that the oracle restatement can be
checked against the real thing while /root/reference is mounted (build
container only; the GPU box never has it).

Shims applied:
  1. stub modules for kornia / timm / skimage (imported at module scope by
     layers.py:8, networks/Unet.py:3, datasets/kitti_dataset.py:10, never used
     on the hot path);
  2. torchvision resnet constructors patched to weights=None (no network);
  3. Tensor.cuda / Module.cuda made no-ops when no GPU is visible
     (trainer.py:517 calls .cuda() on the tie-break noise);
  4. Trainer built with Trainer.__new__ and the handful of attributes the
     hot-path methods read (trainer.py:386-549).
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference where it is mounted (the build container); otherwise oracle/_ref, the unmodified copy that
# oracle/build_ref.py makes (git-ignored, shipped to the GPU box like the built .so files)
REFERENCE_ROOT = os.environ.get("SQLX_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/trainer.py") else os.path.join(_HERE, "_ref"))


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "trainer.py"))


_loaded = {}
_orig_cuda = {}


def cuda_noop(on):
    """Make Tensor.cuda / Module.cuda no-ops (on=True) or restore them: the reference calls .cuda() unconditionally
    (trainer.py:517 on the tie-break noise), which must stay on the host for the CPU arm."""
    import torch
    if on and not _orig_cuda:
        _orig_cuda.update(t=torch.Tensor.cuda, m=torch.nn.Module.cuda)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    elif not on and _orig_cuda:
        torch.Tensor.cuda, torch.nn.Module.cuda = _orig_cuda.pop("t"), _orig_cuda.pop("m")


def load(force_cpu=False):
    """Returns a namespace with the reference modules: layers, networks, trainer, options.
    force_cpu: make Tensor.cuda / Module.cuda no-ops even when a GPU is visible (the CPU arm of bench.py on the GPU
    box: trainer.py:517 calls .cuda() on the tie-break noise)."""
    import torch
    cuda_noop(force_cpu or not torch.cuda.is_available())
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not found at %s (run oracle/build_ref.py where /root/reference is mounted)"
                           % REFERENCE_ROOT)

    def stub(name, **attrs):
        if name in sys.modules:
            return
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    stub("kornia"); stub("kornia.geometry"); stub("kornia.geometry.depth", depth_to_3d=None)
    stub("timm", create_model=None)
    stub("skimage"); stub("skimage.transform")
    try:
        import tensorboardX  # noqa: F401
    except Exception:
        stub("tensorboardX", SummaryWriter=object)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torchvision.models as tvm
    import layers as ref_layers
    import networks as ref_networks
    import trainer as ref_trainer
    import options as ref_options
    import networks.resnet_encoder as RE
    import networks.lite_res_encoder as LRE
    for mod in (RE, LRE):
        for n in ("resnet18", "resnet34", "resnet50", "resnet101", "resnet152"):
            setattr(mod.models, n,
                    (lambda ctor: (lambda pretrained=False, **k: ctor(weights=None)))(getattr(tvm, n)))
    _loaded.update(layers=ref_layers, networks=ref_networks, trainer=ref_trainer, options=ref_options)
    return types.SimpleNamespace(**_loaded)


def make_trainer(batch_size, height, width, scales=(0,), frame_ids=(0, -1, 1), use_stereo=False,
                 extra_args=(), device="cpu"):
    """A reference Trainer with only the attributes the loss path needs (no models, no data)."""
    import torch
    ref = load(force_cpu=(str(device) == "cpu"))
    argv = ["--height", str(height), "--width", str(width), "--batch_size", str(batch_size),
            "--scales"] + [str(s) for s in scales] + ["--frame_ids"] + [str(f) for f in frame_ids if f != "s"]
    if use_stereo:
        argv.append("--use_stereo")
    argv += list(extra_args)
    opt = ref.options.MonodepthOptions().parser.parse_args(argv)
    if use_stereo and "s" not in opt.frame_ids:
        opt.frame_ids.append("s")          # trainer.py:52-53
    T = ref.trainer.Trainer.__new__(ref.trainer.Trainer)
    T.opt = opt
    T.device = torch.device(device)
    T.num_scales = len(opt.scales)
    T.num_input_frames = len(opt.frame_ids)
    T.num_pose_frames = 2
    T.use_pose_net = True
    T.models = {}
    if not opt.no_ssim:
        T.ssim = ref.layers.SSIM()
    T.backproject_depth = {0: ref.layers.BackprojectDepth(batch_size, height, width).to(T.device)}
    T.project_3d = {0: ref.layers.Project3D(batch_size, height, width).to(T.device)}
    return T
