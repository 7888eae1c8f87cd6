"""Import shim that lets the UNMODIFIED reference tree staged under baseline/_ref (git-ignored; __graft_entry__.build()
copies /root/reference there, the gpurun snapshot carries it to the GPU box) be imported and driven without its
constructor (Trainer.__init__ needs NCCL + LOCAL_RANK + KITTI + network access).

Used by: bench.py --impl reference (CPU arm), bench.py's reference_gpu leg, tests/test_gpu_trainer_gate.py (the PR1 gate:
reference Trainer.process_batch unpatched vs patched with planedepth_b200), scratch/ref_gpu_probe.py.  Nothing of the
reference is copied into the repo; this file only stubs missing third-party modules and builds a bare Trainer object."""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "trainer.py"))


_loaded = None
_saved_cuda = None


def enter_cpu_mode() -> None:
    """The reference calls ``.cuda()`` on helper tensors (layers.py:140,196; depth_decoder.py:148): for a CPU run those calls
    become identities.  Process-wide monkeypatch: leave_cpu_mode() undoes it (bench.py runs its GPU legs before / after)."""
    global _saved_cuda
    import torch
    import torch.nn as nn

    if _saved_cuda is None:
        _saved_cuda = (torch.Tensor.cuda, nn.Module.cuda)
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self


def leave_cpu_mode() -> None:
    global _saved_cuda
    import torch
    import torch.nn as nn

    if _saved_cuda is not None:
        torch.Tensor.cuda, nn.Module.cuda = _saved_cuda
        _saved_cuda = None


def load(cpu_only: bool = False):
    """Returns (trainer module, layers module, networks package, options module) of the reference, or None."""
    global _loaded
    if _loaded is not None:
        if cpu_only:
            enter_cpu_mode()
        return _loaded
    if not available():
        return None
    import torch
    import torch.nn as nn

    for name in ["tensorboardX", "IPython", "skimage", "skimage.transform", "matplotlib"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["IPython"].embed = lambda *a, **k: None
    sys.modules["matplotlib"].scale = None
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    six = types.ModuleType("torch._six")
    six.string_classes = (str, bytes)
    sys.modules["torch._six"] = six
    torch._six = six
    if not torch.cuda.is_available():
        cpu_only = True
    if cpu_only:
        enter_cpu_mode()
    import PIL.Image

    if not hasattr(PIL.Image, "ANTIALIAS"):
        PIL.Image.ANTIALIAS = PIL.Image.LANCZOS
    import torchvision.models as tvm

    if not hasattr(tvm.resnet, "model_urls"):
        tvm.resnet.model_urls = {}  # pose_net.py:57 (only dereferenced with pretrained=True)
    # no network: `pretrained=True` constructors (layers.py:381, resnet_encoder.py:35) get seeded random weights instead
    for fn in ("vgg19", "resnet18", "resnet34", "resnet50"):
        orig = getattr(tvm, fn)
        if getattr(orig, "_pd_offline", False):
            continue

        def make(orig):
            def offline(pretrained=False, *a, **k):
                k.pop("weights", None)
                return orig(weights=None, *a, **k)

            offline._pd_offline = True
            return offline

        setattr(tvm, fn, make(orig))
    sys.path.insert(0, REF)
    import layers as ref_layers
    import networks as ref_networks
    import options as ref_options
    import trainer as ref_trainer

    _loaded = (ref_trainer, ref_layers, ref_networks, ref_options)
    return _loaded


def default_options(**over):
    """The reference's own argparse defaults (options.py) with overrides."""
    _, _, _, ref_options = load()
    argv, sys.argv = sys.argv, [sys.argv[0]]
    try:
        opt = ref_options.MonodepthOptions().parser.parse_args([])
    finally:
        sys.argv = argv
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def bare_trainer(opt, device, models=None, pc_net=None):
    """A reference Trainer without its constructor: exactly the attributes process_batch / predict_poses /
    pred_novel_images / compute_losses / generate_post_process_disp read (trainer.py:46-176)."""
    import torch.nn as nn

    ref_trainer, ref_layers, _, _ = load()
    t = object.__new__(ref_trainer.Trainer)
    t.opt = opt
    t.device = device
    t.target_sides = ([] if opt.no_stereo else ["r"]) + list(opt.novel_frame_ids)
    t.models = models or {}
    t.softmax = nn.Softmax(1)
    t.ssim = ref_layers.SSIM().to(device)
    t.backproject_depth = ref_layers.BackprojectDepth(opt.height, opt.width).to(device)
    t.project_3d = ref_layers.Project3D(opt.height, opt.width).to(device)
    t.homography_warp = ref_layers.HomographyWarp(opt.height, opt.width).to(device)
    t.pc_net = pc_net
    return t


def build_models(opt, device, seed=0):
    """Seeded ResNet encoder + DepthDecoder (+ pose networks for monocular frames) as Trainer.create_models builds them
    (trainer.py:186-204, 95-97), un-wrapped (no DDP), random weights."""
    import torch

    _, _, nets, _ = load()
    torch.manual_seed(seed)
    models = {}
    models["encoder"] = nets.ResnetEncoder(opt.num_layers, False)
    models["depth"] = nets.DepthDecoder(models["encoder"].num_ch_enc, opt.disp_levels, opt.disp_min, opt.disp_max, opt.num_ep,
                                        pe_type=opt.pe_type, use_denseaspp=opt.use_denseaspp, xz_levels=opt.xz_levels,
                                        yz_levels=opt.yz_levels, use_mixture_loss=opt.use_mixture_loss,
                                        render_probability=opt.render_probability, plane_residual=opt.plane_residual)
    if len(opt.novel_frame_ids) > 0 and not opt.use_colmap:
        models["pose_encoder"] = nets.ResnetPoseEncoder(18, False, 2)
        models["pose"] = nets.PoseDecoder(models["pose_encoder"].num_ch_enc, num_input_features=1, num_frames_to_predict_for=1, num_ep=8)
    for k in models:
        models[k] = models[k].to(device)
    return models


def patch_trainer_class(Trainer):
    """the reference Trainer with the mixin in front (its pred_novel_images / generate_images_pred /
    compute_losses / perceptual_loss / generate_post_process_disp win the method resolution) and the x-constancy promise."""
    from planedepth_b200.boundary import HotPathMixin

    class Patched(HotPathMixin, Trainer):
        disp_rowwise = property(lambda self: self.opt.yz_levels == 0 and self.opt.net_type == "ResNet")

    return Patched
