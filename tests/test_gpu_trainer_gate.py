"""GPU (-m gpu): BASELINE.json configs[0], the PR1 parity gate — the reference's own ``Trainer.process_batch`` +
``backward`` (trainer.py:325-356, 300) with seeded ResNet-18 + DepthDecoder on synthetic random images, run

    (a) unpatched: the unmodified reference code staged under baseline/_ref (F.grid_sample + autograd), and
    (b) with INTEGRATION.md §2's patch (HotPathMixin in front of the reference Trainer, disp_rowwise promise),

on the same weights and inputs, on the GPU.  Every ``losses`` entry must agree to 1e-4 and the encoder / decoder (/ pose
network) parameter gradients to 1e-4 of the largest gradient magnitude of their network.  Skipped when baseline/_ref did not
travel with the snapshot (it is git-ignored; ``__graft_entry__.build()`` stages it)."""
import copy
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import shim  # noqa: E402

from helpers import REPORT, bounded_check  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not shim.available(), reason="baseline/_ref not staged")]

TOL = 1e-4

CASES = {
    # BASELINE configs[0]: batch 2, 640x192, 49 vertical planes, stereo warp, ResNet-18 encoder, L1 photometric term
    "stereo_l1": dict(xz_levels=0, novel_frame_ids=[]),
    # Laplacian mixture + learned plane residual, 49 + 14 planes (the dense cat layout of depth_decoder.py:181-182)
    "mixture_residual": dict(xz_levels=14, use_mixture_loss=True, plane_residual=True, novel_frame_ids=[]),
    # monocular frames through homography_warp + pose network, automask min-reprojection
    "homography_automask": dict(xz_levels=14, warp_type="homography_warp", novel_frame_ids=[-1, 1], automask=True, use_colmap=False),
}


def make_inputs(opt, B, seed, device):
    from planedepth_b200.synthetic import make_batch, make_opt

    o = make_opt(novel_frame_ids=list(opt.novel_frame_ids))
    b = make_batch(B, opt.height, opt.width, o, seed=seed, device="cpu", requires_grad=False)
    g = torch.Generator().manual_seed(seed + 5)
    inputs = {k: v.clone() for k, v in b.inputs.items()}
    for s in ["l", "r"] + list(opt.novel_frame_ids):  # colour-jittered copy: what the networks see (mono_dataset.py:162-171)
        inputs[("color_aug", s)] = (inputs[("color", s)] * (0.9 + 0.2 * torch.rand(B, 3, 1, 1, generator=g))).clamp(0, 1)
    return {k: v.to(device) for k, v in inputs.items()}


def run(trainer_cls, opt, models, inputs, pc_net):
    dev = torch.device("cuda")
    t = shim.bare_trainer(opt, dev, models=models, pc_net=pc_net)
    t.__class__ = trainer_cls
    for m in models.values():
        m.train()
        m.zero_grad(set_to_none=True)
    outputs, losses = t.process_batch({k: v.clone() for k, v in inputs.items()})
    losses["loss/total_loss"].backward()
    torch.cuda.synchronize()
    grads = {}
    for name, m in models.items():
        for pn, prm in m.named_parameters():
            if prm.grad is not None:
                grads[name + "." + pn] = prm.grad.detach().clone()
    return {k: float(v) for k, v in losses.items()}, grads, outputs


@pytest.mark.parametrize("case", list(CASES))
def test_process_batch_patched_matches_reference(case):
    tr, layers, nets, _ = shim.load()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    dev = torch.device("cuda")
    opt = shim.default_options(num_layers=18, height=192, width=640, batch_size=2, **CASES[case])
    models = shim.build_models(opt, dev, seed=3)
    torch.manual_seed(11)
    pc_net = layers.Vgg19_pc().to(dev).eval()  # VGG19 architecture, seeded random weights (no network for the checkpoint)
    inputs = make_inputs(opt, 2, 21, dev)
    Patched = shim.patch_trainer_class(tr.Trainer)
    # identical weights AND identical BatchNorm running statistics for both runs
    models_b = {k: copy.deepcopy(m) for k, m in models.items()}
    lr, gr, out_r = run(tr.Trainer, opt, models, inputs, pc_net)
    lp, gp, out_p = run(Patched, opt, models_b, inputs, pc_net)
    assert set(lr) == set(lp), (sorted(lr), sorted(lp))
    for k in lr:
        bounded_check(torch.tensor(lp[k]), torch.tensor(lr[k]), TOL * max(1.0, abs(lr[k])), "%s %s" % (case, k))
    for s in (["r"] + list(opt.novel_frame_ids)):
        bounded_check(out_p[("rgb_rec", s)], out_r[("rgb_rec", s)].detach(), TOL, "%s rgb_rec@%s" % (case, s), allow_frac=2e-4)
    assert set(gr) == set(gp)
    scale = {}
    for k, g in gr.items():
        net = k.split(".")[0]
        scale[net] = max(scale.get(net, 0.0), float(g.abs().max()))
    worst = {}
    for k, g in gr.items():
        net = k.split(".")[0]
        st = bounded_check(gp[k], g, TOL * scale[net], "%s grad %s" % (case, k))
        worst[net] = max(worst.get(net, 0.0), st["max_err"] / scale[net])
    print(case, "losses", lp, "worst gradient error / network max:", worst)
