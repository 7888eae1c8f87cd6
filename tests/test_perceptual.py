"""Perceptual-term scheduling (planedepth_b200/perceptual.py) against the reference formula
(/root/reference/trainer.py:672-685): same value and gradient in fp32 (1e-4 gate, measured ~1e-6), bf16 autocast inside
BASELINE.json's 2e-2 gate, and the cache really removes passes (source features are shared by the target sides)."""
import pytest
import torch
import torch.nn as nn

from oracle import pd_oracle as O
from planedepth_b200.perceptual import PerceptualSchedule, reference_perceptual_loss


class TinyPc(nn.Module):
    """Three-level frozen feature pyramid with the structure of Vgg19_pc (layers.py:378-422): conv / relu / pool slices."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.s1 = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, padding=1), nn.ReLU())
        self.s2 = nn.Sequential(nn.MaxPool2d(2), nn.Conv2d(8, 16, 3, padding=1), nn.ReLU())
        self.s3 = nn.Sequential(nn.MaxPool2d(2), nn.Conv2d(16, 32, 3, padding=1), nn.ReLU())
        for p in self.parameters():
            p.requires_grad = False
        self.calls = 0

    def forward(self, x):
        self.calls += 1
        a = self.s1(x)
        b = self.s2(a)
        return a, b, self.s3(b)


def _data(dev, B=2, H=32, W=48):
    g = torch.Generator().manual_seed(0)
    pred = torch.rand(B, 3, H, W, generator=g).to(dev).requires_grad_(True)
    return pred, torch.rand(B, 3, H, W, generator=g).to(dev), torch.rand(B, 3, H, W, generator=g).to(dev)


@pytest.mark.parametrize("automask", [False, True])
def test_scheduled_equals_reference_formula_cpu(automask):
    net = TinyPc()
    pred, tgt, src = _data("cpu")
    want = O.perceptual_loss(net, pred, tgt, src if automask else None)
    (gw,) = torch.autograd.grad(want, pred)
    assert torch.allclose(reference_perceptual_loss(net, pred, tgt, src if automask else None), want, atol=1e-7)
    sched = PerceptualSchedule(net)
    got = sched(pred, tgt, src if automask else None)
    (gg,) = torch.autograd.grad(got, pred)
    assert abs(float(got) - float(want)) <= 1e-6 * max(1.0, abs(float(want)))
    assert (gg - gw).abs().max() <= 1e-5 * gw.abs().max()


def test_cache_shares_constant_features_across_sides():
    net = TinyPc()
    pred, tgt, src = _data("cpu")
    tgt2 = tgt.flip(-1).contiguous()
    sched = PerceptualSchedule(net)
    sched.new_batch()
    net.calls = 0
    a = sched(pred, tgt, src) + sched(pred, tgt2, src) + sched(pred, tgt, src)
    # reference: 3 sides x 3 passes = 9 network calls; scheduled: 3 pred passes + 2 constant passes (tgt+src batched, then tgt2)
    assert net.calls == 5 and sched.stats["cache_hits"] == 3
    b = sum(reference_perceptual_loss(net, pred, t, src) for t in (tgt, tgt2, tgt))
    assert abs(float(a) - float(b)) <= 1e-6 * abs(float(b))
    # an in-place update of a cached tensor invalidates its entry (the key carries the version counter)
    tgt.mul_(0.5)
    sched(pred, tgt, src)
    assert sched.stats["cache_hits"] == 4  # only src hit
    sched.new_batch()
    assert not sched._cache


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [("scheduled", 1e-4), ("scheduled_bf16", 2e-2)])
def test_boundary_perceptual_modes_on_gpu(mode, tol):
    """Through HotPathMixin.compute_losses with a VGG-style network on the GPU: total loss and d loss / d logits against the
    reference scheduling of the same network."""
    import sys
    import os

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_parity import CONFIGS, build_on
    from planedepth_b200.boundary import HotPath

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = TinyPc().cuda()
    res = {}
    for m in ("reference", mode):
        cg = build_on("cuda", CONFIGS[16], seed=900)  # disp_warp, automask, mask_novel, xz planes
        hp = HotPath(cg.opt, cg.target_sides, pc_net=net, perceptual_mode=m)
        losses = hp.process(cg.inputs, cg.outputs)
        losses["loss/total_loss"].backward()
        res[m] = ({k: float(v) for k, v in losses.items()}, cg.leaves["logits"].grad.clone())
    for k, v in res["reference"][0].items():
        assert abs(res[mode][0][k] - v) <= tol * max(1.0, abs(v)), (k, res[mode][0][k], v)
    gw, gg = res["reference"][1], res[mode][1]
    assert (gg - gw).abs().max() <= tol * gw.abs().max()
