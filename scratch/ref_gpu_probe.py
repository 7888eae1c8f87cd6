#!/usr/bin/env python
"""Reference GPU path on the same B200: the UNMODIFIED reference code (baseline/_ref through the import shim; F.grid_sample +
autograd) doing the same work as one bench.py step — pred_novel_images + photometric term, forward + backward — on the same
synthetic batch.  The denominator of north_star's ">= 10x the reference GPU grid_sample + SSIM path".
usage: ref_gpu_probe.py [cfg2 cfg3 ...]   -> one JSON object on stdout"""
import json, os, sys, types
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from planedepth_b200.synthetic import make_batch, make_opt  # noqa: E402


def load_ref():
    ref = os.path.join(ROOT, "baseline", "_ref")
    for name in ["tensorboardX", "IPython", "skimage", "skimage.transform", "matplotlib"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["IPython"].embed = lambda *a, **k: None
    sys.modules["matplotlib"].scale = None
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    six = types.ModuleType("torch._six"); six.string_classes = (str, bytes)
    sys.modules["torch._six"] = six; torch._six = six
    import PIL.Image
    if not hasattr(PIL.Image, "ANTIALIAS"):
        PIL.Image.ANTIALIAS = PIL.Image.LANCZOS
    sys.path.insert(0, ref)
    import layers, trainer
    return trainer, layers


def main():
    names = sys.argv[1:] or ["cfg2", "cfg3", "cfg4", "cfg5"]
    tr, layers = load_ref()
    res = {}
    for name in names:
        B, H, W, over, photometric, desc = bench.CONFIGS[name]
        opt = make_opt(**over)
        mnov = opt.self_distillation > 0
        opt.self_distillation = 0.0
        batch = make_batch(B, H, W, opt, seed=1234, device="cuda", layout="reference", mask_novel=mnov)
        leaves = list(batch.leaves.values())
        t = object.__new__(tr.Trainer)
        t.opt = types.SimpleNamespace(**vars(opt))
        t.opt.use_ssim = photometric == "ssim_l1"
        t.target_sides = batch.target_sides
        t.softmax = nn.Softmax(1)
        t.ssim = layers.SSIM().cuda()
        t.homography_warp = layers.HomographyWarp(H, W)
        t.backproject_depth = layers.BackprojectDepth(H, W)
        t.project_3d = layers.Project3D(H, W)

        def step():
            out = batch.attach(dict(batch.outputs))
            tr.Trainer.pred_novel_images(t, batch.inputs, out)
            total = 0
            for s in batch.target_sides:
                if photometric == "ssim_l1":
                    total = total + tr.Trainer.compute_reprojection_loss(t, out[("rgb_rec", s)], batch.inputs[("color", s)]).mean()
                elif opt.use_mixture_loss:
                    err = torch.abs(out[("rgb_rec_layered", s)] - batch.inputs[("color", s)][:, None]).mean(2)
                    total = total + layers.multimodal_loss(err, out[("sigma_rec", s)], out[("pi_rec", s)], dist="lap").mean()
                else:
                    total = total + torch.abs(out[("rgb_rec", s)] - batch.inputs[("color", s)]).mean()
            torch.autograd.grad(total, leaves, allow_unused=True)
            return total

        torch.cuda.reset_peak_memory_stats()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = 10
        for _ in range(n):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        res[name] = {"workload": desc, "fwd_bwd_ms": ms, "images_per_s": B / ms * 1e3, "peak_GB": torch.cuda.max_memory_allocated() / 1e9,
                     "sides": len(batch.target_sides)}
        del batch, leaves, t
        torch.cuda.empty_cache()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
