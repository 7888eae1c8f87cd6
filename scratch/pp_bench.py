#!/usr/bin/env python
"""Times generate_post_process_disp's post-network part (trainer.py:421-466) on the GPU: this library
(pd_occlusion_masks_fwd) against the unmodified reference code (baseline/_ref through the import shim, F.grid_sample
on the same GPU).  usage: pp_bench.py [B N H W]"""
import json, os, sys, types
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_ref():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref, "trainer.py")):
        return None
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
    import trainer
    return trainer


def timeit(fn, n=10, w=3):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B, N, H, W = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else (12, 49, 192, 640)
    from planedepth_b200.functional import occlusion_masks
    g = torch.Generator().manual_seed(0)
    n_xz = 14 if N > 49 else 0
    n_v = N - n_xz
    lev = torch.arange(n_v, dtype=torch.float32)[None] + torch.rand(2 * B, n_v, generator=g) - 0.5
    base = (300.0 * (2.0 / 300.0) ** (lev / (n_v - 1))).reshape(2 * B, n_v, 1, 1).cuda()
    disp_layered = base.expand(2 * B, n_v, H, W)
    if n_xz:
        gy = torch.linspace(-1, 1, H)[None, None, :, None].expand(2 * B, 1, H, W).cuda()
        h = (0.1852 + 0.1852 * torch.rand(2 * B, n_xz, generator=g)).cuda()
        disp_layered = torch.cat([disp_layered, 0.1 * 0.58 * W / (h[:, :, None, None] * 1.92 / (gy.clamp_min(1e-7) / 2.0))], 1)
    logits = torch.randn(2 * B, N, H, W, device="cuda")
    outs = {"logits": logits, "probability": torch.softmax(logits, 1), "disp_layered": disp_layered, "disp": 1 + 20 * torch.rand(2 * B, 1, H, W, device="cuda")}
    res = {"shape": [B, N, H, W], "layout": "expand" if not n_xz else "dense cat"}
    res["ours_ms"] = timeit(lambda: occlusion_masks(outs["logits"], outs["probability"], outs["disp_layered"], outs["disp"]))
    if n_xz:
        rw = outs["disp_layered"][..., :1].expand(-1, -1, -1, W)
        res["ours_rowwise_ms"] = timeit(lambda: occlusion_masks(outs["logits"], outs["probability"], rw, outs["disp"]))
    res["ours_exact_ms"] = timeit(lambda: occlusion_masks(outs["logits"], outs["probability"], outs["disp_layered"], outs["disp"], exact_coords=True))
    x1 = H * W * 4
    res["algorithmic_bytes"] = B * x1 * (2 * N + 3 * N + 2 * N + N + 6)  # A: read N write N; B: read N; x2 halves; mask_novel: read N
    res["ours_GBps"] = res["algorithmic_bytes"] / res["ours_ms"] / 1e6
    tr = load_ref()
    if tr is not None:
        t = object.__new__(tr.Trainer)
        t.opt = types.SimpleNamespace(num_ep=8, net_type="ResNet")
        t.softmax = nn.Softmax(1)
        t.fixed_models = {"encoder": lambda x: x, "depth": lambda f, grids: outs}
        xs = torch.linspace(-1, 1, W)[None, None, None, :].expand(B, 1, H, W)
        ys = torch.linspace(-1, 1, H)[None, None, :, None].expand(B, 1, H, W)
        inputs = {("color_aug", "l"): torch.rand(B, 3, H, W).cuda(), "grid": torch.cat([xs, ys], 1).contiguous().cuda()}
        with torch.no_grad():
            res["reference_gpu_ms"] = timeit(lambda: tr.Trainer.generate_post_process_disp(t, inputs), n=5, w=2)
            want = tr.Trainer.generate_post_process_disp(t, inputs)
        got = occlusion_masks(outs["logits"], outs["probability"], outs["disp_layered"], outs["disp"])
        res["max_err_disp_pp"] = float((got[0] - want[0]).abs().max())
        res["max_err_mask_novel"] = float((got[1] - want[1]).abs().max())
        res["speedup_vs_reference_gpu"] = res["reference_gpu_ms"] / res["ours_ms"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
