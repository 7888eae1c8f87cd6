#!/usr/bin/env python
"""disp_warp with a per-pixel (x-varying) disp_layered -- yz planes (depth_decoder.py:209-236) / PladeNet-style geometry
(missing #8): pd_warp_composite_fwd/bwd (general kernels) against the reference's op chain
(trainer.py:540-603: grid, F.grid_sample of the stacked tensor, softmax, compositing) in eager PyTorch, forward + backward.
usage: dense_bench.py B N H W"""
import json, os, sys
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planedepth_b200 import _lib as L
from planedepth_b200.functional import WarpConfig, warp_composite


def eager(src, logits, disp, sign):
    """trainer.py:540-603 for one side, no mixture, as the reference issues it."""
    B, N, H, W = logits.shape
    gx, gy = torch.meshgrid(torch.linspace(-1, 1, W, device=src.device), torch.linspace(-1, 1, H, device=src.device), indexing="xy")
    grid = torch.stack([gx, gy], -1)[None, None].expand(B, N, H, W, 2)
    off = torch.zeros_like(grid)
    off = torch.stack([sign * disp / (W - 1) * 2.0, torch.zeros_like(disp)], -1)
    g = (grid + off).reshape(B * N, H, W, 2)
    feat = torch.cat([src[:, None].expand(-1, N, -1, -1, -1), logits[:, :, None]], 2).reshape(B * N, 4, H, W)
    rec = F.grid_sample(feat, g, mode="bilinear", padding_mode="zeros", align_corners=True).reshape(B, N, 4, H, W)
    p = torch.softmax(rec[:, :, 3], 1)
    return (rec[:, :, :3] * p[:, :, None]).sum(1)


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
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    src = torch.rand(B, 3, H, W, device=dev, generator=g)
    logits = torch.randn(B, N, H, W, device=dev, generator=g, requires_grad=True)
    base = 300.0 * (2.0 / 300.0) ** (torch.arange(N, device=dev) / (N - 1.0))
    x = torch.linspace(0.8, 1.2, W, device=dev)
    disp = (base[None, :, None, None] * x[None, None, None, :]).expand(B, N, H, W).contiguous().requires_grad_(True)
    up = torch.randn(B, 3, H, W, device=dev, generator=g)
    cfg = WarpConfig(L.PD_WARP_DISP, False, False, disp_sign=-1.0, shape=(B, N, H, W))
    mask = None

    def ours(with_disp):
        rec = warp_composite(cfg, src, None, logits, None, disp if with_disp else disp.detach(), mask)[0]
        torch.autograd.grad([rec], [logits] + ([disp] if with_disp else []), [up])

    def ref(with_disp):
        rec = eager(src, logits, disp if with_disp else disp.detach(), -1.0)
        torch.autograd.grad([rec], [logits] + ([disp] if with_disp else []), [up])

    res = {"shape": [B, N, H, W]}
    rec_o = warp_composite(cfg, src, None, logits, None, disp.detach(), mask)[0]
    rec_r = eager(src, logits, disp.detach(), -1.0)
    res["max_abs_diff_rgb_rec"] = float((rec_o - rec_r).abs().max())
    res["ours_fwd_bwd_ms"] = timeit(lambda: ours(False))
    res["ours_fwd_bwd_with_disp_grad_ms"] = timeit(lambda: ours(True))
    res["eager_torch_fwd_bwd_ms"] = timeit(lambda: ref(False), n=3, w=1)
    res["eager_torch_fwd_bwd_with_disp_grad_ms"] = timeit(lambda: ref(True), n=3, w=1)
    res["speedup"] = res["eager_torch_fwd_bwd_ms"] / res["ours_fwd_bwd_ms"]
    res["img_per_s"] = B / res["ours_fwd_bwd_ms"] * 1e3
    xb = B * N * H * W * 4
    res["algorithmic_bytes"] = xb * 2 + xb * 3  # fwd: logits + disparity; bwd: logits + disparity + g_logits
    res["ours_GBps"] = res["algorithmic_bytes"] / res["ours_fwd_bwd_ms"] / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
