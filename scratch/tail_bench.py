#!/usr/bin/env python
"""Decoder tail (depth_decoder.py:258-291) on the GPU: pd_plane_tail_fwd/bwd against the same formulas in eager PyTorch
(what the reference's decoder executes), forward + backward.  usage: tail_bench.py B N H W mixture(0/1)"""
import json, os, sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planedepth_b200.boundary import decoder_tail


def eager_tail(logits_raw, sigma_raw, padding_mask, disp_layered, mixture):
    """The operations of networks/depth_decoder.py:258-291 as the reference's decoder issues them (eager PyTorch baseline)."""
    W = logits_raw.shape[-1]
    out = {"logits": logits_raw * padding_mask}
    out["probability"] = torch.softmax(out["logits"], 1)
    if mixture:
        sigma = torch.clamp(torch.sigmoid(sigma_raw), 0.01, 1.0)
        out["sigma"] = sigma
        weights = out["probability"] / sigma
        weights = weights * padding_mask
        out["probability"] = weights / weights.sum(1, True)
    out["disp"] = (out["probability"] * disp_layered).sum(1, True)
    out["depth"] = 0.1 * 0.58 * W / out["disp"]
    return out


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
    B, N, H, W, mix = [int(v) for v in sys.argv[1:6]] if len(sys.argv) >= 6 else (12, 49, 192, 640, 0)
    mix = bool(mix)
    lr = torch.randn(B, N, H, W, device="cuda", requires_grad=True)
    sr = torch.randn(B, N, H, W, device="cuda", requires_grad=True) if mix else None
    base = (300.0 * (2.0 / 300.0) ** (torch.arange(N, device="cuda") / (N - 1.0))).reshape(1, N, 1, 1).repeat(B, 1, 1, 1).requires_grad_(True)
    mask = torch.ones(B, N, H, W, device="cuda")
    gl, gs, gd = torch.randn(B, N, H, W, device="cuda"), torch.randn(B, N, H, W, device="cuda"), torch.randn(B, 1, H, W, device="cuda")

    def step(fn):
        out = fn(lr, sr, mask, base.expand(B, N, H, W), mix)
        outs = [out["logits"], out["disp"]] + ([out["sigma"]] if mix else [])
        grads = [gl, gd] + ([gs] if mix else [])
        torch.autograd.grad(outs, [lr, base] + ([sr] if mix else []), grads)

    res = {"shape": [B, N, H, W], "mixture": mix}
    res["ours_fwd_bwd_ms"] = timeit(lambda: step(decoder_tail))
    res["eager_torch_fwd_bwd_ms"] = timeit(lambda: step(eager_tail), n=5, w=2)
    res["speedup"] = res["eager_torch_fwd_bwd_ms"] / res["ours_fwd_bwd_ms"]
    x = B * N * H * W * 4
    m = 1 if mix else 0
    res["algorithmic_bytes"] = x * ((1 + m) + (2 + m) + (1 + m) * 2 + (1 + m))  # fwd: read raw(,sraw) write logits, prob(, sigma); bwd: read saved + upstream, write grads
    res["ours_GBps"] = res["algorithmic_bytes"] / res["ours_fwd_bwd_ms"] / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
