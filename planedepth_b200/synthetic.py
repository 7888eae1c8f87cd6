"""Synthetic inputs for the hot path in the reference's dict layout.

Everything is generated on the CPU from a seeded generator and then moved, so a CPU reference run and
a GPU run see identical data.  ``layout="reference"`` reproduces what the reference DepthDecoder hands
over (depth_decoder.py:147-260): dense fp32 ``logits``/``sigma``, ``disp_layered`` as a stride-0
expand of a [B,N,1,1] tensor (vertical planes) or a dense cat (with xz planes), dense fp32
``padding_mask``.  ``layout="compact"`` hands the mask over as a stride-0 expand as well (what a fused
decoder tail would emit)."""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Sequence

import torch


def make_opt(**over):
    o = dict(warp_type="disp_warp", match_aug=False, use_mixture_loss=False, render_probability=False, automask=False,
             alpha_self=0.0, self_distillation=0.0, alpha_pc=0.1, alpha_smooth=0.04, gamma_smooth=2, use_ssim=False,
             novel_frame_ids=[], no_stereo=False, use_colmap=True, plane_residual=False,
             disp_levels=49, disp_min=2.0, disp_max=300.0, xz_levels=0)
    o.update(over)
    return SimpleNamespace(**o)


def intrinsics(B, H, W):
    K = torch.tensor([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32)
    K = K[None].repeat(B, 1, 1)
    return K, torch.linalg.pinv(K)


def mono_pose(B, f, gen):
    """Small rigid motion of the PoseDecoder's scale (pose_net.py:150: 0.01 * net output)."""
    aa = 0.01 * torch.randn(B, 3, generator=gen)
    tr = 0.01 * torch.randn(B, 3, generator=gen) + torch.tensor([0.0, 0.0, 0.05 * f])
    th = aa.norm(dim=1, keepdim=True).clamp_min(1e-8)
    k = aa / th
    Kx = torch.zeros(B, 3, 3)
    Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
    Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
    R = torch.eye(3)[None] + torch.sin(th)[:, :, None] * Kx + (1 - torch.cos(th))[:, :, None] * (Kx @ Kx)
    T = torch.eye(4)[None].repeat(B, 1, 1)
    T[:, :3, :3] = R
    T[:, :3, 3] = tr
    return T


def make_batch(B: int, H: int, W: int, opt, seed: int = 0, device="cpu", layout: str = "reference",
               requires_grad: bool = True, mask_novel: bool = False):
    """Returns SimpleNamespace(inputs, outputs, leaves, target_sides)."""
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.rand(*s, generator=g)
    frames: List = list(opt.novel_frame_ids)
    n_v, n_xz = opt.disp_levels, opt.xz_levels
    N = n_v + n_xz
    dev = torch.device(device)
    inputs = {}
    for s in ["l", "r"] + frames:
        inputs[("color", s)] = rnd(B, 3, H, W).to(dev)
        inputs[("color_aug", s)] = inputs[("color", s)]
    K, iK = intrinsics(B, H, W)
    inputs["K"], inputs["inv_K"] = K.to(dev), iK.to(dev)
    for s, tx in (("l", 0.1), ("r", -0.1)):
        T = torch.eye(4)[None].repeat(B, 1, 1)
        T[:, 0, 3] = tx
        inputs[("Rt", s)] = T.to(dev)
    gridx, gridy = torch.meshgrid(torch.linspace(-1, 1, W), torch.linspace(-1, 1, H), indexing="xy")
    inputs["grid"] = torch.stack([gridx, gridy], 0)[None].expand(B, -1, -1, -1).contiguous().to(dev)
    # plane geometry, depth_decoder.py:148-183 ------------------------------------------------------
    leaves = {}
    lev = torch.arange(n_v, dtype=torch.float32)[None, :, None, None].expand(B, -1, -1, -1)
    if opt.plane_residual:
        lev = lev + (rnd(B, n_v, 1, 1) - 0.5)
    base = (opt.disp_max * (opt.disp_min / opt.disp_max) ** (lev / (n_v - 1))).contiguous().to(dev)
    if opt.plane_residual and requires_grad:
        base.requires_grad_(True)
        leaves["disp_base"] = base
    # Derived tensors are built from the detached leaf here; `attach()` (below) installs the differentiable
    # views inside the step, the way the decoder re-derives them every step.  (A view made here would
    # create the leaf's gradient accumulator on the legacy default stream, which a later CUDA-graph capture
    # of the step is not allowed to synchronise with.)
    based = base.detach()
    distance = 0.1 * 0.58 * W / based[:, :, 0, 0]
    norm = torch.tensor([0.0, 0.0, 1.0], device=dev)[None, None].expand(B, n_v, 3)
    disp_layered = based.expand(-1, -1, H, W)
    ones = torch.ones(B, n_v, 1, 1, device=dev)
    padding_mask = ones.expand(-1, -1, H, W) if layout == "compact" else torch.ones(B, n_v, H, W, device=dev)
    if n_xz:
        gy = inputs["grid"][:, 1:, :, :]
        hl = torch.arange(n_xz, dtype=torch.float32)[None, :, None, None].expand(B, -1, -1, -1)
        if opt.plane_residual:
            hl = hl + (rnd(B, n_xz, 1, 1) - 0.5)
        h = (0.1852 + (0.3704 - 0.1852) * hl / (n_xz - 1)).to(dev)
        if opt.plane_residual and requires_grad:
            h.requires_grad_(True)
            leaves["xz_h"] = h
        if layout == "compact":
            # one value per (image, plane, row), expanded along x with a zero stride
            gyc = gy[..., :1]
            xz_mask = (gyc >= 1e-7).expand(-1, n_xz, -1, -1).float()
            Z = h.detach().expand(-1, -1, H, 1) * 1.92 / (gyc.clamp_min(1e-7) / 2.0)
            disp_layered = torch.cat([based.expand(-1, -1, H, 1), 0.1 * 0.58 * W / Z], 1).expand(-1, -1, -1, W)
            padding_mask = torch.cat([ones.expand(-1, -1, H, 1), xz_mask], 1).expand(-1, -1, -1, W)
        else:
            xz_mask = (gy >= 1e-7).expand(-1, n_xz, -1, -1)
            Z = h.detach().expand(-1, -1, H, W) * 1.92 / (gy.clamp_min(1e-7) / 2.0)
            disp_layered = torch.cat([disp_layered, 0.1 * 0.58 * W / Z], 1)
            padding_mask = torch.cat([padding_mask, xz_mask], 1)
        norm = torch.cat([norm, torch.tensor([0.0, 1.0, 0.0], device=dev)[None, None].expand(B, n_xz, 3)], 1)
        distance = torch.cat([distance, h.detach()[:, :, 0, 0]], 1)
    logits = (torch.randn(B, N, H, W, generator=g).to(dev) * padding_mask).contiguous()
    logits.requires_grad_(requires_grad)
    leaves["logits"] = logits
    outputs = {"logits": logits, "disp_layered": disp_layered, "padding_mask": padding_mask, "distance": distance, "norm": norm,
               "probability": logits.detach()}
    if opt.use_mixture_loss:
        sigma = torch.sigmoid(torch.randn(B, N, H, W, generator=g)).clamp(0.01, 1.0).to(dev)
        sigma.requires_grad_(requires_grad)
        outputs["sigma"] = sigma
        leaves["sigma"] = sigma
    if mask_novel:
        outputs["mask_novel"] = rnd(B, 1, H, W).to(dev)
    outputs[("Rt", "r")] = inputs[("Rt", "r")]
    for f in frames:
        T = mono_pose(B, f, g).to(dev)
        inputs[("Rt", f)] = T
        outputs[("Rt", f)] = T
    target_sides = ([] if opt.no_stereo else ["r"]) + frames

    def attach(out):
        """Install the differentiable plane-geometry views (depth_decoder.py:153-183) into `out`."""
        if "disp_base" not in leaves:
            return out
        b = leaves["disp_base"]
        dl = b.expand(-1, -1, H, W)
        dist = 0.1 * 0.58 * W / b[:, :, 0, 0]
        if n_xz:
            hh = leaves["xz_h"]
            if layout == "compact":
                gyc = inputs["grid"][:, 1:, :, :1]
                Zg = hh.expand(-1, -1, H, 1) * 1.92 / (gyc.clamp_min(1e-7) / 2.0)
                dl = torch.cat([b.expand(-1, -1, H, 1), 0.1 * 0.58 * W / Zg], 1).expand(-1, -1, -1, W)
            else:
                Zg = hh.expand(-1, -1, H, W) * 1.92 / (inputs["grid"][:, 1:, :, :].clamp_min(1e-7) / 2.0)
                dl = torch.cat([dl, 0.1 * 0.58 * W / Zg], 1)
            dist = torch.cat([dist, hh[:, :, 0, 0]], 1)
        out["disp_layered"], out["distance"] = dl, dist
        return out

    return SimpleNamespace(inputs=inputs, outputs=outputs, leaves=leaves, target_sides=target_sides, shape=(B, N, H, W), attach=attach)
