"""CPU oracle for the PlaneDepth photometric-reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``planedepth_b200/`` imports this file; the only
callers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.  The product path is the CUDA library behind ``include/planedepth_b200.h`` and
fails loudly when that library is missing.

What it is: a plain fp32 PyTorch-on-CPU restatement (explicit gathers, no ``F.grid_sample``) of

* ``Trainer.pred_novel_images``      /root/reference/trainer.py:523-603
* ``Trainer.compute_losses``         /root/reference/trainer.py:701-773
* ``Trainer.compute_reprojection_loss``  trainer.py:687-699
* ``HomographyWarp.forward``         /root/reference/layers.py:206-234
* ``BackprojectDepth`` / ``Project3D``   layers.py:150-156 / 169-182
* ``SSIM.forward``                   layers.py:292-306
* ``laplacian`` / ``multimodal_loss``    layers.py:454-455 / 465-466
* ``get_smooth_loss_disp``           layers.py:243-256
* the bilinear sampler the reference borrows from torch (third-party, not under /root/reference):
  ``torch.nn.functional.grid_sample(..., padding_mode="zeros", align_corners=True)`` as pinned by the
  installed wheel torch 2.11.0 (``ATen/native/GridSampler.h`` ``grid_sampler_unnormalize`` /
  ``grid_sampler_compute_source_index``; the reference README pins pytorch==1.10.0, same semantics).

Parity pin: the reference ships **no** tests, golden vectors or fixtures for this path (SURVEY.md §4,
§8c), so the oracle is pinned against outputs of the reference itself, executed in the build container
through an import shim: ``tests/golden/make_golden.py`` (committed) calls the *unbound reference
methods* on seeded inputs and stores inputs + outputs + gradients in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them through this file.  Gradients come from autograd over
these explicit formulas, which is the same chain rule ATen's ``grid_sampler_2d_backward`` implements.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Callable, Dict, Optional, Sequence

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# coordinates
# --------------------------------------------------------------------------------------------


def normalise(p: torch.Tensor, size: int) -> torch.Tensor:
    """pixel -> [-1,1]; the reference's in-place ``/= (size-1)`` then ``(p - 0.5) * 2``
    (trainer.py:549-551, layers.py:179-181, layers.py:231-233)."""
    return (p / (size - 1) - 0.5) * 2


def unnormalise(g: torch.Tensor, size: int) -> torch.Tensor:
    """[-1,1] -> pixel, ``align_corners=True`` (ATen grid_sampler_unnormalize)."""
    return ((g + 1) / 2) * (size - 1)


def pixel_centres(H: int, W: int, device=None):
    ys, xs = torch.meshgrid(
        torch.arange(H, dtype=torch.float32, device=device),
        torch.arange(W, dtype=torch.float32, device=device),
        indexing="ij",
    )
    return xs, ys


def disp_warp_coords(disp_layered: torch.Tensor, target_side):
    """trainer.py:540-554.  ``u = x + D`` for target 'r', ``x - D`` for 'l', unchanged otherwise;
    ``v = y``.  Returns source-pixel coordinates [B,N,H,W] each (before normalisation)."""
    B, N, H, W = disp_layered.shape
    xs, ys = pixel_centres(H, W, disp_layered.device)
    xs = xs.expand(B, N, H, W)
    ys = ys.expand(B, N, H, W)
    if target_side == "r":
        u = xs + disp_layered
    elif target_side == "l":
        u = xs - disp_layered
    else:
        u = xs.clone()
    return u, ys


def homography_matrices(distance, norm, T, K, inv_K):
    """layers.py:211-220.  distance [B,N], norm [B,N,3], T/K/inv_K [B,4,4] (per image; the
    reference expands them to B*N).  Returns H_t2s [B,N,3,3]."""
    B, N = distance.shape
    R = T[:, None, :3, :3]
    t = T[:, None, :3, 3:4]
    n = norm.to(torch.float32).reshape(B, N, 1, 3)
    Rtnd = R + torch.matmul(t, n) / distance.reshape(B, N, 1, 1)
    K3 = K[:, None, :3, :3]
    iK3 = inv_K[:, None, :3, :3]
    H_s2t = torch.matmul(K3, torch.matmul(Rtnd, iK3))
    return torch.inverse(H_s2t)


def homography_coords(distance, norm, T, K, inv_K, H: int, W: int):
    """layers.py:206-234 up to (not including) the normalisation.  Returns u, v [B,N,H,W] and the
    validity mask [B,N,H,W] (bool)."""
    B, N = distance.shape
    Ht2s = homography_matrices(distance, norm, T, K, inv_K)  # [B,N,3,3]
    xs, ys = pixel_centres(H, W, distance.device)
    hom = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, device=xs.device)], 0)
    q = torch.matmul(Ht2s, hom)  # [B,N,3,HW]
    rays = torch.matmul(inv_K[:, None, :3, :3], hom)  # [B,1,3,HW]
    Rn = torch.matmul(T[:, None, :3, :3], norm.to(torch.float32).reshape(B, N, 3, 1))  # [B,N,3,1]
    facing = (rays * Rn).sum(2) > 0.0
    z = q[:, :, 2]
    mask = facing & (z > 1e-7)
    z = torch.where(z < 1e-7, torch.full_like(z, 1e-7), z)
    u = q[:, :, 0] / z
    v = q[:, :, 1] / z
    return u.reshape(B, N, H, W), v.reshape(B, N, H, W), mask.reshape(B, N, H, W)


def depth_warp_coords(disp_layered, T, K, inv_K, eps: float = 1e-7):
    """trainer.py:533-538 with layers.py:150-156 and 169-182 (before normalisation)."""
    B, N, H, W = disp_layered.shape
    depth = 0.1 * 0.58 * W / disp_layered
    xs, ys = pixel_centres(H, W, disp_layered.device)
    hom = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, device=xs.device)], 0)
    rays = torch.matmul(inv_K[:, None, :3, :3], hom)  # [B,1,3,HW]
    pts = depth.reshape(B, N, 1, H * W) * rays  # [B,N,3,HW]
    pts = torch.cat([pts, torch.ones(B, N, 1, H * W, device=pts.device)], 2)
    P = torch.matmul(K, T)[:, None, :3, :]  # [B,1,3,4]
    cam = torch.matmul(P, pts)  # [B,N,3,HW]
    u = cam[:, :, 0] / (cam[:, :, 2] + eps)
    v = cam[:, :, 1] / (cam[:, :, 2] + eps)
    return u.reshape(B, N, H, W), v.reshape(B, N, H, W)


# --------------------------------------------------------------------------------------------
# bilinear sampler (zeros padding, align_corners=True) — explicit restatement of ATen
# --------------------------------------------------------------------------------------------


def bilinear_sample(feat: torch.Tensor, gx: torch.Tensor, gy: torch.Tensor, padding: str = "zeros"):
    """feat [M,C,H,W]; gx, gy [M,H,W] normalised coordinates.  Returns [M,C,H,W].

    Follows ATen ``grid_sampler_2d`` (bilinear): unnormalise, floor, weights
    ``nw=(x1-x)(y1-y) ne=(x-x0)(y1-y) sw=(x1-x)(y-y0) se=(x-x0)(y-y0)``; a tap outside the image
    contributes 0 ("zeros") — or coordinates are clipped first ("border", trainer.py:624-628)."""
    M, C, H, W = feat.shape
    x = unnormalise(gx, W)
    y = unnormalise(gy, H)
    if padding == "border":
        x = x.clamp(0, W - 1)
        y = y.clamp(0, H - 1)
    x0 = torch.floor(x)
    y0 = torch.floor(y)
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - x) * (y1 - y)
    w_ne = (x - x0) * (y1 - y)
    w_sw = (x1 - x) * (y - y0)
    w_se = (x - x0) * (y - y0)
    flat = feat.reshape(M, C, H * W)

    def tap(xi, yi, w):
        inside = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
        xi_c = xi.clamp(0, W - 1).to(torch.int64)
        yi_c = yi.clamp(0, H - 1).to(torch.int64)
        idx = (yi_c * W + xi_c).reshape(M, 1, H * W).expand(-1, C, -1)
        val = torch.gather(flat, 2, idx).reshape(M, C, *xi.shape[1:])
        return val * (w * inside.to(w.dtype))[:, None]

    return tap(x0, y0, w_nw) + tap(x1, y0, w_ne) + tap(x0, y1, w_sw) + tap(x1, y1, w_se)


# --------------------------------------------------------------------------------------------
# pred_novel_images
# --------------------------------------------------------------------------------------------


def _opt(opt, name, default):
    return getattr(opt, name, default)


def novel_view(opt, src, logits, sigma, u, v, mask):
    """trainer.py:567-603 for one target side, given source coordinates (u, v) [B,N,H,W] and the
    validity mask [B,N,H,W] (float or bool).  Returns a dict with the reference's tensors."""
    B, N, H, W = logits.shape
    gx = normalise(u, W).reshape(B * N, H, W)
    gy = normalise(v, H).reshape(B * N, H, W)
    chans = [src[:, None].expand(-1, N, -1, -1, -1).reshape(B * N, 3, H, W), logits.reshape(B * N, 1, H, W)]
    if _opt(opt, "use_mixture_loss", False):
        chans.append(sigma.reshape(B * N, 1, H, W))
    feats = torch.cat(chans, 1)
    rec = bilinear_sample(feats, gx, gy).reshape(B, N, -1, H, W)
    rec = rec * mask.reshape(B, N, 1, H, W).to(rec.dtype)
    out = {}
    out["rgb_rec_layered"] = rec[:, :, :3]
    out["logit_rec"] = rec[:, :, 3]
    prob = torch.softmax(out["logit_rec"], dim=1)
    if _opt(opt, "use_mixture_loss", False):
        sigma_rec = rec[:, :, 4].clamp(0.01, 1.0)
        out["sigma_rec"] = sigma_rec
        out["pi_rec"] = prob
        wts = prob / sigma_rec
        prob = wts / wts.sum(1, keepdim=True)
    out["probability_rec"] = prob
    out["rgb_rec"] = (out["rgb_rec_layered"] * prob[:, :, None]).sum(1)
    return out


def pred_novel_images(opt, target_sides: Sequence, inputs: Dict, outputs: Dict) -> None:
    """Restatement of Trainer.pred_novel_images (trainer.py:523-603); mutates ``outputs``.

    Deviation from upstream (documented defect D1): ``depth_warp`` uses
    ``outputs["padding_mask"]`` where the reference dereferences an unbound local."""
    B, N, H, W = outputs["probability"].shape
    color = "color_aug" if _opt(opt, "match_aug", False) else "color"
    src = inputs[(color, "l")]
    for s in target_sides:
        wt = _opt(opt, "warp_type", "disp_warp")
        if wt == "disp_warp":
            u, v = disp_warp_coords(outputs["disp_layered"], s)
            mask = outputs["padding_mask"]
        elif wt == "homography_warp":
            u, v, mask = homography_coords(
                outputs["distance"], outputs["norm"], outputs[("Rt", s)], inputs["K"], inputs["inv_K"], H, W
            )
        elif wt == "depth_warp":
            u, v = depth_warp_coords(outputs["disp_layered"], inputs[("Rt", s)], inputs["K"], inputs["inv_K"])
            mask = outputs["padding_mask"]
        else:
            raise ValueError(wt)
        res = novel_view(opt, src, outputs["logits"], outputs.get("sigma"), u, v, mask)
        for k, val in res.items():
            outputs[(k, s)] = val


# --------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------


def laplacian_mixture_nll(err, sigma, pi):
    """layers.py:454-455 + 465-466 with dist='lap'.  err/sigma/pi [B,N,H,W] -> [B,1,H,W]."""
    dens = 0.5 * torch.exp(-(err.abs() / sigma)) / sigma
    return -torch.log((pi * dens).sum(1, keepdim=True) + 1e-7)


def ssim_map(x, y):
    """layers.py:276-306: reflect-pad 1, 3x3 box means, biased variances, clamp((1-n/d)/2,0,1)."""
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    yp = F.pad(y, (1, 1, 1, 1), mode="reflect")

    def box(t):
        return F.avg_pool2d(t, 3, 1)

    mx, my = box(xp), box(yp)
    sx = box(xp * xp) - mx * mx
    sy = box(yp * yp) - my * my
    sxy = box(xp * yp) - mx * my
    num = (2 * mx * my + C1) * (2 * sxy + C2)
    den = (mx * mx + my * my + C1) * (sx + sy + C2)
    return torch.clamp((1 - num / den) / 2, 0, 1)


def reprojection_loss(pred, target, use_ssim: bool):
    """trainer.py:687-699."""
    l1 = (target - pred).abs().mean(1, keepdim=True)
    if not use_ssim:
        return l1
    return 0.85 * ssim_map(pred, target).mean(1, keepdim=True) + 0.15 * l1


def smooth_loss_disp(disp, img, gamma: float = 1.0):
    """layers.py:243-256."""
    dx = (disp[..., :, :-1] - disp[..., :, 1:]).abs()
    dy = (disp[..., :-1, :] - disp[..., 1:, :]).abs()
    ix = (img[..., :, :-1] - img[..., :, 1:]).abs().mean(1, keepdim=True)
    iy = (img[..., :-1, :] - img[..., 1:, :]).abs().mean(1, keepdim=True)
    return (dx * torch.exp(-gamma * ix)).mean() + (dy * torch.exp(-gamma * iy)).mean()


def perceptual_loss(pc_net: Callable, pred, target, source=None):
    """trainer.py:672-685 (the feature network itself is out of scope; any callable returning three
    feature maps)."""
    fp, ft = pc_net(pred), pc_net(target)
    fs = pc_net(source) if source is not None else None
    total = 0
    for i in range(3):
        lp = ((fp[i] - ft[i]) ** 2).mean(1, keepdim=True)
        if fs is not None:
            la = ((fs[i] - ft[i]) ** 2).mean(1, keepdim=True)
            lp = torch.minimum(lp, la)
        total = total + lp.mean()
    return total


def photometric_map(opt, inputs, outputs, s, loss_mode: Optional[str] = None):
    """Per-pixel photometric term of trainer.py:717-742 for target side ``s`` -> (ph [B,1,H,W], pred).

    ``loss_mode``: None = what the reference does (mixture NLL if opt.use_mixture_loss else L1);
    "ssim_l1" = BASELINE.json north_star: ``compute_reprojection_loss`` with use_ssim (trainer.py:687-699)
    applied to (pred, target), automask against the identity reprojection of the source."""
    color = "color_aug" if _opt(opt, "match_aug", False) else "color"
    target = inputs[(color, s)]
    src = inputs[(color, "l")]
    pred = outputs[("rgb_rec", s)]
    m = outputs.get("mask_novel")
    if m is not None:
        pred = pred * m + target * (1.0 - m)
    if loss_mode is None:
        loss_mode = "mixture" if _opt(opt, "use_mixture_loss", False) else "l1"
    if loss_mode == "mixture":
        err = (outputs[("rgb_rec_layered", s)] - target[:, None]).abs().mean(2)
        ph = laplacian_mixture_nll(err, outputs[("sigma_rec", s)], outputs[("pi_rec", s)])
        if _opt(opt, "automask", False):
            err_a = (src[:, None] - target[:, None]).abs().mean(2)
            ph_a = laplacian_mixture_nll(err_a, outputs[("sigma_rec", s)].detach(), outputs[("pi_rec", s)].detach())
            ph = torch.minimum(ph, ph_a)
        if m is not None:
            ph = ph * m
    elif loss_mode == "l1":
        ph = (pred - target).abs().mean(1, keepdim=True)
        if _opt(opt, "automask", False):
            ph = torch.minimum(ph, (src - target).abs().mean(1, keepdim=True))
    elif loss_mode == "ssim_l1":
        ph = reprojection_loss(pred, target, True)
        if _opt(opt, "automask", False):
            ph = torch.minimum(ph, reprojection_loss(src, target, True))
    else:
        raise ValueError(loss_mode)
    return ph, pred


def compute_losses(opt, target_sides, inputs, outputs, pc_net: Optional[Callable], loss_mode: Optional[str] = None):
    """Restatement of Trainer.compute_losses (trainer.py:701-773).  The ``alpha_self`` branch is
    unreachable upstream (defect D2) and is omitted."""
    B, N, H, W = outputs["probability"].shape
    color = "color_aug" if _opt(opt, "match_aug", False) else "color"
    losses = {"loss/ph_loss": 0, "loss/pc_loss": 0, "loss/total_loss": 0}
    for s in target_sides:
        ph, pred = photometric_map(opt, inputs, outputs, s, loss_mode)
        ph = ph.mean()
        total = ph
        losses["loss/ph_loss"] = losses["loss/ph_loss"] + ph
        if pc_net is not None:
            src = inputs[(color, "l")] if _opt(opt, "automask", False) else None
            pc = perceptual_loss(pc_net, pred, inputs[(color, s)], src)
            losses["loss/pc_loss"] = losses["loss/pc_loss"] + pc
            total = total + _opt(opt, "alpha_pc", 0.1) * pc
        if _opt(opt, "self_distillation", 0.0) > 0:
            dl = (outputs["disp"] - outputs["disp_pp"]).abs().mean()
            losses["loss/disp_loss"] = dl
            total = total + opt.self_distillation * dl
        losses["loss/total_loss"] = losses["loss/total_loss"] + total
    # trainer.py:765-766 divides every entry present at that point (incl. disp_loss) by #targets
    for k in list(losses.keys()):
        losses[k] = losses[k] / len(target_sides)
    x0 = int(0.2 * W)
    sm = smooth_loss_disp(outputs["disp"][..., x0:], inputs[("color", "l")][..., x0:], _opt(opt, "gamma_smooth", 2))
    losses["loss/smooth_loss"] = sm
    losses["loss/total_loss"] = losses["loss/total_loss"] + _opt(opt, "alpha_smooth", 0.04) * sm
    return losses


# --------------------------------------------------------------------------------------------
# convenience: one call = what bench.py's cpu arm and the parity tests time / compare
# --------------------------------------------------------------------------------------------


def default_opt(**over):
    o = dict(
        warp_type="disp_warp", match_aug=False, use_mixture_loss=False, automask=False, alpha_pc=0.1,
        alpha_smooth=0.04, gamma_smooth=2, self_distillation=0.0, use_ssim=False,
    )
    o.update(over)
    return SimpleNamespace(**o)


def hot_path(opt, target_sides, inputs, outputs, pc_net=None, loss_mode=None):
    """pred_novel_images + compute_losses, returning the losses dict (outputs is mutated)."""
    pred_novel_images(opt, target_sides, inputs, outputs)
    return compute_losses(opt, target_sides, inputs, outputs, pc_net, loss_mode)


# --------------------------------------------------------------------------------------------
# generate_post_process_disp: occlusion masks from the flipped pass
# --------------------------------------------------------------------------------------------


def _shift_sample(feat, disp_layered, sign):
    """F.grid_sample(feat.reshape(B*N,1,H,W), grid(x + sign*disp, y)) of trainer.py:424-445, per plane."""
    B, N, H, W = feat.shape
    xs, ys = pixel_centres(H, W, feat.device)
    u = xs.expand(B, N, H, W) + sign * disp_layered
    v = ys.expand(B, N, H, W)
    gx = normalise(u, W).reshape(B * N, H, W)
    gy = normalise(v, H).reshape(B * N, H, W)
    return bilinear_sample(feat.reshape(B * N, 1, H, W), gx, gy).reshape(B, N, H, W)


def post_process_disp(outputs: Dict):
    """trainer.py:421-466 — everything after the flipped forward pass of ``generate_post_process_disp``.

    ``outputs`` are the decoder outputs for the batch ``cat([img, img.flip(-1)])`` (2B images):
    ``probability``, ``logits``, ``disp_layered`` [2B,N,H,W] and ``disp`` [2B,1,H,W].  Returns
    ``(disp_pp, mask_novel)`` [B,1,H,W] each (detached), plus the two occlusion maps for tests."""
    B2, N, H, W = outputs["probability"].shape
    B = B2 // 2
    dl, df = outputs["disp_layered"][:B], outputs["disp_layered"][B:]
    # trainer.py:441-447: left logits -> right view, softmax over planes, back to the left view, sum, clip at 1
    plr = torch.softmax(_shift_sample(outputs["logits"][:B], dl, +1.0), 1)
    o_l = _shift_sample(plr, df, -1.0).sum(1, True).clamp(max=1.0)
    # trainer.py:449-454: the flipped half, un-flipped, the other way round
    pfrl = torch.softmax(_shift_sample(outputs["logits"][B:].flip(-1), df, -1.0), 1)
    o_fr = _shift_sample(pfrl, dl, +1.0).sum(1, True).clamp(max=1.0)
    d_l, d_f = outputs["disp"][:B], outputs["disp"][B:].flip(-1)
    mean_disp = d_l * 0.5 + d_f * 0.5  # :456
    disp_pp = mean_disp * o_fr + d_l * (1 - o_fr)  # :458
    disp_pp = disp_pp * o_l + d_f * (1 - o_l)  # :459
    mask_novel = _shift_sample(outputs["probability"][:B], dl, +1.0).sum(1, True).clamp(max=1.0)  # :461-463
    return disp_pp.detach(), mask_novel.detach(), o_l.detach(), o_fr.detach()


# --------------------------------------------------------------------------------------------
# decoder tail: depth_decoder.py:258-291 after the dispconv / sigmaconv convolutions
# --------------------------------------------------------------------------------------------


def decoder_tail(logits_raw, sigma_raw, padding_mask, disp_layered, mixture: bool):
    """``logits_raw`` / ``sigma_raw`` = outputs of ``convs["dispconv"]`` / ``convs["sigmaconv"]`` [B,N,H,W].
    Returns the dict entries the decoder writes: logits, probability, [sigma, pi], disp, depth (``render_probability`` off)."""
    W = logits_raw.shape[-1]
    out = {}
    out["logits"] = logits_raw * padding_mask  # :259
    out["probability"] = torch.softmax(out["logits"], 1)  # :276
    if mixture:
        sigma = torch.clamp(torch.sigmoid(sigma_raw), 0.01, 1.0)  # :279-280
        out["sigma"] = sigma
        out["pi"] = pi = out["probability"]
        weights = pi / sigma
        weights = weights * padding_mask
        weights = weights / weights.sum(1, True)  # :282-284
        out["probability"] = weights
    out["disp"] = (out["probability"] * disp_layered).sum(1, True)  # :288
    out["depth"] = 0.1 * 0.58 * W / out["disp"]  # :290
    return out


# --------------------------------------------------------------------------------------------
# input staging: the loader's float resize, datasets/pair_transforms.py:28-48, 63-78
# --------------------------------------------------------------------------------------------


def _cubic_weights(t):
    """ATen cubic_convolution1 / cubic_convolution2 with A = -0.75 (UpSample.h get_cubic_upsample_coefficients)."""
    A = -0.75
    x0, x2, x3 = t + 1.0, 1.0 - t, 2.0 - t
    w0 = ((A * x0 - 5.0 * A) * x0 + 8.0 * A) * x0 - 4.0 * A
    w1 = ((A + 2.0) * t - (A + 3.0)) * t * t + 1.0
    w2 = ((A + 2.0) * x2 - (A + 3.0)) * x2 * x2 + 1.0
    w3 = ((A * x3 - 5.0 * A) * x3 + 8.0 * A) * x3 - 4.0 * A
    return [w0, w1, w2, w3]


def resize_frames_u8(frames_u8, size, full_size=None, crop=(0, 0)):
    """What the reference's loader does to one decoded frame, restated with explicit gathers: ``ToTensor`` (uint8 / 255),
    ``F.interpolate(x, full_size, mode="bicubic", align_corners=True)`` (third-party: ATen upsample_bicubic2d — source
    coordinate ``o * (in - 1) / (out - 1)``, taps floor - 1 .. floor + 2 clamped to the image, A = -0.75), the crop window of
    RandomResizeCrop and ``.clamp(0, 1)``.  ``frames_u8``: [B,3,Hs,Ws] uint8 -> fp32 [B,3,H,W]."""
    x = frames_u8.to(torch.float32) / 255.0
    B, C, Hs, Ws = x.shape
    H, W = size
    Hf, Wf = (H, W) if full_size is None else full_size
    sy = torch.tensor((Hs - 1) / (Hf - 1) if Hf > 1 else 0.0, dtype=torch.float32)
    sx = torch.tensor((Ws - 1) / (Wf - 1) if Wf > 1 else 0.0, dtype=torch.float32)
    fy = sy * (torch.arange(H, dtype=torch.float32) + crop[0])
    fx = sx * (torch.arange(W, dtype=torch.float32) + crop[1])
    iy, ix = fy.floor(), fx.floor()
    wy, wx = _cubic_weights(fy - iy), _cubic_weights(fx - ix)
    iy, ix = iy.long(), ix.long()
    out = torch.zeros(B, C, H, W, dtype=torch.float32)
    for j in range(4):
        rows = x[:, :, (iy - 1 + j).clamp(0, Hs - 1), :]  # [B,C,H,Ws]
        acc = torch.zeros(B, C, H, W, dtype=torch.float32)
        for k in range(4):
            acc = acc + rows[:, :, :, (ix - 1 + k).clamp(0, Ws - 1)] * wx[k][None, None, None, :]
        out = out + acc * wy[j][None, None, :, None]
    return out.clamp(0.0, 1.0)
