"""Drop-in replacements for the two reference methods that make up the photometric-reconstruction path:

    Trainer.pred_novel_images(inputs, outputs)  -> None      /root/reference/trainer.py:523-603
    Trainer.compute_losses(inputs, outputs)     -> dict      /root/reference/trainer.py:701-773

Same names (plus the Monodepth2 spelling ``generate_images_pred``), same ``inputs`` / ``outputs`` dict
contract, same ``self`` attributes (``opt``, ``target_sides``, ``pc_net``), same
error behaviour (Python exceptions).  Use either as a mixin in front of the reference Trainer or patch
the two methods onto it.

What changes underneath: the grid build, the N-plane bilinear warp, the validity mask, the softmax /
Laplacian-mixture weighting, the compositing and the photometric term run inside hand-written sm_100a
kernels (planedepth_b200/csrc) behind the C ABI in include/planedepth_b200.h; the N warped tensors are
never materialised.  Consequently ``outputs[("rgb_rec_layered"|"logit_rec"|"probability_rec"|
"sigma_rec"|"pi_rec", side)]`` are only produced when ``self.materialize_layered`` is true (they are
detached copies for inspection; no live reference code path consumes them — ``mirror_occlusion_mask``
is broken upstream).  ``outputs[("rgb_rec", side)]`` is always produced and is
differentiable (``log_img`` and the perceptual term read it).

What stays PyTorch (out of scope): the perceptual network, the smoothness term, the
self-distillation |disp - disp_pp| term and the tiny 3x3 pose / homography algebra.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib as L
from .functional import WarpConfig, photometric_loss, warp_composite

_WARP = {"disp_warp": L.PD_WARP_DISP, "homography_warp": L.PD_WARP_HOMOGRAPHY, "depth_warp": L.PD_WARP_DEPTH}


def _flag(opt, name, default):
    return getattr(opt, name, default)


def homography_params(distance, norm, T, K, inv_K):
    """Host prologue of HomographyWarp.forward (layers.py:211-223): ``H_t2s = inv(K (R + t n^T / d) K^-1)``
    per (image, plane) and the plane normal in the target frame ``R n``.  Tiny batched 3x3 algebra kept in
    PyTorch so autograd carries d loss / d H_t2s (returned by the kernel) back to ``distance`` / ``T``.
    Returns hmat [B*N,12] (H row-major | R n) and cam [B,9] (inv_K 3x3)."""
    B, N = distance.shape
    R = T[:, None, :3, :3]
    t = T[:, None, :3, 3:4]
    n = norm.to(torch.float32).reshape(B, N, 1, 3)
    H_s2t = K[:, None, :3, :3] @ ((R + (t @ n) / distance.reshape(B, N, 1, 1)) @ inv_K[:, None, :3, :3])
    H_t2s = torch.inverse(H_s2t)
    Rn = (R @ n.transpose(-1, -2)).reshape(B * N, 3)
    hmat = torch.cat([H_t2s.reshape(B * N, 9), Rn.detach()], 1)
    cam = inv_K[:, :3, :3].reshape(B, 9)
    return hmat, cam


def depth_warp_params(T, K, inv_K):
    """BackprojectDepth / Project3D constants (layers.py:152, 172): cam [B,21] = inv_K 3x3 | (K T)[:3,:]."""
    B = K.shape[0]
    P = (K @ T)[:, :3, :]
    return torch.cat([inv_K[:, :3, :3].reshape(B, 9), P.reshape(B, 12)], 1)


class HotPathMixin:
    """Provides ``pred_novel_images`` / ``generate_images_pred`` / ``compute_losses`` on any object that
    carries the reference Trainer's attributes."""

    #: also materialise the per-plane tensors of trainer.py:582-602 (costs 5*N extra planes of HBM writes)
    materialize_layered: bool = False
    #: reproduce the fp32 rounding of the reference's coordinate normalise / un-normalise round trip bit for bit
    #: (PD_FLAG_EXACT_COORDS); the default stereo fast path samples at the exact positions instead
    exact_coords: bool = False
    #: photometric term: None = reference behaviour (mixture NLL if opt.use_mixture_loss else L1);
    #: "ssim_l1" = 0.85*SSIM + 0.15*L1 (compute_reprojection_loss, trainer.py:687-699) on the novel view
    photometric: Optional[str] = None

    # ------------------------------------------------------------------------------------------
    def pred_novel_images(self, inputs: Dict, outputs: Dict) -> None:
        opt = self.opt
        B, N, H, W = outputs["probability"].shape
        color = "color_aug" if _flag(opt, "match_aug", False) else "color"
        src = inputs[(color, "l")]
        mixture = bool(_flag(opt, "use_mixture_loss", False))
        automask = bool(_flag(opt, "automask", False))
        if _flag(opt, "render_probability", False):
            raise NotImplementedError("render_probability is unreachable upstream (shape error at depth_decoder.py:259)")
        wt = _flag(opt, "warp_type", "disp_warp")
        if wt not in _WARP:
            raise ValueError("unknown warp_type %r" % (wt,))
        for side in self.target_sides:
            disp = mask = hmat = cam = None
            sign = 0.0
            if wt == "disp_warp":
                disp = outputs["disp_layered"]
                mask = outputs["padding_mask"]
                sign = 1.0 if side == "r" else (-1.0 if side == "l" else 0.0)
            elif wt == "homography_warp":
                hmat, cam = homography_params(outputs["distance"], outputs["norm"], outputs[("Rt", side)], inputs["K"], inputs["inv_K"])
            else:
                disp = outputs["disp_layered"]
                mask = outputs["padding_mask"]  # upstream dereferences an unbound local here (defect D1)
                cam = depth_warp_params(inputs[("Rt", side)], inputs["K"], inputs["inv_K"])
            cfg = WarpConfig(warp_type=_WARP[wt], mixture=mixture, automask=automask, disp_sign=sign, shape=(B, N, H, W),
                             layered=bool(self.materialize_layered), exact_coords=bool(self.exact_coords))
            tgt = inputs[(color, side)] if mixture else None
            rgb_rec, nll, nll_auto, layered = warp_composite(
                cfg, src, tgt, outputs["logits"], outputs.get("sigma") if mixture else None, disp, mask, hmat, cam)
            outputs[("rgb_rec", side)] = rgb_rec
            if mixture:
                outputs[("nll_rec", side)] = nll
                if automask:
                    outputs[("nll_auto_rec", side)] = nll_auto
            if layered is not None:
                for k, v in layered.items():
                    outputs[(k, side)] = v

    # Monodepth2 name used by BASELINE.json's north_star
    generate_images_pred = pred_novel_images

    # ------------------------------------------------------------------------------------------
    def perceptual_loss(self, pred, target, source=None):
        """trainer.py:672-685 — stays PyTorch (the feature network is cuDNN territory)."""
        pv, tv = self.pc_net(pred), self.pc_net(target)
        sv = self.pc_net(source) if source is not None else None
        total = 0
        for i in range(3):
            lp = ((pv[i] - tv[i]) ** 2).mean(1, True)
            if sv is not None:
                la = ((sv[i] - tv[i]) ** 2).mean(1, True)
                lp, _ = torch.cat([lp, la], dim=1).min(1, True)
            total = total + lp.mean()
        return total

    def _photometric_mode(self) -> int:
        mode = self.photometric
        if mode is None:
            mode = "mixture" if _flag(self.opt, "use_mixture_loss", False) else "l1"
        return {"l1": L.PD_LOSS_L1, "mixture": L.PD_LOSS_MIXTURE, "ssim_l1": L.PD_LOSS_SSIM_L1}[mode]

    def compute_losses(self, inputs: Dict, outputs: Dict) -> Dict[str, torch.Tensor]:
        opt = self.opt
        B, N, H, W = outputs["probability"].shape
        color = "color_aug" if _flag(opt, "match_aug", False) else "color"
        automask = bool(_flag(opt, "automask", False))
        mode = self._photometric_mode()
        if mode == L.PD_LOSS_MIXTURE and not _flag(opt, "use_mixture_loss", False):
            raise ValueError("photometric='mixture' needs opt.use_mixture_loss (sigma channel)")
        pc_net = getattr(self, "pc_net", None)
        losses = {"loss/ph_loss": 0, "loss/pc_loss": 0, "loss/total_loss": 0}
        src = inputs[(color, "l")]
        mask_novel = outputs.get("mask_novel")
        for side in self.target_sides:
            target = inputs[(color, side)]
            ph_sum, pred, _ = photometric_loss(
                mode, automask, outputs[("rgb_rec", side)], target, src, mask_novel,
                outputs.get(("nll_rec", side)), outputs.get(("nll_auto_rec", side)))
            ph_loss = ph_sum / float(B * H * W)
            losses["loss/ph_loss"] = losses["loss/ph_loss"] + ph_loss
            total = ph_loss
            if pc_net is not None:
                pc = self.perceptual_loss(pred, target, src if automask else None)
                losses["loss/pc_loss"] = losses["loss/pc_loss"] + pc
                total = total + _flag(opt, "alpha_pc", 0.1) * pc
            if _flag(opt, "self_distillation", 0.0) > 0:
                dl = torch.abs(outputs["disp"] - outputs["disp_pp"]).mean()
                losses["loss/disp_loss"] = dl
                total = total + opt.self_distillation * dl
            losses["loss/total_loss"] = losses["loss/total_loss"] + total
        n_t = len(self.target_sides)
        for k in list(losses.keys()):  # trainer.py:765-766
            losses[k] = losses[k] / n_t
        if "disp" in outputs:
            x0 = int(0.2 * W)
            sm = smooth_loss_disp(outputs["disp"][..., x0:], inputs[("color", "l")][..., x0:], _flag(opt, "gamma_smooth", 2))
            losses["loss/smooth_loss"] = sm
            losses["loss/total_loss"] = losses["loss/total_loss"] + _flag(opt, "alpha_smooth", 0.04) * sm
        return losses


def smooth_loss_disp(disp, img, gamma=1.0):
    """Edge-aware first-order smoothness (layers.py:243-256) — out of the hot path, kept in PyTorch."""
    gdx = (disp[:, :, :, :-1] - disp[:, :, :, 1:]).abs()
    gdy = (disp[:, :, :-1, :] - disp[:, :, 1:, :]).abs()
    gix = (img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, keepdim=True)
    giy = (img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, keepdim=True)
    return (gdx * torch.exp(-gamma * gix)).mean() + (gdy * torch.exp(-gamma * giy)).mean()


class HotPath(HotPathMixin):
    """Stand-alone carrier of the attributes the two methods read from ``self`` (what tests, bench.py
    and smoke() instantiate instead of the full Trainer, whose constructor needs NCCL + KITTI)."""

    def __init__(self, opt, target_sides=None, pc_net=None, photometric: Optional[str] = None, materialize_layered: bool = False,
                 exact_coords: bool = False):
        self.opt = opt
        if target_sides is None:
            target_sides = ([] if _flag(opt, "no_stereo", False) else ["r"]) + list(_flag(opt, "novel_frame_ids", []))
        self.target_sides = target_sides
        self.pc_net = pc_net
        self.photometric = photometric
        self.materialize_layered = materialize_layered
        self.exact_coords = exact_coords

    def process(self, inputs, outputs):
        self.pred_novel_images(inputs, outputs)
        return self.compute_losses(inputs, outputs)
