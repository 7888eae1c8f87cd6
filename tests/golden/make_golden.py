"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run once in the build container (CPU, no GPU needed):

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

The reference (svip-lab/PlaneDepth @ /root/reference) has no tests and no golden vectors for the
photometric-reconstruction path, so parity is pinned against *outputs of the reference's
own functions*: this script imports the reference through a small compatibility shim (missing
tensorboardX / IPython / skimage, the removed ``torch._six``, ``.cuda()`` on a CPU-only box), builds a
bare ``Trainer`` object without running its constructor (which needs NCCL + GPU + network), and calls
the unbound reference methods

    Trainer.pred_novel_images      trainer.py:523-603
    Trainer.compute_losses         trainer.py:701-773
    Trainer.compute_reprojection_loss  trainer.py:687-699   (SSIM + L1, layers.py:276-306)
    HomographyWarp / BackprojectDepth / Project3D   layers.py:128-234

on small seeded inputs, then stores inputs, outputs and autograd gradients as ``.npz``.
Nothing here runs on the GPU box; tests only read the ``.npz`` files.

The only deviation from upstream behaviour: the ``norm`` tensor is passed as float32 (upstream emits
int64 when only vertical planes exist and then crashes — defect D5 in SURVEY.md §8a).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("PLANEDEPTH_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def install_shim():
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
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    import PIL.Image

    if not hasattr(PIL.Image, "ANTIALIAS"):
        PIL.Image.ANTIALIAS = PIL.Image.LANCZOS
    sys.path.insert(0, REF)


def pyramid_features(x):
    """Deterministic stand-in for the VGG19 feature net (its weights need the network); exercises
    the external gradient d loss / d rgb_rec that the fused backward must consume."""
    return [x, F.avg_pool2d(x, 2), F.avg_pool2d(x, 4)]


def bare_trainer(Trainer, layers, H, W, **flags):
    opt = dict(
        net_type="ResNet", warp_type="disp_warp", match_aug=False, use_mixture_loss=False,
        render_probability=False, automask=False, alpha_self=0.0, self_distillation=0.0, alpha_pc=0.1,
        alpha_smooth=0.04, gamma_smooth=2, use_ssim=False, use_mom=False, novel_frame_ids=[],
        no_stereo=False, use_colmap=True, height=H, width=W,
    )
    opt.update(flags)
    t = object.__new__(Trainer)
    t.opt = types.SimpleNamespace(**opt)
    t.device = torch.device("cpu")
    t.target_sides = ([] if t.opt.no_stereo else ["r"]) + list(t.opt.novel_frame_ids)
    t.softmax = nn.Softmax(1)
    t.ssim = layers.SSIM()
    t.backproject_depth = layers.BackprojectDepth(H, W)
    t.project_3d = layers.Project3D(H, W)
    t.homography_warp = layers.HomographyWarp(H, W)
    t.pc_net = pyramid_features
    return t


def intrinsics(B, H, W):
    K = torch.tensor([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32)
    K = K[None].repeat(B, 1, 1)
    return K, torch.linalg.pinv(K)


def small_pose(B, gen, scale=1.0):
    """4x4 rigid transforms from small axis-angle / translation draws (Rodrigues)."""
    aa = 0.03 * scale * torch.randn(B, 3, generator=gen)
    tr = 0.05 * scale * torch.randn(B, 3, generator=gen)
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


def to_np(d):
    out = {}
    for k, v in d.items():
        if v is None:
            continue
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    return out


def key(name, s):
    return "%s@%s" % (name, s)


def run_case(Trainer, layers, name, B, N, H, W, seed, warp, mixture, automask, frames, mask_novel, sd,
             n_xz=0, dense_disp=False):
    g = torch.Generator().manual_seed(seed)
    t = bare_trainer(Trainer, layers, H, W, warp_type=warp, use_mixture_loss=mixture, automask=automask,
                     novel_frame_ids=frames, self_distillation=sd)
    rnd = lambda *s: torch.rand(*s, generator=g)
    inputs = {}
    for s in ["l", "r"] + frames:
        inputs[("color", s)] = rnd(B, 3, H, W)
        inputs[("color_aug", s)] = inputs[("color", s)].clone()
    K, iK = intrinsics(B, H, W)
    inputs["K"], inputs["inv_K"] = K, iK
    Tr = torch.eye(4)[None].repeat(B, 1, 1)
    Tr[:, 0, 3] = -0.1
    inputs[("Rt", "r")] = Tr
    # plane parameters (leaves) -------------------------------------------------------------
    n_v = N - n_xz
    lev = (torch.arange(n_v, dtype=torch.float32)[None, :] + rnd(B, n_v) - 0.5).requires_grad_(True)
    dmax, dmin = 0.4 * W, 0.6
    disp_v = dmax * (dmin / dmax) ** (lev / max(n_v - 1, 1))  # [B,n_v]
    distance = 0.1 * 0.58 * W / disp_v
    norm = torch.tensor([0.0, 0.0, 1.0])[None, None].expand(B, n_v, 3)
    disp_layered = disp_v[:, :, None, None].expand(-1, -1, H, W)
    padding_mask = torch.ones(B, n_v, H, W)
    leaves = {"lev": lev}
    if n_xz:
        gy = torch.linspace(-1, 1, H)[None, None, :, None].expand(B, 1, H, W)
        hlev = (torch.arange(n_xz, dtype=torch.float32)[None, :] + rnd(B, n_xz) - 0.5).requires_grad_(True)
        leaves["hlev"] = hlev
        h = 0.1852 + (0.3704 - 0.1852) * hlev / max(n_xz - 1, 1)
        xz_mask = (gy >= 1e-7).expand(-1, n_xz, -1, -1)
        Z = h[:, :, None, None] * 1.92 / (gy.clamp_min(1e-7) / 2.0)
        disp_xz = 0.1 * 0.58 * W / Z
        disp_layered = torch.cat([disp_layered, disp_xz], 1)
        padding_mask = torch.cat([padding_mask, xz_mask], 1)
        xz_n = torch.tensor([0.0, 1.0, 0.0])[None, None].expand(B, n_xz, 3)
        norm = torch.cat([norm, xz_n], 1)
        distance = torch.cat([distance, h], 1)
    if dense_disp:
        bump = (0.3 * torch.randn(B, N, H, W, generator=g)).requires_grad_(True)
        leaves["bump"] = bump
        disp_layered = disp_layered + bump
    logits = (1.5 * torch.randn(B, N, H, W, generator=g) * padding_mask).requires_grad_(True)
    leaves["logits"] = logits
    outputs = {"logits": logits, "disp_layered": disp_layered, "padding_mask": padding_mask,
               "distance": distance, "norm": norm, "probability": torch.softmax(logits, 1).detach()}
    if mixture:
        sigma = torch.sigmoid(1.5 * torch.randn(B, N, H, W, generator=g)).clamp(0.01, 1.0).requires_grad_(True)
        leaves["sigma"] = sigma
        outputs["sigma"] = sigma
    disp = (1.0 + 20 * rnd(B, 1, H, W)).requires_grad_(True)
    leaves["disp"] = disp
    outputs["disp"] = disp
    if mask_novel:
        outputs["mask_novel"] = rnd(B, 1, H, W)
    if sd > 0:
        outputs["disp_pp"] = 1.0 + 20 * rnd(B, 1, H, W)
    for f in frames:
        Tf = small_pose(B, g).requires_grad_(True)
        leaves["T%d" % f] = Tf
        inputs[("Rt", f)] = Tf
    outputs.update({("Rt", "r"): Tr})
    for f in frames:
        outputs[("Rt", f)] = inputs[("Rt", f)]
    # ---- the reference ---------------------------------------------------------------------
    Trainer.pred_novel_images(t, inputs, outputs)
    losses = Trainer.compute_losses(t, inputs, outputs)
    losses["loss/total_loss"].backward()
    rec = {"meta_BNHW": np.array([B, N, H, W]), "meta_nxz": np.array(n_xz)}
    rec.update(to_np({"in_" + ("%s@%s" % k if isinstance(k, tuple) else k): v for k, v in inputs.items()}))
    rec.update(to_np({"leaf_" + k: v for k, v in leaves.items()}))
    rec.update(to_np({"grad_" + k: v.grad for k, v in leaves.items()}))
    rec.update(to_np({"pre_disp_layered": disp_layered, "pre_padding_mask": padding_mask.float(),
                      "pre_distance": distance, "pre_norm": norm}))
    for opt_k in ("mask_novel", "disp_pp"):
        if opt_k in outputs:
            rec["pre_" + opt_k] = outputs[opt_k].numpy()
    for s in t.target_sides:
        for nm in ("rgb_rec", "rgb_rec_layered", "logit_rec", "probability_rec", "sigma_rec", "pi_rec"):
            if (nm, s) in outputs:
                rec["out_" + key(nm, s)] = outputs[(nm, s)].detach().numpy()
    for k, v in losses.items():
        rec["loss_" + k.split("/")[1]] = np.asarray(float(v))
    rec["meta_flags"] = np.array([warp, str(int(mixture)), str(int(automask)), ",".join(map(str, frames)),
                                  str(int(mask_novel)), str(sd)])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, {k: float(v) for k, v in losses.items()})


def run_primitives(Trainer, layers):
    g = torch.Generator().manual_seed(7)
    B, N, H, W = 2, 5, 32, 64
    t = bare_trainer(Trainer, layers, H, W, use_ssim=True)
    pred = torch.rand(B, 3, H, W, generator=g).requires_grad_(True)
    tgt = torch.rand(B, 3, H, W, generator=g)
    # a smoother pair as well: SSIM on pure noise sits in the clamp
    base = F.interpolate(torch.rand(B, 3, H // 4, W // 4, generator=g), size=(H, W), mode="bilinear", align_corners=False)
    pred2 = (base + 0.05 * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1).requires_grad_(True)
    rl = Trainer.compute_reprojection_loss(t, pred, tgt)
    rl2 = Trainer.compute_reprojection_loss(t, pred2, base)
    (rl.mean() + rl2.mean()).backward()
    ssim_raw = layers.SSIM()(pred2.detach(), base)
    K, iK = intrinsics(B, H, W)
    dist = (0.5 + 5 * torch.rand(B, N, generator=g))
    nrm = F.normalize(torch.tensor([0.0, 0.0, 1.0]) + 0.3 * torch.randn(B, N, 3, generator=g), dim=-1)
    T = small_pose(B, g, 2.0)
    ex = lambda M: M[:, None].expand(-1, N, -1, -1).reshape(B * N, 4, 4)
    grid, mask = t.homography_warp(dist, nrm, ex(T), ex(K), ex(iK))
    depth = 1.0 + 10 * torch.rand(B, 1, H, W, generator=g)
    cam = t.backproject_depth(depth, iK)
    grid_d = t.project_3d(cam, K, T)
    np.savez_compressed(
        os.path.join(HERE, "primitives.npz"),
        pred=pred.detach().numpy(), tgt=tgt.numpy(), pred2=pred2.detach().numpy(), base=base.numpy(),
        reproj=rl.detach().numpy(), reproj2=rl2.detach().numpy(), ssim2=ssim_raw.numpy(),
        grad_pred=pred.grad.numpy(), grad_pred2=pred2.grad.numpy(),
        K=K.numpy(), inv_K=iK.numpy(), dist=dist.numpy(), nrm=nrm.numpy(), T=T.numpy(),
        homo_grid=grid.numpy(), homo_mask=mask.numpy(), depth=depth.numpy(), depth_grid=grid_d.numpy(),
    )
    print("primitives ok")


def main():
    install_shim()
    import layers  # noqa  (reference)
    import trainer  # noqa (reference)

    Trainer = trainer.Trainer
    torch.manual_seed(0)
    torch.set_num_threads(4)
    #                      name                 B  N  H   W  seed warp               mix   auto  frames  mnov  sd
    run_case(Trainer, layers, "disp_l1",          2, 7, 32, 64, 1, "disp_warp",       False, False, [],     False, 0.0)
    run_case(Trainer, layers, "disp_l1_auto_xz",  2, 8, 32, 64, 2, "disp_warp",       False, True,  [],     False, 0.0, n_xz=3)
    run_case(Trainer, layers, "disp_mix_mask_sd", 2, 7, 32, 64, 3, "disp_warp",       True,  True,  [],     True,  1.0, n_xz=2)
    run_case(Trainer, layers, "disp_dense_mix",   1, 6, 32, 64, 4, "disp_warp",       True,  False, [],     False, 0.0, dense_disp=True)
    run_case(Trainer, layers, "homo_l1_auto",     2, 6, 32, 64, 5, "homography_warp", False, True,  [-1, 1], False, 0.0, n_xz=2)
    run_case(Trainer, layers, "homo_mix",         1, 6, 32, 64, 6, "homography_warp", True,  True,  [1],    True,  0.0, n_xz=2)
    run_primitives(Trainer, layers)


if __name__ == "__main__":
    main()
