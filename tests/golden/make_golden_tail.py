"""Golden fixture for the decoder tail (networks/depth_decoder.py:258-291), generated FROM THE REFERENCE.

    python tests/golden/make_golden_tail.py        # needs /root/reference (read-only); CPU only

A real (small, randomly initialised) reference ``DepthDecoder`` runs its unmodified ``forward`` on random feature
pyramids; forward hooks capture the outputs of ``convs["dispconv"]`` / ``convs["sigmaconv"]`` (the inputs of the tail).
Stored: those raw tensors, ``padding_mask`` / ``disp_layered`` as the decoder built them, the tail's outputs, and the
gradients of a fixed random linear functional of (logits, sigma, disp) w.r.t. the raw tensors and ``disp_layered``."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import install_shim  # noqa: E402


def run(name, seed, mixture, xz_levels, H, W, B=2, no_levels=7):
    from networks.depth_decoder import DepthDecoder

    torch.manual_seed(seed)
    num_ch_enc = [8, 8, 16, 16, 32]
    dec = DepthDecoder(num_ch_enc, no_levels=no_levels, num_ep=0, use_denseaspp=False, xz_levels=xz_levels,
                       use_mixture_loss=mixture, plane_residual=True)
    feats = [torch.randn(B, c, H // (2 ** (i + 1)), W // (2 ** (i + 1))) for i, c in enumerate(num_ch_enc)]
    xs = torch.linspace(-1, 1, W)[None, None, None, :].expand(B, 1, H, W)
    ys = torch.linspace(-1, 1, H)[None, None, :, None].expand(B, 1, H, W)
    grids = torch.cat([xs, ys], 1).contiguous()
    cap = {}

    def hook(key):
        def f(mod, inp, out):
            out.retain_grad()
            cap[key] = out
        return f

    dec.convs["dispconv"].register_forward_hook(hook("logits_raw"))
    if mixture:
        dec.convs["sigmaconv"].register_forward_hook(hook("sigma_raw"))
    # a livelier logit / sigma range than a fresh conv gives
    with torch.no_grad():
        for k in ("dispconv",) + (("sigmaconv",) if mixture else ()):
            for prm in dec.convs[k].parameters():
                prm.mul_(8.0)
    out = dec(feats, grids)
    out["disp_layered"].retain_grad()
    g = torch.Generator().manual_seed(seed + 1)
    N = out["logits"].shape[1]
    A = torch.randn(B, N, H, W, generator=g)
    Bm = torch.randn(B, N, H, W, generator=g)
    Cd = torch.randn(B, 1, H, W, generator=g)
    L = (out["logits"] * A).sum() + (out["disp"] * Cd).sum()
    if mixture:
        L = L + (out["sigma"] * Bm).sum()
    L.backward()
    rec = dict(meta_BNHW=np.array([B, N, H, W]), mixture=np.array(int(mixture)),
               logits_raw=cap["logits_raw"].detach().numpy(), padding_mask=out["padding_mask"].float().numpy(),
               disp_layered=out["disp_layered"].detach().numpy(), A=A.numpy(), Bm=Bm.numpy(), Cd=Cd.numpy(),
               out_logits=out["logits"].detach().numpy(), out_probability=out["probability"].detach().numpy(),
               out_disp=out["disp"].detach().numpy(), out_depth=out["depth"].detach().numpy(),
               grad_logits_raw=cap["logits_raw"].grad.numpy(), grad_disp_layered=out["disp_layered"].grad.numpy())
    if mixture:
        rec.update(sigma_raw=cap["sigma_raw"].detach().numpy(), out_sigma=out["sigma"].detach().numpy(),
                   out_pi=out["pi"].detach().numpy(), grad_sigma_raw=cap["sigma_raw"].grad.numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, float(out["disp"].mean()), float(out["probability"].max()))


def main():
    install_shim()
    run("tail_plain", 21, False, 3, 64, 96, B=1)
    run("tail_mix", 22, True, 3, 64, 96, B=1)


if __name__ == "__main__":
    main()
