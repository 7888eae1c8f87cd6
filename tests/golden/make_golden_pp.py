"""Golden fixture for ``Trainer.generate_post_process_disp`` (trainer.py:404-466), generated FROM THE REFERENCE.

    python tests/golden/make_golden_pp.py         # needs /root/reference (read-only); CPU only

Same shim and bare-Trainer technique as make_golden.py.  The networks the method runs internally
(``fixed_models["encoder"]`` / ``["depth"]``, out of scope here) are replaced by stubs that hand back prepared,
seeded decoder outputs for the 2B-image batch ``cat([img, img.flip(-1)])``; everything after them — the part
this repository implements — is the unmodified reference code.  Stored: the decoder outputs and the two
results ``disp_pp`` / ``mask_novel``."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import bare_trainer, install_shim  # noqa: E402


def run(Trainer, layers, name, B, N, H, W, seed, n_xz):
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.rand(*s, generator=g)
    t = bare_trainer(Trainer, layers, H, W)
    t.opt.num_ep = 8
    n_v = N - n_xz
    lev = torch.arange(n_v, dtype=torch.float32)[None, :] + rnd(2 * B, n_v) - 0.5
    disp_layered = ((0.3 * W) * (1.5 / (0.3 * W)) ** (lev / max(n_v - 1, 1)))[:, :, None, None].expand(-1, -1, H, W)
    mask = torch.ones(2 * B, n_v, H, W)
    if n_xz:
        gy = torch.linspace(-1, 1, H)[None, None, :, None].expand(2 * B, 1, H, W)
        h = 0.1852 + 0.1852 * rnd(2 * B, n_xz)
        Z = h[:, :, None, None] * 1.92 / (gy.clamp_min(1e-7) / 2.0)
        disp_layered = torch.cat([disp_layered, 0.1 * 0.58 * W / Z], 1)
        mask = torch.cat([mask, (gy >= 1e-7).expand(-1, n_xz, -1, -1).float()], 1)
    logits = 1.5 * torch.randn(2 * B, N, H, W, generator=g) * mask
    outputs = {"logits": logits, "probability": torch.softmax(logits, 1), "disp_layered": disp_layered.contiguous(),
               "disp": 1.0 + 20 * rnd(2 * B, 1, H, W)}
    t.fixed_models = {"encoder": lambda x: x, "depth": lambda f, grids: outputs}
    xs = torch.linspace(-1, 1, W)[None, None, None, :].expand(B, 1, H, W)
    ys = torch.linspace(-1, 1, H)[None, None, :, None].expand(B, 1, H, W)
    inputs = {("color_aug", "l"): rnd(B, 3, H, W), "grid": torch.cat([xs, ys], 1).contiguous()}
    with torch.no_grad():
        disp_pp, mask_novel = Trainer.generate_post_process_disp(t, inputs)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta_BNHW=np.array([B, N, H, W]),
                        logits=logits.numpy(), probability=outputs["probability"].numpy(), disp_layered=outputs["disp_layered"].numpy(),
                        disp=outputs["disp"].numpy(), disp_pp=disp_pp.numpy(), mask_novel=mask_novel.numpy())
    print(name, float(disp_pp.mean()), float(mask_novel.mean()))


def main():
    install_shim()
    import layers  # noqa (reference)
    import trainer  # noqa (reference)

    torch.manual_seed(0)
    run(trainer.Trainer, layers, "pp_vertical", 2, 7, 24, 64, 11, 0)
    run(trainer.Trainer, layers, "pp_xz", 1, 8, 16, 96, 12, 3)


if __name__ == "__main__":
    main()
