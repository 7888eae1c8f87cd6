"""CPU: the oracle (oracle/pd_oracle.py) replays the golden fixtures generated FROM THE REFERENCE
(tests/golden/make_golden.py) — this is what pins the oracle (the reference has no tests of its own)."""
import numpy as np
import pytest
import torch

from helpers import CASES, assert_close, load_case, pyramid_features
from oracle import pd_oracle as O

# fp32 restatement vs the reference's own fp32 run of the same math on the same CPU
FWD_ATOL = 2e-5
GRAD_RTOL = 2e-4


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_outputs_and_grads(name):
    c = load_case(name)
    losses = O.hot_path(c.opt, c.target_sides, c.inputs, c.outputs, pyramid_features)
    for k, want in c.expect.items():
        if k.startswith("out_"):
            nm, s = k[4:].split("@")
            s = s if s in ("l", "r") else int(s)
            assert_close(c.outputs[(nm, s)], want, FWD_ATOL, 1e-5, what=k)
        elif k.startswith("loss_"):
            assert_close(losses["loss/" + k[5:]], want, 1e-5, 1e-5, what=k)
    losses["loss/total_loss"].backward()
    for k, want in c.expect.items():
        if k.startswith("grad_"):
            g = c.leaves[k[5:]].grad
            scale = float(np.abs(want).max())
            assert_close(g, want, GRAD_RTOL * scale + 1e-9, 0.0, what=k)


def test_oracle_sampler_equals_torch_grid_sample():
    # the reference's sampler is third-party: torch.nn.functional.grid_sample (trainer.py:573-577)
    g = torch.Generator().manual_seed(0)
    feat = torch.randn(3, 4, 17, 23, generator=g, requires_grad=True)
    gx = (torch.rand(3, 17, 23, generator=g) * 2.6 - 1.3).requires_grad_(True)
    gy = (torch.rand(3, 17, 23, generator=g) * 2.6 - 1.3).requires_grad_(True)
    a = O.bilinear_sample(feat, gx, gy)
    b = torch.nn.functional.grid_sample(feat, torch.stack([gx, gy], -1), padding_mode="zeros", align_corners=True)
    assert_close(a, b, 1e-6, what="sample")
    w = torch.randn(a.shape, generator=g)
    ga = torch.autograd.grad((a * w).sum(), [feat, gx, gy])
    gb = torch.autograd.grad((b * w).sum(), [feat, gx, gy])
    for x, y, nm in zip(ga, gb, ["d_feat", "d_gx", "d_gy"]):
        assert_close(x, y, 2e-5, 1e-5, what=nm)


def test_oracle_primitives_match_reference():
    import os

    from helpers import GOLDEN

    z = np.load(os.path.join(GOLDEN, "primitives.npz"))
    T = lambda k: torch.from_numpy(z[k])
    pred = T("pred").requires_grad_(True)
    pred2 = T("pred2").requires_grad_(True)
    rl = O.reprojection_loss(pred, T("tgt"), True)
    rl2 = O.reprojection_loss(pred2, T("base"), True)
    assert_close(rl, z["reproj"], 1e-6, what="reproj")
    assert_close(rl2, z["reproj2"], 1e-6, what="reproj2")
    assert_close(O.ssim_map(pred2.detach(), T("base")), z["ssim2"], 1e-6, what="ssim")
    (rl.mean() + rl2.mean()).backward()
    assert_close(pred.grad, z["grad_pred"], 1e-8, 1e-4, what="grad_pred")
    assert_close(pred2.grad, z["grad_pred2"], 1e-8, 1e-4, what="grad_pred2")
    B, N = z["dist"].shape
    H, W = z["depth"].shape[-2:]
    u, v, mask = O.homography_coords(T("dist"), T("nrm"), T("T"), T("K"), T("inv_K"), H, W)
    grid = torch.stack([O.normalise(u, W), O.normalise(v, H)], -1).reshape(B * N, H, W, 2)
    want = torch.from_numpy(z["homo_grid"])
    # projective coordinates can be huge where z -> 1e-7; compare where the reference grid is sane
    sane = want.abs().amax(-1) < 50
    assert_close(grid[sane], want[sane], 2e-4, 1e-4, what="homography grid")
    assert (mask.reshape(z["homo_mask"].shape).numpy() == z["homo_mask"]).mean() > 0.9999
    disp = 0.1 * 0.58 * W / T("depth")
    ud, vd = O.depth_warp_coords(disp, T("T"), T("K"), T("inv_K"))
    gd = torch.stack([O.normalise(ud, W), O.normalise(vd, H)], -1).reshape(B, H, W, 2)
    assert_close(gd, z["depth_grid"], 2e-4, 1e-4, what="depth grid")


@pytest.mark.parametrize("name", ["pp_vertical", "pp_xz"])
def test_oracle_post_process_disp_matches_reference(name):
    """Trainer.generate_post_process_disp (trainer.py:404-466) run by tests/golden/make_golden_pp.py vs the restatement."""
    import os

    from helpers import GOLDEN

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    outputs = {k: torch.from_numpy(z[k]) for k in ("logits", "probability", "disp_layered", "disp")}
    disp_pp, mask_novel, o_l, o_fr = O.post_process_disp(outputs)
    assert_close(disp_pp, z["disp_pp"], 1e-5, what="disp_pp")
    assert_close(mask_novel, z["mask_novel"], 1e-6, what="mask_novel")
    assert float(o_l.max()) <= 1.0 and float(o_fr.max()) <= 1.0 and float(mask_novel.max()) <= 1.0


@pytest.mark.parametrize("name", ["tail_plain", "tail_mix"])
def test_oracle_decoder_tail_matches_reference(name):
    """A real reference DepthDecoder's forward/backward (tests/golden/make_golden_tail.py) vs the restated tail."""
    import os

    from helpers import GOLDEN

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    mix = bool(z["mixture"])
    T = lambda k: torch.from_numpy(z[k])
    lr = T("logits_raw").requires_grad_(True)
    sr = T("sigma_raw").requires_grad_(True) if mix else None
    dl = T("disp_layered").requires_grad_(True)
    out = O.decoder_tail(lr, sr, T("padding_mask"), dl, mix)
    L = (out["logits"] * T("A")).sum() + (out["disp"] * T("Cd")).sum()
    if mix:
        L = L + (out["sigma"] * T("Bm")).sum()
    L.backward()
    for k in ["logits", "probability", "disp", "depth"] + (["sigma", "pi"] if mix else []):
        assert_close(out[k], z["out_" + k], 1e-6, 1e-6, what=k)
    assert_close(lr.grad, z["grad_logits_raw"], 1e-6, 1e-5, what="grad logits_raw")
    assert_close(dl.grad, z["grad_disp_layered"], 1e-6, 1e-5, what="grad disp_layered")
    if mix:
        assert_close(sr.grad, z["grad_sigma_raw"], 1e-6, 1e-5, what="grad sigma_raw")
