"""Shared test helpers: load a golden fixture (tests/golden/*.npz, produced from the reference by
tests/golden/make_golden.py) back into the `inputs` / `outputs` dict contract of the reference
(SURVEY.md Appendix B), with fresh autograd leaves."""
import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["disp_l1", "disp_l1_auto_xz", "disp_mix_mask_sd", "disp_dense_mix", "homo_l1_auto", "homo_mix"]


def pyramid_features(x):
    # same deterministic stand-in for the perceptual network as make_golden.py
    return [x, F.avg_pool2d(x, 2), F.avg_pool2d(x, 4)]


def _side(s):
    return s if s in ("l", "r") else int(s)


def load_case(name, device="cpu"):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, N, H, W = [int(v) for v in z["meta_BNHW"]]
    n_xz = int(z["meta_nxz"])
    warp, mix, auto, frames, mnov, sd = [str(v) for v in z["meta_flags"]]
    frames = [int(f) for f in frames.split(",") if f]
    opt = SimpleNamespace(
        warp_type=warp, use_mixture_loss=bool(int(mix)), automask=bool(int(auto)), novel_frame_ids=frames,
        self_distillation=float(sd), match_aug=False, alpha_pc=0.1, alpha_smooth=0.04, gamma_smooth=2,
        no_stereo=False, use_ssim=False, render_probability=False, alpha_self=0.0,
    )
    T = lambda k: torch.from_numpy(z[k]).to(device)
    inputs = {}
    for k in z.files:
        if not k.startswith("in_"):
            continue
        nm = k[3:]
        if "@" in nm:
            a, s = nm.split("@")
            inputs[(a, _side(s))] = T(k)
        else:
            inputs[nm] = T(k)
    leaves = {k[5:]: T(k).requires_grad_(True) for k in z.files if k.startswith("leaf_")}
    # rebuild the plane geometry from the leaves exactly as make_golden.py did -------------------
    n_v = N - n_xz
    lev = leaves["lev"]
    dmax, dmin = 0.4 * W, 0.6
    disp_v = dmax * (dmin / dmax) ** (lev / max(n_v - 1, 1))
    distance = 0.1 * 0.58 * W / disp_v
    norm = torch.tensor([0.0, 0.0, 1.0], device=device)[None, None].expand(B, n_v, 3)
    disp_layered = disp_v[:, :, None, None].expand(-1, -1, H, W)
    padding_mask = torch.ones(B, n_v, H, W, device=device)
    if n_xz:
        gy = torch.linspace(-1, 1, H, device=device)[None, None, :, None].expand(B, 1, H, W)
        h = 0.1852 + (0.3704 - 0.1852) * leaves["hlev"] / max(n_xz - 1, 1)
        xz_mask = (gy >= 1e-7).expand(-1, n_xz, -1, -1)
        Z = h[:, :, None, None] * 1.92 / (gy.clamp_min(1e-7) / 2.0)
        disp_layered = torch.cat([disp_layered, 0.1 * 0.58 * W / Z], 1)
        padding_mask = torch.cat([padding_mask, xz_mask], 1)
        norm = torch.cat([norm, torch.tensor([0.0, 1.0, 0.0], device=device)[None, None].expand(B, n_xz, 3)], 1)
        distance = torch.cat([distance, h], 1)
    if "bump" in leaves:
        disp_layered = disp_layered + leaves["bump"]
    outputs = {
        "logits": leaves["logits"], "disp_layered": disp_layered, "padding_mask": padding_mask,
        "distance": distance, "norm": norm, "probability": torch.softmax(leaves["logits"].detach(), 1),
        "disp": leaves["disp"],
    }
    if "sigma" in leaves:
        outputs["sigma"] = leaves["sigma"]
    for k in ("mask_novel", "disp_pp"):
        if "pre_" + k in z.files:
            outputs[k] = T("pre_" + k)
    outputs[("Rt", "r")] = inputs[("Rt", "r")]
    for f in frames:
        inputs[("Rt", f)] = leaves["T%d" % f]
        outputs[("Rt", f)] = leaves["T%d" % f]
    target_sides = ["r"] + frames
    expect = {k: z[k] for k in z.files if k.startswith(("out_", "grad_", "loss_", "pre_"))}
    return SimpleNamespace(opt=opt, inputs=inputs, outputs=outputs, leaves=leaves, target_sides=target_sides,
                           expect=expect, shape=(B, N, H, W), name=name)


def max_err(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0


def assert_close(got, want, atol, rtol=0.0, what=""):
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    want = want.detach().cpu().numpy() if torch.is_tensor(want) else np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    tol = atol + rtol * np.abs(want)
    bad = err > tol
    assert not bad.any(), "%s: %d/%d off, max err %.3e (atol %.1e rtol %.1e, max |want| %.3e)" % (
        what, int(bad.sum()), bad.size, float(err.max()), atol, rtol, float(np.abs(want).max()))


# --------------------------------------------------------------------------------------------------
# bounded comparison used by the GPU parity tests
# --------------------------------------------------------------------------------------------------
REPORT = []  # filled when PD_TEST_REPORT is set: (what, stats) of every bounded_check call


def error_stats(got, want, atol):
    g = got.detach().cpu().double().numpy() if torch.is_tensor(got) else np.asarray(got, dtype=np.float64)
    w = want.detach().cpu().double().numpy() if torch.is_tensor(want) else np.asarray(want, dtype=np.float64)
    assert g.shape == w.shape, (g.shape, w.shape)
    err = np.abs(g - w)
    bad = err > atol
    st = {"max_err": float(err.max()) if err.size else 0.0, "atol": float(atol), "frac_bad": float(bad.mean()) if err.size else 0.0,
          "n": int(err.size), "worst_row_frac": 0.0, "worst_col_frac": 0.0, "bad_lines": 0, "finite": bool(np.isfinite(g).all())}
    if err.ndim >= 2 and err.shape[-1] >= 8 and err.shape[-2] >= 8 and bad.any():
        b3 = bad.reshape(-1, err.shape[-2], err.shape[-1])
        rows, cols = b3.mean(2), b3.mean(1)
        st["worst_row_frac"] = float(rows.max())
        st["worst_col_frac"] = float(cols.max())
        st["bad_lines"] = int((rows > 0.5).sum() + (cols > 0.5).sum())  # image rows / columns that are mostly wrong
    return st


def bounded_check(got, want, atol, what, allow_frac=0.0, cap=100.0, max_bad_lines=0):
    """|got - want| <= atol except for a fraction `allow_frac` of knife-edge elements (|x| sign kinks, clamp / min / floor
    decisions that flip under rounding-level differences).  The exemptions are bounded too: no element may be off by more
    than cap * atol, and at most `max_bad_lines` image rows / columns of the whole tensor may be mostly (> 50 %) beyond
    tolerance — a wrong row or column of a large tensor stays below any fraction gate; a border / pad / ring-phase bug
    hits a line of EVERY plane and image, far more than the one or two lines a legitimate knife edge can touch (a plane
    whose border tap weight is ~1e-4, so that the sigma clamp gate of its border column hinges on the position noise)."""
    st = error_stats(got, want, atol)
    if os.environ.get("PD_TEST_REPORT"):
        REPORT.append((what, st))
        return st
    assert st["finite"], "%s: non-finite values" % what
    assert st["frac_bad"] <= allow_frac, "%s: %.3g of %d elements off by more than %.1e (max err %.3e)" % (
        what, st["frac_bad"], st["n"], atol, st["max_err"])
    assert st["max_err"] <= cap * atol, "%s: an exempt element is off by %.3e > %g x tolerance %.1e" % (what, st["max_err"], cap, atol)
    assert st["bad_lines"] <= max_bad_lines, "%s: %d image rows / columns are mostly beyond tolerance (worst row %.0f%%, worst column %.0f%%)" % (
        what, st["bad_lines"], 100 * st["worst_row_frac"], 100 * st["worst_col_frac"])
    return st
