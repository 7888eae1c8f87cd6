"""GPU (-m gpu): parity at the BASELINE.json shapes, with the persistent loops of the streamed kernels iterating.

bench.py's own synthetic batches (planedepth_b200.synthetic.make_batch, the seeds bench.py uses) go through the CPU oracle
and through the CUDA path; forward tensors, the per-pixel NLL maps, every loss entry and every gradient are compared.

* cfg2 full: B=12, 640x192, N=49, 0.85 SSIM + 0.15 L1, dense all-ones mask, with and without the x-constancy promise,
  eagerly and through the CUDA graph bench.py replays (592 / 444 CTAs over 2304 rows: 4 - 6 row groups per CTA).
* cfg3: B=4, 1280x384, N=49, Laplacian mixture + plane_residual (nll map, d/d disparity).
* cfg4 per-GPU shape: N=49+14, homography warp, target sides [r, -1, 1], automask L1.
* cfg5 per-GPU shape: N=49+14, 1280x384, mixture + mask_novel blend.
* a cheap variant: one CTA per SM (pd_set_tuning) at B=2, so that every CTA walks >= 3 row groups in seconds.
* the same shapes with SMOOTH (network-like, low-pass) fields instead of iid noise.

Gates (north_star: 1e-4 fp32).  Two classes of kernels:

EXACT (PD_FLAG_EXACT_COORDS: bit-faithful row kernels; general kernels): everything at 1e-4 — forward tensors and the
per-pixel NLL absolute, losses 1e-5, every gradient (plane parameters included) 1e-4 of its tensor's maximum.  Exempt:
at most 1e-5 of the gradient elements (|pred - tgt| sign kinks of the L1 term where the residual is below the forward
rounding noise), each bounded by 2.5 x the tensor's maximum (a sign flip), never half a row or column.

DEFAULT (streamed kernels; sample positions u = x + d exact instead of the reference's fp32 normalise / un-normalise round
trip, DESIGN.md deviations: <= 6e-5 px at W = 640, 1.2e-4 px at W = 1280).  On the iid-noise batches of bench.py that
position noise meets the steepest possible fields (colour slope up to 1 / px, logit slope ~ 2 / px, sigma down to 0.01):
forward 1e-4 on all but 1e-4 of the pixels (none beyond 1e-3), per-pixel NLL within delta_u * slope / sigma_min = 2e-2,
gradients 1e-4 of their maximum on all but 1e-4 of the elements, plane-parameter gradients (sums over every pixel, the
1/sigma-amplified ones included) 5e-3; measured values are in profiles/r2_parity_fullsize.json.  On smooth (network-like)
fields the default kernels meet the EXACT gates except where 1/sigma amplifies the position noise: per-pixel NLL 1e-3
(sigma_min = 0.01 turns 1.2e-4 px x slope 0.05 into 6e-4), plane-parameter gradients 5e-4."""
import json
import os

import pytest
import torch

from helpers import REPORT, bounded_check
from oracle import pd_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-4

SHAPES = {
    # name: (H, W, opt overrides, photometric, mask_novel)
    "cfg2": (192, 640, dict(), "ssim_l1", False),
    "cfg3": (384, 1280, dict(use_mixture_loss=True, plane_residual=True), None, False),
    "cfg4": (192, 640, dict(warp_type="homography_warp", xz_levels=14, novel_frame_ids=[-1, 1], automask=True), None, False),
    "cfg5": (384, 1280, dict(use_mixture_loss=True, plane_residual=True, xz_levels=14), None, True),
    # narrow stand-ins that keep the persistent-loop test cheap
    "cfg2_l1": (192, 640, dict(), None, False),
    "cfg3_small": (192, 1280, dict(use_mixture_loss=True, plane_residual=True, automask=True), None, False),
}


def _build(name, B, seed, device):
    from planedepth_b200.synthetic import make_batch, make_opt

    H, W, over, photometric, mnov = SHAPES[name]
    opt = make_opt(**over)
    b = make_batch(B, H, W, opt, seed=seed, device=device, mask_novel=mnov)
    g = torch.Generator().manual_seed(seed + 1)
    disp = (1 + 20 * torch.rand(B, 1, H, W, generator=g)).to(device).requires_grad_(True)
    b.outputs["disp"] = disp
    b.leaves["disp"] = disp
    return opt, b, photometric


def _smooth(t, g, amp, offset=0.0, lo=None, hi=None):
    """Low-pass field with the statistics of a network output: bicubic upsampling of coarse (1/16) noise."""
    B, C, H, W = t.shape
    coarse = torch.randn(B, C, max(H // 16, 2), max(W // 16, 2), generator=g)
    f = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=True) * amp + offset
    return f.clamp(lo, hi) if lo is not None else f


def build(name, B, seed, device, smooth=False):
    opt, b, photometric = _build(name, B, seed, device)
    if smooth:
        g = torch.Generator().manual_seed(seed + 2)
        for k in list(b.inputs):
            if isinstance(k, tuple) and k[0] in ("color", "color_aug"):
                b.inputs[k] = _smooth(b.inputs[k].cpu(), torch.Generator().manual_seed(seed + 3 + sum(map(ord, str(k[1])))), 0.25, 0.5, 0.0, 1.0).to(device)
        mask = b.outputs["padding_mask"].detach().cpu().float()
        lg = (_smooth(b.outputs["logits"].detach().cpu(), g, 1.5) * mask).contiguous().to(device).requires_grad_(True)
        b.outputs["logits"] = b.leaves["logits"] = lg
        b.outputs["probability"] = lg.detach()
        if "sigma" in b.outputs:
            sg = torch.sigmoid(_smooth(b.outputs["sigma"].detach().cpu(), g, 1.0)).clamp(0.01, 1.0).to(device).requires_grad_(True)
            b.outputs["sigma"] = b.leaves["sigma"] = sg
    return opt, b, photometric


def oracle_run(name, B, seed, smooth=False):
    opt, b, photometric = build(name, B, seed, "cpu", smooth)
    out = b.attach(dict(b.outputs))
    losses = O.hot_path(opt, b.target_sides, b.inputs, out, None, loss_mode=photometric)
    grads = torch.autograd.grad(losses["loss/total_loss"], list(b.leaves.values()), allow_unused=True)
    res = {"losses": {k: float(v) for k, v in losses.items()}, "grads": dict(zip(b.leaves.keys(), grads)), "sides": b.target_sides}
    for s in b.target_sides:
        res[("rgb_rec", s)] = out[("rgb_rec", s)].detach()
        if opt.use_mixture_loss:
            tgt = b.inputs[("color", s)]
            err = (out[("rgb_rec_layered", s)] - tgt[:, None]).abs().mean(2)
            res[("nll", s)] = O.laplacian_mixture_nll(err, out[("sigma_rec", s)], out[("pi_rec", s)]).detach()
            res["sigma_min"] = float(out[("sigma_rec", s)].min())
    return res


def cuda_run(name, B, seed, exact=False, rowwise=True, graph=False, smooth=False, bf16=False):
    from planedepth_b200.boundary import HotPath
    from planedepth_b200.graph import GraphedStep, make_step

    opt, b, photometric = build(name, B, seed, "cuda", smooth)
    if bf16:  # network outputs stored as bf16 (leaves included: the gradients come back as bf16)
        for k in ("logits", "sigma"):
            if k in b.leaves:
                b.leaves[k] = b.leaves[k].detach().to(torch.bfloat16).requires_grad_(True)
                b.outputs[k] = b.leaves[k]
    hp = HotPath(opt, b.target_sides, pc_net=None, photometric=photometric, exact_coords=exact, disp_rowwise=rowwise)
    keys = list(b.leaves.keys())
    leaves = [b.leaves[k] for k in keys]
    if graph:
        step = make_step(hp, b.inputs, b.outputs, leaves, b.attach)
        gs = GraphedStep(step, warmup=2)
        for _ in range(2):
            r = gs.replay()
        torch.cuda.synchronize()
        grads = [r.get("grad%d" % i) for i in range(len(leaves))]
        return {"losses": {"loss/total_loss": float(r["loss"])}, "grads": dict(zip(keys, grads)), "sides": b.target_sides}
    out = b.attach(dict(b.outputs))
    losses = hp.process(b.inputs, out)
    grads = torch.autograd.grad(losses["loss/total_loss"], leaves, allow_unused=True)
    torch.cuda.synchronize()
    res = {"losses": {k: float(v) for k, v in losses.items()}, "grads": dict(zip(keys, grads)), "sides": b.target_sides}
    for s in b.target_sides:
        res[("rgb_rec", s)] = out[("rgb_rec", s)].detach()
        if opt.use_mixture_loss:
            res[("nll", s)] = out[("nll_rec", s)].detach()
    return res


REDUCED = ("disp_base", "xz_h")  # plane-parameter gradients: sums over all H*W pixels


def compare(tag, want, got, exact, noise=True):
    """`exact`: EXACT-class gates; otherwise DEFAULT-class gates on iid noise (`noise`) or on smooth fields."""
    strict = exact or not noise
    nll_tol = TOL if exact else (2e-2 if noise else 1e-3)
    red_tol = TOL if exact else (5e-3 if noise else 5e-4)
    for s in want["sides"]:
        if ("rgb_rec", s) in got:
            bounded_check(got[("rgb_rec", s)], want[("rgb_rec", s)], TOL, "%s rgb_rec@%s" % (tag, s), allow_frac=(1e-5 if strict else 1e-4), cap=10)
        if ("nll", s) in want and ("nll", s) in got:
            bounded_check(got[("nll", s)], want[("nll", s)], nll_tol, "%s nll@%s" % (tag, s), allow_frac=0.0)
    for k, v in got["losses"].items():
        bounded_check(torch.tensor(v), torch.tensor(want["losses"][k]), 1e-5, "%s %s" % (tag, k))
    for k, gw in want["grads"].items():
        if gw is None:
            continue
        gg = got["grads"][k]
        assert gg is not None, "%s: no CUDA gradient for %s" % (tag, k)
        scale = float(gw.abs().max()) + 1e-12
        if k in REDUCED:
            bounded_check(gg, gw, red_tol * scale, "%s grad_%s" % (tag, k))
        else:
            bounded_check(gg, gw, TOL * scale, "%s grad_%s" % (tag, k), allow_frac=(1e-5 if strict else 1e-4), cap=2.5 / TOL,
                          max_bad_lines=(0 if exact else 2))


@pytest.fixture(scope="module", autouse=True)
def dump_report():
    yield
    if os.environ.get("PD_TEST_REPORT"):
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.path.join("gpurun_out", os.environ["PD_TEST_REPORT"]), "w") as f:
            json.dump([[w, s] for w, s in REPORT], f, indent=1)


def test_cfg2_full_size_all_gradients():
    """BASELINE configs[1] exactly as bench.py runs it (seed 1234, B=12): eager and CUDA-graph replay, promise on / off,
    and the bit-faithful kernels."""
    want = oracle_run("cfg2", 12, 1234)
    compare("cfg2 fused", want, cuda_run("cfg2", 12, 1234), exact=False)
    compare("cfg2 fused graph", want, cuda_run("cfg2", 12, 1234, graph=True), exact=False)
    compare("cfg2 fused nopromise", want, cuda_run("cfg2", 12, 1234, rowwise=False), exact=False)
    compare("cfg2 fused nopromise graph", want, cuda_run("cfg2", 12, 1234, rowwise=False, graph=True), exact=False)
    compare("cfg2 exact", want, cuda_run("cfg2", 12, 1234, exact=True), exact=True)


def test_cfg3_full_size_mixture_residual():
    want = oracle_run("cfg3", 4, 1234)
    compare("cfg3 fused", want, cuda_run("cfg3", 4, 1234), exact=False)
    compare("cfg3 fused graph", want, cuda_run("cfg3", 4, 1234, graph=True), exact=False)
    compare("cfg3 exact", want, cuda_run("cfg3", 4, 1234, exact=True), exact=True)


def test_cfg4_shape_homography_three_sides():
    want = oracle_run("cfg4", 2, 1234)
    compare("cfg4 fast", want, cuda_run("cfg4", 2, 1234), exact=False)
    # the general kernels use the reference's arithmetic, but the 3x3 inverse (LU on the CPU, LU on the GPU) and the fp32
    # homography products differ in the last bits between the two devices: forward noise ~5e-5, DEFAULT-class gates
    compare("cfg4 general", want, cuda_run("cfg4", 2, 1234, exact=True), exact=False)


def test_cfg5_shape_mixture_mask_novel():
    want = oracle_run("cfg5", 1, 1234)
    compare("cfg5 fused", want, cuda_run("cfg5", 1, 1234), exact=False)
    compare("cfg5 fused nopromise", want, cuda_run("cfg5", 1, 1234, rowwise=False), exact=False)
    compare("cfg5 exact", want, cuda_run("cfg5", 1, 1234, exact=True), exact=True)


@pytest.mark.parametrize("name,B", [("cfg2_l1", 2), ("cfg2", 2), ("cfg3_small", 2)])
def test_persistent_loop_iterates(name, B):
    """One CTA per SM: 148 CTAs over B*H rows -> every CTA walks >= 3 row groups (ring phases wrap, the coefficient /
    source double buffers and the backward's exchange buffers are reused across groups)."""
    from planedepth_b200 import _lib

    want = oracle_run(name, B, 77)
    H = SHAPES[name][0]
    assert B * H >= 2 * 148
    with _lib.tuned(stream_ctas_per_sm=1):
        got = cuda_run(name, B, 77)
        got_np = cuda_run(name, B, 77, rowwise=False)
    compare(name + " 1cta/sm", want, got, exact=False)
    compare(name + " 1cta/sm nopromise", want, got_np, exact=False)
    with _lib.tuned(stream_ctas_per_sm=1, stream_nst=2, stream_hs=3):
        got = cuda_run(name, B, 77)
    compare(name + " 1cta/sm ring 2x3", want, got, exact=False)


@pytest.mark.parametrize("name,B", [("cfg2", 4), ("cfg3", 1), ("cfg5", 1), ("cfg4", 1)])
def test_smooth_fields_meet_the_exact_gates_in_default_mode(name, B):
    """Network-like (low-pass) logits / sigma / images: the default kernels' exact sample positions are then
    indistinguishable from the reference's round trip at the 1e-4 gate, plane-parameter gradients included."""
    want = oracle_run(name, B, 4321, smooth=True)
    compare(name + " smooth default", want, cuda_run(name, B, 4321, smooth=True), exact=False, noise=False)


@pytest.mark.parametrize("name,B,one_cta", [("cfg2", 4, False), ("cfg2", 2, True), ("cfg3", 1, False), ("cfg3_small", 2, True)])
def test_bf16_storage_at_full_size(name, B, one_cta):
    """pd_warp_desc.dtype = PD_DTYPE_BF16 at the BASELINE widths, with the persistent loops iterating: the bf16-storage kernels
    against the SAME kernels fed through the upcast path (identical fp32 arithmetic on identical values: forward equal to
    rounding noise, gradients equal up to their final bf16 rounding)."""
    from planedepth_b200 import _lib, functional

    def run():
        if one_cta:
            with _lib.tuned(stream_ctas_per_sm=1):
                return cuda_run(name, B, 55, bf16=True)
        return cuda_run(name, B, 55, bf16=True)

    seen = []
    orig = functional._supports
    functional._supports = lambda *a, **k: (seen.append(orig(*a, **k)) or seen[-1])
    try:
        got = run()
    finally:
        functional._supports = orig
    assert seen and all(seen), "the library did not take the bf16 storage path"
    functional.BF16_STORAGE = False
    try:
        want = run()
    finally:
        functional.BF16_STORAGE = True
    for s_ in want["sides"]:
        bounded_check(got[("rgb_rec", s_)], want[("rgb_rec", s_)], 1e-6, "%s bf16 rgb_rec@%s" % (name, s_), allow_frac=1e-5, cap=100)
        if ("nll", s_) in want:
            bounded_check(got[("nll", s_)], want[("nll", s_)], 1e-5, "%s bf16 nll@%s" % (name, s_), allow_frac=1e-5, cap=100)
    for k, v in want["losses"].items():
        bounded_check(torch.tensor(got["losses"][k]), torch.tensor(v), 1e-6, "%s bf16 %s" % (name, k))
    for k, gw in want["grads"].items():
        if gw is None:
            continue
        gg = got["grads"][k]
        assert gg.dtype == gw.dtype
        scale = float(gw.float().abs().max()) + 1e-12
        # one bf16 ulp (2^-8 relative to the tensor's maximum at most) where the two fp32 values straddle a rounding boundary
        bounded_check(gg.float(), gw.float(), 2.0 ** -8 * scale, "%s bf16 grad_%s" % (name, k), allow_frac=1e-5, cap=4)
