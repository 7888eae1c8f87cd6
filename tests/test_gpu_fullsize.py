"""GPU (-m gpu): parity at the BASELINE.json shapes, with the persistent loops of the streamed kernels iterating.

bench.py's own synthetic batches (planedepth_b200.synthetic.make_batch, the seeds bench.py uses) go through the CPU oracle
and through the CUDA path; forward tensors, the per-pixel NLL maps, every loss entry and every gradient are compared.

* cfg2 full: B=12, 640x192, N=49, 0.85 SSIM + 0.15 L1, dense all-ones mask, with and without the x-constancy promise,
  eagerly and through the CUDA graph bench.py replays (592 / 444 CTAs over 2304 rows: 4 - 6 row groups per CTA).
* cfg3: B=4, 1280x384, N=49, Laplacian mixture + plane_residual (nll map, d/d disparity).
* cfg4 per-GPU shape: N=49+14, homography warp, target sides [r, -1, 1], automask L1.
* cfg5 per-GPU shape: N=49+14, 1280x384, mixture + mask_novel blend.
* a cheap variant: one CTA per SM (pd_set_tuning) at B=2, so that every CTA walks >= 3 row groups in seconds.

Tolerances (north_star: 1e-4 fp32): forward tensors 1e-4 absolute, losses 1e-4, gradients 1e-4 of the tensor's maximum.
``bounded_check`` additionally caps the knife-edge exemptions (no element beyond 1e-2 of the scale, no row mostly wrong).
The bit-faithful mode (PD_FLAG_EXACT_COORDS) must meet the same gates with tighter exemptions; the default mode samples at
the exact positions u = x + d (DESIGN.md deviations): its per-pixel NLL inherits delta_u * |d colour / du| / sigma, so the
NLL map of the default mode is gated at 1e-4 / sigma_min of the case and reported, the exact mode at 1e-4."""
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


def build(name, B, seed, device):
    from planedepth_b200.synthetic import make_batch, make_opt

    H, W, over, photometric, mnov = SHAPES[name]
    opt = make_opt(**over)
    b = make_batch(B, H, W, opt, seed=seed, device=device, mask_novel=mnov)
    g = torch.Generator().manual_seed(seed + 1)
    disp = (1 + 20 * torch.rand(B, 1, H, W, generator=g)).to(device).requires_grad_(True)
    b.outputs["disp"] = disp
    b.leaves["disp"] = disp
    return opt, b, photometric


def oracle_run(name, B, seed):
    opt, b, photometric = build(name, B, seed, "cpu")
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


def cuda_run(name, B, seed, exact=False, rowwise=True, graph=False):
    from planedepth_b200.boundary import HotPath
    from planedepth_b200.graph import GraphedStep, make_step

    opt, b, photometric = build(name, B, seed, "cuda")
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


# plane / pose parameter gradients are sums over all H*W pixels (knife-edge pixels included): 5e-4 of their maximum
REDUCED = ("disp_base", "xz_h")


def compare(tag, want, got, exact, nll_tol=None):
    for s in want["sides"]:
        if ("rgb_rec", s) in got:
            bounded_check(got[("rgb_rec", s)], want[("rgb_rec", s)], TOL, "%s rgb_rec@%s" % (tag, s), allow_frac=(2e-5 if exact else 2e-4))
        if ("nll", s) in want and ("nll", s) in got:
            tol = TOL if exact else (nll_tol or TOL)
            bounded_check(got[("nll", s)], want[("nll", s)], tol, "%s nll@%s" % (tag, s), allow_frac=(2e-5 if exact else 2e-4))
    for k, v in got["losses"].items():
        bounded_check(torch.tensor(v), torch.tensor(want["losses"][k]), TOL, "%s %s" % (tag, k))
    for k, gw in want["grads"].items():
        if gw is None:
            continue
        gg = got["grads"][k]
        assert gg is not None, "%s: no CUDA gradient for %s" % (tag, k)
        scale = float(gw.abs().max()) + 1e-12
        tol = (5e-4 if k in REDUCED else TOL) * scale
        bounded_check(gg, gw, tol, "%s grad_%s" % (tag, k), allow_frac=(2e-4 if exact else 2e-3))


@pytest.fixture(scope="module", autouse=True)
def dump_report():
    yield
    if os.environ.get("PD_TEST_REPORT"):
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.path.join("gpurun_out", os.environ["PD_TEST_REPORT"]), "w") as f:
            json.dump([[w, s] for w, s in REPORT], f, indent=1)


def nll_gate(want):
    # default mode: positions differ from the reference's round trip by <= 1.2e-4 px (W = 1280); the per-pixel NLL moves by
    # that times the colour slope (<= 1 for colours in [0,1]) over sigma
    return 1.2e-4 / max(want.get("sigma_min", 1.0), 0.01) + TOL


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
    compare("cfg3 fused", want, cuda_run("cfg3", 4, 1234), exact=False, nll_tol=nll_gate(want))
    compare("cfg3 fused graph", want, cuda_run("cfg3", 4, 1234, graph=True), exact=False)
    compare("cfg3 exact", want, cuda_run("cfg3", 4, 1234, exact=True), exact=True)


def test_cfg4_shape_homography_three_sides():
    want = oracle_run("cfg4", 2, 1234)
    compare("cfg4 fast", want, cuda_run("cfg4", 2, 1234), exact=False)
    compare("cfg4 general", want, cuda_run("cfg4", 2, 1234, exact=True), exact=True)


def test_cfg5_shape_mixture_mask_novel():
    want = oracle_run("cfg5", 1, 1234)
    compare("cfg5 fused", want, cuda_run("cfg5", 1, 1234), exact=False, nll_tol=nll_gate(want))
    compare("cfg5 fused nopromise", want, cuda_run("cfg5", 1, 1234, rowwise=False), exact=False, nll_tol=nll_gate(want))
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
    compare(name + " 1cta/sm", want, got, exact=False, nll_tol=nll_gate(want))
    compare(name + " 1cta/sm nopromise", want, got_np, exact=False, nll_tol=nll_gate(want))
    with _lib.tuned(stream_ctas_per_sm=1, stream_nst=2, stream_hs=3):
        got = cuda_run(name, B, 77)
    compare(name + " 1cta/sm ring 2x3", want, got, exact=False, nll_tol=nll_gate(want))
