"""CPU: host-side logic of the boundary that needs no GPU — the module imports, the drop-in surface is complete, and the
product path refuses CPU tensors instead of falling back."""
import inspect
from types import SimpleNamespace

import pytest
import torch


def test_boundary_imports_and_exposes_the_reference_surface():
    from planedepth_b200 import boundary as B

    for name in ("pred_novel_images", "generate_images_pred", "compute_losses", "perceptual_loss", "generate_post_process_disp",
                 "post_process_disp"):
        assert callable(getattr(B.HotPathMixin, name)), name
    # same positional signatures as trainer.py:523, :701, :404
    assert list(inspect.signature(B.HotPathMixin.pred_novel_images).parameters) == ["self", "inputs", "outputs"]
    assert list(inspect.signature(B.HotPathMixin.compute_losses).parameters) == ["self", "inputs", "outputs"]
    assert list(inspect.signature(B.HotPathMixin.generate_post_process_disp).parameters) == ["self", "inputs"]
    hp = B.HotPath(SimpleNamespace(warp_type="disp_warp"), ["r"])
    assert hp.target_sides == ["r"] and hp.disp_rowwise is False and hp.exact_coords is False


def test_no_cpu_fallback():
    """A CPU tensor must raise, not run somewhere else."""
    from planedepth_b200 import _lib
    from planedepth_b200.functional import smooth_loss

    with pytest.raises(_lib.PlaneDepthLibraryError):
        smooth_loss(torch.rand(1, 1, 4, 8), torch.rand(1, 3, 4, 8))


def test_compact_expand_base_recovers_the_decoders_stride0_expand():
    """depth_decoder.py:156 hands out ``disp_layered`` as a stride-0 expand of [B,N,1,1]: the boundary finds that base so that
    the kernels reduce the gradient into the compact shape; anything else is left alone."""
    from planedepth_b200.functional import _strides4, compact_expand_base

    base = torch.rand(2, 5, 1, 1)
    ex = base.expand(2, 5, 6, 8)
    got = compact_expand_base(ex)
    assert got.shape == (2, 5, 1, 1) and got.data_ptr() == base.data_ptr()
    s = _strides4(got)
    assert (s.y, s.x) == (0, 0) and s.n == 1 and s.b == 5
    dense = torch.rand(2, 5, 6, 8)
    assert compact_expand_base(dense) is dense
    rows = torch.rand(2, 5, 6, 1).expand(2, 5, 6, 8)  # per-row values (xz planes)
    r = compact_expand_base(rows)
    assert r.shape == (2, 5, 6, 1) and _strides4(r).x == 0
    assert compact_expand_base(dense[:, :, ::2]).shape == (2, 5, 3, 8)  # a strided slice is not an expand


def test_rowwise_view_gradient_sums_to_the_x_reduced_gradient():
    """``disp_rowwise``: the dense, x-constant ``cat`` is read at column 0 and the gradient comes back spread over x, so that
    any x-constant producer receives exactly the x-reduced gradient."""
    from planedepth_b200.boundary import rowwise_view

    h = torch.rand(2, 3, 4, 1, requires_grad=True)
    dense = (h * 2.0).expand(2, 3, 4, 8).contiguous()  # what a cat of x-constant planes looks like
    v = rowwise_view(dense)
    assert v.stride(3) == 0 and torch.equal(v, dense)
    w = torch.rand(2, 3, 4, 8)
    (v * w).sum().backward()
    assert torch.allclose(h.grad, 2.0 * w.sum(3, keepdim=True), atol=1e-6)


def test_homography_params_match_the_reference_formula():
    """boundary.homography_params = layers.py:206-226 up to the projective divide: H_t2s = inv(K (R + t n^T / d) K^-1), R n."""
    from planedepth_b200.boundary import homography_params
    from oracle import pd_oracle as O

    g = torch.Generator().manual_seed(3)
    B, N, H, W = 2, 4, 24, 32
    K = torch.tensor([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])[None].repeat(B, 1, 1)
    iK = torch.linalg.pinv(K)
    T = torch.eye(4)[None].repeat(B, 1, 1)
    T[:, :3, 3] = 0.05 * torch.randn(B, 3, generator=g)
    T[:, 0, 1], T[:, 1, 0] = 0.02, -0.02
    dist = 0.5 + 5 * torch.rand(B, N, generator=g)
    nrm = torch.nn.functional.normalize(torch.tensor([0.0, 0.0, 1.0]) + 0.2 * torch.randn(B, N, 3, generator=g), dim=-1)
    hmat, cam = homography_params(dist, nrm, T, K, iK)
    assert hmat.shape == (B * N, 12) and cam.shape == (B, 9)
    u, v, mask = O.homography_coords(dist, nrm, T, K, iK, H, W)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    p = torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(3, -1)
    q = hmat[:, :9].reshape(B * N, 3, 3) @ p
    uu = (q[:, 0] / q[:, 2].clamp_min(1e-7)).reshape(B, N, H, W)
    vv = (q[:, 1] / q[:, 2].clamp_min(1e-7)).reshape(B, N, H, W)
    sane = u.abs() < 1e4
    assert torch.allclose(uu[sane], u[sane], rtol=1e-4, atol=2e-3) and torch.allclose(vv[sane], v[sane], rtol=1e-4, atol=2e-3)


def test_compute_losses_raises_on_missing_terms_like_the_reference():
    """a mis-wired integration (no pc_net, no outputs["disp"]) must not silently train on a different total loss;
    the stand-alone HotPath carrier opts out explicitly."""
    from types import SimpleNamespace

    import pytest
    import torch

    from planedepth_b200.boundary import HotPath, HotPathMixin

    class Carrier(HotPathMixin):
        pass

    c = Carrier()
    c.opt = SimpleNamespace(warp_type="disp_warp")
    c.target_sides = ["r"]
    outputs = {"probability": torch.zeros(1, 2, 4, 8)}
    with pytest.raises(AttributeError, match="pc_net"):
        c.compute_losses({}, outputs)
    c.pc_net = lambda x: [x, x, x]
    with pytest.raises(KeyError, match="disp"):
        c.compute_losses({}, outputs)
    assert HotPath(c.opt, ["r"]).skip_missing_terms is True and HotPathMixin.skip_missing_terms is False
