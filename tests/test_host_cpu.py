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
