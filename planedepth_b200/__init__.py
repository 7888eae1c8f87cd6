"""planedepth_b200 — B200-native photometric-reconstruction path of PlaneDepth.

Public surface (mirrors the reference seam):
    HotPathMixin.pred_novel_images / generate_images_pred / compute_losses   (boundary.py)
    warp_composite(...) / photometric_loss(...)                             (functional.py)
    HomographyWarp / BackprojectDepth / Project3D / SSIM                    (layers.py)
"""
from ._lib import PlaneDepthLibraryError, build_library, lib  # noqa: F401

__all__ = ["PlaneDepthLibraryError", "build_library", "lib"]
__version__ = "0.1.0"
