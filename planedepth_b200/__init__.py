"""planedepth_b200 — B200-native photometric-reconstruction path of PlaneDepth.

Public surface (mirrors the reference seam):
    HotPathMixin.pred_novel_images / generate_images_pred / compute_losses /
    generate_post_process_disp, decoder_tail                                (boundary.py)
    warp_composite / photometric_loss / smooth_loss / plane_tail /
    occlusion_masks / resize_frames_u8                                      (functional.py, autograd over the C ABI)
    PerceptualSchedule                                                      (perceptual.py)
    GraphedStep / make_step                                                 (graph.py)
    shard helpers, freeze_unused_parameters                                 (dist.py)

The reference's layers.py classes on the path (HomographyWarp, BackprojectDepth, Project3D, SSIM; trainer.py:152-162) have
no stand-alone counterpart: their grids are never materialised -- the kernels form the sample positions in registers.
"""
from ._lib import PlaneDepthLibraryError, build_library, lib  # noqa: F401

__all__ = ["PlaneDepthLibraryError", "build_library", "lib"]
__version__ = "0.1.0"
