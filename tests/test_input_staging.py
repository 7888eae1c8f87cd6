"""Device-side input staging: pd_resize_bicubic_u8 against the reference loader's transform
(/root/reference/datasets/pair_transforms.py:63-78 Resize, :28-48 RandomResizeCrop = ToTensor + F.interpolate(bicubic,
align_corners=True) + crop + clamp).  The oracle's explicit-gather restatement is pinned against torch's F.interpolate on
the CPU (third-party arithmetic, like grid_sample); the CUDA kernel is compared with the oracle on the GPU."""
import pytest
import torch
import torch.nn.functional as F

from oracle import pd_oracle as O

CASES = [
    # Hs, Ws, (H, W), full_size, crop
    (375, 1242, (192, 640), None, (0, 0)),       # KITTI frame -> stage-1 / stage-2 size (Resize)
    (375, 1242, (384, 1280), None, (0, 0)),      # HR fine-tuning: up-sampling
    (375, 1242, (192, 640), (233, 771), (17, 64)),  # RandomResizeCrop: factor 0.621, window inside the resized frame
    (37, 53, (20, 48), None, (0, 0)),            # ragged small sizes
    (24, 40, (24, 40), None, (0, 0)),            # identity size
]


def reference_transform(frames_u8, size, full_size, crop):
    x = frames_u8.float() / 255.0  # ToTensor
    H, W = size
    Hf, Wf = (H, W) if full_size is None else full_size
    y = F.interpolate(x, size=(Hf, Wf), mode="bicubic", align_corners=True).clamp(0.0, 1.0)
    return y[:, :, crop[0]:crop[0] + H, crop[1]:crop[1] + W]


@pytest.mark.parametrize("case", CASES)
def test_oracle_resize_matches_torch_interpolate(case):
    Hs, Ws, size, full, crop = case
    g = torch.Generator().manual_seed(Hs + Ws)
    frames = torch.randint(0, 256, (2, 3, Hs, Ws), generator=g, dtype=torch.uint8)
    want = reference_transform(frames, size, full, crop)
    got = O.resize_frames_u8(frames, size, full, crop)
    assert (got - want).abs().max().item() <= 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["hwc", "chw"])
@pytest.mark.parametrize("case", CASES)
def test_cuda_resize_matches_oracle(case, layout):
    from planedepth_b200.functional import resize_frames_u8

    Hs, Ws, size, full, crop = case
    g = torch.Generator().manual_seed(Hs * 3 + Ws)
    frames = torch.randint(0, 256, (2, 3, Hs, Ws), generator=g, dtype=torch.uint8)
    want = O.resize_frames_u8(frames, size, full, crop)
    dev = frames.cuda()
    if layout == "hwc":
        dev = dev.permute(0, 2, 3, 1).contiguous()
    got = resize_frames_u8(dev, size, full, crop)
    assert got.shape == want.shape and got.dtype == torch.float32
    err = (got.cpu() - want).abs().max().item()
    assert err <= 1e-5, err  # fp32 sums of 16 taps in a different association; values in [0, 1]
    # the clamp is part of the transform: bicubic overshoot never leaves [0, 1]
    assert got.min().item() >= 0.0 and got.max().item() <= 1.0
