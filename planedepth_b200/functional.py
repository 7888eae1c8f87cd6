"""torch.autograd bridges over the C ABI (include/planedepth_b200.h).

PyTorch is used here for device memory, streams and autograd bookkeeping only; every tensor handed to
the library is a raw device pointer.  There is no eager fallback."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, replace
from typing import Optional, Tuple

import torch

from . import _lib as L


#: when set to a list, every C-ABI call appends (name, start_event, end_event) recorded on the launch stream
KERNEL_TIMELINE = None

#: Fold pd_photometric_bwd into the prologue of pd_warp_composite_bwd (pd_warp_grad_out's fused form): the photometric
#: node's backward hands its operands (d loss / d ph_sum, the unit gradient saved by the forward, the perceptual term's
#: gradient) to the warp node that produced its input instead of launching a kernel and writing g_rgb_rec to HBM.
#: True: for the stereo (disp_warp) kernels; "all": every warp type (the ABI supports it everywhere; tests); False: never.
FUSE_PHOTOMETRIC_BWD = True


#: hand bf16 ``logits`` / ``sigma`` to the kernels as bf16 (pd_warp_desc.dtype) where the library serves it — the streamed stereo
#: kernels; gradients come back as bf16 — instead of upcasting them with torch first
BF16_STORAGE = True


class _Link:
    """Side channel between the autograd node that produced ``rgb_rec`` (_WarpComposite) and the photometric node that
    consumes it: autograd runs the consumer's backward first; it parks its operands here."""

    __slots__ = ("pending",)

    def __init__(self):
        self.pending = None


_PLACEHOLDER = {}


def _placeholder(shape, device):
    """Stride-0 zero standing in for a gradient whose value travels through a _Link (no kernel, no memory)."""
    z = _PLACEHOLDER.get(device)
    if z is None:
        z = _PLACEHOLDER[device] = torch.zeros((), device=device)
    return z.expand(shape)


def _is_placeholder(g) -> bool:
    return g is not None and g.numel() > 1 and all(st == 0 for st in g.stride())


def _call(name, fn, *args):
    """Invoke one C-ABI entry point, optionally bracketed by CUDA events on the current stream."""
    tl = KERNEL_TIMELINE
    if tl is None:
        L.check(fn(*args), name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    L.check(fn(*args), name)
    e.record()
    tl.append((name, s, e))


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    """fp32, contiguous, on the device.  Other dtypes (e.g. bf16 network outputs under autocast) are UPCAST here with torch: a
    convenience at the boundary, not a storage format — the kernels compute and stream fp32 only."""
    if t.device.type != "cuda":
        raise L.PlaneDepthLibraryError("%s must be a CUDA tensor (planedepth_b200 has no CPU path)" % what)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _strides4(t: torch.Tensor) -> L.Strides4:
    # size-1 dimensions broadcast: give them stride 0 so the kernel may index them with any coordinate
    s = [0 if t.size(i) == 1 else t.stride(i) for i in range(4)]
    return L.Strides4(*s)


def compact_expand_base(t: torch.Tensor) -> torch.Tensor:
    """If ``t`` is an ``expand`` view of a smaller 4-d tensor (depth_decoder.py:156 hands out
    ``disp_layered`` that way), return that base so gradients arrive in the compact shape instead of a
    dense [B,N,H,W] buffer that autograd would have to sum afterwards.  Otherwise return ``t``."""
    base = t._base if t._is_view() else None
    if base is None or base.dim() != 4 or base.data_ptr() != t.data_ptr() or base.dtype != t.dtype:
        return t
    for i in range(4):
        same = base.size(i) == t.size(i) and (t.size(i) == 1 or base.stride(i) == t.stride(i))
        bcast = base.size(i) == 1 and (t.stride(i) == 0 or t.size(i) == 1)
        if not (same or bcast):
            return t
    return base


def _workspace(lib, desc, dev) -> Optional[torch.Tensor]:
    """Scratch buffer of pd_warp_composite_workspace_bytes(desc) (caller-owned; the caching allocator recycles it)."""
    n = int(lib.pd_warp_composite_workspace_bytes(C.byref(desc)))
    return torch.empty((n + 3) // 4, device=dev, dtype=torch.float32) if n else None


@dataclass(frozen=True)
class WarpConfig:
    warp_type: int
    mixture: bool
    automask: bool
    disp_sign: float = 0.0
    shape: Tuple[int, int, int, int] = (0, 0, 0, 0)  # B,N,H,W
    layered: bool = False  # also materialise the per-plane tensors of trainer.py:582-602 (detached)
    exact_coords: bool = False  # PD_FLAG_EXACT_COORDS: bit-faithful coordinate round trip (slower stereo path)
    bf16: bool = False  # pd_warp_desc.dtype = PD_DTYPE_BF16: logits / sigma (and their gradients) are stored as bf16


class _WarpComposite(torch.autograd.Function):
    """pd_warp_composite_fwd / _bwd: trainer.py:533-603 for one target side."""

    @staticmethod
    def forward(ctx, cfg: WarpConfig, link, src, tgt, logits, sigma, disp, mask, hmat, cam):
        lib = L.lib()
        ctx.link = link
        ctx.set_materialize_grads(False)  # an unused output's gradient arrives as None, not as a zero-filled tensor
        B, N, H, W = cfg.shape
        dev = logits.device
        desc = _warp_desc(cfg, disp, mask)
        tin = L.WarpIn(src=_ptr(src), tgt=_ptr(tgt), logits=_ptr(logits), sigma=_ptr(sigma), disp=_ptr(disp), mask=_ptr(mask),
                       hmat=_ptr(hmat), cam=_ptr(cam))
        rgb_rec = torch.empty(B, 3, H, W, device=dev, dtype=torch.float32)
        stats = torch.empty((lib.pd_warp_composite_stats_bytes(C.byref(desc)) + 3) // 4, device=dev, dtype=torch.float32)
        nll = torch.empty(B, 1, H, W, device=dev, dtype=torch.float32) if cfg.mixture else None
        nll_auto = torch.empty(B, 1, H, W, device=dev, dtype=torch.float32) if (cfg.mixture and cfg.automask) else None
        out = L.WarpOut(rgb_rec=_ptr(rgb_rec), stats=_ptr(stats), nll=_ptr(nll), nll_auto=_ptr(nll_auto))
        layered = None
        if cfg.layered:
            layered = {
                "rgb_rec_layered": torch.empty(B, N, 3, H, W, device=dev),
                "logit_rec": torch.empty(B, N, H, W, device=dev),
                "probability_rec": torch.empty(B, N, H, W, device=dev),
            }
            if cfg.mixture:
                layered["sigma_rec"] = torch.empty(B, N, H, W, device=dev)
                layered["pi_rec"] = torch.empty(B, N, H, W, device=dev)
            for k, v in layered.items():
                setattr(out, k, v.data_ptr())
        _call("pd_warp_composite_fwd", lib.pd_warp_composite_fwd, C.byref(desc), C.byref(tin), C.byref(out), _ptr(_workspace(lib, desc, dev)), _stream())
        ctx.cfg = cfg
        ctx.desc = desc
        ctx.save_for_backward(src, tgt, logits, sigma, disp, mask, hmat, cam, rgb_rec, stats)
        ctx.layered = layered
        outs = [rgb_rec,
                nll if nll is not None else torch.empty(0, device=dev),
                nll_auto if nll_auto is not None else torch.empty(0, device=dev)]
        ctx.mark_non_differentiable(outs[2])
        if not cfg.mixture:
            ctx.mark_non_differentiable(outs[1])
        names = []
        if layered is not None:
            for k, v in layered.items():
                names.append(k)
                outs.append(v)
                ctx.mark_non_differentiable(v)
        ctx.layer_names = names
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_rgb, g_nll, g_nll_auto, *unused):
        lib = L.lib()
        cfg = ctx.cfg
        B, N, H, W = cfg.shape
        src, tgt, logits, sigma, disp, mask, hmat, cam, rgb_rec, stats = ctx.saved_tensors
        dev = logits.device
        need = ctx.needs_input_grad[1:]  # (cfg, src, tgt, logits, sigma, disp, mask, hmat, cam)
        fused = ctx.link.pending if ctx.link is not None else None
        if ctx.link is not None:
            ctx.link.pending = None
        if fused is not None:
            # the photometric node parked its operands: placeholders carry no value, anything else is a genuine extra gradient
            g_rgb = None if (g_rgb is None or _is_placeholder(g_rgb)) else _f32c(g_rgb, "grad rgb_rec")
            if g_nll is not None and (_is_placeholder(g_nll) or not g_nll.numel()):
                g_nll = None
        else:
            g_rgb = torch.zeros_like(rgb_rec) if g_rgb is None else _f32c(g_rgb, "grad rgb_rec")
        if cfg.mixture and g_nll is not None and g_nll.numel():
            g_nll = _f32c(g_nll, "grad nll")
        else:
            g_nll = None
        tin = L.WarpIn(src=_ptr(src), tgt=_ptr(tgt), logits=_ptr(logits), sigma=_ptr(sigma), disp=_ptr(disp), mask=_ptr(mask),
                       hmat=_ptr(hmat), cam=_ptr(cam))
        saved = L.WarpOut(rgb_rec=_ptr(rgb_rec), stats=_ptr(stats))
        gout = L.WarpGradOut(g_rgb_rec=_ptr(g_rgb), g_nll=_ptr(g_nll))
        if fused is not None:
            gout.g_ph_sum, gout.ph_scale = _ptr(fused["g_ph_sum"]), float(fused["ph_scale"])
            gout.g_unit, gout.g_unit_nll = _ptr(fused["g_unit"]), _ptr(fused["g_unit_nll"])
            gout.g_pred, gout.mask_novel = _ptr(fused["g_pred"]), _ptr(fused["mask_novel"])
        g_logits = torch.empty_like(logits) if need[3] else None
        g_sigma = torch.empty_like(sigma) if (cfg.mixture and sigma is not None and need[4]) else None
        g_disp = g_hmat = None
        gin = L.WarpGradIn(g_logits=_ptr(g_logits), g_sigma=_ptr(g_sigma))
        spread = 1
        if disp is not None and need[5]:
            # broadcast dimensions of the input (size 1 or stride 0) are reduced inside the kernel: the buffer is
            # compact there, and the result is handed back spread evenly over the broadcast extent (a stride-0
            # view), so that autograd's expand-backward sums it to the reduced value again
            shape = [1 if (disp.size(i) == 1 or disp.stride(i) == 0) else disp.size(i) for i in range(4)]
            for i in range(4):
                if shape[i] == 1:
                    spread *= disp.size(i)
            g_disp = torch.empty(shape, device=dev, dtype=torch.float32)
            gin.g_disp = g_disp.data_ptr()
            gin.g_disp_stride = _strides4(g_disp)
        if hmat is not None and need[7]:
            g9 = torch.empty(B * N, 9, device=dev, dtype=torch.float32)
            gin.g_hmat = g9.data_ptr()
        _call("pd_warp_composite_bwd", lib.pd_warp_composite_bwd, C.byref(ctx.desc), C.byref(tin), C.byref(saved), C.byref(gout),
              C.byref(gin), _ptr(_workspace(lib, ctx.desc, dev)), _stream())
        if hmat is not None and need[7]:
            g_hmat = torch.cat([g9, torch.zeros(B * N, 3, device=dev)], 1)
        if g_disp is not None and tuple(g_disp.shape) != tuple(disp.shape):
            g_disp = (g_disp / spread if spread > 1 else g_disp).expand(disp.shape)
        return (None, None, None, None, g_logits, g_sigma, g_disp, None, g_hmat, None)


class _WarpCompositeSides(torch.autograd.Function):
    """pd_warp_composite_fwd / _bwd for ALL target sides of a homography warp in one autograd node (trainer.py:528-603: the loop
    over ``self.target_sides``).  The forward is the per-side call; the backward zero-fills ``g_logits`` / ``g_sigma`` once and
    lets every side's scatter kernel add into them (``PD_FLAG_ACCUMULATE``) -- per-side nodes leave that sum to autograd, which
    costs one [B,N,H,W] zero-fill and one [B,N,H,W] add per extra side."""

    @staticmethod
    def forward(ctx, cfg: WarpConfig, nsides: int, src, logits, sigma, cam, *per_side):
        lib = L.lib()
        ctx.set_materialize_grads(False)
        B, N, H, W = cfg.shape
        dev = logits.device
        tgts, hmats = per_side[:nsides], per_side[nsides:]
        desc = _warp_desc(cfg, None, None)
        ws = _workspace(lib, desc, dev)  # the homography fast path's packed copy of `src`: built by the first side, reused by the rest
        ready = L.WarpDesc.from_buffer_copy(desc)
        ready.flags |= L.PD_FLAG_WORKSPACE_READY
        saved, outs = [], []
        for s_ in range(nsides):
            tin = L.WarpIn(src=_ptr(src), tgt=_ptr(tgts[s_]), logits=_ptr(logits), sigma=_ptr(sigma), hmat=_ptr(hmats[s_]), cam=_ptr(cam))
            rgb_rec = torch.empty(B, 3, H, W, device=dev, dtype=torch.float32)
            stats = torch.empty((lib.pd_warp_composite_stats_bytes(C.byref(desc)) + 3) // 4, device=dev, dtype=torch.float32)
            nll = torch.empty(B, 1, H, W, device=dev, dtype=torch.float32) if cfg.mixture else torch.empty(0, device=dev)
            nll_auto = torch.empty(B, 1, H, W, device=dev, dtype=torch.float32) if (cfg.mixture and cfg.automask) else torch.empty(0, device=dev)
            out = L.WarpOut(rgb_rec=_ptr(rgb_rec), stats=_ptr(stats), nll=_ptr(nll if cfg.mixture else None),
                            nll_auto=_ptr(nll_auto if (cfg.mixture and cfg.automask) else None))
            _call("pd_warp_composite_fwd", lib.pd_warp_composite_fwd, C.byref(desc if s_ == 0 else ready), C.byref(tin), C.byref(out), _ptr(ws), _stream())
            saved += [rgb_rec, stats]
            outs += [rgb_rec, nll, nll_auto]
            ctx.mark_non_differentiable(nll_auto)
            if not cfg.mixture:
                ctx.mark_non_differentiable(nll)
        ctx.cfg, ctx.desc, ctx.nsides, ctx.ws = cfg, desc, nsides, ws  # ws stays private to this node until its backward has run
        ctx.save_for_backward(src, logits, sigma, cam, *tgts, *hmats, *saved)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        lib = L.lib()
        cfg, S = ctx.cfg, ctx.nsides
        B, N, H, W = cfg.shape
        t = ctx.saved_tensors
        src, logits, sigma, cam = t[:4]
        tgts, hmats, saved = t[4:4 + S], t[4 + S:4 + 2 * S], t[4 + 2 * S:]
        dev = logits.device
        need = ctx.needs_input_grad  # (cfg, nsides, src, logits, sigma, cam, tgt x S, hmat x S)
        g_logits = torch.zeros_like(logits) if need[3] else None
        g_sigma = torch.zeros_like(sigma) if (cfg.mixture and sigma is not None and need[4]) else None
        desc = L.WarpDesc.from_buffer_copy(ctx.desc)
        desc.flags |= L.PD_FLAG_ACCUMULATE | (L.PD_FLAG_WORKSPACE_READY if ctx.ws is not None else 0)
        g_hmats = [None] * S
        for s_ in range(S):
            g_rgb, g_nll = grads[3 * s_], grads[3 * s_ + 1]
            g_nll = _f32c(g_nll, "grad nll") if (cfg.mixture and g_nll is not None and g_nll.numel()) else None
            if g_rgb is None and g_nll is None:
                continue  # this side's outputs did not reach the loss
            rgb_rec, stats = saved[2 * s_], saved[2 * s_ + 1]
            g_rgb = torch.zeros_like(rgb_rec) if g_rgb is None else _f32c(g_rgb, "grad rgb_rec")
            tin = L.WarpIn(src=_ptr(src), tgt=_ptr(tgts[s_]), logits=_ptr(logits), sigma=_ptr(sigma), hmat=_ptr(hmats[s_]), cam=_ptr(cam))
            sv = L.WarpOut(rgb_rec=_ptr(rgb_rec), stats=_ptr(stats))
            gout = L.WarpGradOut(g_rgb_rec=_ptr(g_rgb), g_nll=_ptr(g_nll))
            gin = L.WarpGradIn(g_logits=_ptr(g_logits), g_sigma=_ptr(g_sigma))
            g9 = None
            if need[6 + S + s_]:
                g9 = torch.empty(B * N, 9, device=dev, dtype=torch.float32)
                gin.g_hmat = g9.data_ptr()
            _call("pd_warp_composite_bwd", lib.pd_warp_composite_bwd, C.byref(desc), C.byref(tin), C.byref(sv), C.byref(gout), C.byref(gin),
                  _ptr(ctx.ws), _stream())
            if g9 is not None:
                g_hmats[s_] = torch.cat([g9, torch.zeros(B * N, 3, device=dev)], 1)
        return (None, None, None, g_logits, g_sigma, None) + (None,) * S + tuple(g_hmats)


def warp_composite_sides(cfg: WarpConfig, src, tgts, logits, sigma, hmats, cam):
    """Homography warp of several target sides through one autograd node (see _WarpCompositeSides).
    Returns a list of (rgb_rec, nll | None, nll_auto | None) per side."""
    S = len(hmats)
    src = _f32c(src, "src").detach()
    logits = _f32c(logits, "logits")
    sigma = _f32c(sigma, "sigma") if cfg.mixture else None
    tg = [(_f32c(t_, "tgt").detach() if (cfg.mixture and t_ is not None) else None) for t_ in tgts]
    hm = [_f32c(h_, "hmat") for h_ in hmats]
    outs = _WarpCompositeSides.apply(cfg, S, src, logits, sigma, _f32c(cam, "cam").detach(), *tg, *hm)
    res = []
    for s_ in range(S):
        rgb_rec, nll, nll_auto = outs[3 * s_:3 * s_ + 3]
        res.append((rgb_rec, nll if cfg.mixture else None, nll_auto if (cfg.mixture and cfg.automask) else None))
    return res


def _warp_desc(cfg: WarpConfig, disp, mask):
    B, N, H, W = cfg.shape
    desc = L.WarpDesc(B=B, N=N, H=H, W=W, warp_type=cfg.warp_type, mixture=int(cfg.mixture), automask=int(cfg.automask),
                      mask_dtype=L.PD_MASK_NONE, disp_sign=float(cfg.disp_sign), flags=(L.PD_FLAG_EXACT_COORDS if cfg.exact_coords else 0),
                      dtype=(L.PD_DTYPE_BF16 if cfg.bf16 else L.PD_DTYPE_F32))
    if disp is not None:
        desc.disp_stride = _strides4(disp)
    if mask is not None:
        desc.mask_stride = _strides4(mask)
        desc.mask_dtype = L.PD_MASK_F32 if mask.dtype == torch.float32 else L.PD_MASK_U8
    return desc


def _supports(cfg: WarpConfig, src, tgt, logits, sigma, disp, mask, hmat, cam) -> bool:
    desc = _warp_desc(cfg, disp, mask)
    tin = L.WarpIn(src=_ptr(src), tgt=_ptr(tgt), logits=_ptr(logits), sigma=_ptr(sigma), disp=_ptr(disp), mask=_ptr(mask), hmat=_ptr(hmat), cam=_ptr(cam))
    return bool(L.lib().pd_warp_composite_supports(C.byref(desc), C.byref(tin)))


def warp_composite(cfg: WarpConfig, src, tgt, logits, sigma=None, disp=None, mask=None, hmat=None, cam=None):
    """Returns (rgb_rec, nll | None, nll_auto | None, layered dict | None)."""
    src = _f32c(src, "src")
    want_bf16 = (BF16_STORAGE and logits.dtype == torch.bfloat16 and logits.is_cuda and not cfg.layered and not cfg.exact_coords
                 and (not cfg.mixture or (sigma is not None and sigma.dtype == torch.bfloat16)))
    if want_bf16:
        logits = logits.contiguous()
        sigma = sigma.contiguous() if cfg.mixture else None
    else:
        logits = _f32c(logits, "logits")
        sigma = _f32c(sigma, "sigma") if cfg.mixture else None
    tgt = _f32c(tgt, "tgt") if (cfg.mixture and tgt is not None) else None
    if disp is not None:
        if disp.dtype != torch.float32:
            disp = disp.float()
        disp = compact_expand_base(disp)
    if mask is not None:
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        elif mask.dtype not in (torch.float32, torch.uint8):
            mask = mask.float()
        if mask.dim() != 4:
            raise ValueError("padding_mask must be 4-d [B,N,H,W] (broadcastable)")
        mask = mask.detach()
    if hmat is not None:
        hmat = _f32c(hmat, "hmat")
    if cam is not None:
        cam = _f32c(cam, "cam").detach()
    if want_bf16:
        # bf16 storage is served by the streamed stereo kernels only: ask the library, otherwise upcast like any other dtype
        cfg = replace(cfg, bf16=True)
        if not _supports(cfg, src, tgt, logits, sigma, disp, mask, hmat, cam):
            cfg = replace(cfg, bf16=False)
            logits = _f32c(logits, "logits")
            sigma = _f32c(sigma, "sigma") if cfg.mixture else None
    # the side channel to the photometric node exists for the stereo kernels only: their row prologue absorbs the fused form
    # for free, while the thread-per-pixel homography / general backward kernels lose 10 % to it (measured, cfg 4)
    link = _Link() if (cfg.warp_type == L.PD_WARP_DISP or FUSE_PHOTOMETRIC_BWD == "all") else None
    outs = _WarpComposite.apply(cfg, link, src.detach(), None if tgt is None else tgt.detach(), logits, sigma, disp, mask, hmat, cam)
    rgb_rec, nll, nll_auto = outs[:3]
    # photometric_loss() finds the producer through these attributes (the dict contract hands the same objects over)
    if link is not None:
        rgb_rec._pd_link = link
        if cfg.mixture:
            nll._pd_link = link
    layered = None
    if cfg.layered:
        names = ["rgb_rec_layered", "logit_rec", "probability_rec"] + (["sigma_rec", "pi_rec"] if cfg.mixture else [])
        layered = dict(zip(names, outs[3:]))
    return rgb_rec, (nll if cfg.mixture else None), (nll_auto if (cfg.mixture and cfg.automask) else None), layered


class _Photometric(torch.autograd.Function):
    """pd_photometric_fwd / _bwd: trainer.py:720-742 (+687-699) for one target side.  The forward also produces the
    unit gradient d ph_sum / d rgb_rec (d ph_sum / d nll in mixture mode) while its tiles are on chip; the backward
    is a streaming scale-and-add."""

    @staticmethod
    def forward(ctx, mode: int, automask: bool, want_map: bool, scale: float, link, rgb_rec, tgt, src, mask_novel, nll, nll_auto):
        lib = L.lib()
        ctx.link = link
        ctx.set_materialize_grads(False)  # pred without a consumer (no perceptual term): g_pred is None, no fill, no read
        ctx.scale = float(scale)
        B, _, H, W = rgb_rec.shape
        dev = rgb_rec.device
        mixture = mode == L.PD_LOSS_MIXTURE
        desc = L.LossDesc(B=B, H=H, W=W, loss_mode=mode, automask=int(automask), has_mask_novel=int(mask_novel is not None),
                          out_scale=float(scale))
        tin = L.LossIn(rgb_rec=_ptr(rgb_rec), tgt=_ptr(tgt), src=_ptr(src), mask_novel=_ptr(mask_novel), nll=_ptr(nll), nll_auto=_ptr(nll_auto))
        pred = torch.empty_like(rgb_rec) if mask_novel is not None else None
        ph_map = torch.empty(B, 1, H, W, device=dev) if want_map else None
        ph_sum = torch.empty((), device=dev, dtype=torch.float32)
        need_grad = any(ctx.needs_input_grad)
        g_unit = torch.empty_like(rgb_rec) if (need_grad and not mixture) else None
        g_unit_nll = torch.empty(B, 1, H, W, device=dev) if (need_grad and mixture) else None
        ws = torch.empty(lib.pd_photometric_workspace_bytes(C.byref(desc)) // 4, device=dev, dtype=torch.float32)
        out = L.LossOut(pred=_ptr(pred), ph_map=_ptr(ph_map), ph_sum=_ptr(ph_sum), g_unit=_ptr(g_unit), g_unit_nll=_ptr(g_unit_nll))
        _call("pd_photometric_fwd", lib.pd_photometric_fwd, C.byref(desc), C.byref(tin), C.byref(out), ws.data_ptr(), _stream())
        ctx.desc = desc
        ctx.has_pred = pred is not None
        ctx.mixture = mixture
        ctx.shape = (B, H, W)
        ctx.save_for_backward(mask_novel, g_unit, g_unit_nll)
        if pred is None:
            # without the mask_novel blend pred IS rgb_rec; with a link it is still handed out as an output of this node (an
            # alias, no copy) so that the perceptual term's gradient arrives here and rgb_rec keeps a single consumer
            pred = rgb_rec.detach() if link is not None else torch.empty(0, device=dev)
            ctx.has_pred = link is not None
        outs = (ph_sum, pred, ph_map if ph_map is not None else torch.empty(0, device=dev))
        ctx.mark_non_differentiable(outs[2])
        return outs

    @staticmethod
    def backward(ctx, g_sum, g_pred, g_map):
        lib = L.lib()
        mask_novel, g_unit, g_unit_nll = ctx.saved_tensors
        B, H, W = ctx.shape
        dev = (g_unit if g_unit is not None else g_unit_nll).device
        g_sum = torch.zeros((), device=dev) if g_sum is None else _f32c(g_sum, "grad ph_sum")
        g_pred = _f32c(g_pred, "grad pred") if (ctx.has_pred and g_pred is not None) else None
        if ctx.link is not None:
            # fused: the warp node's backward kernel forms g_ph_sum * unit + g_pred * mask_novel in its prologue
            ctx.link.pending = {"g_ph_sum": g_sum, "ph_scale": ctx.scale, "g_unit": g_unit, "g_unit_nll": g_unit_nll, "g_pred": g_pred,
                                "mask_novel": mask_novel}
            return (None, None, None, None, None, _placeholder((B, 3, H, W), dev), None, None, None,
                    _placeholder((B, 1, H, W), dev) if ctx.mixture else None, None)
        tin = L.LossIn(mask_novel=_ptr(mask_novel))
        saved = L.LossOut(g_unit=_ptr(g_unit), g_unit_nll=_ptr(g_unit_nll))
        g_rgb = torch.empty(B, 3, H, W, device=dev, dtype=torch.float32)
        g_nll = torch.empty(B, 1, H, W, device=dev, dtype=torch.float32) if ctx.mixture else None
        gout = L.LossGradOut(g_ph_sum=_ptr(g_sum), g_pred=_ptr(g_pred))
        gin = L.LossGradIn(g_rgb_rec=_ptr(g_rgb), g_nll=_ptr(g_nll))
        _call("pd_photometric_bwd", lib.pd_photometric_bwd, C.byref(ctx.desc), C.byref(tin), C.byref(saved), C.byref(gout), C.byref(gin), None,
              _stream())
        return (None, None, None, None, None, g_rgb, None, None, None, g_nll, None)


def photometric_loss(mode: int, automask: bool, rgb_rec, tgt, src=None, mask_novel=None, nll=None, nll_auto=None, want_map=False,
                     scale: float = 1.0):
    """Returns (scale * ph_sum 0-dim, pred [B,3,H,W], ph_map | None); ``scale = 1/(B*H*W)`` makes the first ``ph.mean()``.  ``pred`` is ``rgb_rec`` itself (same
    autograd node) when there is no ``mask_novel`` blend."""
    rgb_rec = _f32c(rgb_rec, "rgb_rec")
    tgt = _f32c(tgt, "tgt").detach()
    need_src = automask and mode != L.PD_LOSS_MIXTURE
    src = _f32c(src, "src").detach() if need_src else None
    mask_novel = _f32c(mask_novel, "mask_novel").detach() if mask_novel is not None else None
    if mode == L.PD_LOSS_MIXTURE:
        nll = _f32c(nll, "nll")
        nll_auto = _f32c(nll_auto, "nll_auto").detach() if automask else None
    else:
        nll = nll_auto = None
    # fuse with the producing warp node when rgb_rec (and, in mixture mode, nll) come straight from it
    link = getattr(rgb_rec, "_pd_link", None) if FUSE_PHOTOMETRIC_BWD else None
    if link is not None and mode == L.PD_LOSS_MIXTURE and getattr(nll, "_pd_link", None) is not link:
        link = None
    ph_sum, pred, ph_map = _Photometric.apply(mode, automask, want_map, float(scale), link, rgb_rec, tgt, src, mask_novel, nll, nll_auto)
    return ph_sum, (pred if (mask_novel is not None or link is not None) else rgb_rec), (ph_map if want_map else None)


def occlusion_masks(logits, probability, disp_layered, disp, exact_coords: bool = False):
    """pd_occlusion_masks_fwd: trainer.py:421-466 on the decoder outputs of the 2B-image batch ``cat([img, img.flip(-1)])``.
    Returns (disp_pp, mask_novel, o_l, o_fr), [B,1,H,W] each; no gradients (the reference detaches them, :464)."""
    lib = L.lib()
    with torch.no_grad():
        logits = _f32c(logits.detach(), "logits")
        probability = _f32c(probability.detach(), "probability")
        disp = _f32c(disp.detach(), "disp")
        dl = disp_layered.detach()
        if dl.dtype != torch.float32:
            dl = dl.float()
        dl = compact_expand_base(dl)
        B2, N, H, W = logits.shape
        if B2 % 2 or tuple(probability.shape) != (B2, N, H, W) or tuple(disp.shape) != (B2, 1, H, W):
            raise ValueError("occlusion_masks expects the outputs of the 2B-image batch: logits / probability [2B,N,H,W], disp [2B,1,H,W]")
        B = B2 // 2
        dev = logits.device
        desc = L.OcclDesc(B=B, N=N, H=H, W=W, flags=(L.PD_FLAG_EXACT_COORDS if exact_coords else 0), disp_stride=_strides4(dl))
        outs = [torch.empty(B, 1, H, W, device=dev, dtype=torch.float32) for _ in range(4)]
        o_l, o_fr, mask_novel, disp_pp = outs
        ws = torch.empty(lib.pd_occlusion_masks_workspace_bytes(C.byref(desc)) // 4, device=dev, dtype=torch.float32)
        tin = L.OcclIn(logits=_ptr(logits), probability=_ptr(probability), disp_layered=_ptr(dl), disp=_ptr(disp))
        out = L.OcclOut(o_l=_ptr(o_l), o_fr=_ptr(o_fr), mask_novel=_ptr(mask_novel), disp_pp=_ptr(disp_pp))
        _call("pd_occlusion_masks_fwd", lib.pd_occlusion_masks_fwd, C.byref(desc), C.byref(tin), C.byref(out), ws.data_ptr(), _stream())
    return disp_pp, mask_novel, o_l, o_fr


def resize_frames_u8(frames: torch.Tensor, size, full_size=None, crop=(0, 0), out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pd_resize_bicubic_u8: raw uint8 frames ([B,Hs,Ws,3] interleaved or [B,3,Hs,Ws] planar, on the device) ->
    fp32 [B,3,H,W] = ``F.interpolate(frames / 255, full_size, mode="bicubic", align_corners=True)[..., crop window].clamp(0, 1)``
    (datasets/pair_transforms.py:63-78; with ``full_size`` / ``crop``: RandomResizeCrop, :28-48).  No gradients (data)."""
    lib = L.lib()
    if frames.device.type != "cuda":
        raise L.PlaneDepthLibraryError("frames must be a CUDA tensor (planedepth_b200 has no CPU path)")
    if frames.dtype != torch.uint8 or frames.dim() != 4:
        raise ValueError("frames must be uint8 [B,Hs,Ws,3] or [B,3,Hs,Ws]")
    hwc = frames.shape[-1] == 3 and frames.shape[1] != 3
    frames = frames.contiguous()
    B = frames.shape[0]
    Hs, Ws = (frames.shape[1], frames.shape[2]) if hwc else (frames.shape[2], frames.shape[3])
    H, W = int(size[0]), int(size[1])
    Hf, Wf = (H, W) if full_size is None else (int(full_size[0]), int(full_size[1]))
    if out is None:
        out = torch.empty(B, 3, H, W, device=frames.device, dtype=torch.float32)
    desc = L.ResizeDesc(B=B, Hs=Hs, Ws=Ws, Hf=Hf, Wf=Wf, y0=int(crop[0]), x0=int(crop[1]), H=H, W=W, src_layout=0 if hwc else 1)
    _call("pd_resize_bicubic_u8", lib.pd_resize_bicubic_u8, C.byref(desc), frames.data_ptr(), out.data_ptr(), _stream())
    return out


class _SmoothLoss(torch.autograd.Function):
    """pd_smooth_loss_fwd / _bwd: get_smooth_loss_disp (layers.py:243-256) on the crop [..., x0:] (trainer.py:768-771)."""

    @staticmethod
    def forward(ctx, x0: int, gamma: float, disp, img):
        lib = L.lib()
        B, _, H, W = disp.shape
        desc = L.SmoothDesc(B=B, H=H, W=W, x0=int(x0), gamma=float(gamma))
        loss = torch.empty((), device=disp.device, dtype=torch.float32)
        ws = torch.empty(lib.pd_smooth_loss_workspace_bytes(C.byref(desc)) // 4, device=disp.device, dtype=torch.float32)
        _call("pd_smooth_loss_fwd", lib.pd_smooth_loss_fwd, C.byref(desc), disp.data_ptr(), img.data_ptr(), loss.data_ptr(), ws.data_ptr(), _stream())
        ctx.desc = desc
        ctx.save_for_backward(disp, img)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        disp, img = ctx.saved_tensors
        g = _f32c(g, "grad smooth loss")
        g_disp = torch.empty_like(disp)
        _call("pd_smooth_loss_bwd", lib.pd_smooth_loss_bwd, C.byref(ctx.desc), disp.data_ptr(), img.data_ptr(), g.data_ptr(), g_disp.data_ptr(), _stream())
        return None, None, g_disp, None


def smooth_loss(disp, img, x0: int = 0, gamma: float = 1.0):
    """``get_smooth_loss_disp(disp[..., x0:], img[..., x0:], gamma)`` for ``disp`` [B,1,H,W], ``img`` [B,3,H,W] (0-dim tensor)."""
    disp = _f32c(disp, "disp")
    img = _f32c(img, "img").detach()
    if disp.dim() != 4 or disp.shape[1] != 1 or img.dim() != 4 or img.shape[1] != 3 or disp.shape[-2:] != img.shape[-2:]:
        raise ValueError("smooth_loss expects disp [B,1,H,W] and img [B,3,H,W]")
    return _SmoothLoss.apply(int(x0), float(gamma), disp, img)


class _PlaneTail(torch.autograd.Function):
    """pd_plane_tail_fwd / _bwd: networks/depth_decoder.py:258-291 after the dispconv / sigmaconv convolutions."""

    @staticmethod
    def forward(ctx, mixture: bool, logits_raw, sigma_raw, disp_layered, mask):
        lib = L.lib()
        B, N, H, W = logits_raw.shape
        dev = logits_raw.device
        desc = L.TailDesc(B=B, N=N, H=H, W=W, mixture=int(mixture), mask_dtype=L.PD_MASK_NONE, disp_stride=_strides4(disp_layered))
        if mask is not None:
            desc.mask_stride = _strides4(mask)
            desc.mask_dtype = L.PD_MASK_F32 if mask.dtype == torch.float32 else L.PD_MASK_U8
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        logits, prob, disp, depth, stats = f(B, N, H, W), f(B, N, H, W), f(B, 1, H, W), f(B, 1, H, W), f(B, 3, H, W)
        sigma = f(B, N, H, W) if mixture else None
        tin = L.TailIn(logits_raw=_ptr(logits_raw), sigma_raw=_ptr(sigma_raw), disp_layered=_ptr(disp_layered), mask=_ptr(mask))
        out = L.TailOut(logits=_ptr(logits), sigma=_ptr(sigma), probability=_ptr(prob), pi=None, disp=_ptr(disp), depth=_ptr(depth), stats=_ptr(stats))
        _call("pd_plane_tail_fwd", lib.pd_plane_tail_fwd, C.byref(desc), C.byref(tin), C.byref(out), _stream())
        ctx.desc, ctx.mixture = desc, mixture
        ctx.save_for_backward(disp_layered, mask, logits, sigma, disp, stats)
        return logits, (sigma if mixture else torch.empty(0, device=dev)), prob, disp, depth

    @staticmethod
    def backward(ctx, g_logits, g_sigma, g_prob, g_disp, g_depth):
        lib = L.lib()
        disp_layered, mask, logits, sigma, disp, stats = ctx.saved_tensors
        dev = logits.device
        need = ctx.needs_input_grad  # (mixture, logits_raw, sigma_raw, disp_layered, mask)
        c = lambda g, what: None if g is None else _f32c(g, what)
        g_logits, g_prob, g_disp, g_depth = c(g_logits, "grad logits"), c(g_prob, "grad probability"), c(g_disp, "grad disp"), c(g_depth, "grad depth")
        g_sigma = c(g_sigma, "grad sigma") if (ctx.mixture and g_sigma is not None and g_sigma.numel()) else None
        tin = L.TailIn(disp_layered=_ptr(disp_layered), mask=_ptr(mask))
        saved = L.TailOut(logits=_ptr(logits), sigma=_ptr(sigma), disp=_ptr(disp), stats=_ptr(stats))
        gout = L.TailGradOut(g_logits=_ptr(g_logits), g_sigma=_ptr(g_sigma), g_probability=_ptr(g_prob), g_disp=_ptr(g_disp), g_depth=_ptr(g_depth))
        g_raw = torch.empty_like(logits) if need[1] else None
        g_sraw = torch.empty_like(logits) if (ctx.mixture and need[2]) else None
        gin = L.TailGradIn(g_logits_raw=_ptr(g_raw), g_sigma_raw=_ptr(g_sraw))
        g_dl, spread = None, 1
        if need[3]:
            # broadcast dimensions are reduced inside the kernel and handed back spread evenly (see _WarpComposite.backward)
            shape = [1 if (disp_layered.size(i) == 1 or disp_layered.stride(i) == 0) else disp_layered.size(i) for i in range(4)]
            for i in range(4):
                if shape[i] == 1:
                    spread *= disp_layered.size(i)
            g_dl = torch.empty(shape, device=dev, dtype=torch.float32)
            gin.g_disp_layered = g_dl.data_ptr()
            gin.g_disp_stride = _strides4(g_dl)
        _call("pd_plane_tail_bwd", lib.pd_plane_tail_bwd, C.byref(ctx.desc), C.byref(tin), C.byref(saved), C.byref(gout), C.byref(gin), _stream())
        if g_dl is not None and tuple(g_dl.shape) != tuple(disp_layered.shape):
            g_dl = (g_dl / spread if spread > 1 else g_dl).expand(disp_layered.shape)
        return None, g_raw, g_sraw, g_dl, None


def plane_tail(logits_raw, sigma_raw, padding_mask, disp_layered, mixture: bool):
    """The decoder's outputs after its last convolutions (depth_decoder.py:258-291): returns a dict with ``logits``,
    ``probability``, ``disp``, ``depth`` and, with ``mixture``, ``sigma`` and ``pi`` (= softmax before the reweighting is
    not materialised: nothing reads it; the key maps to ``None``)."""
    logits_raw = _f32c(logits_raw, "logits_raw")
    sigma_raw = _f32c(sigma_raw, "sigma_raw") if mixture else None
    dl = disp_layered if disp_layered.dtype == torch.float32 else disp_layered.float()
    dl = compact_expand_base(dl)
    mask = padding_mask
    if mask is not None:
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        elif mask.dtype not in (torch.float32, torch.uint8):
            mask = mask.float()
        mask = mask.detach()
    logits, sigma, prob, disp, depth = _PlaneTail.apply(bool(mixture), logits_raw, sigma_raw, dl, mask)
    out = {"logits": logits, "probability": prob, "disp": disp, "depth": depth}
    if mixture:
        out["sigma"], out["pi"] = sigma, None
    return out
