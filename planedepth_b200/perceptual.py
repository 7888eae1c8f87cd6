"""Scheduling of the perceptual term of ``Trainer.compute_losses``.

The reference evaluates ``perceptual_loss(pred, target[, source])`` (/root/reference/trainer.py:672-685) once per target
side: three separate passes through the frozen feature network ``pc_net`` (``Vgg19_pc`` / ``Resnet18_pc``,
layers.py:378-450), everything in fp32 NCHW.  After the warp / loss path is fused this term dominates the step
(54.8 GFLOP per image and pass at 640x192).  The convolutions stay cuDNN's; what changes is how they are fed:

* the constant passes (target, source: data, no gradient) run under ``no_grad`` as ONE batched call, and their features
  are cached per batch: the source features are identical for every target side (``source`` is always colour "l"), the
  target features of a side are reused if the same tensor comes back (e.g. the self-distillation and main stages);
* ``channels_last`` activations (cuDNN's native layout for tensor-core convolutions);
* optional bf16 autocast of the feature network with the squared differences and every reduction in fp32
  (``dtype=torch.bfloat16``; BASELINE.json's 2e-2 bf16 gate);
* the formula itself is unchanged: sum over the three feature levels of ``mean(min(mean_c (p - t)^2, mean_c (s - t)^2))``.

fp32 results equal the reference formula up to cuDNN algorithm choice (<= 1e-6 relative; tests/test_perceptual.py)."""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple

import torch


def reference_perceptual_loss(pc_net: Callable, pred, target, source=None):
    """trainer.py:672-685 as written (three passes, caller's dtype / layout)."""
    pv, tv = pc_net(pred), pc_net(target)
    sv = pc_net(source) if source is not None else None
    total = 0
    for i in range(3):
        lp = ((pv[i] - tv[i]) ** 2).mean(1, True)
        if sv is not None:
            la = ((sv[i] - tv[i]) ** 2).mean(1, True)
            lp, _ = torch.cat([lp, la], dim=1).min(1, True)
        total = total + lp.mean()
    return total


class PerceptualSchedule:
    """Callable with the signature of ``Trainer.perceptual_loss``; holds the per-batch feature cache."""

    def __init__(self, pc_net: Callable, dtype: Optional[torch.dtype] = None, channels_last: bool = True, cache: bool = True):
        self.pc_net = pc_net
        self.dtype = dtype
        self.channels_last = channels_last
        self.use_cache = cache
        self._cache: Dict[Tuple, Sequence[torch.Tensor]] = {}
        self.stats = {"const_passes": 0, "const_images": 0, "cache_hits": 0, "pred_passes": 0}

    # -- cache ---------------------------------------------------------------------------------------------------------
    @staticmethod
    def _key(t: torch.Tensor) -> Tuple:
        return (t.data_ptr(), t._version, tuple(t.shape), t.dtype, str(t.device))

    def new_batch(self) -> None:
        """Forget cached features (call once per training step; ``compute_losses`` does)."""
        self._cache.clear()

    # -- feature passes ------------------------------------------------------------------------------------------------
    def _features(self, x: torch.Tensor):
        if self.channels_last and x.dim() == 4:
            x = x.contiguous(memory_format=torch.channels_last)
        if self.dtype is not None and x.is_cuda:
            with torch.autocast("cuda", dtype=self.dtype):
                f = self.pc_net(x)
        else:
            f = self.pc_net(x)
        return [v.float() for v in list(f)[:3]]

    def _constant_features(self, tensors: Sequence[torch.Tensor]):
        """Features of data tensors (no gradient), one batched pass over those not cached yet."""
        todo = [t for t in tensors if not (self.use_cache and self._key(t) in self._cache)]
        self.stats["cache_hits"] += len(tensors) - len(todo)
        if todo:
            with torch.no_grad():
                feats = self._features(torch.cat([t.detach() for t in todo], 0) if len(todo) > 1 else todo[0].detach())
            self.stats["const_passes"] += 1
            self.stats["const_images"] += sum(t.shape[0] for t in todo)
            o = 0
            for t in todo:
                b = t.shape[0]
                self._cache[self._key(t)] = [f[o:o + b] for f in feats]
                o += b
        out = [self._cache[self._key(t)] for t in tensors]
        if not self.use_cache:
            self._cache.clear()
        return out

    # -- the loss ------------------------------------------------------------------------------------------------------
    def __call__(self, pred, target, source=None):
        consts = self._constant_features([target] + ([source] if source is not None else []))
        tv = consts[0]
        sv = consts[1] if source is not None else None
        pv = self._features(pred)
        self.stats["pred_passes"] += 1
        total = 0
        for i in range(3):
            lp = ((pv[i] - tv[i]) ** 2).mean(1, True)
            if sv is not None:  # automask: per-pixel min against the identity reprojection (trainer.py:680-683)
                la = ((sv[i] - tv[i]) ** 2).mean(1, True)
                lp = torch.minimum(lp, la)  # same values and gradient routing as cat + min(1): ties go to the first operand
            total = total + lp.mean()
        return total
