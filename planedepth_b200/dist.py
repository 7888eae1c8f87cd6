"""Host-side data-parallel plumbing of the path.

Units are (image, target side); a rank owns a contiguous shard of the global batch and the path needs no
data-path collective.  What crosses ranks: a barrier on both sides of a timed region, the MAX over ranks of
device-measured time, and (for parity checks) a SUM of per-rank loss numerators.  All of it goes through
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of the images rank owns; shards differ by at most one image and cover the batch exactly once
    (the DistributedSampler partition of trainer.py:139 without its padding duplicates)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_seed(seed: int, rank: int) -> int:
    """Distinct, reproducible synthetic data per rank."""
    return seed + 7919 * rank


def barrier(world_size: int, cuda: bool) -> None:
    if world_size > 1:
        dist.barrier()
    if cuda:
        torch.cuda.synchronize()


def max_over_ranks(value: float, world_size: int, device="cpu") -> float:
    """MAX all-reduce of a device-measured duration (never a wall clock): the step ends when the slowest rank ends."""
    if world_size == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def global_mean_loss(local_sum: torch.Tensor, local_count: int, world_size: int) -> torch.Tensor:
    """ph_loss.mean() over the GLOBAL batch from per-rank numerators (what DDP's gradient averaging makes of the
    per-rank means when shards are equal; exact also for ragged shards)."""
    num = local_sum.detach().to(torch.float64).reshape(1).clone()
    den = torch.tensor([float(local_count)], dtype=torch.float64, device=num.device)
    if world_size > 1:
        dist.all_reduce(num, op=dist.ReduceOp.SUM)
        dist.all_reduce(den, op=dist.ReduceOp.SUM)
    return (num / den).to(torch.float32)[0]


def aggregate_throughput(units_per_rank: int, world_size: int, ms_per_step: float) -> float:
    """Whole-job units/s (weak scaling: every rank processes units_per_rank per step)."""
    return units_per_rank * world_size / (ms_per_step * 1e-3)


def freeze_unused_parameters(module: torch.nn.Module, dry_run_loss) -> int:
    """Freeze the parameters a training step never touches, found by one dry run (``dry_run_loss()`` returns the step's
    loss for the un-wrapped ``module``).  The reference wraps its networks with ``find_unused_parameters=True``
    (trainer.py:99) because the ResNet encoder's classifier head never contributes to the loss: a graph traversal and a
    bitmap all-reduce every step.  With the dead parameters frozen once, DistributedDataParallel keeps a static bucket
    plan and only the gradient buckets cross devices.  Returns the number of parameters frozen."""
    for p in module.parameters():
        p.grad = None
    dry_run_loss().backward()
    frozen = 0
    for p in module.parameters():
        if p.requires_grad and p.grad is None:
            p.requires_grad_(False)
            frozen += p.numel()
        p.grad = None
    return frozen
