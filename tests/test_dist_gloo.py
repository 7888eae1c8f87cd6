"""CPU, world_size 2, gloo: the host-side data-parallel logic of the path (planedepth_b200/dist.py) — shard
partition, per-rank seeds, MAX-over-ranks timing, global-mean loss from per-rank numerators checked against
the oracle on the un-sharded batch."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(ws), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from oracle import pd_oracle as O
    from planedepth_b200 import dist as D
    from planedepth_b200.synthetic import make_batch, make_opt

    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        assert D.env() == (rank, rank, ws)
        GB, H, W = 5, 8, 32  # ragged: 3 + 2 images
        lo, hi = D.shard_range(GB, rank, ws)
        opt = make_opt(disp_levels=5, disp_max=8.0, disp_min=1.0)
        full = make_batch(GB, H, W, opt, seed=11, device="cpu", requires_grad=False)
        # the shard is a slice of the global batch: run the oracle on it and reduce numerators
        sl = lambda t: t[lo:hi] if torch.is_tensor(t) and t.dim() > 0 and t.shape[0] == GB else t
        inputs = {k: sl(v) for k, v in full.inputs.items()}
        outputs = {k: sl(v) for k, v in full.outputs.items()}
        O.pred_novel_images(opt, full.target_sides, inputs, outputs)
        ph, _ = O.photometric_map(opt, inputs, outputs, "r", None)
        g = D.global_mean_loss(ph.sum(), ph.numel(), ws)
        # reference value: the un-sharded batch
        fo = dict(full.outputs)
        O.pred_novel_images(opt, full.target_sides, full.inputs, fo)
        ph_full, _ = O.photometric_map(opt, full.inputs, fo, "r", None)
        assert abs(float(g) - float(ph_full.mean())) < 1e-6
        # timing reduction: the slowest rank defines the step
        assert D.max_over_ranks(1.0 + rank, ws) == float(ws)
        D.barrier(ws, cuda=False)
        q.put((rank, lo, hi, D.shard_seed(5, rank), float(g)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_reductions():
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, "rank exited with %s" % p.exitcode
    got = sorted(q.get(timeout=5) for _ in range(ws))
    assert [(g[1], g[2]) for g in got] == [(0, 3), (3, 5)]
    assert got[0][3] != got[1][3]
    assert abs(got[0][4] - got[1][4]) < 1e-9


def _ddp_worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(ws), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    import torch.nn as nn

    from planedepth_b200 import dist as D

    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        torch.manual_seed(0)  # identical replicas

        class Net(nn.Module):  # a used trunk and a never-used head, like the ResNet encoder's fc layer
            def __init__(self):
                super().__init__()
                self.trunk = nn.Conv2d(3, 4, 3, padding=1)
                self.fc = nn.Linear(4, 10)

            def forward(self, x):
                return self.trunk(x)

        net = Net()
        g = torch.Generator().manual_seed(100 + rank)  # distinct shard per rank
        x = torch.rand(2, 3, 8, 8, generator=g)
        frozen = D.freeze_unused_parameters(net, lambda: net(x).square().mean())
        assert frozen == 4 * 10 + 10 and not net.fc.weight.requires_grad and net.trunk.weight.requires_grad
        ddp = nn.parallel.DistributedDataParallel(net)  # no find_unused_parameters: the static plan must hold for several steps
        for _ in range(3):
            for p in net.parameters():
                p.grad = None
            ddp(x).square().mean().backward()
        # the all-reduce averaged the per-rank gradients: identical on both ranks, equal to the mean of the local ones
        gsum = net.trunk.weight.grad.clone()
        ref = [torch.zeros_like(gsum) for _ in range(ws)]
        dist.all_gather(ref, gsum)
        assert all(torch.equal(r, ref[0]) for r in ref)
        q.put((rank, float(gsum.abs().sum())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_ddp_with_frozen_unused_parameters():
    """The bench's ddp leg on CPU: DistributedDataParallel over gloo, world size 2, with the never-used parameters frozen by
    dist.freeze_unused_parameters instead of find_unused_parameters=True."""
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, "rank exited with %s" % p.exitcode
    got = sorted(q.get(timeout=5) for _ in range(ws))
    assert abs(got[0][1] - got[1][1]) < 1e-12 and got[0][1] > 0


def test_shard_range_partitions_exactly():
    from planedepth_b200.dist import aggregate_throughput, shard_range

    for gb in (1, 7, 8, 12, 96):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard_range(gb, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)
    assert aggregate_throughput(12, 8, 0.5) == 12 * 8 / 0.5e-3
