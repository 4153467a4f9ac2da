"""CPU (gloo, world_size 2): the host-side logic of the data-parallel trainer — flat parameter/gradient
buffers, bucket split, sum all-reduce + 1/world scaling — gives every rank the averaged-gradient update
that a single process computes on the concatenated batch.  Mirrors what DDP does for the reference
(train_unet.py:384-386, 221-226).  The fused CUDA optimiser is replaced by a plain torch AdamW over the
same flat buffers through `optimizer_factory`; kernels are not involved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_segmentation_project_b200.dp import DataParallelTrainer, FlatParams
from multimodal_segmentation_project_b200.models.unet import UNet3D


class _TorchFlatAdamW:
    def __init__(self, fp, lr=1e-2):
        self.fp = fp
        self.opt = torch.optim.AdamW([fp.flat], lr=lr, weight_decay=1e-2)

    def step(self, grad_scale=1.0):
        self.fp.flat.grad = self.fp.grad * grad_scale
        self.opt.step()


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = torch.nn.ModuleList([torch.nn.Conv3d(1, 4, 3, padding=1), torch.nn.Conv3d(4, 4, 3, padding=1)])
        self.bottleneck = torch.nn.Conv3d(4, 4, 3, padding=1)
        self.final_conv = torch.nn.Conv3d(4, 3, 1)

    def forward(self, x):
        for e in self.encoder:
            x = torch.relu(e(x))
        return self.final_conv(torch.relu(self.bottleneck(x)))


def _loss(logits, y):
    return torch.nn.functional.cross_entropy(logits, y.squeeze(1))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = _Tiny()
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(2, 1, 6, 6, 6, generator=g)
    y = torch.randint(0, 3, (2, 1, 6, 6, 6), generator=g)
    tr = DataParallelTrainer(model, _loss, autocast_dtype=None, optimizer_factory=lambda fp: _TorchFlatAdamW(fp))
    assert tr.world == 2
    for _ in range(3):
        tr.step(x, y)
    out[rank] = tr.fp.flat.detach().clone()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process_average():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert torch.equal(a, b), "ranks diverged after the all-reduce"
    # single process: average of the two per-rank gradients == gradient of the mean loss over both batches
    torch.manual_seed(0)
    model = _Tiny()
    fp = FlatParams(model)
    opt = _TorchFlatAdamW(fp)
    data = []
    for r in range(2):
        g = torch.Generator().manual_seed(100 + r)
        data.append((torch.randn(2, 1, 6, 6, 6, generator=g), torch.randint(0, 3, (2, 1, 6, 6, 6), generator=g)))
    for _ in range(3):
        fp.zero_grad()
        for x, y in data:
            (_loss(model(x), y) / 2).backward()
        opt.step(1.0)
    assert torch.allclose(fp.flat, a, rtol=1e-5, atol=1e-7)


def test_flat_params_layout_and_buckets():
    torch.manual_seed(0)
    net = UNet3D(1, 4)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    fp = FlatParams(net)
    assert fp.total >= 5647908 and fp.total % 4 == 0
    after = net.state_dict()
    assert list(after.keys()) == list(before.keys())
    for k in before:
        assert torch.equal(after[k], before[k]), k
    # parameters and gradients alias the flat buffers
    lo, hi = fp.flat.data_ptr(), fp.flat.data_ptr() + fp.flat.numel() * 4
    for _, p in net.named_parameters():
        assert lo <= p.data_ptr() < hi and p.grad is not None and p.grad.shape == p.shape
    # bucket 0 = final conv + decoder + upconvs + bottleneck: the parameter-heavy, FLOP-light layers; then encoder levels 3+2+1
    # (reduced under the level-0 backward) and level 0 alone (30 KB: the only all-reduce left after backward, hidden under AdamW)
    b = fp.buckets()
    assert len(b) == 3 and 0.80 < b[0].numel() / 5647908 < 0.90
    assert sum(t.numel() for t in b) == fp.total and b[2].numel() * 4 < 32 * 2 ** 10
    assert all(e % 4 == 0 for e in fp.bucket_elems)          # 16-byte aligned bucket ranges (FlatAdamW.apply_range)
    names = dict(net.named_parameters())
    assert fp.early_sentinel is names["encoder.3.double_conv.4.weight"]
    assert len(fp.sentinels) == 2 and fp.sentinels[0] is names["encoder.3.double_conv.4.weight"] and fp.sentinels[1] is names["encoder.0.double_conv.4.weight"]
    assert [n for n, _ in fp.order[fp.bucket_params[1]:fp.bucket_params[2]]][0].startswith("encoder.")
    assert all(n.startswith("encoder.0.") for n, _ in fp.order[fp.bucket_params[2]:])
    # two-level encoders keep one level per bucket; a single level has no tail bucket
    small = FlatParams(UNet3D(1, 2, features=[8, 16]))
    assert len(small.buckets()) == 3 and all(n.startswith("encoder.0.") for n, _ in small.order[small.bucket_params[2]:])
    # writing through the flat buffer is visible in the module (what the fused optimiser relies on)
    fp.flat.zero_()
    assert float(net.final_conv.weight.detach().abs().sum()) == 0.0
    # frozen encoder: no late bucket sentinel needed
    net2 = UNet3D(1, 4)
    for p in net2.encoder.parameters():
        p.requires_grad = False
    fp2 = FlatParams(net2)
    assert fp2.early_sentinel is None and len(fp2.buckets()) == 1 and fp2.sentinels == []


def _accum_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = _Tiny()
    tr = DataParallelTrainer(model, _loss, autocast_dtype=None, optimizer_factory=lambda fp: _TorchFlatAdamW(fp), accumulation_steps=2)
    before = tr.fp.flat.detach().clone()
    for micro in range(4):                                   # two optimiser steps of two micro-batches each
        g = torch.Generator().manual_seed(1000 + 10 * micro + rank)
        x = torch.randn(2, 1, 6, 6, 6, generator=g)
        y = torch.randint(0, 3, (2, 1, 6, 6, 6), generator=g)
        tr.step(x, y)
        if micro == 0:
            assert torch.equal(tr.fp.flat, before), "no optimiser step (and no all-reduce) on a non-boundary micro-step"
    out[rank] = tr.fp.flat.detach().clone()
    dist.destroy_process_group()


def test_gradient_accumulation_two_ranks_gloo():
    """accelerator.accumulate semantics (train_unet.py:221) on two ranks: per optimiser step the update uses the mean over
    ranks of the mean over micro-batches; ranks stay identical."""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_accum_worker, args=(2, port, out), nprocs=2, join=True)
    assert torch.equal(out[0], out[1])
    torch.manual_seed(0)
    model = _Tiny()
    fp = FlatParams(model)
    opt = _TorchFlatAdamW(fp)
    for step in range(2):
        fp.zero_grad()
        for micro in (2 * step, 2 * step + 1):
            for rank in range(2):
                g = torch.Generator().manual_seed(1000 + 10 * micro + rank)
                x = torch.randn(2, 1, 6, 6, 6, generator=g)
                y = torch.randint(0, 3, (2, 1, 6, 6, 6), generator=g)
                (_loss(model(x), y) / 4).backward()
        opt.step(1.0)
    assert torch.allclose(fp.flat, out[0], rtol=1e-5, atol=1e-7)


class _TinyBN(_Tiny):
    def __init__(self):
        super().__init__()
        self.norm = torch.nn.BatchNorm3d(4)

    def forward(self, x):
        for e in self.encoder:
            x = torch.relu(e(x))
        return self.final_conv(torch.relu(self.norm(self.bottleneck(x))))


def _unseeded_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1234 + 77 * rank)              # the reference seeds only with --seed: ranks build DIFFERENT models
    model = _TinyBN()
    with torch.no_grad():
        model.norm.running_mean.fill_(float(rank + 1))
        model.norm.num_batches_tracked.fill_(5 * (rank + 1))
    init = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    tr = DataParallelTrainer(model, _loss, autocast_dtype=None, optimizer_factory=lambda fp: _TorchFlatAdamW(fp))
    after_ctor = tr.fp.flat.detach().clone()
    bufs = (model.norm.running_mean.clone(), model.norm.num_batches_tracked.clone())
    g = torch.Generator().manual_seed(100 + rank)
    tr.step(torch.randn(2, 1, 6, 6, 6, generator=g), torch.randint(0, 3, (2, 1, 6, 6, 6), generator=g))
    out[rank] = (init, after_ctor, bufs, tr.fp.flat.detach().clone())
    dist.destroy_process_group()


def test_constructor_broadcasts_rank0_state():
    """DDP / Accelerate.prepare broadcast rank 0's parameters and buffers at construction (train_unet.py:384-386); without it
    unseeded ranks would train permanently different replicas."""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_unseeded_worker, args=(2, port, out), nprocs=2, join=True)
    init0, ctor0, bufs0, step0 = out[0]
    init1, ctor1, bufs1, step1 = out[1]
    assert not torch.equal(init0, init1), "the test needs ranks that start from different weights"
    n = init0.numel()
    # FlatParams orders early-bucket parameters first, so compare as multisets through sorting
    assert torch.equal(ctor0, ctor1) and torch.equal(torch.sort(ctor0[:n]).values, torch.sort(init0).values)
    assert torch.equal(bufs0[0], bufs1[0]) and float(bufs1[0][0]) == 1.0
    assert int(bufs0[1]) == int(bufs1[1]) == 5
    assert torch.equal(step0, step1), "replicas diverged after one step"
