"""Synthetic volumes for parity tests and benchmarks: a structured problem whose
labels are spatially coherent (class fractions 55/20/15/10 %) so that gradients carry signal, and a
worst-case variant (uniform noise, uniform labels). Generated on the CPU with a seeded generator so
every implementation sees identical inputs; this is input generation, not part of the hot path."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _kth_quantiles(flat: torch.Tensor, qs):
    n = flat.numel()
    out = []
    for q in qs:
        k = min(max(int(round(q * (n - 1))) + 1, 1), n)
        out.append(torch.kthvalue(flat, k).values)
    return torch.stack(out)


def structured_volume(batch: int, size, seed: int = 1234, classes: int = 4):
    """returns (image fp32 [B,1,*size] in ~[0,1], labels int64 [B,1,*size])"""
    if isinstance(size, int):
        size = (size, size, size)
    g = torch.Generator().manual_seed(seed)
    n = torch.randn((batch, 1, *size), generator=g)
    sm = F.avg_pool3d(n, 5, 1, 2, count_include_pad=True)
    sm = F.avg_pool3d(sm, 5, 1, 2, count_include_pad=True)
    lo, hi = sm.min(), sm.max()
    sm = (sm - lo) / (hi - lo)
    fracs = [0.55, 0.75, 0.90][: classes - 1] if classes <= 4 else [i / classes for i in range(1, classes)]
    edges = _kth_quantiles(sm.flatten(), fracs)
    labels = torch.bucketize(sm, edges).to(torch.int64)
    image = sm + 0.05 * torch.randn(sm.shape, generator=g)
    return image.contiguous(), labels.contiguous()


def worst_case_volume(batch: int, size, seed: int = 1234, classes: int = 4):
    if isinstance(size, int):
        size = (size, size, size)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((batch, 1, *size), generator=g)
    y = torch.randint(0, classes, (batch, 1, *size), generator=g, dtype=torch.int64)
    return x, y
