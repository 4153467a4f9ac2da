"""Drop-in for the DANN-specific components of the reference's ``train_dann.py:22-49``:
``GradientReversal`` / ``grad_reverse`` and ``DomainDiscriminator`` (same ``net.{0,3,6,8}``
state_dict keys; ``hidden_dim`` is accepted and unused, as in the reference)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as F


class GradientReversal(torch.autograd.Function):
    """forward: identity view; backward: ``-lambda_ * grad`` (reference train_dann.py:22-29)."""

    @staticmethod
    def forward(ctx, x, lambda_):
        ctx.lambda_ = lambda_
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        return F.scale(grad_output, -float(ctx.lambda_)), None


def grad_reverse(x, lambda_):
    return GradientReversal.apply(x, lambda_)


class DomainDiscriminator(nn.Module):
    """Linear(in,256)-ReLU-Drop(.2)-Linear(256,128)-ReLU-Drop(.2)-Linear(128,64)-ReLU-Linear(64,2)."""

    def __init__(self, in_features, hidden_dim=128):
        super().__init__()
        self.net = nn.Sequential(
            nn.Linear(in_features, 256),
            nn.ReLU(),
            nn.Dropout(0.2),
            nn.Linear(256, 128),
            nn.ReLU(),
            nn.Dropout(0.2),
            nn.Linear(128, 64),
            nn.ReLU(),
            nn.Linear(64, 2),
        )

    def _mask(self, drop, shape, device):
        if not self.training or drop.p <= 0.0:
            return None
        return torch.empty(shape, dtype=torch.float32, device=device).bernoulli_(1.0 - drop.p).div_(1.0 - drop.p)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("DomainDiscriminator (b200) needs a CUDA tensor: there is no CPU fallback")
        net = self.net
        b = x.shape[0]
        h = F.linear_act(x, net[0].weight, net[0].bias, self._mask(net[2], (b, 256), x.device), relu=True)
        h = F.linear_act(h, net[3].weight, net[3].bias, self._mask(net[5], (b, 128), x.device), relu=True)
        h = F.linear_act(h, net[6].weight, net[6].bias, None, relu=True)
        return F.linear_act(h, net[8].weight, net[8].bias, None, relu=False)


def domain_cross_entropy(logits, labels):
    """nn.CrossEntropyLoss() on the discriminator output (reference train_dann.py:258)."""
    return F.cross_entropy_rows(logits, labels)
