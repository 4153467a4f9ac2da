"""Restatement of train_dann.py:22-49 and the DANN step arithmetic :243-260 (TEST INFRASTRUCTURE)."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Reverse(torch.autograd.Function):
    """train_dann.py:22-29: identity forward, -lambda * g backward"""

    @staticmethod
    def forward(ctx, x, lam):
        ctx.lam = lam
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return -ctx.lam * g, None


def grad_reverse(x, lam):
    return _Reverse.apply(x, lam)


def init_discriminator(in_features=256, seed=None):
    """train_dann.py:34-46: Linear(in,256) Linear(256,128) Linear(128,64) Linear(64,2); keys net.{0,3,6,8}"""
    if seed is not None:
        torch.manual_seed(seed)
    sd = OrderedDict()
    for idx, (i, o) in zip((0, 3, 6, 8), ((in_features, 256), (256, 128), (128, 64), (64, 2))):
        lin = nn.Linear(i, o)
        sd[f"net.{idx}.weight"] = lin.weight.detach().clone()
        sd[f"net.{idx}.bias"] = lin.bias.detach().clone()
    return sd


def discriminator_forward(sd, x, masks=(None, None)):
    """masks: the two Dropout(0.2) multipliers ([B,256], [B,128]) or None (eval / p=0)"""
    h = F.relu(F.linear(x, sd["net.0.weight"], sd["net.0.bias"]))
    if masks[0] is not None:
        h = h * masks[0]
    h = F.relu(F.linear(h, sd["net.3.weight"], sd["net.3.bias"]))
    if masks[1] is not None:
        h = h * masks[1]
    h = F.relu(F.linear(h, sd["net.6.weight"], sd["net.6.bias"]))
    return F.linear(h, sd["net.8.weight"], sd["net.8.bias"])


def domain_loss(disc_sd, source_feat, target_feat, lam):
    """train_dann.py:248-258"""
    s = discriminator_forward(disc_sd, grad_reverse(source_feat, lam))
    t = discriminator_forward(disc_sd, grad_reverse(target_feat, lam))
    labels = torch.cat([torch.zeros(s.shape[0], dtype=torch.long), torch.ones(t.shape[0], dtype=torch.long)]).to(s.device)
    return F.cross_entropy(torch.cat([s, t]), labels)
