"""Functional restatement of the reference U-Net (TEST INFRASTRUCTURE — see oracle/__init__.py).

The network is expressed as a pure function of a ``state_dict``-shaped mapping so that the same
weights can be fed to the oracle and to the CUDA implementation, dropout masks are explicit inputs
 and intermediate activations can be tapped layer by layer.

reference: models/unet.py:6-22 (DoubleConv), :34-62 (constructor / init order), :64-90 (forward);
models/unet_dann.py:65-98 (forward with pooled bottleneck features).
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm3d default, models/unet.py:12
BN_MOMENTUM = 0.1   # nn.BatchNorm3d default


def block_names(n_levels: int = 4):
    enc = [f"encoder.{i}" for i in range(n_levels)]
    dec = [f"decoder.{i}" for i in range(n_levels)]
    return enc + ["bottleneck"] + dec


def init_state_dict(in_channels=1, out_channels=1, features=(16, 32, 64, 128), seed=None) -> "OrderedDict[str, torch.Tensor]":
    """Creates parameters with the same RNG consumption order as the reference constructor
    (models/unet.py:44-62: encoder blocks, bottleneck, then per level upconv + decoder block, final conv),
    each layer using torch's default Conv3d / ConvTranspose3d / BatchNorm3d initialisation."""
    if seed is not None:
        torch.manual_seed(seed)

    def conv_entries(prefix, layer):
        return [(prefix + ".weight", layer.weight.detach().clone()), (prefix + ".bias", layer.bias.detach().clone())]

    def bn_entries(prefix, c):
        return [(prefix + ".weight", torch.ones(c)), (prefix + ".bias", torch.zeros(c)),
                (prefix + ".running_mean", torch.zeros(c)), (prefix + ".running_var", torch.ones(c)),
                (prefix + ".num_batches_tracked", torch.tensor(0, dtype=torch.long))]

    def block_entries(prefix, cin, cout):
        e = conv_entries(f"{prefix}.double_conv.0", nn.Conv3d(cin, cout, 3, padding=1))
        e += bn_entries(f"{prefix}.double_conv.1", cout)
        e += conv_entries(f"{prefix}.double_conv.4", nn.Conv3d(cout, cout, 3, padding=1))
        e += bn_entries(f"{prefix}.double_conv.5", cout)
        return e

    enc, ups, decs = [], [], []
    cin = in_channels
    for i, f in enumerate(features):                       # RNG order: encoder blocks ...
        enc += block_entries(f"encoder.{i}", cin, f)
        cin = f
    bott = block_entries("bottleneck", features[-1], features[-1] * 2)   # ... bottleneck ...
    for i, f in enumerate(reversed(features)):             # ... then (upconv, decoder block) per level ...
        ups += conv_entries(f"upconvs.{i}", nn.ConvTranspose3d(f * 2, f, 2, stride=2))
        decs += block_entries(f"decoder.{i}", f * 2, f)
    fin = conv_entries("final_conv", nn.Conv3d(features[0], out_channels, 1))  # ... final conv last
    # key order = module registration order: encoder, bottleneck, upconvs, decoder, final_conv
    return OrderedDict(enc + bott + ups + decs + fin)


def n_levels_of(sd) -> int:
    return sum(1 for k in sd if k.startswith("encoder.") and k.endswith("double_conv.0.weight"))


def dropout3d_masks(sd, batch: int, p: float, generator=None):
    """One [B, C] multiplier per Dropout3d in module order (18 for the default net): bernoulli(1-p)/(1-p)."""
    masks = OrderedDict()
    for blk in block_names(n_levels_of(sd)):
        for conv_idx, drop_idx in ((0, 3), (4, 7)):
            c = sd[f"{blk}.double_conv.{conv_idx}.weight"].shape[0]
            m = torch.empty(batch, c).bernoulli_(1.0 - p, generator=generator) / (1.0 - p) if p > 0 else None
            masks[f"{blk}.double_conv.{drop_idx}"] = m
    return masks


def _conv_bn_relu_drop(sd, blk, conv_idx, x, training, mask, taps):
    bn_idx = conv_idx + 1
    x = F.conv3d(x, sd[f"{blk}.double_conv.{conv_idx}.weight"], sd[f"{blk}.double_conv.{conv_idx}.bias"], padding=1)
    if taps is not None:
        taps[f"{blk}.double_conv.{conv_idx}"] = x
    p = f"{blk}.double_conv.{bn_idx}"
    x = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], training,
                     BN_MOMENTUM, BN_EPS)
    if training and p + ".num_batches_tracked" in sd:
        sd[p + ".num_batches_tracked"] += 1
    x = F.relu(x)
    if mask is not None:
        x = x * mask.to(x.dtype)[:, :, None, None, None]
    if taps is not None:
        taps[f"{blk}.double_conv.{conv_idx + 2}"] = x
    return x


def double_conv(sd, blk, x, training=True, masks=None, taps=None):
    """reference models/unet.py:9-22"""
    masks = masks or {}
    x = _conv_bn_relu_drop(sd, blk, 0, x, training, masks.get(f"{blk}.double_conv.3"), taps)
    return _conv_bn_relu_drop(sd, blk, 4, x, training, masks.get(f"{blk}.double_conv.7"), taps)


def unet3d_forward(sd, x, training=True, masks=None, return_features=None, taps=None):
    """reference models/unet.py:64-90; with ``return_features`` not None behaves like
    models/unet_dann.py:65-98 and returns ``(logits, gap_or_None)``."""
    L = n_levels_of(sd)
    skips = []
    for i in range(L):
        x = double_conv(sd, f"encoder.{i}", x, training, masks, taps)
        skips.append(x)
        x = F.max_pool3d(x, 2, 2)
    bott = double_conv(sd, "bottleneck", x, training, masks, taps)
    x = bott
    skips = skips[::-1]
    for i in range(L):
        x = F.conv_transpose3d(x, sd[f"upconvs.{i}.weight"], sd[f"upconvs.{i}.bias"], stride=2)
        s = skips[i]
        if x.shape != s.shape:
            x = F.interpolate(x, size=s.shape[2:])
        x = torch.cat((s, x), dim=1)
        x = double_conv(sd, f"decoder.{i}", x, training, masks, taps)
    logits = F.conv3d(x, sd["final_conv.weight"], sd["final_conv.bias"])
    if return_features is None:
        return logits
    return logits, (torch.mean(bott, dim=[2, 3, 4]) if return_features else None)


def trainable(sd):
    """names of the entries the reference registers as nn.Parameters"""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


def clone_for_autograd(sd, device="cpu"):
    out = OrderedDict()
    names = set(trainable(sd))
    for k, v in sd.items():
        t = v.detach().clone().to(device)
        if k in names:
            t.requires_grad_(True)
        out[k] = t
    return out


def train_step_grads(sd, x, y, loss_fn, masks=None, autocast_dtype=None):
    """forward + loss + backward (reference train_unet.py:222-225); returns (loss, logits, grads dict).
    ``autocast_dtype`` emulates Accelerate mixed precision: forward under autocast, logits cast to fp32,
    loss outside autocast."""
    p = clone_for_autograd(sd, x.device)
    if autocast_dtype is not None:
        with torch.autocast(x.device.type, dtype=autocast_dtype):
            logits = unet3d_forward(p, x, True, masks)
        logits = logits.float()
    else:
        logits = unet3d_forward(p, x, True, masks)
    loss = loss_fn(logits, y)
    loss.backward()
    grads = OrderedDict((k, p[k].grad.detach()) for k in trainable(sd))
    # running statistics were updated in the clones; hand them back
    buffers = OrderedDict((k, p[k].detach()) for k in p if k not in grads)
    return loss.detach(), logits.detach(), grads, buffers
