"""Drop-in for the reference's ``models/unet.py`` (DoubleConv :6-22, UNet3D :24-90).

Same constructor arguments, attribute names (``encoder``, ``pool``, ``bottleneck``, ``upconvs``,
``decoder``, ``final_conv``, ``dropout_rate``, ``output_activation``) and the same 136-entry
``state_dict``: the sub-modules are ordinary ``torch.nn`` parameter
containers, but ``forward`` never calls them — it runs the hand-written sm_100a kernels of
libb200unet through :mod:`..functional`.

Compute dtype: bf16 (fp32 accumulate) inside ``torch.autocast`` — the reference's Accelerate
mixed-precision path — or when ``model.compute_dtype = torch.bfloat16`` is set; fp32 otherwise.
Logits are always returned as fp32 NCDHW, which is what Accelerate's output cast hands the loss.
CUDA only: a CPU tensor raises (no fallback).
"""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn

from .. import functional as F


class DoubleConv(nn.Module):
    """[Conv3d => BatchNorm3d => ReLU => Dropout3d] x 2 — reference models/unet.py:6-22."""

    def __init__(self, in_channels, out_channels, dropout_rate=0.1):
        super().__init__()
        # Indices 0,1,4,5 own the state; 2,3,6,7 are stateless (same numbering as the reference).
        self.double_conv = nn.Sequential(
            nn.Conv3d(in_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm3d(out_channels),
            nn.ReLU(inplace=True),
            nn.Dropout3d(p=dropout_rate),
            nn.Conv3d(out_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm3d(out_channels),
            nn.ReLU(inplace=True),
            nn.Dropout3d(p=dropout_rate),
        )

    def _mask(self, drop: nn.Dropout3d, n: int, c: int, device):
        # Dropout3d == x * bernoulli_(1-p)/(1-p) with noise shape [N, C, 1, 1, 1]
        p = drop.p
        if not self.training or p <= 0.0:
            return None
        if p >= 1.0:
            return torch.zeros((n, c), dtype=torch.float32, device=device)
        return torch.empty((n, c), dtype=torch.float32, device=device).bernoulli_(1.0 - p).div_(1.0 - p)

    def forward_cl(self, x0, x1=None, impl=0, defer_last_norm=False, pool=False):
        """Channels-last forward; ``x1`` is the second half of a virtual concat (skip first).
        ``defer_last_norm``: stop after the second convolution and its batch statistics and return ``(conv_out, stats)`` — the
        caller (the fused head) applies BatchNorm + ReLU on the fly instead of materialising them.
        ``pool``: return ``(skip, MaxPool3d(2,2)(skip))`` — the encoder hand-off (reference models/unet.py:69-71) as part of the
        block: the last BatchNorm apply pass then writes the pooled tensor too."""
        seq = self.double_conv
        n = x0.shape[0]
        m0 = self._mask(seq[3], n, seq[0].out_channels, x0.device)
        h, hctx = F.conv_bn_act(x0, x1, seq[0], seq[1], m0, seq[1].training, impl, return_ctx=True)
        # without dropout h == relu(bn(conv_out)): the second conv's data-gradient kernel can then carry the first BatchNorm's
        # backward reduction (functional._conv_bwd_from_dconv)
        prev = hctx if (m0 is None and seq[1].training) else None
        if defer_last_norm:
            return F.conv_batch_stats(h, None, seq[4], seq[5], impl, prev=prev)
        m1 = self._mask(seq[7], n, seq[4].out_channels, x0.device)
        if pool:
            if F.bn_pool_fusable(h, seq[4].out_channels, m1):
                return F.conv_bn_act_skip_pool(h, None, seq[4], seq[5], seq[5].training, impl, prev=prev)
            # skip connection + MaxPool3d(2,2); their gradients meet in one kernel
            return F.skip_and_pool(F.conv_bn_act(h, None, seq[4], seq[5], m1, seq[5].training, impl, prev=prev))
        return F.conv_bn_act(h, None, seq[4], seq[5], m1, seq[5].training, impl, prev=prev)

    def forward(self, x):
        """NCDHW in, NCDHW out (fp32) — the stand-alone module contract of the reference."""
        dtype = _compute_dtype(self, x)
        h = F._ToChannelsLast.apply(x, dtype)
        y = self.forward_cl(h)
        return _ToChannelsFirst.apply(y)


class _ToChannelsFirst(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.dtype = x.dtype
        return F.to_channels_first_f32(x)

    @staticmethod
    def backward(ctx, g):
        return F.to_channels_last(g, ctx.dtype)


_warned_fp16 = False


def _compute_dtype(module, x):
    global _warned_fp16
    forced = getattr(module, "compute_dtype", None)
    if forced is not None:
        return forced
    if torch.is_autocast_enabled():
        if torch.get_autocast_dtype("cuda") == torch.float16 and not _warned_fp16:
            warnings.warn("fp16 autocast requested: libb200unet computes in bf16 (fp32 accumulate) instead")
            _warned_fp16 = True
        return torch.bfloat16
    if x.dtype == torch.bfloat16:
        return torch.bfloat16
    return torch.float32


class UNet3D(nn.Module):
    """3D U-Net for volumetric segmentation — reference models/unet.py:24-90.

    Args mirror the reference: in_channels, out_channels, features, output_activation, dropout_rate.
    """

    def __init__(self, in_channels=1, out_channels=1, features=[16, 32, 64, 128], output_activation=None,
                 dropout_rate=0.1):
        super().__init__()
        self.encoder = nn.ModuleList()
        self.pool = nn.MaxPool3d(kernel_size=2, stride=2)
        self.output_activation = output_activation
        self.dropout_rate = dropout_rate
        for feature in features:
            self.encoder.append(DoubleConv(in_channels, feature, dropout_rate))
            in_channels = feature
        self.bottleneck = DoubleConv(features[-1], features[-1] * 2, dropout_rate)
        self.upconvs = nn.ModuleList()
        self.decoder = nn.ModuleList()
        for feature in reversed(features):
            self.upconvs.append(nn.ConvTranspose3d(feature * 2, feature, kernel_size=2, stride=2))
            self.decoder.append(DoubleConv(feature * 2, feature, dropout_rate))
        self.final_conv = nn.Conv3d(features[0], out_channels, kernel_size=1)
        self.compute_dtype = None  # None = follow autocast; torch.bfloat16 / torch.float32 to force
        self.conv_impl = 0         # 0 auto, 1 CUDA-core implicit GEMM, 2 tcgen05

    # -- the network body on channels-last tensors; returns (logits NCDHW fp32, bottleneck NDHWC)
    def _body(self, x, head=None):
        if not x.is_cuda:
            raise RuntimeError("UNet3D (b200) needs a CUDA tensor: there is no CPU fallback "
                               "(use the reference implementation or the oracle for CPU runs)")
        if x.dim() != 5:
            raise ValueError(f"expected [B, C, D, H, W] input, got {tuple(x.shape)}")
        dtype = _compute_dtype(self, x)
        impl = self.conv_impl
        h = F._ToChannelsLast.apply(x, dtype)
        skips = []
        for down in self.encoder:
            skip, h = down.forward_cl(h, None, impl, pool=True)  # skip connection + MaxPool3d(2,2)
            skips.append(skip)
        bott = self.bottleneck.forward_cl(h, None, impl)
        h = bott
        skips = skips[::-1]
        for idx in range(len(self.upconvs)):
            up = self.upconvs[idx]
            h = F.conv_transpose2(h, up.weight, up.bias)
            skip = skips[idx]
            if h.shape[1:4] != skip.shape[1:4]:
                h = F.nearest_resize(h, skip.shape[1:4])  # models/unet.py:81-83
            if head is not None and idx == len(self.upconvs) - 1:
                # fused head: the last BatchNorm + ReLU, the final conv, the loss and the confusion counts read the last
                # convolution's output once (csrc/head_fused.cu)
                conv_out, stats = self.decoder[idx].forward_cl(skip, h, impl, defer_last_norm=True)
                mode, alpha, beta, want_conf, target = head
                loss, logits, conf = F.fused_head(conv_out, stats, self.decoder[idx].double_conv[5], self.final_conv, target, mode, alpha, beta,
                                                  want_conf, round_bf16=True)
                return logits, bott, loss, conf
            h = self.decoder[idx].forward_cl(skip, h, impl)  # virtual cat((skip, up), 1)
        logits = F.final_conv1x1(h, self.final_conv.weight, self.final_conv.bias, round_bf16=(dtype == torch.bfloat16))
        if self.output_activation is not None:
            logits = self.output_activation(logits)
        return logits, bott

    def forward(self, x):
        return self._body(x)[0]

    def fused_head_available(self, x) -> bool:
        """The single-pass head serves the training configuration of the reference's scripts: train mode, bf16 compute, 16 features at
        the top level, 2..4 classes, no output activation, no dropout after the last convolution."""
        last = self.decoder[-1].double_conv if len(self.decoder) else None
        return bool(self.training and last is not None and len(self.upconvs) > 0 and x.is_cuda and _compute_dtype(self, x) == torch.bfloat16
                    and last[4].out_channels == 16 and 2 <= self.final_conv.out_channels <= 4 and self.output_activation is None
                    and last[7].p <= 0.0 and last[5].training and last[5].momentum is not None and self.final_conv.in_channels == 16)

    def forward_with_loss(self, x, target, loss_fn, want_confusion=False):
        """``(logits, loss, confusion)`` with ``loss == loss_fn(model(x).float(), target)`` (``loss_fn`` one of this package's
        ``utils.metrics`` losses) and ``confusion`` the int64 ``[C, C]`` counts ``conf[target, argmax]`` (or None).  When the
        configuration allows (``fused_head_available``) everything after the last convolution runs as one kernel per direction;
        otherwise this is exactly the unfused sequence.  ``target`` may be int64 (the reference's contract) or uint8."""
        spec = getattr(loss_fn, "_b200_spec", None)
        if spec is not None and self.fused_head_available(x):
            mode, alpha, beta = spec
            logits, _, loss, conf = self._body(x, head=(mode, alpha, beta, bool(want_confusion), target))
            return logits, loss, conf
        logits = self(x)
        t = target if target.dtype == torch.int64 else target.long()
        loss = loss_fn(logits.float(), t)
        conf = F.confusion_counts(logits.detach(), t) if want_confusion else None
        return logits, loss, conf
