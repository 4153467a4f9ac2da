"""Drop-in for the reference's ``models/unet_dann.py``: identical constructor and state_dict,
``forward(x, return_features=False)`` ALWAYS returns a tuple ``(logits, gap_or_None)`` where
``gap = mean(bottleneck, dim=[2,3,4])`` -> ``[B, 2*features[-1]]`` (reference :65-98)."""
from __future__ import annotations

from .. import functional as F
from .unet import DoubleConv, UNet3D as _UNet3D  # noqa: F401


class UNet3D(_UNet3D):
    def forward(self, x, return_features=False):
        logits, bott = self._body(x)
        if return_features:
            return logits, F.global_avg_pool(bott)
        return logits, None
