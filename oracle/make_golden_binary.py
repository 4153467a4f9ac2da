"""Golden vectors for the reference's binary helpers (utils/metrics.py:6-12 dice_loss, :42-63 dice_score / iou_score /
accuracy_score, :131-135 calculate_metrics) — dead code in the reference's scripts, but part of the module's API.

    python -m oracle.make_golden_binary        (build container only: imports /root/reference)

TEST INFRASTRUCTURE (see oracle/__init__.py): every number written is an output of the unmodified reference.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .make_golden import OUT, _import_reference, _np


def main():
    _, _, RM, _ = _import_reference()
    out = {}
    cases = (("a", (2, 1, 6, 7, 8), 1, 0.5), ("b", (3, 1, 4, 4, 4), 2, 0.2), ("c", (1, 1, 16, 16, 16), 3, 0.9),
             ("empty_target", (2, 1, 4, 4, 4), 4, 0.0), ("full_target", (2, 1, 4, 4, 4), 5, 1.0))
    for name, shape, seed, frac in cases:
        g = torch.Generator().manual_seed(seed)
        logits = (torch.randn(shape, generator=g) * 2).requires_grad_(True)
        prob = torch.rand(shape, generator=g)                    # scores take probabilities thresholded at 0.5
        prob.view(-1)[::7] = 0.5                                   # exactly on the threshold: (pred > 0.5) is False
        tgt = (torch.rand(shape, generator=g) < frac).float()
        out[f"{name}/logits"], out[f"{name}/prob"], out[f"{name}/target"] = _np(logits), _np(prob), _np(tgt)
        out[f"{name}/dice_score"] = np.float64(RM.dice_score(prob, tgt))
        out[f"{name}/iou_score"] = np.float64(RM.iou_score(prob, tgt))
        out[f"{name}/accuracy_score"] = np.float64(RM.accuracy_score(prob, tgt))
        d, i, a = RM.calculate_metrics(prob, tgt)
        out[f"{name}/calculate_metrics"] = np.array([d, i, a], dtype=np.float64)
        loss = RM.dice_loss(logits, tgt)
        loss.backward()
        out[f"{name}/dice_loss"] = np.float32(loss.item())
        out[f"{name}/dice_loss_grad"] = _np(logits.grad)
    path = os.path.join(OUT, "binary_helpers.npz")
    np.savez_compressed(path, **out)
    print("written", path, f"{os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
