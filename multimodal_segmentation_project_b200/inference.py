"""Sliding-window inference + per-organ metrics (BASELINE config #5).

The reference evaluates whole volumes with one forward (test_model.py:248) and has no sliding-window
code; this wrapper is the build-defined tiling of that evaluation: windows are
independent forwards of ``model.eval()``, overlapping logits are averaged, the stitched logits go
through the fused argmax + confusion-count kernel and the evaluator's Dice/IoU rules
(test_model.py:265-285) are applied to the integer counts.  Windows are sharded round-robin across
ranks when torch.distributed is initialised (no data-path collective other than the final sum of the
accumulators)."""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check
from .functional import _ptr, _require_cuda, _stream, confusion_counts


def window_starts(size: int, window: int, stride: int):
    if size <= window:
        return [0]
    starts = list(range(0, size - window + 1, stride))
    if starts[-1] != size - window:
        starts.append(size - window)
    return starts


def _accumulate(acc, cnt, logits, origin):
    L = _lib.load()
    _, C, D, H, W = acc.shape
    wd, wh, ww = logits.shape[2:]
    check(L.b200_window_accumulate(_ptr(acc), _ptr(cnt), _ptr(logits), C, D, H, W, origin[0], origin[1], origin[2], wd, wh, ww, _stream()),
          "window_accumulate")


@torch.no_grad()
def sliding_window_logits(model, volume: torch.Tensor, window=128, stride=64, process_group=None) -> torch.Tensor:
    """volume [1, Cin, D, H, W] (CUDA) -> averaged logits [1, C, D, H, W] fp32."""
    _require_cuda(volume)
    if volume.dim() != 5 or volume.shape[0] != 1:
        raise ValueError(f"expected a [1, C, D, H, W] volume, got {tuple(volume.shape)}")
    if isinstance(window, int):
        window = (window,) * 3
    if isinstance(stride, int):
        stride = (stride,) * 3
    L = _lib.load()
    D, H, W = volume.shape[2:]
    win = tuple(min(w, s) for w, s in zip(window, (D, H, W)))
    origins = [(d, h, w) for d in window_starts(D, win[0], stride[0]) for h in window_starts(H, win[1], stride[1])
               for w in window_starts(W, win[2], stride[2])]
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
    was_training = model.training
    model.eval()
    acc = cnt = None
    for i, (d0, h0, w0) in enumerate(origins):
        if i % world != rank:
            continue
        patch = volume[:, :, d0:d0 + win[0], h0:h0 + win[1], w0:w0 + win[2]].contiguous()
        out = model(patch)
        logits = (out[0] if isinstance(out, tuple) else out).float().contiguous()
        if acc is None:
            acc = torch.zeros((1, logits.shape[1], D, H, W), dtype=torch.float32, device=volume.device)
            cnt = torch.zeros((D, H, W), dtype=torch.float32, device=volume.device)
        _accumulate(acc, cnt, logits, (d0, h0, w0))
    if acc is None:  # more ranks than windows
        probe = model(volume[:, :, :win[0], :win[1], :win[2]].contiguous())
        c = (probe[0] if isinstance(probe, tuple) else probe).shape[1]
        acc = torch.zeros((1, c, D, H, W), dtype=torch.float32, device=volume.device)
        cnt = torch.zeros((D, H, W), dtype=torch.float32, device=volume.device)
    if world > 1:
        dist.all_reduce(acc, group=process_group)
        dist.all_reduce(cnt, group=process_group)
    check(L.b200_window_finalize(_ptr(acc), _ptr(cnt), acc.shape[1], D * H * W, _stream()), "window_finalize")
    model.train(was_training)
    return acc


def organ_metrics_from_confusion(conf, classes=(1, 2, 3)):
    """The evaluator's rule (reference test_model.py:265-285): Dice = 2I/(P+T), IoU = I/(P+T-I), 0.0 when the
    class is absent from the target (or the denominator is 0); means over ALL listed classes."""
    out = {}
    for k in classes:
        T, P, I = int(conf[k, :].sum()), int(conf[:, k].sum()), int(conf[k, k])
        dice = (2.0 * I / (P + T)) if (T > 0 and P + T > 0) else 0.0
        iou = (I / (P + T - I)) if (T > 0 and P + T - I > 0) else 0.0
        out[k] = {"dice": dice, "iou": iou}
    out["mean_dice"] = sum(out[k]["dice"] for k in classes) / len(classes)
    out["mean_iou"] = sum(out[k]["iou"] for k in classes) / len(classes)
    return out


@torch.no_grad()
def evaluate_volume(model, volume, labels, window=128, stride=64, classes=(1, 2, 3), process_group=None):
    """returns (stitched logits, int64 confusion matrix on the host, per-organ metrics dict)"""
    logits = sliding_window_logits(model, volume, window, stride, process_group)
    conf = confusion_counts(logits, labels).cpu().numpy()
    return logits, conf, organ_metrics_from_confusion(conf, classes)
