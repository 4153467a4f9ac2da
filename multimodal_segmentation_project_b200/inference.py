"""Sliding-window inference + per-organ metrics (BASELINE config #5).

The reference evaluates whole volumes with one forward (test_model.py:248) and has no sliding-window
code; this wrapper is the build-defined tiling of that evaluation: windows are
independent forwards of ``model.eval()``, overlapping logits are averaged, the stitched logits go
through the fused argmax + confusion-count kernel and the evaluator's Dice/IoU rules
(test_model.py:265-285) are applied to the integer counts.  With torch.distributed initialised the OUTPUT
volume is sharded into D-slabs, one owner per slab: no accumulator crosses NVLink, the only collective of an evaluation is one
C x C int64 all-reduce of the confusion counts."""
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


def owned_slab(D: int, rank: int, world: int):
    """Output planes [d_lo, d_hi) owned by `rank`: the volume is cut along D into `world` contiguous slabs (ceil division)."""
    per = (D + world - 1) // world
    return min(rank * per, D), min((rank + 1) * per, D)


@torch.no_grad()
def sliding_window_logits(model, volume: torch.Tensor, window=128, stride=64, process_group=None, gather=True):
    """volume [1, Cin, D, H, W] (CUDA) -> averaged logits fp32.

    Multi-GPU (torch.distributed initialised): ownership of the OUTPUT is sharded, not the windows — rank r owns a slab of D planes
    and runs every window that touches it (windows straddling two slabs are computed by both owners), accumulating only inside
    its slab.  No accumulator ever crosses NVLink: overlap averaging never crosses an ownership boundary.  With
    ``gather=True`` the finished slabs are all-gathered into the full ``[1, C, D, H, W]`` volume; with ``gather=False`` the call
    returns ``(slab_logits [1, C, d_hi - d_lo, H, W], (d_lo, d_hi))`` and nothing but the caller's own reductions is exchanged
    (evaluate_volume() all-reduces the C x C int64 confusion counts: 16 values)."""
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
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
    d_lo, d_hi = owned_slab(D, rank, world)
    acc = _slab_logits(model, volume, win, stride, d_lo, d_hi)
    C = None if acc is None else acc.shape[1]
    if world > 1:   # agree on the class count (a rank with an empty slab has run no window)
        t = torch.tensor([0 if C is None else C], device=volume.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=process_group)
        C = int(t.item())
    if acc is None:
        acc = torch.zeros((1, C, d_hi - d_lo, H, W), dtype=torch.float32, device=volume.device)
    if world == 1:
        return acc if gather else (acc, (d_lo, d_hi))
    if not gather:
        return acc, (d_lo, d_hi)
    per = (D + world - 1) // world
    pad = torch.zeros((1, C, per, H, W), dtype=torch.float32, device=volume.device)
    pad[:, :, : d_hi - d_lo] = acc
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=process_group)
    return torch.cat(parts, dim=2)[:, :, :D].contiguous()


@torch.no_grad()
def _slab_logits(model, volume, win, stride, d_lo, d_hi):
    """Averaged logits of output planes [d_lo, d_hi): every window touching the slab, accumulated inside it only.  None if no window ran."""
    L = _lib.load()
    D, H, W = volume.shape[2:]
    origins = [(d, h, w) for d in window_starts(D, win[0], stride[0]) if d < d_hi and d + win[0] > d_lo
               for h in window_starts(H, win[1], stride[1]) for w in window_starts(W, win[2], stride[2])]
    was_training = model.training
    model.eval()
    acc = cnt = None
    for (d0, h0, w0) in origins:
        patch = volume[:, :, d0:d0 + win[0], h0:h0 + win[1], w0:w0 + win[2]].contiguous()
        out = model(patch)
        logits = (out[0] if isinstance(out, tuple) else out).float().contiguous()
        if acc is None:
            acc = torch.zeros((1, logits.shape[1], d_hi - d_lo, H, W), dtype=torch.float32, device=volume.device)
            cnt = torch.zeros((d_hi - d_lo, H, W), dtype=torch.float32, device=volume.device)
        # clip the window to the owned planes: a view of the window's logits, accumulated at slab-relative coordinates
        a, b = max(d0, d_lo), min(d0 + win[0], d_hi)
        _accumulate(acc, cnt, logits[:, :, a - d0:b - d0].contiguous() if (a > d0 or b < d0 + win[0]) else logits, (a - d_lo, h0, w0))
    model.train(was_training)
    if acc is not None and d_hi > d_lo:
        check(L.b200_window_finalize(_ptr(acc), _ptr(cnt), acc.shape[1], (d_hi - d_lo) * H * W, _stream()), "window_finalize")
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
    """returns (stitched logits, int64 confusion matrix on the host, per-organ metrics dict).  Multi-GPU: every rank counts its own
    slab and ONE C x C int64 all-reduce makes the counts global; the returned logits are then the rank's own slab (use
    sliding_window_logits(gather=True) for the whole volume)."""
    world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        logits = sliding_window_logits(model, volume, window, stride, process_group)
        conf = confusion_counts(logits, labels).cpu().numpy()
        return logits, conf, organ_metrics_from_confusion(conf, classes)
    logits, (d_lo, d_hi) = sliding_window_logits(model, volume, window, stride, process_group, gather=False)
    C = logits.shape[1]
    if d_hi > d_lo:
        conf_t = confusion_counts(logits, labels[:, :, d_lo:d_hi].contiguous())
    else:
        conf_t = torch.zeros((C, C), dtype=torch.int64, device=volume.device)
    dist.all_reduce(conf_t, op=dist.ReduceOp.SUM, group=process_group)
    conf = conf_t.cpu().numpy()
    return logits, conf, organ_metrics_from_confusion(conf, classes)
