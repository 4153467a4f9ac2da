"""Restatement of the reference's utils/metrics.py (TEST INFRASTRUCTURE — see oracle/__init__.py).

Losses follow SURVEY.md Appendix D (three batch-global sums per class), metrics follow Appendix E
(integer counts -> fp32 scalar recipe); both are checked against the reference's own functions by
tests/test_oracle_golden.py via the committed golden vectors.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ---- losses ---------------------------------------------------------------------------------
def _sums(pred, target):
    """CE mean and per-class I, P, T — the quantities of utils/metrics.py:24-34 / :145-152."""
    t = target.squeeze(1)
    logp = F.log_softmax(pred, dim=1)
    ce = -logp.gather(1, t.unsqueeze(1)).mean()                       # nn.CrossEntropyLoss(): :24
    p = logp.exp()
    C = pred.shape[1]
    onehot = F.one_hot(t, C).movedim(-1, 1).to(p.dtype)
    dims = [0] + list(range(2, pred.dim()))
    I = (p * onehot).sum(dims)
    P = p.sum(dims)
    T = onehot.sum(dims)
    return ce, I, P, T


def combined_loss(pred, target):
    """utils/metrics.py:14-40: CE + mean_{k>=1} [1 - (2 I_k + 1e-5)/(P_k + T_k + 1e-5)]"""
    ce, I, P, T = _sums(pred, target)
    dice = 1 - (2.0 * I[1:] + 1e-5) / (P[1:] + T[1:] + 1e-5)
    return ce + dice.sum() / (pred.shape[1] - 1)


def dice_only_loss(pred, target):
    """train_unet.py:185-199 ('dice' closure)"""
    _, I, P, T = _sums(pred, target)
    dice = 1 - (2.0 * I[1:] + 1e-5) / (P[1:] + T[1:] + 1e-5)
    return dice.sum() / (pred.shape[1] - 1)


def tversky_loss(pred, target, alpha=0.5, beta=0.5, epsilon=1e-6):
    """utils/metrics.py:137-156: tp=I, fp=P-I, fn=T-I"""
    _, I, P, T = _sums(pred, target)
    tp, fp, fn = I[1:], (P - I)[1:], (T - I)[1:]
    tv = (tp + epsilon) / (tp + alpha * fp + beta * fn + epsilon)
    return (1 - tv).sum() / (pred.shape[1] - 1)


def combined_ce_tversky_loss(pred, target, alpha=0.7, beta=0.3):
    """utils/metrics.py:158-167 — the 0.3 / 0.7 weights are hard-coded there"""
    ce, _, _, _ = _sums(pred, target)
    return 0.3 * ce + 0.7 * tversky_loss(pred, target, alpha, beta)


def distillation_loss(student, teacher, target, alpha=0.7, temperature=2.0):
    """utils/metrics.py:169-190 — KL is the mean over ALL B*C*... elements, times T^2"""
    seg = combined_ce_tversky_loss(student, target)
    ls = F.log_softmax(student / temperature, dim=1)
    q = F.softmax(teacher / temperature, dim=1)
    kl = torch.where(q > 0, q * (q.clamp_min(1e-45).log() - ls), torch.zeros_like(q)).mean() * temperature ** 2
    return alpha * seg + (1 - alpha) * kl


# ---- metrics --------------------------------------------------------------------------------
def confusion_counts(pred, target) -> np.ndarray:
    """conf[t, p] int64 with torch.argmax semantics (first max, NaN = max) — utils/metrics.py:74,101,128"""
    C = pred.shape[1]
    am = torch.argmax(pred, dim=1).reshape(-1).cpu().numpy().astype(np.int64)
    t = target.reshape(-1).cpu().numpy().astype(np.int64)
    ok = (t >= 0) & (t < C)
    return np.bincount(t[ok] * C + am[ok], minlength=C * C).reshape(C, C).astype(np.int64)


def dice_iou_accuracy(pred, target):
    """Appendix E: the reference's scalars from integer counts, fp32, unfused.  Returns python/np values:
    (dice, iou, accuracy) where dice/iou are the int 0 when no foreground class is present."""
    f32 = np.float32
    conf = confusion_counts(pred, target)
    C = conf.shape[0]
    first_spatial = pred.shape[2]          # the class-loop bound bug, utils/metrics.py:78,105
    eps = f32(1e-5)
    dice = iou = f32(0)
    valid = 0
    for k in range(1, min(C, first_spatial)):
        T = int(conf[k].sum())
        if T > 0:
            P = int(conf[:, k].sum())
            I = f32(int(conf[k, k]))
            dice = f32(dice + f32(f32(f32(2.0) * I) + eps) / f32(f32(P + T) + eps))
            iou = f32(iou + f32(I + eps) / f32(f32(f32(P + T) - I) + eps))
            valid += 1
    acc = f32(f32(int(np.trace(conf))) / f32(target.numel()))
    if valid == 0:
        return 0, 0, acc
    return f32(dice / f32(valid)), f32(iou / f32(valid)), acc


# ---- binary helpers (dead code in the reference's scripts, part of the module API) -----------
def binary_counts(pred, target):
    """Per-sample integer counts behind utils/metrics.py:42-63: I = #(pred>0.5 & t), P = #(pred>0.5), T = sum(t), correct."""
    B = pred.shape[0]
    p = (pred.reshape(B, -1) > 0.5).cpu().numpy()
    t = target.reshape(B, -1).cpu().numpy() != 0
    return [(int((p[b] & t[b]).sum()), int(p[b].sum()), int(t[b].sum()), int((p[b] == t[b]).sum())) for b in range(B)]


def dice_score(pred, target, epsilon=1e-6):
    """utils/metrics.py:42-48 — fp32 per-sample ratios, then their fp32 mean"""
    f32 = np.float32
    v = [f32(f32(f32(2.0) * f32(I)) + f32(epsilon)) / f32(f32(f32(P) + f32(T)) + f32(epsilon)) for I, P, T, _ in binary_counts(pred, target)]
    return float(np.mean(np.asarray(v, dtype=np.float32), dtype=np.float32))


def iou_score(pred, target, epsilon=1e-6):
    """utils/metrics.py:50-56"""
    f32 = np.float32
    v = [f32(f32(I) + f32(epsilon)) / f32(f32(f32(f32(P) + f32(T)) - f32(I)) + f32(epsilon)) for I, P, T, _ in binary_counts(pred, target)]
    return float(np.mean(np.asarray(v, dtype=np.float32), dtype=np.float32))


def accuracy_score(pred, target):
    """utils/metrics.py:58-63"""
    f32 = np.float32
    return float(f32(sum(c[3] for c in binary_counts(pred, target))) / f32(target.numel()))


def dice_loss(pred, target, epsilon=1e-6):
    """utils/metrics.py:6-12: 1 - (2 sum(sigmoid(x) t) + eps) / (sum(sigmoid(x)) + sum(t) + eps)"""
    p = torch.sigmoid(pred).reshape(-1)
    t = target.reshape(-1).to(p.dtype)
    return 1 - (2.0 * (p * t).sum() + epsilon) / (p.sum() + t.sum() + epsilon)
