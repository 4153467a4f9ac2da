"""CPU property tests (hypothesis): the oracle's restatements against second, independent formulations of the same reference
lines, and the host-side logic of the drop-in (flat parameter layout, sliding-window tiling, slab ownership) on random shapes.
SURVEY.md §4 "Unit (CPU, no GPU)": odd sizes, absent classes, ties, NaN, class-loop-bound quirk.  No GPU, no /root/reference."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as TF
from hypothesis import given, settings, strategies as st

from multimodal_segmentation_project_b200.dp import FlatParams
from multimodal_segmentation_project_b200.inference import organ_metrics_from_confusion, owned_slab, window_starts
from multimodal_segmentation_project_b200.models.unet import UNet3D
from oracle import metrics_oracle as OM

SETTINGS = dict(max_examples=40, deadline=None, derandomize=True)   # the same examples on every run


def _case(seed, B, C, D, H, W, absent, ties, nan):
    g = torch.Generator().manual_seed(seed)
    pred = torch.randn(B, C, D, H, W, generator=g)
    if ties:                                   # exact ties between classes: argmax must take the first maximum
        pred = torch.round(pred)
    if nan:                                    # NaN counts as the maximum for torch.argmax
        pred.view(-1)[:: 7] = float("nan")
    target = torch.randint(0, C, (B, 1, D, H, W), generator=g)
    if absent and C > 2:
        target[target == C - 1] = 0            # a foreground class that never occurs in the target
    return pred, target


# ------------------------------------------------------------------------------------------------ metrics (utils/metrics.py:65-129)
def _masks_metrics(pred, target):
    """The reference's formulation restated independently of the oracle: arg-max mask, per-class boolean masks, torch fp32 sums;
    the class loop runs to pred.size(1) AFTER the arg-max, i.e. to the first spatial size."""
    t = target.squeeze(1)
    am = torch.argmax(pred, dim=1)
    dice = iou = 0
    valid = 0
    for k in range(1, am.size(1)):
        pm, tm = am == k, t == k
        if tm.sum() > 0:
            inter = (pm & tm).sum().float()
            both = pm.sum() + tm.sum()
            dice = dice + (2.0 * inter + 1e-5) / (both + 1e-5)
            iou = iou + (inter + 1e-5) / (both - inter + 1e-5)
            valid += 1
    acc = (am == t).float().mean()
    return dice / max(valid, 1), iou / max(valid, 1), acc


@settings(**SETTINGS)
@given(seed=st.integers(0, 2 ** 20), B=st.integers(1, 2), C=st.integers(2, 5), D=st.integers(1, 6), H=st.integers(1, 5), W=st.integers(1, 5),
       absent=st.booleans(), ties=st.booleans(), nan=st.booleans())
def test_metric_recipe_from_counts_equals_mask_formulation(seed, B, C, D, H, W, absent, ties, nan):
    pred, target = _case(seed, B, C, D, H, W, absent, ties, nan)
    d0, i0, a0 = OM.dice_iou_accuracy(pred, target)
    d1, i1, a1 = _masks_metrics(pred, target)
    # integer counts -> the same fp32 operations in the same order: bit-equal (counts here are far below 2^24)
    assert np.float32(d0) == np.float32(float(d1)) and np.float32(i0) == np.float32(float(i1)) and np.float32(a0) == np.float32(float(a1))
    conf = OM.confusion_counts(pred, target)
    assert conf.sum() == target.numel() and (conf >= 0).all()
    assert int(np.trace(conf)) == int((torch.argmax(pred, dim=1) == target.squeeze(1)).sum())


@settings(**SETTINGS)
@given(seed=st.integers(0, 2 ** 20), C=st.integers(2, 5), n=st.integers(1, 200))
def test_organ_metrics_rule(seed, C, n):
    """inference.organ_metrics_from_confusion — test_model.py:265-285: Dice 2I/(P+T), IoU I/(P+T-I), 0.0 for a class absent from the target."""
    rng = np.random.default_rng(seed)
    conf = rng.integers(0, n, size=(C, C)).astype(np.int64)
    if C > 2:
        conf[C - 1, :] = 0                    # class C-1 absent from the target
    classes = tuple(range(1, C))
    m = organ_metrics_from_confusion(conf, classes)
    for k in classes:
        T, P, I = conf[k].sum(), conf[:, k].sum(), conf[k, k]
        want_d = 2.0 * I / (P + T) if T > 0 else 0.0
        want_i = I / (P + T - I) if (T > 0 and P + T - I > 0) else 0.0
        assert math.isclose(m[k]["dice"], want_d, rel_tol=1e-12, abs_tol=0) and math.isclose(m[k]["iou"], want_i, rel_tol=1e-12, abs_tol=0)
        assert 0.0 <= m[k]["iou"] <= m[k]["dice"] <= 1.0
    assert math.isclose(m["mean_dice"], sum(m[k]["dice"] for k in classes) / len(classes), rel_tol=1e-12)


# ------------------------------------------------------------------------------------------------ losses (utils/metrics.py:14-40, 137-190)
def _dice_ce_direct(pred, target):
    """combined_loss written the long way round (torch's own cross_entropy, one-hot targets, explicit soft-Dice per foreground
    class with the reference's 1e-5), fp64."""
    p = pred.double()
    t = target.squeeze(1)
    C = p.shape[1]
    ce = TF.cross_entropy(p, t)
    sm = torch.softmax(p, dim=1)
    oh = TF.one_hot(t, C).movedim(-1, 1).double()
    dice = 0.0
    for k in range(1, C):                      # background skipped, utils/metrics.py:28
        inter, card = (sm[:, k] * oh[:, k]).sum(), sm[:, k].sum() + oh[:, k].sum()
        dice = dice + (1.0 - (2.0 * inter + 1e-5) / (card + 1e-5))
    return ce, dice / (C - 1)


@settings(**SETTINGS)
@given(seed=st.integers(0, 2 ** 20), B=st.integers(1, 2), C=st.integers(2, 4), D=st.integers(1, 4), H=st.integers(1, 4), W=st.integers(1, 4),
       absent=st.booleans())
def test_loss_restatements_are_consistent(seed, B, C, D, H, W, absent):
    pred, target = _case(seed, B, C, D, H, W, absent, False, False)
    pred = pred.double().requires_grad_(True)
    # softmax shift invariance: adding a per-voxel constant to every class logit changes neither the losses nor their gradients
    shift = torch.randn(B, 1, D, H, W, dtype=torch.float64, generator=torch.Generator().manual_seed(seed + 1))
    for fn in (OM.combined_loss, OM.dice_only_loss, lambda p, t: OM.tversky_loss(p, t, 0.3, 0.7), lambda p, t: OM.combined_ce_tversky_loss(p, t, 0.7, 0.3)):
        a = fn(pred, target)
        b = fn(pred + shift, target)
        assert torch.isfinite(a) and abs(a.item() - b.item()) <= 1e-9 * max(1.0, abs(a.item()))
        (ga,) = torch.autograd.grad(a, pred, retain_graph=True)
        # every per-voxel gradient sums to zero over the classes (the softmax Jacobian annihilates constants)
        assert ga.sum(dim=1).abs().max().item() <= 1e-12
    # the three-sums form against the long way round
    ce, dice = _dice_ce_direct(pred, target)
    assert abs(OM.combined_loss(pred, target).item() - (ce + dice).item()) <= 1e-12
    assert abs(OM.dice_only_loss(pred, target).item() - dice.item()) <= 1e-12
    # Tversky with alpha = beta = 0.5 is the Dice ratio with epsilon 2e-6 instead of 1e-5 (utils/metrics.py:137-156 against :14-40)
    _, I, P, T = OM._sums(pred, target)
    want = (1.0 - (2.0 * I[1:] + 2e-6) / (P[1:] + T[1:] + 2e-6)).sum() / (C - 1)
    assert abs(OM.tversky_loss(pred, target, 0.5, 0.5).item() - want.item()) <= 1e-12
    # distillation: a student that equals its teacher has zero KL, so the loss is alpha x the segmentation loss (distill_unet.py:107-119)
    kd = OM.distillation_loss(pred, pred.detach(), target, alpha=0.7, temperature=2.0)
    seg = OM.combined_ce_tversky_loss(pred, target)
    assert abs(kd.item() - 0.7 * seg.item()) <= 1e-9
    # and the KL term is non-negative for any teacher
    teacher = torch.randn_like(pred)
    kd2 = OM.distillation_loss(pred, teacher, target, alpha=0.0, temperature=1.5)
    assert kd2.item() >= -1e-12


# ------------------------------------------------------------------------------------------------ host logic of the drop-in
@settings(**SETTINGS)
@given(size=st.integers(1, 600), window=st.integers(1, 200), stride=st.integers(1, 200))
def test_window_starts_cover_the_axis(size, window, stride):
    starts = window_starts(size, window, stride)
    w = min(window, size)
    assert starts[0] == 0 and starts == sorted(set(starts))
    assert all(0 <= s and s + w <= size for s in starts) and starts[-1] + w == size
    if stride <= w:                            # overlapping or abutting windows leave no gap
        covered = np.zeros(size, dtype=bool)
        for s in starts:
            covered[s:s + w] = True
        assert covered.all()


@settings(**SETTINGS)
@given(D=st.integers(1, 500), world=st.integers(1, 16))
def test_owned_slabs_partition_the_volume(D, world):
    slabs = [owned_slab(D, r, world) for r in range(world)]
    assert slabs[0][0] == 0 and slabs[-1][1] == D
    assert all(lo <= hi for lo, hi in slabs) and all(slabs[r][1] == slabs[r + 1][0] for r in range(world - 1))
    assert sum(hi - lo for lo, hi in slabs) == D


@settings(max_examples=8, deadline=None, derandomize=True)
@given(levels=st.integers(1, 4), base=st.sampled_from([4, 8]), cin=st.integers(1, 2), cout=st.integers(1, 4), freeze=st.booleans())
def test_flat_params_layout_properties(levels, base, cin, cout, freeze):
    """dp.FlatParams on arbitrary `features`: parameters keep their values and order, slices are disjoint, every bucket starts on a
    16-byte boundary, buckets partition the buffer, the last bucket is the top encoder level (when there are at least two)."""
    torch.manual_seed(0)
    net = UNet3D(cin, cout, features=[base * 2 ** i for i in range(levels)])
    if freeze:
        for p in net.encoder.parameters():
            p.requires_grad = False
    before = {k: v.clone() for k, v in net.state_dict().items()}
    fp = FlatParams(net)
    after = net.state_dict()
    assert list(after) == list(before) and all(torch.equal(after[k], before[k]) for k in before)
    spans = sorted((o, o + k) for o, k in fp.offsets.values())
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= fp.total
    assert fp.total % 4 == 0 and all(e % 4 == 0 for e in fp.bucket_elems) and fp.bucket_elems[0] == 0 and fp.bucket_elems[-1] == fp.total
    assert sum(b.numel() for b in fp.buckets()) == fp.total
    trainable = [n for n, p in net.named_parameters() if p.requires_grad]
    assert sorted(fp.offsets) == sorted(trainable) and len(fp.sentinels) in (0, fp.n_buckets - 1)
    if not freeze and levels >= 2 and fp.n_buckets >= 2:
        assert all(n.startswith("encoder.0.") for n, _ in fp.order[fp.bucket_params[-2]:])
    # a gradient written through the flat buffer is the parameter's .grad (what the gradient sinks rely on)
    fp.grad.fill_(1.0)
    assert all(p.grad is not None and float(p.grad.min()) == 1.0 for p in net.parameters() if p.requires_grad)
