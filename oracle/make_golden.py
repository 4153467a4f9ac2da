"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
on seeded inputs.  Run in the build container only (the reference does not travel to the GPU box):

    python -m oracle.make_golden

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every array written here is an output of the
reference's own code: models/unet.py, models/unet_dann.py, utils/metrics.py, train_dann.py:22-49.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("B200_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF)
    # train_dann.py imports plotting / IO packages at module top that are absent here; its hot-path
    # classes (GradientReversal, DomainDiscriminator) need none of them.
    class _Stub(types.ModuleType):
        __path__ = []

        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return type(item, (), {"__init__": lambda self, *a, **k: None})

    for name in ("matplotlib", "matplotlib.pyplot", "nibabel", "accelerate", "accelerate.utils", "monai", "monai.transforms",
                 "monai.data", "monai.utils", "seaborn"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    import models.unet as ref_unet
    import models.unet_dann as ref_unet_dann
    import utils.metrics as ref_metrics
    try:
        import train_dann as ref_dann
    except Exception as e:  # pragma: no cover
        print("train_dann import failed:", e)
        ref_dann = None
    return ref_unet, ref_unet_dann, ref_metrics, ref_dann


def _np(t):
    return t.detach().cpu().numpy()


def main():
    from multimodal_segmentation_project_b200.synthetic import structured_volume, worst_case_volume

    ref_unet, ref_unet_dann, RM, ref_dann = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- A: train-mode forward/backward, B=2, 16^3, default architecture, seed-0 init -------------
    torch.manual_seed(0)
    net = ref_unet.UNet3D(in_channels=1, out_channels=4, dropout_rate=0.0)
    init_sums = {k: np.array([v.double().sum().item(), v.double().abs().sum().item()]) for k, v in net.state_dict().items()}
    x, y = structured_volume(2, 16, seed=1234)
    net.train()
    logits = net(x)
    loss = RM.combined_loss(logits, y)
    loss.backward()
    a = {"logits": _np(logits), "loss": np.float32(loss.item())}
    for k, v in init_sums.items():
        a["init/" + k] = v
    for k, p in net.named_parameters():
        a["gradnorm/" + k] = np.float64(p.grad.double().norm().item())
        if p.numel() <= 8192:
            a["grad/" + k] = _np(p.grad)
    for k, v in net.state_dict().items():
        if "running" in k or "num_batches" in k:
            a["buf/" + k] = _np(v)
    np.savez_compressed(os.path.join(OUT, "unet_train_b2_s16.npz"), **a)

    # ---- B: eval-mode forward with the running statistics produced by A ---------------------------
    net.eval()
    x2, _ = structured_volume(1, (16, 32, 16), seed=77)
    with torch.no_grad():
        np.savez_compressed(os.path.join(OUT, "unet_eval_s16x32x16.npz"), logits=_np(net(x2)))

    # ---- C: odd sizes -> F.interpolate fix-up path (models/unet.py:81-83) -------------------------
    torch.manual_seed(0)
    net = ref_unet.UNet3D(in_channels=1, out_channels=4, dropout_rate=0.0)
    x3, y3 = structured_volume(2, (20, 18, 22), seed=5)
    net.train()
    lg = net(x3)
    ls = RM.combined_ce_tversky_loss(lg, y3, alpha=0.5, beta=0.5)
    ls.backward()
    c = {"logits": _np(lg), "loss": np.float32(ls.item())}
    for k, p in net.named_parameters():
        c["gradnorm/" + k] = np.float64(p.grad.double().norm().item())
    np.savez_compressed(os.path.join(OUT, "unet_train_odd_20x18x22.npz"), **c)

    # ---- D: DANN variant: tuple return + pooled bottleneck ----------------------------------------
    torch.manual_seed(0)
    dnet = ref_unet_dann.UNet3D(in_channels=1, out_channels=4, dropout_rate=0.0)
    dnet.train()
    lg, gap = dnet(x, return_features=True)
    lg2, none = dnet(x, return_features=False)
    assert none is None
    np.savez_compressed(os.path.join(OUT, "unet_dann_b2_s16.npz"), logits=_np(lg), gap=_np(gap), logits_second_call=_np(lg2))

    # ---- E: small 2-channel-in / 3-class / custom features net (generic constructor path) ----------
    torch.manual_seed(3)
    snet = ref_unet.UNet3D(in_channels=2, out_channels=3, features=[8, 16], dropout_rate=0.0)
    g = torch.Generator().manual_seed(11)
    xs = torch.randn(2, 2, 8, 12, 8, generator=g)
    ys = torch.randint(0, 3, (2, 1, 8, 12, 8), generator=g)
    snet.train()
    lgs = snet(xs)
    lss = RM.combined_loss(lgs, ys)
    lss.backward()
    e = {"logits": _np(lgs), "loss": np.float32(lss.item())}
    for k, p in snet.named_parameters():
        e["grad/" + k] = _np(p.grad)
    np.savez_compressed(os.path.join(OUT, "unet_small_in2_out3_f8_16.npz"), **e)

    # ---- F: losses on random logits -----------------------------------------------------------------
    out = {}
    for name, shape, C, seed in (("a", (2, 4, 5, 6, 7), 4, 1), ("b", (3, 3, 4, 4, 4), 3, 2), ("c", (1, 2, 3, 3, 3), 2, 3)):
        g = torch.Generator().manual_seed(seed)
        z = (torch.randn(shape, generator=g) * 2).requires_grad_(True)
        t = torch.randn(shape, generator=g) * 3
        yy = torch.randint(0, C, (shape[0], 1, *shape[2:]), generator=g)
        out[f"{name}/logits"], out[f"{name}/teacher"], out[f"{name}/target"] = _np(z), _np(t), _np(yy)
        cases = {
            "combined_loss": lambda: RM.combined_loss(z, yy),
            "tversky_05_05": lambda: RM.tversky_loss(z, yy),
            "tversky_07_03": lambda: RM.tversky_loss(z, yy, alpha=0.7, beta=0.3),
            "ce_tversky_07_03": lambda: RM.combined_ce_tversky_loss(z, yy),
            "ce_tversky_05_05": lambda: RM.combined_ce_tversky_loss(z, yy, alpha=0.5, beta=0.5),
            "distill_a07_t2": lambda: RM.distillation_loss(z, t, yy),
            "distill_a05_t4": lambda: RM.distillation_loss(z, t, yy, alpha=0.5, temperature=4.0),
        }
        for cname, fn in cases.items():
            z.grad = None
            l = fn()
            l.backward()
            out[f"{name}/{cname}/loss"] = np.float32(l.item())
            out[f"{name}/{cname}/grad"] = _np(z.grad)
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)

    # ---- G: metrics (incl. the class-loop-bound quirk, absent classes, ties, NaN) -------------------
    m = {}

    def metric_case(name, pred, tgt, store=True):
        d, i, acc = RM.calculate_dice(pred, tgt), RM.calculate_iou(pred, tgt), RM.calculate_accuracy(pred, tgt)
        if store:
            m[f"{name}/pred"], m[f"{name}/target"] = _np(pred), _np(tgt)
        m[f"{name}/dice"] = np.float32(d.item() if torch.is_tensor(d) else d)
        m[f"{name}/iou"] = np.float32(i.item() if torch.is_tensor(i) else i)
        m[f"{name}/acc"] = np.float32(acc.item())
        m[f"{name}/dice_is_tensor"] = np.bool_(torch.is_tensor(d))

    g = torch.Generator().manual_seed(21)
    metric_case("normal", torch.randn(2, 4, 8, 8, 8, generator=g), torch.randint(0, 4, (2, 1, 8, 8, 8), generator=g))
    metric_case("absent_class", torch.randn(1, 4, 6, 6, 6, generator=g), torch.randint(0, 2, (1, 1, 6, 6, 6), generator=g))
    metric_case("first_spatial_2", torch.randn(1, 4, 2, 8, 8, generator=g), torch.randint(0, 4, (1, 1, 2, 8, 8), generator=g))
    metric_case("first_spatial_3", torch.randn(2, 4, 3, 8, 8, generator=g), torch.randint(0, 4, (2, 1, 3, 8, 8), generator=g))
    metric_case("no_foreground", torch.randn(1, 4, 4, 4, 4, generator=g), torch.zeros(1, 1, 4, 4, 4, dtype=torch.long))
    metric_case("ties", torch.zeros(1, 4, 4, 4, 4), torch.randint(0, 4, (1, 1, 4, 4, 4), generator=g))
    pn = torch.randn(1, 4, 4, 4, 4, generator=g)
    pn[0, 2, 1] = float("nan")
    pn[0, 1, 2, 0] = float("nan")
    pn[0, 3, 2, 0] = float("nan")
    metric_case("nan", pn, torch.randint(0, 4, (1, 1, 4, 4, 4), generator=g))
    xw, yw = worst_case_volume(2, 16, seed=9)
    metric_case("structured_like", torch.randn(2, 4, 16, 16, 16, generator=g) + 2.0 * torch.nn.functional.one_hot(yw.squeeze(1), 4).movedim(-1, 1), yw)
    # counts above 2^24: regenerated from the seed by the test (inputs are not stored)
    g = torch.Generator().manual_seed(33)
    big_pred = torch.randn(1, 4, 150, 400, 300, generator=g)
    big_pred[:, 1] += 2.5
    big_tgt = (torch.rand(1, 1, 150, 400, 300, generator=g) < 0.97).long()
    metric_case("big_counts", big_pred, big_tgt, store=False)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **m)

    # ---- H: DANN head ------------------------------------------------------------------------------
    if ref_dann is not None:
        torch.manual_seed(0)
        disc = ref_dann.DomainDiscriminator(256)
        disc.eval()
        g = torch.Generator().manual_seed(4)
        fs = torch.randn(3, 256, generator=g).requires_grad_(True)
        ft = torch.randn(3, 256, generator=g).requires_grad_(True)
        lam = 0.2
        so = disc(ref_dann.grad_reverse(fs, lam))
        to = disc(ref_dann.grad_reverse(ft, lam))
        labels = torch.cat([torch.zeros(3, dtype=torch.long), torch.ones(3, dtype=torch.long)])
        dl = torch.nn.CrossEntropyLoss()(torch.cat([so, to]), labels)
        (lam * dl).backward()  # train_dann.py:260 applies lambda a second time
        h = {"fs": _np(fs), "ft": _np(ft), "source_out": _np(so), "target_out": _np(to), "domain_loss": np.float32(dl.item()),
             "grad_fs": _np(fs.grad), "grad_ft": _np(ft.grad), "lambda": np.float32(lam)}
        for k, v in disc.state_dict().items():
            h["init/" + k] = np.array([v.double().sum().item(), v.double().abs().sum().item()])
        for k, p in disc.named_parameters():
            h["grad/" + k] = _np(p.grad)
        np.savez_compressed(os.path.join(OUT, "dann_head.npz"), **h)
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
