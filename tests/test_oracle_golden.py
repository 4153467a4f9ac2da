"""CPU: pins the oracle (oracle/) against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  Bit-exact for initial weights and metrics; tight tolerances for fp32 math."""
import os

import numpy as np
import pytest
import torch

from multimodal_segmentation_project_b200.synthetic import structured_volume, worst_case_volume
from oracle import dann_oracle as OD
from oracle import metrics_oracle as OM
from oracle.unet_oracle import init_state_dict, train_step_grads, unet3d_forward


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_init_matches_reference_constructor(golden_dir):
    g = _load(golden_dir, "unet_train_b2_s16.npz")
    sd = init_state_dict(1, 4, seed=0)
    assert len(sd) == 136
    for k, v in sd.items():
        ref = g["init/" + k]
        assert v.double().sum().item() == ref[0] and v.double().abs().sum().item() == ref[1], k
    assert sum(v.numel() for k, v in sd.items() if "running" not in k and "num_batches" not in k) == 5647908


def test_train_step_matches_reference(golden_dir):
    g = _load(golden_dir, "unet_train_b2_s16.npz")
    sd = init_state_dict(1, 4, seed=0)
    x, y = structured_volume(2, 16, seed=1234)
    loss, logits, grads, bufs = train_step_grads(sd, x, y, OM.combined_loss)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=1e-6)
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    for k, gr in grads.items():
        assert abs(gr.double().norm().item() - float(g["gradnorm/" + k])) <= 1e-5 * max(1e-3, float(g["gradnorm/" + k])) + 1e-7, k
        if "grad/" + k in g.files:
            np.testing.assert_allclose(gr.numpy(), g["grad/" + k], rtol=1e-4, atol=1e-7)
    for k, b in bufs.items():
        np.testing.assert_allclose(b.numpy(), g["buf/" + k], rtol=1e-6, atol=1e-7)
    # eval forward with the updated running statistics
    for k, b in bufs.items():
        sd[k] = b
    ge = _load(golden_dir, "unet_eval_s16x32x16.npz")
    x2, _ = structured_volume(1, (16, 32, 16), seed=77)
    with torch.no_grad():
        out = unet3d_forward(sd, x2, training=False)
    np.testing.assert_allclose(out.numpy(), ge["logits"], rtol=0, atol=2e-6)


def test_odd_sizes_match_reference(golden_dir):
    g = _load(golden_dir, "unet_train_odd_20x18x22.npz")
    sd = init_state_dict(1, 4, seed=0)
    x, y = structured_volume(2, (20, 18, 22), seed=5)
    loss, logits, grads, _ = train_step_grads(sd, x, y, lambda p, t: OM.combined_ce_tversky_loss(p, t, 0.5, 0.5))
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=1e-6)
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    for k, gr in grads.items():
        ref = float(g["gradnorm/" + k])
        assert abs(gr.double().norm().item() - ref) <= 1e-4 * max(ref, 1e-4) + 1e-7, k


def test_dann_variant_and_small_net(golden_dir):
    g = _load(golden_dir, "unet_dann_b2_s16.npz")
    sd = init_state_dict(1, 4, seed=0)
    x, _ = structured_volume(2, 16, seed=1234)
    with torch.no_grad():
        logits, gap = unet3d_forward(sd, x, True, return_features=True)
        logits2, none = unet3d_forward(sd, x, True, return_features=False)
    assert none is None and gap.shape == (2, 256)
    np.testing.assert_allclose(logits.numpy(), g["logits"], atol=1e-6)
    np.testing.assert_allclose(gap.numpy(), g["gap"], atol=1e-6)
    np.testing.assert_allclose(logits2.numpy(), g["logits_second_call"], atol=1e-6)

    gs = _load(golden_dir, "unet_small_in2_out3_f8_16.npz")
    sds = init_state_dict(2, 3, features=(8, 16), seed=3)
    gen = torch.Generator().manual_seed(11)
    xs = torch.randn(2, 2, 8, 12, 8, generator=gen)
    ys = torch.randint(0, 3, (2, 1, 8, 12, 8), generator=gen)
    loss, logits, grads, _ = train_step_grads(sds, xs, ys, OM.combined_loss)
    np.testing.assert_allclose(logits.numpy(), gs["logits"], atol=1e-6)
    for k, gr in grads.items():
        np.testing.assert_allclose(gr.numpy(), gs["grad/" + k], rtol=1e-4, atol=1e-6)


LOSS_CASES = {
    "combined_loss": lambda z, t, y: OM.combined_loss(z, y),
    "tversky_05_05": lambda z, t, y: OM.tversky_loss(z, y),
    "tversky_07_03": lambda z, t, y: OM.tversky_loss(z, y, 0.7, 0.3),
    "ce_tversky_07_03": lambda z, t, y: OM.combined_ce_tversky_loss(z, y),
    "ce_tversky_05_05": lambda z, t, y: OM.combined_ce_tversky_loss(z, y, 0.5, 0.5),
    "distill_a07_t2": lambda z, t, y: OM.distillation_loss(z, t, y),
    "distill_a05_t4": lambda z, t, y: OM.distillation_loss(z, t, y, 0.5, 4.0),
}


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_losses_match_reference(golden_dir, case):
    g = _load(golden_dir, "losses.npz")
    t = torch.from_numpy(g[f"{case}/teacher"])
    y = torch.from_numpy(g[f"{case}/target"])
    for name, fn in LOSS_CASES.items():
        z = torch.from_numpy(g[f"{case}/logits"]).clone().requires_grad_(True)
        l = fn(z, t, y)
        l.backward()
        assert abs(l.item() - float(g[f"{case}/{name}/loss"])) < 2e-6, name
        np.testing.assert_allclose(z.grad.numpy(), g[f"{case}/{name}/grad"], rtol=1e-4, atol=1e-7, err_msg=name)


METRIC_CASES = ["normal", "absent_class", "first_spatial_2", "first_spatial_3", "no_foreground", "ties", "nan", "structured_like"]


@pytest.mark.parametrize("case", METRIC_CASES)
def test_metrics_bit_exact_vs_reference(golden_dir, case):
    g = _load(golden_dir, "metrics.npz")
    pred, tgt = torch.from_numpy(g[f"{case}/pred"]), torch.from_numpy(g[f"{case}/target"])
    d, i, a = OM.dice_iou_accuracy(pred, tgt)
    assert np.float32(d) == g[f"{case}/dice"] and np.float32(i) == g[f"{case}/iou"], case
    assert np.float32(a) == g[f"{case}/acc"], case
    assert bool(g[f"{case}/dice_is_tensor"]) == (not isinstance(d, int))


def test_metrics_counts_above_2_24(golden_dir):
    g = _load(golden_dir, "metrics.npz")
    gen = torch.Generator().manual_seed(33)
    pred = torch.randn(1, 4, 150, 400, 300, generator=gen)
    pred[:, 1] += 2.5
    tgt = (torch.rand(1, 1, 150, 400, 300, generator=gen) < 0.97).long()
    d, i, a = OM.dice_iou_accuracy(pred, tgt)
    assert np.float32(d) == g["big_counts/dice"] and np.float32(i) == g["big_counts/iou"]
    assert abs(float(a) - float(g["big_counts/acc"])) <= 2e-7  # fp32 mean above 2^24 is order dependent: 1 ulp


def test_dann_head_matches_reference(golden_dir):
    g = _load(golden_dir, "dann_head.npz")
    sd = OD.init_discriminator(256, seed=0)
    for k, v in sd.items():
        ref = g["init/" + k]
        assert v.double().sum().item() == ref[0] and v.double().abs().sum().item() == ref[1], k
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    fs = torch.from_numpy(g["fs"]).clone().requires_grad_(True)
    ft = torch.from_numpy(g["ft"]).clone().requires_grad_(True)
    lam = float(g["lambda"])
    so = OD.discriminator_forward(p, OD.grad_reverse(fs, lam))
    np.testing.assert_allclose(so.detach().numpy(), g["source_out"], atol=1e-6)
    dl = OD.domain_loss(p, fs, ft, lam)
    assert abs(dl.item() - float(g["domain_loss"])) < 1e-6
    (lam * dl).backward()
    np.testing.assert_allclose(fs.grad.numpy(), g["grad_fs"], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(ft.grad.numpy(), g["grad_ft"], rtol=1e-4, atol=1e-8)
    for k in p:
        np.testing.assert_allclose(p[k].grad.numpy(), g["grad/" + k], rtol=1e-4, atol=1e-8)


BINARY_CASES = ["a", "b", "c", "empty_target", "full_target"]


@pytest.mark.parametrize("case", BINARY_CASES)
def test_binary_helpers_match_reference(golden_dir, case):
    """utils/metrics.py:6-12, 42-63, 131-135 (oracle/make_golden_binary.py ran the reference)."""
    g = np.load(os.path.join(golden_dir, "binary_helpers.npz"))
    prob, tgt = torch.from_numpy(g[f"{case}/prob"]), torch.from_numpy(g[f"{case}/target"])
    assert OM.dice_score(prob, tgt) == float(g[f"{case}/dice_score"])
    assert OM.iou_score(prob, tgt) == float(g[f"{case}/iou_score"])
    assert OM.accuracy_score(prob, tgt) == float(g[f"{case}/accuracy_score"])
    z = torch.from_numpy(g[f"{case}/logits"]).requires_grad_(True)
    loss = OM.dice_loss(z, tgt)
    loss.backward()
    assert abs(float(loss) - float(g[f"{case}/dice_loss"])) <= 1e-6
    ref = torch.from_numpy(g[f"{case}/dice_loss_grad"])
    assert (z.grad - ref).abs().max() <= 1e-6 * max(1.0, float(ref.abs().max()))
