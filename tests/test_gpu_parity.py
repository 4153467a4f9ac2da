"""GPU parity tests (run with -m gpu on a B200): every kernel goes through the C ABI and is
compared with the oracle (CPU restatement pinned to the reference) or with the golden vectors
produced by the reference itself.  Tolerances are written next to each assertion:
integer / index work bit-exact; fp32 rel-L2 <= 1e-4; bf16 rel-L2 <= 1e-2 against the same
rounding points (north_star)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as TF

from multimodal_segmentation_project_b200 import _lib
from multimodal_segmentation_project_b200 import functional as F
from multimodal_segmentation_project_b200.models.unet import DoubleConv, UNet3D
from multimodal_segmentation_project_b200.models.unet_dann import UNet3D as UNet3DDann
from multimodal_segmentation_project_b200.synthetic import structured_volume
from multimodal_segmentation_project_b200.train_dann import DomainDiscriminator, domain_cross_entropy, grad_reverse
from multimodal_segmentation_project_b200.utils import metrics as M
from oracle import dann_oracle as OD
from oracle import metrics_oracle as OM
from oracle.unet_oracle import init_state_dict, train_step_grads, unet3d_forward

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cl(x, dtype=torch.float32):  # NCDHW cpu -> NDHWC cuda
    return x.permute(0, 2, 3, 4, 1).contiguous().to("cuda", dtype)


def cf(x):  # NDHWC cuda -> NCDHW cpu fp32
    return x.float().permute(0, 4, 1, 2, 3).contiguous().cpu()


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# --------------------------------------------------------------------------------- losses
LOSS_FNS = {
    "combined_loss": (lambda z, t, y: M.combined_loss(z, y)),
    "tversky_05_05": (lambda z, t, y: M.tversky_loss(z, y)),
    "tversky_07_03": (lambda z, t, y: M.tversky_loss(z, y, alpha=0.7, beta=0.3)),
    "ce_tversky_07_03": (lambda z, t, y: M.combined_ce_tversky_loss(z, y)),
    "ce_tversky_05_05": (lambda z, t, y: M.combined_ce_tversky_loss(z, y, alpha=0.5, beta=0.5)),
    "distill_a07_t2": (lambda z, t, y: M.distillation_loss(z, t, y)),
    "distill_a05_t4": (lambda z, t, y: M.distillation_loss(z, t, y, alpha=0.5, temperature=4.0)),
}


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_losses_vs_reference_golden(cuda_dev, golden_dir, case):
    g = _load(golden_dir, "losses.npz")
    t = torch.from_numpy(g[f"{case}/teacher"]).to(cuda_dev)
    y = torch.from_numpy(g[f"{case}/target"]).to(cuda_dev)
    for name, fn in LOSS_FNS.items():
        z = torch.from_numpy(g[f"{case}/logits"]).to(cuda_dev).requires_grad_(True)
        loss = fn(z, t, y)
        loss.backward()
        ref = float(g[f"{case}/{name}/loss"])
        assert abs(loss.item() - ref) <= 1e-5 * max(1.0, abs(ref)), (name, loss.item(), ref)  # fp32: rel 1e-5
        assert rel_l2(z.grad, torch.from_numpy(g[f"{case}/{name}/grad"])) <= 1e-5, name


@pytest.mark.parametrize("shape", [(2, 4, 32, 32, 32), (1, 4, 7, 9, 11), (2, 5, 8, 8, 8)])
def test_losses_vs_oracle_random(cuda_dev, shape):
    gen = torch.Generator().manual_seed(5)
    z0 = torch.randn(shape, generator=gen) * 3
    t0 = torch.randn(shape, generator=gen) * 3
    y0 = torch.randint(0, shape[1], (shape[0], 1, *shape[2:]), generator=gen)
    pairs = [
        (lambda z, t, y: M.combined_loss(z, y), lambda z, t, y: OM.combined_loss(z, y)),
        (lambda z, t, y: M.combined_ce_tversky_loss(z, y, 0.5, 0.5), lambda z, t, y: OM.combined_ce_tversky_loss(z, y, 0.5, 0.5)),
        (lambda z, t, y: M.distillation_loss(z, t, y, 0.7, 4.0), lambda z, t, y: OM.distillation_loss(z, t, y, 0.7, 4.0)),
        (lambda z, t, y: M.dice_only_loss(z, y), lambda z, t, y: OM.dice_only_loss(z, y)),
    ]
    for ours, oracle in pairs:
        z = z0.clone().to(cuda_dev).requires_grad_(True)
        l = ours(z, t0.to(cuda_dev), y0.to(cuda_dev))
        (l * 3.0).backward()  # exercises the upstream-gradient scaling
        zr = z0.clone().double().requires_grad_(True)
        lr = oracle(zr, t0.double(), y0)
        (lr * 3.0).backward()
        assert abs(l.item() - lr.item()) <= 1e-5 * max(1.0, abs(lr.item()))
        assert rel_l2(z.grad, zr.grad) <= 1e-5


def test_loss_error_behaviour(cuda_dev):
    z = torch.randn(1, 4, 4, 4, 4, device=cuda_dev)
    with pytest.raises(RuntimeError):  # float targets raise in the reference's CrossEntropyLoss path too
        M.combined_loss(z, torch.zeros(1, 1, 4, 4, 4, device=cuda_dev))
    with pytest.raises(RuntimeError):  # no CPU fallback
        M.combined_loss(z.cpu(), torch.zeros(1, 1, 4, 4, 4, dtype=torch.long))
    with pytest.raises(ValueError):
        M.combined_loss(z, torch.zeros(1, 1, 4, 4, 5, dtype=torch.long, device=cuda_dev))


# --------------------------------------------------------------------------------- metrics
METRIC_CASES = ["normal", "absent_class", "first_spatial_2", "first_spatial_3", "no_foreground", "ties", "nan", "structured_like"]


@pytest.mark.parametrize("case", METRIC_CASES)
def test_metrics_bit_exact_vs_reference_golden(cuda_dev, golden_dir, case):
    g = _load(golden_dir, "metrics.npz")
    pred = torch.from_numpy(g[f"{case}/pred"]).to(cuda_dev)
    tgt = torch.from_numpy(g[f"{case}/target"]).to(cuda_dev)
    conf = F.confusion_counts(pred, tgt).cpu().numpy()
    assert np.array_equal(conf, OM.confusion_counts(pred.cpu(), tgt.cpu()))  # counts: bit-exact
    d, i, a = M.calculate_dice(pred, tgt), M.calculate_iou(pred, tgt), M.calculate_accuracy(pred, tgt)
    assert np.float32(float(d)) == g[f"{case}/dice"] and np.float32(float(i)) == g[f"{case}/iou"]  # bit-exact
    assert np.float32(float(a)) == g[f"{case}/acc"]
    assert torch.is_tensor(d) == bool(g[f"{case}/dice_is_tensor"])
    mask = F.argmax_mask(pred).cpu()
    assert torch.equal(mask.long(), torch.argmax(pred.cpu(), dim=1))  # argmax masks: bit-exact


def test_metrics_counts_above_2_24(cuda_dev, golden_dir):
    g = _load(golden_dir, "metrics.npz")
    gen = torch.Generator().manual_seed(33)
    pred = torch.randn(1, 4, 150, 400, 300, generator=gen)
    pred[:, 1] += 2.5
    tgt = (torch.rand(1, 1, 150, 400, 300, generator=gen) < 0.97).long()
    pred, tgt = pred.to(cuda_dev), tgt.to(cuda_dev)
    d, i, a = M.calculate_dice(pred, tgt), M.calculate_iou(pred, tgt), M.calculate_accuracy(pred, tgt)
    assert np.float32(float(d)) == g["big_counts/dice"] and np.float32(float(i)) == g["big_counts/iou"]
    assert abs(float(a) - float(g["big_counts/acc"])) <= 2e-7


def test_confusion_ragged_sizes(cuda_dev):
    gen = torch.Generator().manual_seed(8)
    for shape in [(1, 4, 3, 5, 7), (3, 2, 1, 1, 1), (2, 7, 4, 4, 5)]:
        pred = torch.randn(shape, generator=gen)
        tgt = torch.randint(0, shape[1], (shape[0], 1, *shape[2:]), generator=gen)
        conf = F.confusion_counts(pred.to(cuda_dev), tgt.to(cuda_dev)).cpu().numpy()
        assert np.array_equal(conf, OM.confusion_counts(pred, tgt)), shape
        assert conf.sum() == tgt.numel()


# --------------------------------------------------------------------------------- convolution
CONV_CASES = [
    # N, D, H, W, c0, c1, Cout
    (2, 6, 5, 7, 1, 0, 16),
    (1, 9, 17, 41, 1, 0, 16),  # stem: several w / h / d tiles with ragged edges
    (1, 8, 8, 8, 16, 0, 16),
    (2, 4, 6, 10, 16, 0, 32),
    (1, 5, 4, 6, 32, 32, 32),
    (1, 4, 4, 4, 64, 0, 128),
    (2, 3, 3, 3, 3, 0, 8),
    (1, 4, 5, 3, 16, 16, 16),
]


def _conv_inputs(case, seed=0):
    N, D, H, W, c0, c1, Cout = case
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(N, c0 + c1, D, H, W, generator=gen)
    w = torch.randn(Cout, c0 + c1, 3, 3, 3, generator=gen) / (27 * (c0 + c1)) ** 0.5
    b = torch.randn(Cout, generator=gen)
    gy = torch.randn(N, Cout, D, H, W, generator=gen)
    return x, w, b, gy


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv3d_fprop_dgrad_wgrad_direct(cuda_dev, case, dtype):
    N, D, H, W, c0, c1, Cout = case
    x, w, b, gy = _conv_inputs(case)
    if dtype == torch.bfloat16:  # feed bf16-exact values so that only accumulation order differs
        x, w, gy = x.bfloat16().float(), w.bfloat16().float(), gy.bfloat16().float()
    xr = x.clone().double().requires_grad_(True)
    wr = w.clone().double().requires_grad_(True)
    br = b.clone().double().requires_grad_(True)
    yr = TF.conv3d(xr, wr, br, padding=1)
    yr.backward(gy.double())
    tol = 1e-5 if dtype == torch.float32 else 4e-3  # bf16: output rounding 2^-9
    x0 = cl(x[:, :c0], dtype)
    x1 = cl(x[:, c0:], dtype) if c1 else None
    wd = w.to(cuda_dev)
    wp = F.pack_conv3_weights(wd, _lib.PACK_FPROP, dtype)
    y, _ = F.conv3d_k3_raw(x0, x1, wp, b.to(cuda_dev), Cout, 0, impl=1)
    assert rel_l2(cf(y), yr) <= tol
    # data gradient, split over the two concat halves
    wpd = F.pack_conv3_weights(wd, _lib.PACK_DGRAD, dtype)
    dy = cl(gy, dtype)
    dx0, dx1 = F.conv3d_k3_raw(dy, None, wpd, None, c0, c1, impl=1)
    assert rel_l2(cf(dx0), xr.grad[:, :c0]) <= tol
    if c1:
        assert rel_l2(cf(dx1), xr.grad[:, c0:]) <= tol
    # weight / bias gradients (fp32 outputs in torch layout)
    dw, db = F.conv3d_wgrad_raw(x0, x1, dy, want_bias=(Cout % 8 == 0))
    assert rel_l2(dw, wr.grad) <= 2e-5
    if db is not None:
        assert rel_l2(db, br.grad) <= 2e-5


# --------------------------------------------------------------------------------- BN / pool / convT / heads
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("training", [True, False])
def test_conv_bn_relu_dropout_block(cuda_dev, dtype, training):
    torch.manual_seed(0)
    N, C0, C, D, H, W = 2, 16, 32, 6, 5, 8
    conv = torch.nn.Conv3d(C0, C, 3, padding=1)
    bn = torch.nn.BatchNorm3d(C)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
        bn.running_mean.uniform_(-0.2, 0.2)
        bn.running_var.uniform_(0.5, 1.5)
    bn.train(training)
    x = torch.randn(N, C0, D, H, W)
    mask = torch.empty(N, C).bernoulli_(0.7) / 0.7
    gy = torch.randn(N, C, D, H, W)
    # oracle (fp64 math on CPU)
    import copy
    conv_r, bn_r = copy.deepcopy(conv).double(), copy.deepcopy(bn).double()
    xr = x.double().requires_grad_(True)
    yr = TF.relu(bn_r(conv_r(xr))) * mask.double()[:, :, None, None, None]
    yr.backward(gy.double())
    # ours
    conv_c, bn_c = copy.deepcopy(conv).to(cuda_dev), copy.deepcopy(bn).to(cuda_dev)
    xc = cl(x, dtype).requires_grad_(True)
    y = F.conv_bn_act(xc, None, conv_c, bn_c, mask.to(cuda_dev), training, impl=1)
    y.backward(cl(gy, dtype))
    tol = 1e-4 if dtype == torch.float32 else 1.5e-2
    # bf16 gradients on unstructured random data: a bf16-rounded pre-activation flips the ReLU mask of
    # ~0.3 % of the elements against the fp64 oracle, which alone is sqrt(0.003) ~ 5e-2 rel-L2
    # (SURVEY App. F measures 1e-2..9e-2 for the reference against itself) -> 1e-1 here; the 1e-2
    # bf16 bar is checked on the structured problem in test_unet_bf16_vs_cuda_autocast_oracle.
    gtol = tol if dtype == torch.float32 else 1e-1
    assert rel_l2(cf(y), yr) <= tol
    assert rel_l2(cf(xc.grad), xr.grad) <= gtol
    assert rel_l2(conv_c.weight.grad, conv_r.weight.grad) <= gtol
    assert rel_l2(bn_c.weight.grad, bn_r.weight.grad) <= gtol
    assert rel_l2(bn_c.bias.grad, bn_r.bias.grad) <= gtol
    if training:
        # pre-BN conv bias: analytically zero gradient -> absolute tolerance (SURVEY hard part 5);
        # in bf16 it is the rounding noise of sum(dconv): ~2^-9 * |dconv| * sqrt(M)
        assert conv_c.bias.grad.abs().max().item() <= (1e-4 if dtype == torch.float32 else 1.0)
        assert rel_l2(bn_c.running_mean, bn_r.running_mean) <= (1e-5 if dtype == torch.float32 else 5e-3)
        assert rel_l2(bn_c.running_var, bn_r.running_var) <= (1e-5 if dtype == torch.float32 else 5e-3)
        assert int(bn_c.num_batches_tracked) == 1
    else:
        assert rel_l2(conv_c.bias.grad, conv_r.bias.grad) <= gtol
        assert int(bn_c.num_batches_tracked) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 16, 4, 6, 8), (1, 8, 5, 7, 9)])
def test_maxpool_fwd_bwd(cuda_dev, dtype, shape):
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=gen)
    x = torch.relu(x)  # plenty of exact ties at 0, as after ReLU
    x = x.to(dtype).float()
    xr = x.clone().requires_grad_(True)
    yr = TF.max_pool3d(xr, 2, 2)
    gy = torch.randn(yr.shape, generator=gen).to(dtype).float()
    yr.backward(gy)
    xc = cl(x, dtype).requires_grad_(True)
    y = F.maxpool2(xc)
    y.backward(cl(gy, dtype))
    assert torch.equal(cf(y), yr.detach())          # max is exact
    assert torch.equal(cf(xc.grad), xr.grad)        # gradient routing to the first maximum: exact
    # skip + pool with the fused backward: d(skip)/dx + d(pool)/dx, rounded once like torch's accumulation of two grads
    xs = cl(x, dtype).requires_grad_(True)
    skip, yp = F.skip_and_pool(xs)
    gs = torch.randn(x.shape, generator=gen).to(dtype).float()
    torch.autograd.backward([skip, yp], [cl(gs, dtype), cl(gy, dtype)])
    want = (gs.to(dtype).float() + xr.grad).to(dtype).float()
    assert torch.equal(cf(yp), yr.detach()) and torch.equal(cf(skip), x)
    assert torch.equal(cf(xs.grad), want)
    xs2 = cl(x, dtype).requires_grad_(True)
    skip2, yp2 = F.skip_and_pool(xs2)
    yp2.backward(cl(gy, dtype))                      # only the pool branch used
    assert torch.equal(cf(xs2.grad), xr.grad)


@pytest.mark.parametrize("case", [(2, 16, 16, 8, 12, 16, True, torch.bfloat16), (1, 16, 32, 6, 8, 10, True, torch.bfloat16),
                                  (2, 32, 64, 4, 4, 6, True, torch.float32), (1, 16, 16, 8, 8, 8, False, torch.bfloat16),
                                  (1, 64, 128, 2, 4, 4, True, torch.bfloat16), (1, 8, 24, 4, 6, 6, True, torch.float32)])
def test_encoder_handoff_fused_forward_is_bit_identical(cuda_dev, case):
    """DoubleConv -> (skip, MaxPool3d) — models/unet.py:69-71: BatchNorm apply + pool in one pass (bn_act_pool_fwd_kernel) against
    bn_act_fwd followed by maxpool2_fwd: same element arithmetic, so outputs and every gradient are bit-identical."""
    N, Cin, Cout, D, H, W, use_skip, dtype = case
    torch.manual_seed(3)
    blk = DoubleConv(Cin, Cout, dropout_rate=0.0).to(cuda_dev).train()
    x = torch.randn(N, D, H, W, Cin, device=cuda_dev).to(dtype)
    g_skip = torch.randn(N, D, H, W, Cout, device=cuda_dev).to(dtype)
    g_pool = torch.randn(N, D // 2, H // 2, W // 2, Cout, device=cuda_dev).to(dtype)
    res = []
    try:
        for fused in (True, False):
            F.set_fuse_pool_fwd(fused)
            blk.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_(True)
            n0 = _lib.launch_count()
            skip, pooled = blk.forward_cl(xi, None, 0, pool=True)
            # 24 channels = three 8-channel groups do not divide the 256-thread block: that case keeps the two-pass sequence
            assert type(skip.grad_fn).__name__.startswith("_ConvBNActSkipPool") == (fused and 256 % (Cout // 8) == 0)
            if use_skip:
                torch.autograd.backward([skip, pooled], [g_skip, g_pool])
            else:
                pooled.backward(g_pool)
            torch.cuda.synchronize()
            assert _lib.launch_count() > n0
            res.append((skip.detach().clone(), pooled.detach().clone(), xi.grad.clone(), {k: v.grad.clone() for k, v in blk.named_parameters() if v.grad is not None}))
    finally:
        F.set_fuse_pool_fwd(True)
    (s1, p1, gx1, gp1), (s2, p2, gx2, gp2) = res
    assert torch.equal(s1, s2) and torch.equal(p1, p2) and torch.equal(gx1, gx2)
    assert gp1.keys() == gp2.keys() and all(torch.equal(gp1[k], gp2[k]) for k in gp1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 32, 16, 3, 4, 5), (1, 256, 128, 2, 2, 2), (1, 64, 32, 5, 17, 20), (2, 128, 64, 3, 16, 9)])
def test_conv_transpose_fwd_bwd(cuda_dev, dtype, shape):
    gen = torch.Generator().manual_seed(2)
    N, Cin, Cout, D, H, W = shape
    x = torch.randn(N, Cin, D, H, W, generator=gen).to(dtype).float()
    w = (torch.randn(Cin, Cout, 2, 2, 2, generator=gen) / Cin ** 0.5)
    b = torch.randn(Cout, generator=gen)
    gy = torch.randn(N, Cout, 2 * D, 2 * H, 2 * W, generator=gen).to(dtype).float()
    wq = w.to(dtype).float()
    xr, wr, br = x.double().requires_grad_(True), wq.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = TF.conv_transpose3d(xr, wr, br, stride=2)
    yr.backward(gy.double())
    xc = cl(x, dtype).requires_grad_(True)
    wc, bc = w.to(cuda_dev).requires_grad_(True), b.to(cuda_dev).requires_grad_(True)
    y = F.conv_transpose2(xc, wc, bc)
    y.backward(cl(gy, dtype))
    tol = 1e-5 if dtype == torch.float32 else 4e-3
    assert rel_l2(cf(y), yr) <= tol
    assert rel_l2(cf(xc.grad), xr.grad) <= tol
    assert rel_l2(wc.grad, wr.grad) <= 2e-5 and rel_l2(bc.grad, br.grad) <= 2e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_final_conv_and_gap_and_resize(cuda_dev, dtype):
    gen = torch.Generator().manual_seed(3)
    N, Cin, Cout, D, H, W = 2, 16, 4, 5, 6, 7
    x = torch.randn(N, Cin, D, H, W, generator=gen).to(dtype).float()
    w = torch.randn(Cout, Cin, 1, 1, 1, generator=gen) / 4
    b = torch.randn(Cout, generator=gen)
    gy = torch.randn(N, Cout, D, H, W, generator=gen)
    xr, wr, br = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = TF.conv3d(xr, wr, br)
    yr.backward(gy.double())
    xc = cl(x, dtype).requires_grad_(True)
    wc, bc = w.to(cuda_dev).requires_grad_(True), b.to(cuda_dev).requires_grad_(True)
    y = F.final_conv1x1(xc, wc, bc, round_bf16=False)
    y.backward(gy.to(cuda_dev))
    assert y.dtype == torch.float32 and y.shape == yr.shape
    assert rel_l2(y, yr) <= 1e-5
    assert rel_l2(cf(xc.grad), xr.grad) <= (1e-5 if dtype == torch.float32 else 4e-3)
    assert rel_l2(wc.grad, wr.grad) <= 1e-5 and rel_l2(bc.grad, br.grad) <= 1e-5
    # GAP
    xg = cl(x, dtype).requires_grad_(True)
    gap = F.global_avg_pool(xg)
    gg = torch.randn(N, Cin, generator=gen)
    gap.backward(gg.to(cuda_dev))
    assert rel_l2(gap, x.mean(dim=[2, 3, 4])) <= 1e-5
    expect = (gg / (D * H * W))[:, :, None, None, None].expand_as(x)
    assert rel_l2(cf(xg.grad), expect) <= (1e-6 if dtype == torch.float32 else 4e-3)
    # nearest resize (F.interpolate default) and its adjoint
    xs = cl(x, dtype).requires_grad_(True)
    size = (D + 1, H + 1, W)
    up = F.nearest_resize(xs, size)
    xr2 = x.clone().requires_grad_(True)
    upr = TF.interpolate(xr2, size=size)
    assert torch.equal(cf(up), upr.detach())
    g2 = torch.randn(upr.shape, generator=gen).to(dtype).float()
    upr.backward(g2)
    up.backward(cl(g2, dtype))
    assert rel_l2(cf(xs.grad), xr2.grad) <= (1e-6 if dtype == torch.float32 else 8e-3)


# --------------------------------------------------------------------------------- whole network
def _net_from(sd, cls=UNet3D, **kw):
    net = cls(**kw).cuda()
    net.load_state_dict(sd, strict=True)
    return net


def test_unet_fp32_train_step_vs_reference_golden(cuda_dev, golden_dir):
    g = _load(golden_dir, "unet_train_b2_s16.npz")
    sd = init_state_dict(1, 4, seed=0)
    net = _net_from(sd, in_channels=1, out_channels=4, dropout_rate=0.0)
    assert list(net.state_dict().keys()) == list(sd.keys())
    net.train()
    x, y = structured_volume(2, 16, seed=1234)
    logits = net(x.cuda())
    loss = M.combined_loss(logits, y.cuda())
    loss.backward()
    assert logits.dtype == torch.float32 and tuple(logits.shape) == (2, 4, 16, 16, 16)
    assert rel_l2(logits, torch.from_numpy(g["logits"])) <= 1e-4            # fp32 tolerance (north_star)
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * float(g["loss"])
    ref_loss, ref_logits, ref_grads, ref_bufs = train_step_grads(sd, x, y, OM.combined_loss)
    named = dict(net.named_parameters())
    flat = torch.cat([named[k].grad.flatten().cpu() for k in ref_grads])
    rflat = torch.cat([ref_grads[k].flatten() for k in ref_grads])
    assert rel_l2(flat, rflat) <= 1e-4                                      # global gradient rel-L2
    for k, rg in ref_grads.items():
        if k.endswith("double_conv.0.bias") or k.endswith("double_conv.4.bias"):
            assert named[k].grad.abs().max().item() <= 1e-5, k             # pre-BN biases: absolute
        else:
            assert rel_l2(named[k].grad, rg) <= 2e-3, k
    st = net.state_dict()
    for k, b in ref_bufs.items():
        if "num_batches" in k:
            assert int(st[k]) == int(b)
        else:
            assert rel_l2(st[k], b) <= 1e-4, k
    # eval forward with the updated running statistics vs the reference's eval golden
    ge = _load(golden_dir, "unet_eval_s16x32x16.npz")
    net.eval()
    x2, _ = structured_volume(1, (16, 32, 16), seed=77)
    with torch.no_grad():
        out = net(x2.cuda())
    assert rel_l2(out, torch.from_numpy(ge["logits"])) <= 1e-4


def test_unet_odd_sizes_and_small_net_vs_reference_golden(cuda_dev, golden_dir):
    g = _load(golden_dir, "unet_train_odd_20x18x22.npz")
    sd = init_state_dict(1, 4, seed=0)
    net = _net_from(sd, in_channels=1, out_channels=4, dropout_rate=0.0).train()
    x, y = structured_volume(2, (20, 18, 22), seed=5)
    logits = net(x.cuda())
    loss = M.combined_ce_tversky_loss(logits, y.cuda(), alpha=0.5, beta=0.5)
    loss.backward()
    assert rel_l2(logits, torch.from_numpy(g["logits"])) <= 1e-4
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * float(g["loss"])
    for k, p in net.named_parameters():
        ref = float(g["gradnorm/" + k])
        if not (k.endswith("double_conv.0.bias") or k.endswith("double_conv.4.bias")):
            assert abs(p.grad.double().norm().item() - ref) <= 2e-3 * ref + 1e-7, k
    gs = _load(golden_dir, "unet_small_in2_out3_f8_16.npz")
    sds = init_state_dict(2, 3, features=(8, 16), seed=3)
    snet = _net_from(sds, in_channels=2, out_channels=3, features=[8, 16], dropout_rate=0.0).train()
    gen = torch.Generator().manual_seed(11)
    xs = torch.randn(2, 2, 8, 12, 8, generator=gen)
    ys = torch.randint(0, 3, (2, 1, 8, 12, 8), generator=gen)
    lg = snet(xs.cuda())
    ls = M.combined_loss(lg, ys.cuda())
    ls.backward()
    assert rel_l2(lg, torch.from_numpy(gs["logits"])) <= 1e-4
    for k, p in snet.named_parameters():
        if not (k.endswith("double_conv.0.bias") or k.endswith("double_conv.4.bias")):
            assert rel_l2(p.grad, torch.from_numpy(gs["grad/" + k])) <= 2e-3, k


def test_unet_dann_variant_vs_reference_golden(cuda_dev, golden_dir):
    g = _load(golden_dir, "unet_dann_b2_s16.npz")
    sd = init_state_dict(1, 4, seed=0)
    net = _net_from(sd, cls=UNet3DDann, in_channels=1, out_channels=4, dropout_rate=0.0).train()
    x, _ = structured_volume(2, 16, seed=1234)
    logits, gap = net(x.cuda(), return_features=True)
    logits2, none = net(x.cuda())
    assert none is None and tuple(gap.shape) == (2, 256)
    assert rel_l2(logits, torch.from_numpy(g["logits"])) <= 1e-4
    assert rel_l2(gap, torch.from_numpy(g["gap"])) <= 1e-4
    assert rel_l2(logits2, torch.from_numpy(g["logits_second_call"])) <= 1e-4


def test_unet_bf16_vs_cuda_autocast_oracle(cuda_dev):
    """bf16 tolerance of north_star (rel. L2 <= 1e-2) against the oracle run on the GPU under bf16 autocast (same rounding points),
    structured problem, global norms — each quantity asserted by name — plus App. F's sanity bound: no further from the fp32
    truth than 1.25x the reference's own bf16 path.  Measured on B200 (profiles/r02_parity_measured.jsonl): logits 4.9e-3,
    gradients 6.7e-3 against the bf16 oracle; the oracle's own bf16-vs-fp32 distance is 4.7e-3 / 7.9e-3."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = init_state_dict(1, 4, seed=0)
    x, y = structured_volume(2, 32, seed=1234)
    # warm the weights up with a few fp32 AdamW steps of the oracle on the GPU (App. F: trained weights behave better)
    p = {k: v.clone().cuda() for k, v in sd.items()}
    names = [k for k in p if not ("running" in k or "num_batches" in k)]
    for k in names:
        p[k].requires_grad_(True)
    opt = torch.optim.AdamW([p[k] for k in names], lr=1e-3)
    xc, yc = x.cuda(), y.cuda()
    for _ in range(20):
        opt.zero_grad()
        OM.combined_loss(unet3d_forward(p, xc, True), yc).backward()
        opt.step()
    sdw = {k: v.detach().cpu() for k, v in p.items()}
    l32, z32, g32, _ = train_step_grads(sdw, xc, yc, OM.combined_loss)
    l16, z16, g16, _ = train_step_grads(sdw, xc, yc, OM.combined_loss, autocast_dtype=torch.bfloat16)
    net = _net_from(sdw, in_channels=1, out_channels=4, dropout_rate=0.0).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = net(xc)
    loss = M.combined_loss(logits.float(), yc)
    loss.backward()
    named = dict(net.named_parameters())
    keys = [k for k in g32 if not (k.endswith("double_conv.0.bias") or k.endswith("double_conv.4.bias"))]
    ours = torch.cat([named[k].grad.flatten() for k in keys])
    r16 = torch.cat([g16[k].flatten() for k in keys])
    r32 = torch.cat([g32[k].flatten() for k in keys])
    oracle_self = rel_l2(z16, z32)
    print(f"logits: ours-vs-bf16oracle {rel_l2(logits, z16):.3e} ours-vs-fp32 {rel_l2(logits, z32):.3e} oracle16-vs-32 {oracle_self:.3e}")
    print(f"grads : ours-vs-bf16oracle {rel_l2(ours, r16):.3e} ours-vs-fp32 {rel_l2(ours, r32):.3e} oracle16-vs-32 {rel_l2(r16, r32):.3e}")
    logits_vs_bf16, grads_vs_bf16 = rel_l2(logits, z16), rel_l2(ours, r16)
    logits_vs_fp32, grads_vs_fp32 = rel_l2(logits, z32), rel_l2(ours, r32)
    assert logits_vs_bf16 <= 1e-2, logits_vs_bf16
    assert grads_vs_bf16 <= 1e-2, grads_vs_bf16
    assert logits_vs_fp32 <= 1.25 * oracle_self + 1e-3, (logits_vs_fp32, oracle_self)
    assert grads_vs_fp32 <= 1.25 * rel_l2(r16, r32) + 1e-3, (grads_vs_fp32, rel_l2(r16, r32))
    assert abs(loss.item() - l16.item()) <= 1e-2 * abs(l16.item())
    # argmax agreement with the bf16 oracle
    agree = (logits.argmax(1) == z16.argmax(1)).float().mean().item()
    assert agree >= 0.99


def test_dropout_masks_and_eval_mode(cuda_dev):
    sd = init_state_dict(1, 4, seed=0)
    net = _net_from(sd, in_channels=1, out_channels=4, dropout_rate=0.5).train()
    x, _ = structured_volume(2, 16, seed=3)
    torch.manual_seed(1)
    a = net(x.cuda())
    torch.manual_seed(1)
    b = net(x.cuda())
    torch.manual_seed(2)
    c = net(x.cuda())
    assert torch.equal(a, b) and not torch.equal(a, c)  # masks come from torch's RNG stream
    net.eval()
    with torch.no_grad():
        e1, e2 = net(x.cuda()), net(x.cuda())
    assert torch.equal(e1, e2)


def test_state_dict_round_trip_and_errors(cuda_dev):
    sd = init_state_dict(1, 4, seed=0)
    net = _net_from(sd, in_channels=1, out_channels=4)
    out = net.state_dict()
    assert list(out.keys()) == list(sd.keys())
    for k in sd:
        assert out[k].dtype == sd[k].dtype and tuple(out[k].shape) == tuple(sd[k].shape)
        assert torch.equal(out[k].cpu(), sd[k])
    # DDP-prefixed checkpoints (reference test_model.py:384 strips 'module.')
    pref = {"module." + k: v for k, v in sd.items()}
    net.load_state_dict({k.replace("module.", ""): v for k, v in pref.items()})
    with pytest.raises(RuntimeError):
        UNet3D(1, 4)(torch.zeros(1, 1, 16, 16, 16))  # CPU tensor: no fallback
    # frozen encoder (reference train_unet.py:31-36): encoder grads stay None
    for p_ in net.encoder.parameters():
        p_.requires_grad = False
    net.train()
    x, y = structured_volume(2, 16, seed=1)
    M.combined_loss(net(x.cuda()), y.cuda()).backward()
    assert all(p_.grad is None for p_ in net.encoder.parameters())
    assert all(p_.grad is not None for p_ in net.decoder.parameters())


def test_double_conv_module_ncdhw_contract(cuda_dev):
    torch.manual_seed(0)
    dc = DoubleConv(8, 16, dropout_rate=0.0).cuda().train()
    x = torch.randn(2, 8, 4, 6, 8)
    import copy
    ref = copy.deepcopy(dc).cpu().double()
    y = dc(x.cuda())
    yr = ref.double_conv(x.double())
    assert tuple(y.shape) == tuple(yr.shape) and rel_l2(y, yr) <= 1e-4


# --------------------------------------------------------------------------------- DANN head
def test_dann_head_vs_reference_golden(cuda_dev, golden_dir):
    g = _load(golden_dir, "dann_head.npz")
    sd = OD.init_discriminator(256, seed=0)
    disc = DomainDiscriminator(256).cuda()
    disc.load_state_dict(sd)
    assert list(disc.state_dict().keys()) == list(sd.keys())
    disc.eval()
    fs = torch.from_numpy(g["fs"]).cuda().requires_grad_(True)
    ft = torch.from_numpy(g["ft"]).cuda().requires_grad_(True)
    lam = float(g["lambda"])
    so = disc(grad_reverse(fs, lam))
    to = disc(grad_reverse(ft, lam))
    labels = torch.cat([torch.zeros(3, dtype=torch.long), torch.ones(3, dtype=torch.long)]).cuda()
    dl = domain_cross_entropy(torch.cat([so, to]), labels)
    (lam * dl).backward()
    assert rel_l2(so, torch.from_numpy(g["source_out"])) <= 1e-5
    assert abs(dl.item() - float(g["domain_loss"])) <= 1e-5
    assert rel_l2(fs.grad, torch.from_numpy(g["grad_fs"])) <= 1e-4  # encoder side sees -lambda^2 (App. C-7)
    assert rel_l2(ft.grad, torch.from_numpy(g["grad_ft"])) <= 1e-4
    for k, p in disc.named_parameters():
        assert rel_l2(p.grad, torch.from_numpy(g["grad/" + k])) <= 1e-4, k


def test_dann_head_dropout_train_mode(cuda_dev):
    disc = DomainDiscriminator(256).cuda().train()
    x = torch.randn(4, 256, device=cuda_dev, requires_grad=True)
    torch.manual_seed(0)
    out = disc(x)
    out.sum().backward()
    assert out.shape == (4, 2) and torch.isfinite(x.grad).all()


# --------------------------------------------------------------------------------- optimiser
def test_fused_adamw_matches_torch(cuda_dev):
    torch.manual_seed(0)
    n = 10007
    p0 = torch.randn(n)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-2)
    p = p0.clone().cuda()
    g = torch.zeros(n, device=cuda_dev)
    ours = F.FlatAdamW(p, g, lr=1e-3, weight_decay=1e-2)
    for step in range(5):
        gr = torch.randn(n)
        ref.grad = gr.clone()
        opt.step()
        g.copy_(gr * 2.0)
        ours.step(grad_scale=0.5)
    assert rel_l2(p, ref) <= 1e-6


# --------------------------------------------------------------------------------- tcgen05 convolution
TC_CASES = [
    # N, D, H, W, c0, c1, co0, co1
    (1, 8, 16, 16, 16, 0, 16, 0),
    (2, 5, 20, 24, 16, 0, 32, 0),
    (1, 4, 8, 8, 32, 32, 32, 0),
    (1, 9, 16, 32, 32, 0, 16, 16),     # data-gradient style split output
    (1, 4, 16, 16, 64, 0, 128, 0),
    (1, 3, 8, 8, 128, 0, 256, 0),
    (2, 3, 7, 10, 16, 16, 16, 0),
    (1, 17, 33, 18, 16, 0, 16, 0),
    (2, 44, 64, 64, 16, 0, 16, 0),     # 2*6*16 = 352 tiles -> persistent kernel, last d-block ragged (44 = 5*8 + 4)
    (1, 20, 96, 80, 16, 16, 32, 0),    # concat input, n_tile 32 (dseg 4), persistent
    (1, 21, 80, 96, 32, 0, 16, 16),    # split output (dgrad style), 2 slabs, persistent
]


@pytest.fixture(params=[0, 2], ids=["two-cta", "persistent"])
def conv_kernel_mode(request):
    F.set_conv_persistent(request.param)
    yield request.param
    F.set_conv_persistent(1)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv3d_tcgen05_vs_oracle(cuda_dev, case, conv_kernel_mode):
    N, D, H, W, c0, c1, co0, co1 = case
    Cin, Cout = c0 + c1, co0 + co1
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(N, Cin, D, H, W, generator=gen).bfloat16().float()
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=gen) / (27 * Cin) ** 0.5).bfloat16().float()
    b = torch.randn(Cout, generator=gen)
    yr = TF.conv3d(x.double(), w.double(), b.double(), padding=1)
    x0 = cl(x[:, :c0], torch.bfloat16)
    x1 = cl(x[:, c0:], torch.bfloat16) if c1 else None
    wp = F.pack_conv3_weights(w.to(cuda_dev), _lib.PACK_FPROP_TC, torch.bfloat16)
    y0, y1 = F.conv3d_k3_raw(x0, x1, wp, b.to(cuda_dev), co0, co1, impl=2)
    torch.cuda.synchronize()
    y = cf(y0) if y1 is None else torch.cat([cf(y0), cf(y1)], dim=1)
    assert rel_l2(y, yr) <= 4e-3  # bf16 output rounding (2^-9), fp32 accumulation
    # identical inputs through the CUDA-core kernel: only accumulation order differs -> <= 1 bf16 ulp apart
    wd = F.pack_conv3_weights(w.to(cuda_dev), _lib.PACK_FPROP, torch.bfloat16)
    z0, z1 = F.conv3d_k3_raw(x0, x1, wd, b.to(cuda_dev), co0, co1, impl=1)
    z = cf(z0) if z1 is None else torch.cat([cf(z0), cf(z1)], dim=1)
    assert rel_l2(y, z) <= 3e-3
    # dgrad packing: conv of a Cout-channel tensor back to Cin channels with flipped taps
    gy = torch.randn(N, Cout, D, H, W, generator=gen).bfloat16().float()
    xr = x.double().requires_grad_(True)
    TF.conv3d(xr, w.double(), None, padding=1).backward(gy.double())
    if Cin <= 128 or Cin % 128 == 0:
        wpd = F.pack_conv3_weights(w.to(cuda_dev), _lib.PACK_DGRAD_TC, torch.bfloat16)
        dx0, dx1 = F.conv3d_k3_raw(cl(gy, torch.bfloat16), None, wpd, None, c0, c1, impl=2)
        dx = cf(dx0) if dx1 is None else torch.cat([cf(dx0), cf(dx1)], dim=1)
        assert rel_l2(dx, xr.grad) <= 4e-3


WG_CASES = [
    # N, D, H, W, c0, c1, Cout
    (1, 8, 16, 16, 16, 0, 16),
    (2, 5, 20, 24, 16, 0, 32),
    (1, 6, 12, 16, 32, 0, 16),      # Cin > Cout: X goes to the M side (mirrored taps)
    (1, 4, 8, 8, 32, 32, 32),       # concat input
    (1, 4, 16, 16, 64, 0, 128),
    (1, 3, 8, 8, 128, 0, 256),
    (1, 3, 8, 8, 256, 0, 128),
    (1, 19, 33, 18, 16, 0, 16),
    (2, 9, 20, 33, 32, 0, 32),      # wide-row kernel (32-channel rows): ragged tiles, several d-blocks
    (1, 5, 16, 16, 64, 0, 64),      # wide rows, 2 x 2 slab pairs
    (1, 4, 16, 12, 32, 32, 64),     # wide rows, concat input
]


@pytest.mark.parametrize("case", WG_CASES)
def test_conv3d_wgrad_tcgen05_vs_oracle(cuda_dev, case):
    N, D, H, W, c0, c1, Cout = case
    Cin = c0 + c1
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(N, Cin, D, H, W, generator=gen).bfloat16().float()
    gy = torch.randn(N, Cout, D, H, W, generator=gen).bfloat16().float()
    w = torch.zeros(Cout, Cin, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    TF.conv3d(x.double(), w, None, padding=1).backward(gy.double())
    x0 = cl(x[:, :c0], torch.bfloat16)
    x1 = cl(x[:, c0:], torch.bfloat16) if c1 else None
    dy = cl(gy, torch.bfloat16)
    try:
        F.set_wgrad_impl(2)
        dw, db = F.conv3d_wgrad_raw(x0, x1, dy, want_bias=True)
        torch.cuda.synchronize()
    finally:
        F.set_wgrad_impl(0)
    # bf16-exact inputs, fp32 accumulation in TMEM, fp64 fixed-order split-K reduction
    assert rel_l2(dw, w.grad) <= 2e-5
    assert rel_l2(db, gy.double().sum(dim=(0, 2, 3, 4))) <= 2e-5


WG_PAIR_CASES = [
    # N, D, H, W, c0, c1, forced d-run per CTA (0 = planned); Cout = 16
    (1, 8, 16, 32, 16, 0, 0),
    (1, 12, 16, 16, 16, 16, 12),    # concat input; one CTA streams 14 X planes (both operand rings wrap), half-empty tile
    (2, 11, 20, 40, 16, 16, 3),     # ragged tiles in h and w, two w tiles, 4 d-blocks (the last one short), two samples
    (1, 19, 33, 18, 16, 0, 8),
    (1, 6, 12, 64, 16, 0, 2),
    (1, 1, 16, 16, 16, 0, 0),       # a single plane
]


@pytest.mark.parametrize("case", WG_PAIR_CASES)
def test_conv3d_wgrad_plane_pairs_vs_oracle(cuda_dev, case):
    """wgrad_tc4.cu (operand rows = pairs of w-adjacent voxels, M = 128 x N = 64 instructions) against the fp64 oracle and against the
    M = 64 kernel it replaces (wgrad_tc2.cu): same products, fp32 accumulation in a different order."""
    N, D, H, W, c0, c1, dseg = case
    Cout = 16
    Cin = c0 + c1
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(N, Cin, D, H, W, generator=gen).bfloat16().float()
    gy = torch.randn(N, Cout, D, H, W, generator=gen).bfloat16().float()
    w = torch.zeros(Cout, Cin, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    TF.conv3d(x.double(), w, None, padding=1).backward(gy.double())
    x0 = cl(x[:, :c0], torch.bfloat16)
    x1 = cl(x[:, c0:], torch.bfloat16) if c1 else None
    dy = cl(gy, torch.bfloat16)
    try:
        F.set_wgrad_impl(2)
        F.set_wgrad_pair(True, dseg)
        n0 = _lib.launch_count()
        dw, _ = F.conv3d_wgrad_raw(x0, x1, dy, want_bias=False)
        dw_again, _ = F.conv3d_wgrad_raw(x0, x1, dy, want_bias=False)
        torch.cuda.synchronize()
        F.set_wgrad_pair(False)
        dw_v2, _ = F.conv3d_wgrad_raw(x0, x1, dy, want_bias=False)
        torch.cuda.synchronize()
    finally:
        F.set_wgrad_pair(True, 0)
        F.set_wgrad_impl(0)
    assert _lib.launch_count() > n0
    assert rel_l2(dw, w.grad) <= 2e-5
    assert torch.equal(dw, dw_again)            # fixed-order reduction: run-to-run bitwise identical
    assert rel_l2(dw, dw_v2.double()) <= 2e-6


# ----------------------------------------------------------------------------- binary helpers (a13)
BINARY_CASES = ["a", "b", "c", "empty_target", "full_target"]


@pytest.mark.parametrize("case", BINARY_CASES)
def test_binary_helpers_vs_reference_golden(cuda_dev, golden_dir, case):
    """utils/metrics.py:6-12 (dice_loss, incl. its gradient), :42-63 (dice_score / iou_score / accuracy_score),
    :131-135 (calculate_metrics) against the unmodified reference (oracle/make_golden_binary.py)."""
    g = _load(golden_dir, "binary_helpers.npz")
    prob, tgt = torch.from_numpy(g[f"{case}/prob"]).cuda(), torch.from_numpy(g[f"{case}/target"]).cuda()
    # integer counts -> fp32 ratios in the reference's order: bit-exact
    assert M.dice_score(prob, tgt) == float(g[f"{case}/dice_score"])
    assert M.iou_score(prob, tgt) == float(g[f"{case}/iou_score"])
    assert M.accuracy_score(prob, tgt) == float(g[f"{case}/accuracy_score"])
    assert list(M.calculate_metrics(prob, tgt)) == [float(v) for v in g[f"{case}/calculate_metrics"]]
    z = torch.from_numpy(g[f"{case}/logits"]).cuda().requires_grad_(True)
    loss = M.dice_loss(z, tgt)
    loss.backward()
    assert abs(loss.item() - float(g[f"{case}/dice_loss"])) <= 1e-5 * max(1.0, abs(float(g[f"{case}/dice_loss"])))
    ref = torch.from_numpy(g[f"{case}/dice_loss_grad"])
    assert z.grad is not None and z.grad.shape == z.shape
    assert (z.grad.cpu() - ref).abs().max().item() <= 1e-5 * max(float(ref.abs().max()), 1e-12) + 1e-9


def test_wgrad_kernel_failure_surfaces_as_error(cuda_dev):
    """A failing weight-gradient kernel selection must raise, never hand AdamW an uninitialised dw (round-1 advisor finding)."""
    x = torch.randn(1, 8, 16, 16, 32, device="cuda").to(torch.bfloat16)
    dy = torch.randn(1, 8, 16, 16, 32, device="cuda").to(torch.bfloat16)
    dw, _ = F.conv3d_wgrad_raw(x, None, dy, want_bias=False)      # the wide-row tcgen05 path serves this shape
    _lib.check(_lib.load().b200_debug_fail_next_wgrad(1), "debug_fail_next_wgrad")
    try:
        with pytest.raises(ValueError, match="injected failure"):
            F.conv3d_wgrad_raw(x, None, dy, want_bias=False)
    finally:
        _lib.load().b200_debug_fail_next_wgrad(0)
    dw2, _ = F.conv3d_wgrad_raw(x, None, dy, want_bias=False)
    assert torch.equal(dw, dw2)


def test_bn_buffers_must_be_fp32_and_momentum_none_rejected(cuda_dev):
    net = DoubleConv(16, 16, dropout_rate=0.0).cuda().to(torch.bfloat16)      # model.to(bfloat16): buffers become bf16
    with pytest.raises(TypeError, match="running_mean"):
        net(torch.randn(1, 16, 8, 8, 8, device="cuda", dtype=torch.bfloat16))
    net = DoubleConv(16, 16, dropout_rate=0.0).cuda()
    net.double_conv[1].momentum = None
    with pytest.raises(ValueError, match="momentum=None"):
        net(torch.randn(1, 16, 8, 8, 8, device="cuda"))


def test_bias_free_layers_backward(cuda_dev):
    """bias=None layers: autograd must receive None (not a tensor) for the missing bias."""
    x = torch.randn(1, 4, 4, 4, 16, device="cuda", requires_grad=True)
    w = torch.randn(4, 16, 1, 1, 1, device="cuda", requires_grad=True)
    F.final_conv1x1(x, w, None).sum().backward()
    assert w.grad is not None and x.grad is not None
    wt = torch.randn(16, 8, 2, 2, 2, device="cuda", requires_grad=True)
    x2 = torch.randn(1, 4, 4, 4, 16, device="cuda", requires_grad=True)
    F.conv_transpose2(x2, wt, None).sum().backward()
    assert wt.grad is not None and x2.grad is not None


@pytest.mark.parametrize("shape", [(2, 64, 64, 96, 16, 16), (2, 56, 60, 90, 32, 16), (1, 64, 72, 80, 16, 32), (2, 64, 64, 64, 32, 32),
                                   (2, 40, 128, 128, 16, 16), (1, 46, 100, 128, 16, 32), (2, 24, 100, 224, 32, 16)])  # the last three: row-streaming kernel
def test_conv_epilogue_bn_statistics_match_separate_pass(cuda_dev, shape):
    """Conv3d -> BatchNorm3d (models/unet.py:11-12): statistics emitted by the persistent conv kernel's epilogue equal the ones
    the stand-alone bn_stats pass computes on the stored bf16 tensor; ragged tile borders included."""
    N, D, H, W, Cin, Cout = shape
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.randn(N, D, H, W, Cin, device="cuda", generator=g) + 0.3).to(torch.bfloat16)
    conv = torch.nn.Conv3d(Cin, Cout, 3, padding=1).cuda()
    with torch.no_grad():
        conv.bias.add_(torch.linspace(-2, 3, Cout, device="cuda"))       # |mean| >> std on some channels
    outs = []
    F.conv_bn_act(x, None, conv, torch.nn.BatchNorm3d(Cout).cuda(), None, True)      # packs (and caches) the weights: not part of the count
    for fused in (True, False):
        F.set_fuse_bn_stats(fused)
        try:
            bn = torch.nn.BatchNorm3d(Cout).cuda()
            n0 = _lib.launch_count()
            y = F.conv_bn_act(x, None, conv, bn, None, True)
            launches = _lib.launch_count() - n0
        finally:
            F.set_fuse_bn_stats(True)
        outs.append((y, bn.running_mean.clone(), bn.running_var.clone(), launches))
    (yf, mf, vf, lf), (ys, ms, vs, ls) = outs
    assert lf == ls - 1, (lf, ls)                                       # one pass fewer
    assert rel_l2(mf, ms) <= 2e-6 and rel_l2(vf, vs) <= 2e-5, (rel_l2(mf, ms), rel_l2(vf, vs))
    assert rel_l2(yf, ys) <= 1e-3                                       # same bf16 values up to 1-ulp flips from ~1e-6 scale changes
    # and against torch on the same bf16 conv output (the reference's arithmetic)
    w = conv.weight.detach().to(torch.bfloat16)
    ref = torch.nn.functional.conv3d(x.permute(0, 4, 1, 2, 3).float(), w.float(), conv.bias.detach(), padding=1).to(torch.bfloat16).float()
    assert rel_l2(mf, 0.1 * ref.mean(dim=(0, 2, 3, 4))) <= 2e-3
    assert rel_l2(vf - 0.9, 0.1 * ref.var(dim=(0, 2, 3, 4), unbiased=True)) <= 5e-3


ROWSTREAM_CASES = [
    # N, D, H, W, c0, c1, co0, co1   — every case has >= 8 * 148 output rows and W >= 96, so the row-streaming kernel serves it
    (2, 20, 40, 128, 16, 0, 16, 0),     # ragged last d-block (20 = 2*8 + 4), CTA runs crossing (n, d-block) boundaries
    (1, 33, 48, 128, 16, 16, 16, 0),    # virtual concat (2 slabs), D = 4*8 + 1
    (1, 12, 110, 128, 32, 0, 16, 16),   # split output (data gradient of the decoder's first conv), n_tile 32 -> 4 planes per accumulator
    (2, 16, 40, 100, 16, 0, 16, 0),     # W < 128: masked lanes
    (1, 8, 80, 256, 16, 0, 32, 0),      # two w-tiles, Cout 32
    (1, 9, 70, 224, 16, 0, 16, 0),      # W = 128 + 96
]


@pytest.mark.parametrize("case", ROWSTREAM_CASES)
def test_conv3d_rowstream_vs_oracle(cuda_dev, case):
    """conv_tc4.cu (one 128-voxel row per M tile, collector reuse of A across kh) against fp64 and against the
    persistent window kernel it replaces on these shapes (same bf16 inputs, fp32 accumulation: <= 1 bf16 ulp apart)."""
    N, D, H, W, c0, c1, co0, co1 = case
    Cin, Cout = c0 + c1, co0 + co1
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(N, Cin, D, H, W, generator=gen).bfloat16().float()
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=gen) / (27 * Cin) ** 0.5).bfloat16().float()
    b = torch.randn(Cout, generator=gen)
    yr = TF.conv3d(x.double(), w.double(), b.double(), padding=1)
    x0 = cl(x[:, :c0], torch.bfloat16)
    x1 = cl(x[:, c0:], torch.bfloat16) if c1 else None
    wp = F.pack_conv3_weights(w.to(cuda_dev), _lib.PACK_FPROP_TC, torch.bfloat16)
    outs = []
    for rowstream in (True, False):
        F.set_conv_rowstream(rowstream)
        try:
            y0, y1 = F.conv3d_k3_raw(x0, x1, wp, b.to(cuda_dev), co0, co1, impl=2)
            torch.cuda.synchronize()
        finally:
            F.set_conv_rowstream(True)
        outs.append(cf(y0) if y1 is None else torch.cat([cf(y0), cf(y1)], dim=1))
    assert rel_l2(outs[0], yr) <= 4e-3, rel_l2(outs[0], yr)
    assert rel_l2(outs[0], outs[1]) <= 2e-3
    assert (outs[0] - outs[1]).abs().max() <= 2.0 ** -7 * yr.abs().max()      # never more than a couple of bf16 ulps anywhere
    # run-to-run determinism
    y0b, y1b = F.conv3d_k3_raw(x0, x1, wp, b.to(cuda_dev), co0, co1, impl=2)
    assert torch.equal(cf(y0b) if y1b is None else torch.cat([cf(y0b), cf(y1b)], dim=1), outs[0])


def test_unet_out_channels_1_constructor_default(cuda_dev):
    """The reference constructor's default is out_channels=1 (models/unet.py:35); every call site passes 4.  Forward and backward of
    the one-channel head against the oracle, fp32 and bf16 (the fused 2..4-class head does not apply: generic kernels)."""
    sd = init_state_dict(1, 1, seed=2)
    assert tuple(sd["final_conv.weight"].shape) == (1, 16, 1, 1, 1)
    x, _ = structured_volume(2, 16, seed=3)
    tgt = torch.randn(2, 1, 16, 16, 16, generator=torch.Generator().manual_seed(5))
    p = {k: v.clone() for k, v in sd.items()}
    for k in p:
        if p[k].is_floating_point() and "running" not in k:
            p[k].requires_grad_(True)
    ref_logits = unet3d_forward(p, x, True)
    ((ref_logits - tgt) ** 2).mean().backward()
    net = _net_from(sd, in_channels=1, out_channels=1, dropout_rate=0.0).train()
    logits = net(x.cuda())
    assert tuple(logits.shape) == (2, 1, 16, 16, 16)
    ((logits - tgt.cuda()) ** 2).mean().backward()
    assert rel_l2(logits, ref_logits) <= 1e-4
    keys = [k for k, _ in net.named_parameters() if not (k.endswith("double_conv.0.bias") or k.endswith("double_conv.4.bias"))]
    named = dict(net.named_parameters())
    ours = torch.cat([named[k].grad.flatten().cpu() for k in keys])
    ref = torch.cat([p[k].grad.flatten() for k in keys])
    assert rel_l2(ours, ref) <= 5e-4
    net.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lb = net(x.cuda())
    assert rel_l2(lb, ref_logits) <= 3e-2 and lb.dtype == torch.float32
