"""CPU: the C-ABI library loads and exports every symbol include/b200unet.h declares, ctypes
signatures agree with the header, and the host-side mirror keeps the reference's API contract
(constructor, state_dict layout, error behaviour).  No kernel is launched here."""
import re

import numpy as np
import pytest
import torch

from multimodal_segmentation_project_b200 import _lib
from multimodal_segmentation_project_b200.models.unet import DoubleConv, UNet3D
from multimodal_segmentation_project_b200.models.unet_dann import UNet3D as UNet3DDann
from multimodal_segmentation_project_b200.train_dann import DomainDiscriminator, GradientReversal, grad_reverse  # noqa: F401
from multimodal_segmentation_project_b200.utils import metrics as M
from oracle import dann_oracle as OD
from oracle import metrics_oracle as OM
from oracle.unet_oracle import init_state_dict


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200unet.h but not exported"
    assert lib.b200_version() >= 100


def test_ctypes_signatures_match_header():
    lib = _lib.load()
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    seen = 0
    for m in re.finditer(r"\b(b200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        res, argtypes = lib._b200_signatures[name]
        assert len(argtypes) == n, name
        for a, t in zip(args.split(","), argtypes):
            a = a.strip()
            exp = "c_void_p" if "*" in a else ("c_float" if a.startswith("float") else ("c_long" if a.startswith("int64_t") else "c_int"))
            assert t.__name__ == exp, (name, a, t.__name__)
        seen += 1
    assert seen == len(_lib.declared_symbols())


def test_errors_without_gpu_are_loud():
    lib = _lib.load()
    # argument validation happens before any CUDA call: usable without a device
    rc = lib.b200_seg_loss_fwd(None, None, 1, 4, 8, None, None)
    assert rc == -1 and b"null" in lib.b200_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "seg_loss_fwd")
    rc = lib.b200_confusion(None, None, 1, 400, 8, None, None)
    assert rc < 0
    with pytest.raises(RuntimeError):  # CPU tensors are refused: no fallback
        UNet3D(1, 4)(torch.zeros(1, 1, 16, 16, 16))
    with pytest.raises(RuntimeError):
        M.combined_loss(torch.zeros(1, 4, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError):
        M.calculate_dice(torch.zeros(1, 4, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError):
        DomainDiscriminator(256)(torch.zeros(2, 256))


def test_state_dict_layout_matches_reference():
    torch.manual_seed(0)
    net = UNet3D(in_channels=1, out_channels=4)
    sd = init_state_dict(1, 4, seed=0)  # pinned to the reference constructor by test_oracle_golden
    ours = net.state_dict()
    assert list(ours.keys()) == list(sd.keys()) and len(ours) == 136
    for k in sd:
        assert ours[k].dtype == sd[k].dtype and tuple(ours[k].shape) == tuple(sd[k].shape), k
        assert torch.equal(ours[k], sd[k]), k  # same default init, same RNG consumption order
    assert sum(p.numel() for p in net.parameters()) == 5647908
    assert net.dropout_rate == 0.1 and net.output_activation is None
    assert hasattr(net, "encoder") and hasattr(net, "pool") and hasattr(net, "bottleneck")
    assert isinstance(net.encoder[0], DoubleConv)
    d = UNet3DDann(1, 4)
    assert list(d.state_dict().keys()) == list(sd.keys())
    # default constructor arguments (reference models/unet.py:34-37)
    dflt = UNet3D()
    assert dflt.final_conv.out_channels == 1 and dflt.encoder[0].double_conv[0].in_channels == 1
    disc = DomainDiscriminator(256, hidden_dim=77)
    assert list(disc.state_dict().keys()) == list(OD.init_discriminator(256).keys())
    assert sum(p.numel() for p in disc.parameters()) == 107074


def test_metric_recipe_host_side_matches_oracle():
    gen = torch.Generator().manual_seed(0)
    for shape in [(2, 4, 8, 8, 8), (1, 4, 2, 8, 8), (1, 4, 3, 5, 5), (1, 3, 6, 6, 6)]:
        pred = torch.randn(shape, generator=gen)
        tgt = torch.randint(0, shape[1], (shape[0], 1, *shape[2:]), generator=gen)
        conf = OM.confusion_counts(pred, tgt)
        d, i, v = M.dice_iou_from_confusion(conf, shape[2])
        od, oi, _ = OM.dice_iou_accuracy(pred, tgt)
        if v == 0:
            assert od == 0
        else:
            assert np.float32(d / np.float32(v)) == od and np.float32(i / np.float32(v)) == oi


def test_host_side_planners_without_gpu():
    """Pure host functions of the C ABI (no CUDA call): support predicates and workspace planners of the kernels the step picks."""
    lib = _lib.load()
    BF16, F32 = _lib.B200_BF16, _lib.B200_F32
    # BatchNorm apply + pool in one pass: even sizes and an 8-channel group count that divides the 256-thread block
    ok = lib.b200_bn_act_pool_fwd_supported
    assert ok(128, 128, 128, 16) == 1 and ok(8, 8, 8, 256) == 1 and ok(4, 6, 6, 8) == 1
    assert ok(20, 18, 22, 16) == 1 and ok(21, 18, 22, 16) == 0          # odd D: pooled windows would not cover the tensor
    assert ok(8, 8, 8, 24) == 0 and ok(8, 8, 8, 12) == 0                # 3 groups do not divide the block; C must be a multiple of 8
    assert ok(1, 8, 8, 16) == 0
    # weight-gradient workspace: the benchmark's top-level layers run the voxel-pair kernel on B200_WG4_SMS (96) CTAs, each writing
    # two partial slices [16 ci][27 taps][16 co] fp32 per 16-channel input slab
    slice_bytes = 16 * 27 * 16 * 4
    bn = lib.b200_bn_partials_bytes(16)
    for c0, c1, ctas in ((16, 0, 96), (16, 16, 96)):
        ws = lib.b200_conv3d_wgrad_workspace(c0, c1, 16, 2, 128, 128, 128)
        assert ws >= bn + ctas * 2 * slice_bytes and ws % 4 == 0
    # never smaller for a larger volume; positive for every layer shape of the default U-Net
    prev = 0
    for s in (16, 32, 64, 128):
        ws = lib.b200_conv3d_wgrad_workspace(16, 16, 16, 2, s, s, s)
        assert ws >= prev > -1
        prev = ws
    for c0, c1, co, s in ((1, 0, 16, 128), (16, 0, 32, 64), (64, 64, 64, 32), (128, 0, 256, 8), (256, 0, 256, 8), (128, 128, 128, 16)):
        assert lib.b200_conv3d_wgrad_workspace(c0, c1, co, 2, s, s, s) > 0
    # up-convolution weight gradient: dW partials + the bias-gradient partials the same kernel writes (16 floats per CTA and slab)
    for cin, cout, s in ((32, 16, 64), (64, 32, 32), (128, 64, 16), (256, 128, 8)):
        ws = lib.b200_convt2_wgrad_workspace(cin, cout, 2, s, s, s)
        assert ws >= lib.b200_bn_partials_bytes(cout) + 8 * cin * cout * 4
    # fused-statistics row counts: the first layer (Cin = 1, bf16) reports one row per CTA of its 32 x 8 x 4 tiles, fp32 reports none
    rows = lib.b200_conv3d_k3_bnstats_blocks(BF16, 2, 1, 0, 16, 0, 2, 128, 128, 128)
    assert rows == 2 * (128 // 32) * (128 // 8) * (128 // 4)
    assert lib.b200_conv3d_k3_bnstats_blocks(F32, 2, 1, 0, 16, 0, 2, 128, 128, 128) == 0
    assert lib.b200_conv3d_k3_bnstats_blocks(BF16, 2, 16, 0, 16, 0, 2, 128, 128, 128) == 148     # persistent row-streaming kernel: one row per SM
    # packed weight sizes: 27 taps x Cin x Cout bf16 for the tcgen05 layouts
    assert lib.b200_pack_conv3_bytes(_lib.PACK_FPROP_TC, BF16, 32, 16) == 27 * 16 * 32 * 2
    assert lib.b200_head_blocks(2, 128 ** 3) == 2 * 148
