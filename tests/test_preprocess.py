"""Device-side input pipeline against the reference's own CombinedDataset.__getitem__ outputs
(tests/golden/preprocess.npz) — CPU: the numpy oracle is pinned to them; GPU: the CUDA path through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as OP

GOLD = os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz")
CASES = ["amos_ct", "chaos_mri", "ts_ct", "btcv", "amos_mri"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_pipeline(name):
    g = np.load(GOLD)
    img = OP.preprocess(g[name + "/raw_image"], name)
    lab = OP.remap_labels(g[name + "/raw_label"], name)
    assert np.array_equal(img.astype(np.float32)[None], g[name + "/image"])      # bit-exact: same numpy calls
    assert np.array_equal(lab[None], g[name + "/label"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_pipeline_matches_reference(cuda_dev, name):
    from multimodal_segmentation_project_b200.utils import preprocess as P
    g = np.load(GOLD)
    x = torch.from_numpy(g[name + "/raw_image"]).cuda()
    y = P.preprocess(x, name).cpu().numpy()
    ref = g[name + "/image"][0]
    if P.get_modality(name) == "ct":
        assert np.array_equal(y, ref)                                            # clip + scale: bit-exact float32
    else:
        # moments are accumulated in fp64 here and pairwise in float32 by numpy: mean / std differ by ~1e-7 relative
        assert np.abs(y - ref).max() <= 2e-6, np.abs(y - ref).max()
        assert y.min() >= 0.0 and y.max() <= 1.0
    lab = torch.from_numpy(g[name + "/raw_label"]).cuda()
    out = P.remap_labels(lab, name)
    assert out.dtype == torch.int64 and np.array_equal(out.cpu().numpy(), g[name + "/label"][0])
    out8 = P.remap_labels(lab, name, out_dtype=torch.uint8)
    assert out8.dtype == torch.uint8 and np.array_equal(out8.cpu().numpy().astype(np.int64), g[name + "/label"][0])


@pytest.mark.gpu
def test_order_statistics_exact_and_large_volume(cuda_dev):
    """The radix select returns the exact order statistics numpy's partition finds (ties, negatives, -0.0, duplicates);
    the full MRI path on a 128^3 volume stays within float32 round-off of the oracle."""
    from multimodal_segmentation_project_b200 import _lib
    from multimodal_segmentation_project_b200.utils import preprocess as P
    L = _lib.load()
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 1000, 262147):
        a = rng.normal(0, 50, n).astype(np.float32)
        a[rng.integers(0, n, max(1, n // 10))] = 0.0
        a[rng.integers(0, n, max(1, n // 20))] = -0.0
        a[rng.integers(0, n, max(1, n // 10))] = np.float32(3.5)
        ranks = np.unique(np.clip(np.array([0, n // 100, n // 2, n - 1]), 0, n - 1))[:4]
        x = torch.from_numpy(a).cuda()
        r = torch.from_numpy(ranks.astype(np.int64)).cuda()
        v = torch.empty(len(ranks), dtype=torch.float32, device="cuda")
        ws = torch.empty(L.b200_preprocess_workspace_bytes(), dtype=torch.uint8, device="cuda")
        _lib.check(L.b200_select_ranks_f32(P._ptr(x), n, P._ptr(r), len(ranks), P._ptr(v), P._ptr(ws), P._stream()), "select")
        want = np.sort(a)[ranks]
        assert np.array_equal(v.cpu().numpy(), want), (n, v.cpu().numpy(), want)   # -0.0 == 0.0 compares equal: same value
    vol = (rng.gamma(2.0, 120.0, (128, 128, 128)) + 40 * rng.random((128, 128, 128))).astype(np.float32)
    y = P.preprocess_mri(torch.from_numpy(vol).cuda()).cpu().numpy()
    ref = OP.preprocess_mri(vol)
    assert np.abs(y - ref).max() <= 5e-6, np.abs(y - ref).max()
    with pytest.raises(RuntimeError):
        P.preprocess_ct(torch.zeros(4))          # CPU tensor: no fallback
    with pytest.raises(TypeError):
        P.preprocess_ct(torch.zeros(4, dtype=torch.float64, device="cuda"))
