"""Device versions of the reference's combined_transform() (utils/dataloader.py:252-260: five MONAI intensity transforms) against the
numpy restatement of MONAI's published algorithms (oracle/augment_oracle.py), draw for draw."""
import numpy as np
import pytest
import torch

from oracle import augment_oracle as OA


def test_oracle_bias_field_matches_direct_polynomial():
    """CPU: the leggrid3d formulation equals the explicit sum over MONAI's (i, j, k) coefficient order."""
    rng = np.random.RandomState(0)
    img = rng.rand(1, 5, 6, 7).astype(np.float32) + 0.5
    coeff = rng.uniform(0, 0.1, 20)
    out = OA.rand_bias_field(img, coeff)
    P = [lambda x: np.ones_like(x), lambda x: x, lambda x: 0.5 * (3 * x * x - 1), lambda x: 0.5 * (5 * x ** 3 - 3 * x)]
    xs, ys, zs = [np.linspace(-1, 1, n) for n in (5, 6, 7)]
    f = np.zeros((5, 6, 7))
    for c, (i, j, k) in zip(coeff, OA.bias_field_coeff_order(3)):
        f += c * P[i](xs)[:, None, None] * P[j](ys)[None, :, None] * P[k](zs)[None, None, :]
    assert np.allclose(out, img * np.exp(f)[None], rtol=2e-6)
    assert len(OA.bias_field_coeff_order(3)) == 20


def test_oracle_histogram_shift_and_contrast_fixed_points():
    rng = np.random.RandomState(1)
    img = rng.rand(1, 4, 4, 4).astype(np.float32)
    ref = np.linspace(0, 1, 5)
    assert np.allclose(OA.histogram_shift(img, ref, ref), img, atol=1e-7)          # identical control points: identity
    out = OA.adjust_contrast(img, 1.0)
    assert np.allclose(out, img, atol=1e-6)
    g = OA.adjust_contrast(img, 0.7)
    assert g.min() >= img.min() - 1e-6 and abs(g.max() - img.max()) < 1e-5         # range preserved


@pytest.mark.gpu
def test_device_transforms_match_oracle_draw_for_draw(cuda_dev):
    from multimodal_segmentation_project_b200.utils import augment as A
    rng = np.random.RandomState(5)
    img = (rng.rand(1, 24, 40, 33).astype(np.float32) * 1.7 - 0.2)
    lab = rng.randint(0, 4, img.shape).astype(np.int64)
    x, y = torch.from_numpy(img).cuda(), torch.from_numpy(lab).cuda()
    coeff = rng.uniform(0, 0.1, 20)
    assert np.allclose(A.bias_field(x, coeff).cpu().numpy(), OA.rand_bias_field(img, coeff), rtol=2e-6, atol=1e-7)
    z = rng.randn(*img.shape).astype(np.float32)
    assert np.array_equal(A.gaussian_noise(x, torch.from_numpy(z).cuda(), 0.0, 0.0073).cpu().numpy(), OA.rand_gaussian_noise(img, z, 0.0, 0.0073))
    mm = A.minmax(x).cpu().numpy()
    assert mm[0] == img.min() and mm[1] == img.max()
    for gamma in (0.7, 1.0, 1.5):
        assert np.allclose(A.adjust_contrast(x, gamma).cpu().numpy(), OA.adjust_contrast(img, gamma), rtol=3e-6, atol=1e-6)
    ref = np.linspace(0, 1, 5)
    flt = ref.copy()
    for i in range(1, 4):
        flt[i] = rng.uniform(flt[i - 1], flt[i + 1])
    assert np.allclose(A.histogram_shift(x, ref, flt).cpu().numpy(), OA.histogram_shift(img, ref, flt), rtol=1e-6, atol=1e-7)
    const = torch.full((1, 4, 4, 4), 0.25, device="cuda")
    assert torch.equal(A.histogram_shift(const, ref, flt), const)                  # min == max: unchanged (MONAI warns and returns)
    holes = [(3, 19, 10, 26, 0, 16), (8, 24, 24, 40, 17, 33)]
    gi, gl = A.coarse_dropout(x, y, holes, 0.0)
    oi, ol = OA.coarse_dropout(img, lab, holes, 0.0)
    assert np.array_equal(gi.cpu().numpy(), oi) and np.array_equal(gl.cpu().numpy(), ol)


@pytest.mark.gpu
def test_combined_transform_contract(cuda_dev):
    """Same call contract as the reference's combined_transform(): dict in, dict out, image float32 / label int64 of unchanged shape;
    seeded runs repeat; each fired transform equals the oracle applied with the recorded draws (noise excluded: torch generator)."""
    from multimodal_segmentation_project_b200.utils.augment import combined_transform
    rng = np.random.RandomState(2)
    img = rng.rand(1, 32, 32, 32).astype(np.float32)
    lab = rng.randint(0, 4, img.shape).astype(np.int64)
    sample = {"image": torch.from_numpy(img).cuda(), "label": torch.from_numpy(lab).cuda()}
    fired = set()
    for seed in range(12):
        t = combined_transform(seed=seed)
        out = t(sample)
        again = combined_transform(seed=seed)(sample)
        assert torch.equal(out["image"], again["image"]) and torch.equal(out["label"], again["label"])
        assert out["image"].dtype == torch.float32 and out["label"].dtype == torch.int64 and out["image"].shape == sample["image"].shape
        d = t.last_draws
        fired |= set(d)
        if "noise_std" in d:
            continue
        ref_i, ref_l = img, lab
        if "bias_coeff" in d:
            ref_i = OA.rand_bias_field(ref_i, d["bias_coeff"])
        if "gamma" in d:
            ref_i = OA.adjust_contrast(ref_i, d["gamma"])
        if "control_points" in d:
            ref_i = OA.histogram_shift(ref_i, np.asarray(d["control_points"][0]), np.asarray(d["control_points"][1]))
        if "holes" in d:
            ref_i, ref_l = OA.coarse_dropout(ref_i, ref_l, d["holes"])
        assert np.allclose(out["image"].cpu().numpy(), ref_i, rtol=1e-5, atol=2e-6), (seed, d.keys())
        assert np.array_equal(out["label"].cpu().numpy(), ref_l)
    assert {"bias_coeff", "gamma", "control_points", "holes", "noise_std"} <= fired
