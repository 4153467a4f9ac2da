"""Restatement of the five MONAI intensity transforms the reference's combined_transform() applies
(utils/dataloader.py:252-260) — TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED BY REFERENCE VECTORS: MONAI (requirements.txt: monai>=1.2.0) is a third-party dependency that is neither in
/root/reference nor installed in the build image, so no golden output of the real transforms could be generated.  The functions
below restate MONAI's published array transforms (monai/transforms/intensity/array.py, v1.2-1.3) given EXPLICIT random draws;
the CUDA kernels are checked against them, draw for draw.
"""
from __future__ import annotations

import numpy as np


def bias_field_coeff_order(degree=3):
    """MONAI RandBiasField._generate_random_field (rank 3): coefficient t multiplies P_i(x) P_j(y) P_k(z) for the t-th (i, j, k) of
    `for i in range(degree + 1): for j in range(degree + 1 - i): for k in range(degree + 1 - i - j)`."""
    return [(i, j, k) for i in range(degree + 1) for j in range(degree + 1 - i) for k in range(degree + 1 - i - j)]


def rand_bias_field(img, coeff, degree=3):
    """img [C, D, H, W] float32; out = img * exp(leggrid3d(linspace(-1,1,D), linspace(-1,1,H), linspace(-1,1,W), coeff_mat)), float64
    arithmetic, cast back to float32 (the same field for every channel draw: MONAI draws the coefficients once per call)."""
    C, D, H, W = img.shape
    coeff_mat = np.zeros((degree + 1,) * 3)
    for c, (i, j, k) in zip(coeff, bias_field_coeff_order(degree)):
        coeff_mat[i, j, k] = c
    coords = [np.linspace(-1.0, 1.0, n, dtype=np.float32) for n in (D, H, W)]
    field = np.polynomial.legendre.leggrid3d(coords[0], coords[1], coords[2], coeff_mat)
    return (img * np.exp(field)[None]).astype(np.float32)


def rand_gaussian_noise(img, z, mean, std):
    """out = img + noise, noise = mean + std * z with z the standard-normal draw (std already sampled U(0, std_max))."""
    return (img + (np.float32(mean) + np.float32(std) * z)).astype(np.float32)


def adjust_contrast(img, gamma):
    """MONAI AdjustContrast: ((img - min) / float(range + 1e-7)) ** gamma * range + min"""
    eps = 1e-7
    img_min = img.min()
    img_range = img.max() - img_min
    return (((img - img_min) / float(img_range + eps)) ** gamma * img_range + img_min).astype(np.float32)


def histogram_shift(img, reference, floating):
    """MONAI RandHistogramShift.__call__: np.interp between the scaled control points"""
    img_min, img_max = img.min(), img.max()
    if img_min == img_max:
        return img.copy()
    xp = reference * (img_max - img_min) + img_min
    yp = floating * (img_max - img_min) + img_min
    return np.interp(img, xp, yp).astype(np.float32)


def coarse_dropout(img, label, holes, fill=0.0):
    """MONAI RandCoarseDropout(dropout_holes=True): every hole (d0, d1, h0, h1, w0, w1) of every channel is set to fill"""
    img, label = img.copy(), label.copy()
    for d0, d1, h0, h1, w0, w1 in holes:
        img[:, d0:d1, h0:h1, w0:w1] = fill
        label[:, d0:d1, h0:h1, w0:w1] = int(fill)
    return img, label
