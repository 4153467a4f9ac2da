"""Device-side counterpart of the reference's ``combined_transform()`` (utils/dataloader.py:223-261): the five MONAI intensity
augmentations that are active there — RandBiasFieldd(prob .3), RandGaussianNoised(prob .3, std .01), RandAdjustContrastd(prob .3,
gamma (.7, 1.5)), RandHistogramShiftd(prob .3, 5 control points), RandCoarseDropoutd(image + label, 2 holes of 16^3, fill 0,
prob .3) — applied to CUDA tensors ``image`` float32 ``[C, D, H, W]`` and ``label`` int64 ``[C, D, H, W]``.

The random draws follow MONAI's recipes (which transform fires, the coefficients / gamma / control points / hole positions) with a
``numpy.random.RandomState`` (``set_random_state(seed)`` like monai.transforms.Compose); the Gaussian field itself comes from a
torch CUDA generator.  The arithmetic runs in libb200unet (csrc/augment_kernels.cu).  MONAI is not installed in the build image:
stream-for-stream equality with its RNG consumption is not claimed, equality of each transform GIVEN its draw is tested against
oracle/augment_oracle.py.  CUDA only."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib
from .._lib import check


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _img(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CUDA tensor required (no CPU fallback)")
    if t.dtype != torch.float32 or t.dim() != 4:
        raise TypeError(f"{name}: expected a float32 [C, D, H, W] tensor, got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


def bias_field(image, coeff, degree=3):
    x = _img(image, "bias_field")
    C, D, H, W = x.shape
    c = torch.zeros(20, dtype=torch.float64)
    c[: len(coeff)] = torch.as_tensor(np.asarray(coeff, dtype=np.float64))
    c = c.to(x.device)
    y = torch.empty_like(x)
    check(_lib.load().b200_aug_bias_field(_ptr(x), _ptr(y), C, D, H, W, int(degree), _ptr(c), _stream()), "aug_bias_field")
    return y


def gaussian_noise(image, z, mean, std):
    x = _img(image, "gaussian_noise")
    z = z.to(torch.float32).contiguous()
    y = torch.empty_like(x)
    check(_lib.load().b200_aug_gaussian_noise(_ptr(x), _ptr(z), _ptr(y), x.numel(), float(mean), float(std), _stream()), "aug_gaussian_noise")
    return y


def minmax(image):
    x = image.contiguous()
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    ws = torch.empty(2, dtype=torch.int32, device=x.device)
    check(_lib.load().b200_minmax_f32(_ptr(x), x.numel(), _ptr(out), _ptr(ws), _stream()), "minmax_f32")
    return out


def adjust_contrast(image, gamma):
    x = _img(image, "adjust_contrast")
    y = torch.empty_like(x)
    check(_lib.load().b200_aug_adjust_contrast(_ptr(x), _ptr(y), x.numel(), _ptr(minmax(x)), float(gamma), _stream()), "aug_adjust_contrast")
    return y


def histogram_shift(image, reference, floating):
    x = _img(image, "histogram_shift")
    r = torch.as_tensor(np.asarray(reference, dtype=np.float64)).to(x.device)
    f = torch.as_tensor(np.asarray(floating, dtype=np.float64)).to(x.device)
    y = torch.empty_like(x)
    check(_lib.load().b200_aug_histogram_shift(_ptr(x), _ptr(y), x.numel(), _ptr(minmax(x)), _ptr(r), _ptr(f), int(r.numel()), _stream()),
          "aug_histogram_shift")
    return y


def coarse_dropout(image, label, holes, fill=0.0):
    """in place on copies; holes = [(d0, d1, h0, h1, w0, w1), ...]"""
    x = _img(image, "coarse_dropout").clone()
    lab = label.contiguous().clone()
    if lab.dtype != torch.int64 or lab.shape != x.shape:
        raise TypeError("coarse_dropout: label must be int64 with the image's shape")
    C, D, H, W = x.shape
    h = torch.as_tensor(np.asarray(holes, dtype=np.int32).reshape(-1, 6)).to(x.device)
    check(_lib.load().b200_aug_coarse_dropout(_ptr(x), _ptr(lab), C, D, H, W, _ptr(h), int(h.shape[0]), float(fill), _stream()), "aug_coarse_dropout")
    return x, lab


class CombinedTransform:
    """``combined_transform()`` of the reference as a callable on ``{'image': ..., 'label': ...}`` of CUDA tensors."""

    def __init__(self, prob=0.3, noise_std=0.01, gamma=(0.7, 1.5), control_points=5, holes=2, hole_size=(16, 16, 16), degree=3,
                 coeff_range=(0.0, 0.1), seed=None):
        self.prob, self.noise_std, self.gamma, self.ncp, self.holes, self.hole_size = prob, noise_std, gamma, control_points, holes, hole_size
        self.degree, self.coeff_range = degree, coeff_range
        self.R = np.random.RandomState(seed)
        self.gen = None
        self.last_draws = {}

    def set_random_state(self, seed=None):
        self.R = np.random.RandomState(seed)
        self.gen = None
        return self

    def __call__(self, sample):
        img, lab = sample["image"], sample["label"]
        draws = {}
        if self.R.rand() < self.prob:      # RandBiasField: n_coeff = C(degree + 3, 3) coefficients U(coeff_range)
            n = (self.degree + 1) * (self.degree + 2) * (self.degree + 3) // 6
            draws["bias_coeff"] = self.R.uniform(*self.coeff_range, n).tolist()
            img = bias_field(img, draws["bias_coeff"], self.degree)
        if self.R.rand() < self.prob:      # RandGaussianNoise: std' ~ U(0, std), field N(0, std')
            draws["noise_std"] = float(self.R.uniform(0, self.noise_std))
            if self.gen is None:
                self.gen = torch.Generator(device=img.device).manual_seed(int(self.R.randint(0, 2 ** 31 - 1)))
            z = torch.randn(img.shape, device=img.device, generator=self.gen)
            img = gaussian_noise(img, z, 0.0, draws["noise_std"])
        if self.R.rand() < self.prob:      # RandAdjustContrast
            draws["gamma"] = float(self.R.uniform(*self.gamma))
            img = adjust_contrast(img, draws["gamma"])
        if self.R.rand() < self.prob:      # RandHistogramShift: interior control points redrawn between their neighbours
            ref = np.linspace(0, 1, self.ncp)
            flt = np.copy(ref)
            for i in range(1, self.ncp - 1):
                flt[i] = self.R.uniform(flt[i - 1], flt[i + 1])
            draws["control_points"] = (ref.tolist(), flt.tolist())
            img = histogram_shift(img, ref, flt)
        if self.R.rand() < self.prob:      # RandCoarseDropout: fixed-size boxes at uniformly drawn valid positions
            _, D, H, W = img.shape
            size = [min(s, d) for s, d in zip(self.hole_size, (D, H, W))]
            boxes = []
            for _ in range(self.holes):
                o = [int(self.R.randint(0, d - s + 1)) for s, d in zip(size, (D, H, W))]
                boxes.append((o[0], o[0] + size[0], o[1], o[1] + size[1], o[2], o[2] + size[2]))
            draws["holes"] = boxes
            img, lab = coarse_dropout(img, lab, boxes, 0.0)
        self.last_draws = draws
        return {"image": img, "label": lab}


def combined_transform(seed=None):
    """Same name and role as the reference's factory (utils/dataloader.py:223)."""
    return CombinedTransform(seed=seed)
