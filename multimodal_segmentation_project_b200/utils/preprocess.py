"""Device-side input pipeline: the reference's ``utils/dataloader.py`` preprocessing on CUDA tensors.

``preprocess_ct`` / ``preprocess_mri`` / ``get_modality`` mirror ``CombinedDataset.preprocess_ct`` (:111-117),
``preprocess_mri`` (:128-145) and ``get_modality`` (:98-109); ``remap_labels`` mirrors the label handling of
``__getitem__`` (:162-185, AMOS dictionary and CHAOS intensity ranges).  Volumes stay in HBM: no numpy round trip, no
sort (exact order statistics come from a radix select), one host read of six scalars for the MRI path.
CUDA only (no CPU fallback): the numpy originals are the reference.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib
from .._lib import check

CT_WINDOW = (-160.0, 240.0)                      # utils/dataloader.py:114 "Typical abdominal window"
AMOS_MAPPING = {0: 0, 1: 1, 2: 3, 3: 3, 6: 2}    # utils/dataloader.py:43-49
# utils/dataloader.py:52-58 + :169-180 — intensity ranges of the CHAOS label images, in the dictionary's order
CHAOS_RANGES = [(0, 0, 0), (55, 70, 2), (110, 135, 3), (175, 200, 3), (240, 255, 1)]


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require(t, dtype, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CUDA tensor required (no CPU fallback; the reference's numpy code is the CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t.contiguous()


def get_modality(dataset_name: str) -> str:
    """'ct' if the name ends with '_ct', else 'mri' (also for unknown data sets) — utils/dataloader.py:98-109."""
    return "ct" if dataset_name.lower().endswith("_ct") else "mri"


def preprocess_ct(image: torch.Tensor) -> torch.Tensor:
    """clip to [-160, 240] HU and scale to [0, 1]; float32, bit-exact with the numpy original."""
    x = _require(image, torch.float32, "preprocess_ct")
    y = torch.empty_like(x)
    check(_lib.load().b200_ct_window(_ptr(x), _ptr(y), x.numel(), CT_WINDOW[0], CT_WINDOW[1], _stream()), "ct_window")
    return y


def _workspace(device):
    return torch.empty(_lib.load().b200_preprocess_workspace_bytes(), dtype=torch.uint8, device=device)


def preprocess_mri(image: torch.Tensor) -> torch.Tensor:
    """z-score, clip to the 1st..99th percentile (numpy 'linear' interpolation), min-max to [0, 1]; float32 out."""
    L = _lib.load()
    x = _require(image, torch.float32, "preprocess_mri")
    n = x.numel()
    ws = _workspace(x.device)
    # --- moments and the four order statistics that numpy's percentile interpolates between: one batch of launches
    mom = torch.empty(2, dtype=torch.float64, device=x.device)
    check(L.b200_moments_f32(_ptr(x), n, _ptr(ws), _ptr(mom), _stream()), "moments_f32")
    virt = (n - 1) * np.true_divide(np.array([1, 99]), 100)          # numpy's virtual indices for method='linear'
    prev = np.floor(virt).astype(np.int64)
    nxt = np.minimum(prev + 1, n - 1)
    ranks = torch.from_numpy(np.stack([prev, nxt], 1).reshape(-1)).to(x.device)
    vals = torch.empty(4, dtype=torch.float32, device=x.device)
    check(L.b200_select_ranks_f32(_ptr(x), n, _ptr(ranks), 4, _ptr(vals), _ptr(ws), _stream()), "select_ranks_f32")
    mom_h, vals_h = mom.cpu().numpy(), vals.cpu().numpy()            # the pipeline's only host read: six scalars
    # --- the reference's scalar arithmetic, in its dtypes (float32 statistics, float64 percentiles)
    mean = np.float32(mom_h[0])
    std = np.float32(np.sqrt(mom_h[1]))
    denom = std + np.float32(1e-8)
    z = (vals_h - mean) / denom                                      # monotone: order statistics of z = z of order statistics
    gamma = virt - prev                                              # float64

    def lerp(a, b, t):                                               # numpy's _lerp
        d = np.subtract(b, a)
        return np.subtract(b, d * (1 - t)) if t >= 0.5 else np.add(a, d * t)

    low = float(lerp(z[0], z[1], gamma[0]))
    high = float(lerp(z[2], z[3], gamma[1]))
    params = torch.tensor([float(mean), float(denom), low, high, high - low + 1e-8], dtype=torch.float64).to(x.device)
    y = torch.empty_like(x)
    check(L.b200_mri_normalize(_ptr(x), _ptr(y), n, _ptr(params), _stream()), "mri_normalize")
    return y


def preprocess(image: torch.Tensor, dataset_name: str) -> torch.Tensor:
    """Modality dispatch of ``CombinedDataset.__getitem__`` (utils/dataloader.py:153-159)."""
    return preprocess_ct(image) if get_modality(dataset_name) == "ct" else preprocess_mri(image)


def remap_labels(label: torch.Tensor, dataset_name: str, out_dtype: torch.dtype = torch.int64) -> torch.Tensor:
    """AMOS / CHAOS label conventions -> {0 background, 1 spleen, 2 liver, 3 kidneys}; 'ts*' and 'btcv' pass through
    (utils/dataloader.py:162-185).  ``out_dtype=torch.uint8`` writes 1-byte labels."""
    x = _require(label, torch.int64, "remap_labels")
    if dataset_name.startswith("amos"):
        table = [(k, k, v) for k, v in AMOS_MAPPING.items()]
    elif dataset_name.startswith("chaos"):
        table = CHAOS_RANGES
    else:
        return x.to(out_dtype) if out_dtype != torch.int64 else x
    if out_dtype not in (torch.int64, torch.uint8):
        raise TypeError("remap_labels: out_dtype must be int64 or uint8")
    lo = (ctypes.c_int64 * len(table))(*[t[0] for t in table])
    hi = (ctypes.c_int64 * len(table))(*[t[1] for t in table])
    val = (ctypes.c_int64 * len(table))(*[t[2] for t in table])
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    check(_lib.load().b200_label_remap(_ptr(x), _ptr(out), x.numel(), lo, hi, val, len(table), int(out_dtype == torch.uint8), _stream()),
          "label_remap")
    return out
