"""CPU restatement (numpy) of the reference's input pipeline — utils/dataloader.py.

TEST INFRASTRUCTURE (see oracle/__init__.py): only tests may import this.  Pinned against outputs of the reference's own
``CombinedDataset.__getitem__`` (tests/golden/preprocess.npz, made by oracle/make_golden_preprocess.py)."""
from __future__ import annotations

import numpy as np

AMOS_MAPPING = {0: 0, 1: 1, 2: 3, 3: 3, 6: 2}            # utils/dataloader.py:43-49
CHAOS_MAPPING = {0: 0, 63: 2, 126: 3, 189: 3, 252: 1}    # utils/dataloader.py:52-58


def preprocess_ct(image: np.ndarray) -> np.ndarray:
    """utils/dataloader.py:111-117"""
    lo, hi = -160, 240
    image = np.clip(image, lo, hi)
    return (image - lo) / (hi - lo)


def preprocess_mri(image: np.ndarray) -> np.ndarray:
    """utils/dataloader.py:128-145 (z-score, 1st..99th percentile clip, min-max)"""
    mean = np.mean(image)
    std = np.std(image)
    image = (image - mean) / (std + 1e-8)
    low, high = np.percentile(image, [1, 99])
    image = np.clip(image, low, high)
    image = (image - low) / (high - low + 1e-8)
    return image.astype(np.float32)


def preprocess(image: np.ndarray, dataset_name: str) -> np.ndarray:
    """utils/dataloader.py:153-159: '_ct' suffix -> CT window, everything else -> MRI normalisation"""
    return preprocess_ct(image) if dataset_name.lower().endswith("_ct") else preprocess_mri(image)


def remap_labels(label: np.ndarray, dataset_name: str) -> np.ndarray:
    """utils/dataloader.py:162-185"""
    if dataset_name.startswith("amos"):
        new = np.zeros_like(label)
        for old, idx in AMOS_MAPPING.items():
            new[label == old] = idx
        return new
    if dataset_name.startswith("chaos"):
        new = np.zeros_like(label)
        for old, val in CHAOS_MAPPING.items():
            if old == 63:
                mask = (label >= 55) & (label <= 70)
            elif old == 126:
                mask = (label >= 110) & (label <= 135)
            elif old == 189:
                mask = (label >= 175) & (label <= 200)
            elif old == 252:
                mask = (label >= 240) & (label <= 255)
            else:
                mask = label == 0
            new[mask] = val
        return new
    return label
