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


# CHAOS label images store organs as grey levels; the reference accepts a band around each nominal level
# (utils/dataloader.py:169-180): nominal level -> (lowest, highest) accepted value
CHAOS_BANDS = {0: (0, 0), 63: (55, 70), 126: (110, 135), 189: (175, 200), 252: (240, 255)}


def remap_labels(label: np.ndarray, dataset_name: str) -> np.ndarray:
    """utils/dataloader.py:162-185 — table form: every (band -> class) rule is applied in the dictionary's order onto a
    zero-initialised volume, so a later rule would override an earlier one exactly as the reference's masked stores do."""
    if dataset_name.startswith("amos"):
        rules = [((old, old), new) for old, new in AMOS_MAPPING.items()]
    elif dataset_name.startswith("chaos"):
        rules = [(CHAOS_BANDS[old], new) for old, new in CHAOS_MAPPING.items()]
    else:
        return label  # 'ts*' and 'btcv' already use the target convention
    out = np.zeros_like(label)
    for (lo, hi), new in rules:
        out[(label >= lo) & (label <= hi)] = new
    return out
