"""Generates tests/golden/preprocess.npz with the UNMODIFIED reference data pipeline (utils/dataloader.py imported from
/root/reference; nibabel / monai / accelerate are stubbed, the file reader is replaced by in-memory arrays):

    python -m oracle.make_golden_preprocess

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each case is one CombinedDataset.__getitem__ call: modality-specific
intensity preprocessing (preprocess_ct / preprocess_mri) and the AMOS / CHAOS label remaps."""
from __future__ import annotations

import os
import sys

import numpy as np

from .make_golden import OUT, REF, _import_reference


def main():
    _import_reference()  # installs the stubs for the absent third-party packages
    if "scipy" not in sys.modules:
        import scipy.ndimage  # noqa: F401  (real scipy is installed)
    sys.path.insert(0, REF)
    import utils.dataloader as dl

    rng = np.random.default_rng(7)
    vols = {
        "amos_ct": (rng.uniform(-1000, 1500, (20, 24, 18)), rng.integers(0, 16, (20, 24, 18))),
        "chaos_mri": (rng.gamma(2.0, 150.0, (24, 20, 22)), rng.choice([0, 63, 126, 189, 252, 60, 115, 180, 250, 100, 30, 255, 55, 70, 71], (24, 20, 22))),
        "ts_ct": (rng.normal(40, 300, (16, 16, 16)), rng.integers(0, 4, (16, 16, 16))),
        "btcv": (np.abs(rng.normal(300, 120, (18, 16, 20))) + 5 * (rng.random((18, 16, 20)) > 0.99) * 4000, rng.integers(0, 4, (18, 16, 20))),
        "amos_mri": (rng.gamma(1.5, 80.0, (33, 17, 29)), rng.integers(0, 8, (33, 17, 29))),
    }

    class _Img:
        def __init__(self, a):
            self.a = a

        def get_fdata(self):
            return np.asarray(self.a, dtype=np.float64)

    files = {}
    dl.nib.load = lambda path: _Img(files[path])
    ds = object.__new__(dl.CombinedDataset)
    ds.transform = None
    ds.amos_mapping = {0: 0, 1: 1, 2: 3, 3: 3, 6: 2}
    ds.chaos_mapping = {0: 0, 63: 2, 126: 3, 189: 3, 252: 1}
    ds.samples = []
    out = {}
    for name, (img, lab) in vols.items():
        files[name + "/img"] = img
        files[name + "/lab"] = lab
        ds.samples.append({"image_path": name + "/img", "label_path": name + "/lab", "dataset_name": name})
    # the literal dictionaries above must be the reference's own
    src = open(os.path.join(REF, "utils", "dataloader.py")).read()
    assert "2: 3,  # right kidney" in src and "252: 1," in src
    for i, name in enumerate(vols):
        image_t, label_t = ds[i]
        out[name + "/raw_image"] = np.asarray(vols[name][0], dtype=np.float64).astype(np.float32)   # what astype(np.float32) of get_fdata gives
        out[name + "/raw_label"] = np.asarray(vols[name][1], dtype=np.float64).astype(np.int64)
        out[name + "/image"] = image_t.numpy()
        out[name + "/label"] = label_t.numpy()
        assert image_t.dtype.is_floating_point and image_t.shape[0] == 1
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **out)
    print("preprocess golden written:", {k: v.shape for k, v in out.items() if k.endswith("/image")})


if __name__ == "__main__":
    main()
