"""ctypes binding of libb200unet.so (the C ABI declared in include/b200unet.h).

There is deliberately no fallback: if the shared library is missing or a kernel call fails, the
caller gets an exception.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"``
(or ``make -C multimodal_segmentation_project_b200/csrc``).
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200unet.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200unet.h")

B200_F32, B200_BF16 = 0, 1
PACK_FPROP, PACK_DGRAD, PACK_FPROP_TC, PACK_DGRAD_TC = 0, 1, 2, 3
LOSS_DICE_CE, LOSS_TVERSKY, LOSS_CE_TVERSKY, LOSS_DICE, LOSS_CE = 0, 1, 2, 3, 4

_lib = None


class B200Error(RuntimeError):
    """A libb200unet call returned a negative status."""


def declared_symbols(header_path: str = HEADER_PATH) -> list[str]:
    """Names of every function the C header declares (used by the CPU-side ABI test)."""
    with open(header_path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def _declare(lib) -> None:
    P, I, L, F = c_void_p, c_int, c_int64, c_float
    sig = {
        "b200_last_error": (c_char_p, []),
        "b200_version": (I, []),
        "b200_check_device": (I, [I]),
        "b200_launch_count": (L, []),
        "b200_ncdhw_to_ndhwc": (I, [I, P, P, L, L, L, P]),
        "b200_ndhwc_to_ncdhw": (I, [I, P, P, L, L, L, P]),
        "b200_cast_from_f32": (I, [I, P, P, L, P]),
        "b200_cast_to_f32": (I, [I, P, P, L, P]),
        "b200_pack_conv3_weights": (I, [I, I, P, P, I, I, P]),
        "b200_pack_conv3_bytes": (L, [I, I, I, I]),
        "b200_pack_conv3_batched": (I, [P, I, L, P]),
        "b200_conv3d_k3_select": (I, [I, I, I, I, I, I, I, I, I, I]),
        "b200_set_conv_persistent": (I, [I]),
        "b200_set_conv_rowstream": (I, [I]),
        "b200_conv3d_k3": (I, [I, I, P, I, P, I, P, P, P, I, P, I, I, I, I, I, P]),
        "b200_conv3d_k3_bnstats_blocks": (I, [I, I, I, I, I, I, I, I, I, I]),
        "b200_conv3d_k3_bnstats": (I, [I, I, P, I, P, I, P, P, P, I, I, I, I, I, P, P]),
        "b200_conv3d_k3_bnbwd_blocks": (I, [I, I, I, I, I, I, I, I]),
        "b200_conv3d_k3_bnbwd": (I, [I, I, P, I, P, P, I, I, I, I, I, P, P, P, P, P, P, P]),
        "b200_conv3d_wgrad_workspace": (L, [I, I, I, I, I, I, I]),
        "b200_set_wgrad_impl": (I, [I]),
        "b200_set_wgrad_pair": (I, [I, I]),
        "b200_debug_fail_next_wgrad": (I, [I]),
        "b200_conv3d_wgrad": (I, [I, P, I, P, I, P, I, P, P, P, L, I, I, I, I, P]),
        "b200_bn_partials_bytes": (L, [I]),
        "b200_bn_stats": (I, [I, P, L, I, P, P]),
        "b200_bn_finalize": (I, [I, P, P, L, I, P, P, F, F, I, P, P, P, P, P, P, P, P]),
        "b200_bn_finalize_ex": (I, [P, I, P, L, I, P, P, F, F, P, P, P, P, P, P, P, P]),
        "b200_bn_act_fwd": (I, [I, P, P, P, P, P, P, I, L, L, I, P]),
        "b200_bn_act_bwd_reduce": (I, [I, P, P, P, P, P, P, P, I, L, L, I, P, P]),
        "b200_bn_bwd_finalize": (I, [P, L, I, P, P, P, P]),
        "b200_bn_bwd_finalize_ex": (I, [P, I, L, I, P, P, P, P]),
        "b200_head_blocks": (I, [L, L]),
        "b200_head_fwd": (I, [P, P, P, P, P, P, I, P, I, L, L, I, I, P, P, P, P]),
        "b200_head_bwd": (I, [P, P, I, P, P, P, P, P, P, P, P, L, L, I, I, P, P, P, P, P, P]),
        "b200_bn_act_bwd_apply": (I, [I, P, P, P, P, P, P, P, P, I, P, I, L, L, I, P]),
        "b200_channel_sum": (I, [I, P, L, I, P, P, P]),
        "b200_maxpool2_fwd": (I, [I, P, P, I, I, I, I, I, P]),
        "b200_maxpool2_bwd": (I, [I, P, P, P, I, I, I, I, I, P]),
        "b200_bn_act_pool_fwd_supported": (I, [I, I, I, I]),
        "b200_bn_act_pool_fwd": (I, [I, P, P, P, P, P, P, I, I, I, I, I, P]),
        "b200_ct_window": (I, [P, P, L, F, F, P]),
        "b200_preprocess_workspace_bytes": (L, []),
        "b200_moments_f32": (I, [P, L, P, P, P]),
        "b200_select_ranks_f32": (I, [P, L, P, I, P, P, P]),
        "b200_mri_normalize": (I, [P, P, L, P, P]),
        "b200_label_remap": (I, [P, P, L, P, P, P, I, I, P]),
        "b200_maxpool2_bwd_add": (I, [I, P, P, P, P, I, I, I, I, I, P]),
        "b200_convt2_fwd": (I, [I, P, P, P, P, I, I, I, I, I, I, P]),
        "b200_convt2_bwd_data": (I, [I, P, P, P, I, I, I, I, I, I, P]),
        "b200_convt2_wgrad_workspace": (L, [I, I, I, I, I, I]),
        "b200_convt2_bwd_weight": (I, [I, P, P, P, P, P, L, I, I, I, I, I, I, P]),
        "b200_nearest_resize_fwd": (I, [I, P, P, I, I, I, I, I, I, I, I, P]),
        "b200_nearest_resize_bwd": (I, [I, P, P, I, I, I, I, I, I, I, I, P]),
        "b200_conv1x1_fwd": (I, [I, P, P, P, P, L, L, I, I, I, P]),
        "b200_conv1x1_bwd": (I, [I, P, P, P, P, P, P, P, L, L, I, I, P]),
        "b200_conv1x1_partials_bytes": (L, [I, I]),
        "b200_seg_loss_fwd": (I, [P, P, L, I, L, P, P]),
        "b200_kd_loss_fwd": (I, [P, P, P, F, L, I, L, P, P]),
        "b200_seg_loss_finalize": (I, [P, I, F, F, F, F, I, L, I, L, P, P, P]),
        "b200_seg_loss_bwd": (I, [P, P, P, P, P, F, L, I, L, P, P]),
        "b200_confusion": (I, [P, P, L, I, L, P, P]),
        "b200_argmax": (I, [P, L, I, L, P, P]),
        "b200_window_accumulate": (I, [P, P, P, I, I, I, I, I, I, I, I, I, I, P]),
        "b200_window_finalize": (I, [P, P, I, L, P]),
        "b200_gap_fwd": (I, [I, P, P, L, L, I, P]),
        "b200_gap_bwd": (I, [I, P, P, I, L, L, I, P]),
        "b200_scale_f32": (I, [P, P, F, L, P]),
        "b200_linear_fwd": (I, [P, P, P, P, I, P, I, I, I, P]),
        "b200_linear_bwd": (I, [P, P, P, P, P, I, P, P, P, I, I, I, P]),
        "b200_ce_rows": (I, [P, P, I, I, P, P, P]),
        "b200_aug_bias_field": (I, [P, P, I, I, I, I, I, P, P]),
        "b200_aug_gaussian_noise": (I, [P, P, P, L, F, F, P]),
        "b200_minmax_f32": (I, [P, L, P, P, P]),
        "b200_aug_adjust_contrast": (I, [P, P, L, P, F, P]),
        "b200_aug_histogram_shift": (I, [P, P, L, P, P, P, I, P]),
        "b200_aug_coarse_dropout": (I, [P, P, I, I, I, I, P, I, F, P]),
        "b200_adamw_prepare": (I, [P, F, F, P, P]),
        "b200_adamw_flat": (I, [P, P, P, P, L, P, F, F, F, F, F, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._b200_signatures = sig


def load():
    """Load (once) and return the ctypes handle. Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} not found: the CUDA library has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        _declare(lib)
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().b200_last_error().decode("utf-8", "replace")
        err = {-1: "B200_ERR_SHAPE", -2: "B200_ERR_ALIGN", -3: "B200_ERR_ARCH", -4: "B200_ERR_CUDA", -5: "B200_ERR_UNSUPPORTED"}.get(rc, str(rc))
        if rc in (-1, -5):
            raise ValueError(f"libb200unet {what}: {err}: {msg}")
        raise B200Error(f"libb200unet {what}: {err}: {msg}")


def launch_count() -> int:
    return int(load().b200_launch_count())
