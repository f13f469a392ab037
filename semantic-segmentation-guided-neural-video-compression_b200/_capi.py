"""ctypes binding of include/dmc_b200.h (the C ABI of the CUDA engine).

Loading never falls back to anything: if the shared library is missing it is built with
nvcc (build.py); if that fails, or a call is made without a CUDA device, an exception is
raised.  There is no CPU path behind this module.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_void_p

from . import build as _build

VARIANT_IDS = {"old": 0, "performance": 1, "fast": 2, "mask_prop": 3, "intra": 4}
FLAG_SIMT_GEMM = 1
FLAG_KEEP_TAPS = 2
FLAG_RECON_SPLIT3 = 8

# every symbol include/dmc_b200.h declares: name -> (restype, argtypes)
_F = POINTER(c_float)
SIGNATURES = {
    "dmc_create": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "dmc_destroy": (None, [c_void_p]),
    "dmc_last_error": (c_char_p, [c_void_p]),
    "dmc_num_weights": (c_int, [c_void_p]),
    "dmc_weight_key": (c_char_p, [c_void_p, c_int]),
    "dmc_weight_shape": (c_int, [c_void_p, c_int, POINTER(c_int64)]),
    "dmc_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int, c_void_p]),
    "dmc_finalize_weights": (c_int, [c_void_p, c_void_p]),
    "dmc_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dmci_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dmc_decode_begin": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "dmc_decode_sigma": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "dmc_decode_symbols": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "dmc_decode_finish": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "dmc_get_tap": (c_int, [c_void_p, c_char_p, c_void_p, c_int64, POINTER(c_int64), c_void_p]),
    "dmc_frame_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                c_int, c_void_p]),
    "dmc_frames_from_u8": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_void_p]),
    "dmc_mask_from_logits": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "dmc_op_conv2d": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 12 + [c_void_p]),
    "dmc_op_depth_conv_block": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p] +
                                [c_int] * 8 + [c_void_p]),
    "dmc_op_gaussian_bits": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "dmc_dcb_train_create": (c_int, [c_int] * 9 + [POINTER(c_void_p)]),
    "dmc_dcb_train_destroy": (None, [c_void_p]),
    "dmc_dcb_train_last_error": (c_char_p, [c_void_p]),
    "dmc_dcb_train_forward": (c_int, [c_void_p, c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_int, c_void_p]),
    "dmc_dcb_train_backward": (c_int, [c_void_p, c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p,
                                       POINTER(c_void_p), c_void_p, c_int, c_void_p]),
    "dmc_conv1x1_train_create": (c_int, [c_int] * 7 + [POINTER(c_void_p)]),
    "dmc_conv1x1_train_destroy": (None, [c_void_p]),
    "dmc_conv1x1_train_last_error": (c_char_p, [c_void_p]),
    "dmc_conv1x1_train_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dmc_conv1x1_train_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                           c_void_p]),
    "dmc_convkxk_train_create": (c_int, [c_int] * 10 + [POINTER(c_void_p)]),
    "dmc_convkxk_train_destroy": (None, [c_void_p]),
    "dmc_convkxk_train_last_error": (c_char_p, [c_void_p]),
    "dmc_convkxk_train_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dmc_convkxk_train_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                           c_void_p]),
    "dmc_op_quant_train": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "dmc_op_gaussian_bits_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                              c_void_p]),
    "dmc_kernel_launches": (c_int64, []),
    "dmc_profile_enable": (c_int, [c_void_p, c_int]),
    "dmc_profile_read": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64), POINTER(c_double),
                                 POINTER(c_double)]),
    "dmc_bench_gemm": (c_int, [c_int] * 8 + [POINTER(c_float)]),
    "dmc_bench_dwconv": (c_int, [c_int] * 6 + [POINTER(c_float)]),
    "dmc_bench_dcb": (c_int, [c_int] * 7 + [POINTER(c_float)]),
    "dmc_num_sms": (c_int, []),
    "dmc_version": (c_char_p, []),
    "dmc_set_acc_comp": (c_int, [c_float]),
    "dmc_get_acc_comp": (c_float, []),
    "dmc_rans_create": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "dmc_rans_destroy": (None, [c_void_p]),
    "dmc_rans_last_error": (c_char_p, [c_void_p]),
    "dmc_rans_index_gaussian": (c_int, [c_void_p, c_int64, c_float, c_float, c_int, c_void_p, c_void_p]),
    "dmc_rans_index_channels": (c_int, [c_int64, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "dmc_rans_max_bytes": (c_int64, [c_int64]),
    "dmc_rans_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, POINTER(c_int64), c_void_p]),
    "dmc_rans_decode": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
}

# names of the bits of dmc_forward's finite_flag (include/dmc_b200.h)
FINITE_TAGS = ("feature_adaptor", "feature_extractor.ctx", "feature_extractor.ctx_t", "encoder", "hyper_encoder",
               "y_prior_fusion", "y_hat", "decoder")

_lib = None


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """dlopen the engine; builds it first when the .so is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing:
        path = _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: the CUDA extension is required (no fallback)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class EngineError(RuntimeError):
    pass


def check(rc: int, handle=None):
    if rc != 0:
        msg = load().dmc_last_error(handle)
        raise EngineError(f"dmc_b200 error {rc}: {msg.decode() if msg else '?'}")
