"""ctypes binding of libdsrl_b200.so (the C-ABI declared in include/dsrl_b200.h).

There is deliberately no fallback: if the library is missing, or a compute entry point is called without a
B200-class device, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSRL_B200_LIB") or os.path.join(PKG_DIR, "libdsrl_b200.so")

OK, ERR_BAD_SHAPE, ERR_BAD_DTYPE, ERR_UNSUPPORTED, ERR_CUDA, ERR_BAD_ARG = 0, -1, -2, -3, -4, -5
FA_REFERENCE, FA_POSITION = 0, 1
REDUCE_NONE, REDUCE_MEAN, REDUCE_SUM = 0, 1, 2
PREC_FP32, PREC_TF32, PREC_BF16, PREC_F16 = 0, 1, 2, 3
PREC_EXACT_SIGNS = 16      # flag OR-ed onto a position-mode precision
U8, I32, I64 = 0, 1, 2

# every symbol include/dsrl_b200.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = (
    "dsrl_version", "dsrl_last_error", "dsrl_launch_count",
    "dsrl_fa_saved_bytes", "dsrl_fa_workspace_bytes", "dsrl_fa_forward", "dsrl_fa_backward", "dsrl_fa_forward_backward",
    "dsrl_fa_sign_stats", "dsrl_scale_grads",
    "dsrl_seg_counts", "dsrl_seg_counts_from_logits",
    "dsrl_ce_saved_bytes", "dsrl_ce_forward", "dsrl_ce_backward",
    "dsrl_ce_forward_tap", "dsrl_tap_dw_blocks", "dsrl_ce_backward_tap",
    "dsrl_mse_workspace_bytes", "dsrl_mse_forward", "dsrl_mse_backward",
    "dsrl_ft_bn_forward", "dsrl_ft_bn_backward", "dsrl_fa_forward_backward_transformed",
)


class DsrlError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libdsrl_b200 error {code}: {msg}")
        self.code = code


_lock = threading.Lock()
_lib = None


def _declare(lib):
    c = ctypes
    vp, i, sz, i64 = c.c_void_p, c.c_int, c.c_size_t, c.c_int64
    lib.dsrl_version.restype = i
    lib.dsrl_version.argtypes = []
    lib.dsrl_last_error.restype = c.c_char_p
    lib.dsrl_last_error.argtypes = []
    lib.dsrl_launch_count.restype = c.c_uint64
    lib.dsrl_launch_count.argtypes = []
    lib.dsrl_fa_saved_bytes.restype = sz
    lib.dsrl_fa_saved_bytes.argtypes = [i] * 8
    lib.dsrl_fa_workspace_bytes.restype = sz
    lib.dsrl_fa_workspace_bytes.argtypes = [i] * 8
    lib.dsrl_fa_sign_stats.restype = i
    lib.dsrl_fa_sign_stats.argtypes = [vp, c.POINTER(c.c_uint64), vp]
    lib.dsrl_scale_grads.restype = i
    lib.dsrl_scale_grads.argtypes = [vp, vp, i64, vp, i64, vp]
    lib.dsrl_fa_forward.restype = i
    lib.dsrl_fa_forward.argtypes = [i, i, vp, vp, i, i, i, i, i, i, i, i, vp, vp, sz, vp, sz, vp]
    lib.dsrl_fa_backward.restype = i
    lib.dsrl_fa_backward.argtypes = [i, i, vp, vp, vp, sz, vp, vp, vp, i, i, i, i, i, i, i, vp, sz, vp]
    lib.dsrl_fa_forward_backward.restype = i
    lib.dsrl_fa_forward_backward.argtypes = [i, i, vp, vp, i, i, i, i, i, i, i, vp, vp, vp, vp, vp, sz, vp, sz, vp]
    lib.dsrl_seg_counts.restype = i
    lib.dsrl_seg_counts.argtypes = [vp, i, vp, i, vp, i64, i64, i, i, vp, vp]
    lib.dsrl_seg_counts_from_logits.restype = i
    lib.dsrl_seg_counts_from_logits.argtypes = [vp, vp, i, vp, i64, i64, i64, i, i, vp, vp, vp]
    lib.dsrl_ce_saved_bytes.restype = sz
    lib.dsrl_ce_saved_bytes.argtypes = [i, i64]
    lib.dsrl_ce_forward.restype = i
    lib.dsrl_ce_forward.argtypes = [vp, vp, i, i, i, i64, i64, i, vp, vp, sz, vp]
    lib.dsrl_ce_backward.restype = i
    lib.dsrl_ce_backward.argtypes = [vp, vp, i, i, i, i64, i64, i, vp, sz, vp, vp, vp]
    f = c.c_float
    lib.dsrl_ce_forward_tap.restype = i
    lib.dsrl_ce_forward_tap.argtypes = [vp, vp, i, i, i, i, i, i64, i, vp, vp, sz, vp, vp, i, vp]
    lib.dsrl_tap_dw_blocks.restype = i64
    lib.dsrl_tap_dw_blocks.argtypes = [i, i, i, i]
    lib.dsrl_ce_backward_tap.restype = i
    lib.dsrl_ce_backward_tap.argtypes = [vp, vp, i, i, i, i, i, i64, i, vp, sz, vp, vp, vp, vp, vp, i, c.POINTER(c.c_int64), vp]
    lib.dsrl_mse_workspace_bytes.restype = sz
    lib.dsrl_mse_workspace_bytes.argtypes = [i, i, i, i]
    lib.dsrl_mse_forward.restype = i
    lib.dsrl_mse_forward.argtypes = [vp, vp, i, i, i, i, vp, vp, sz, vp, vp, i, vp]
    lib.dsrl_mse_backward.restype = i
    lib.dsrl_mse_backward.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp, vp, i, vp]
    lib.dsrl_ft_bn_forward.restype = i
    lib.dsrl_ft_bn_forward.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, f, f, i, vp, vp]
    lib.dsrl_ft_bn_backward.restype = i
    lib.dsrl_ft_bn_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, i, vp]
    lib.dsrl_fa_forward_backward_transformed.restype = i
    lib.dsrl_fa_forward_backward_transformed.argtypes = [vp, vp, vp, i, i, i, i, i, vp, vp, vp, vp, vp, sz, vp, sz, vp]


def lib():
    """The loaded library; raises if it has not been built (run `python -m dualsuperreslearningforsemseg_b200.build`)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m dualsuperreslearningforsemseg_b200.build` "
                        "(needs nvcc).  There is no CPU or PyTorch fallback for this path.")
                handle = ctypes.CDLL(LIB_PATH)
                _declare(handle)
                _lib = handle
    return _lib


def check(rc: int):
    if rc != OK:
        raise DsrlError(rc, lib().dsrl_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().dsrl_launch_count())


# ---- tracing (the reference wraps its commands in torch.autograd.profiler; here NVTX ranges around the C-ABI calls) ----
NVTX = bool(int(os.environ.get("DSRL_NVTX", "0")))


class nvtx_range:
    """``with nvtx_range("dsrl.fa_forward"):`` -- an NVTX range when DSRL_NVTX=1 (for ``ncu --nvtx`` / timeline tools), free otherwise."""

    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_pop()
