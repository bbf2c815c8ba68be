"""CPU oracle for the mIoU / mean-accuracy counts -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product path never does.

NumPy restatement of ``metrices/mIoU.py:21-35,37-41`` and ``metrices/Accuracy.py:19-24,26-30`` as exact
integer set counts followed by the same float64 finishing arithmetic.  It does
not call ``np.histogram``: the reference's ``np.histogram(x, bins=NC, range=(1, NC))`` on ``x + 1`` is an
exact bincount of the classes 0..NC-1 (the last bin is right-inclusive, everything outside is dropped,
``uint8`` 255+1 wraps to 0 and is dropped too), which ``tests/test_oracle_seg.py`` pins against the golden
vectors generated from the unmodified reference (``tests/golden/seg_golden.npz``) and against the
reference's own fixture ``scratchpad.py:361-363`` (mIoU 66.66666666666666, accuracy 77.77777777777779).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import warnings

import numpy as np

__all__ = ["seg_counts", "iou_from_counts", "accuracy_from_counts", "MIoUOracle", "AccuracyOracle",
           "argmax_first", "build_c", "seg_counts_c"]


def seg_counts(pred, target, mask, num_classes: int):
    """Exact counts of one ``update(pred, target, mask)`` call.

    Returns int64 ``(area_pred[NC], area_inter[NC], area_target[NC], correct, valid)``:
      area_pred[c]   = #{mask & pred == c}                (mIoU.py:21,24,27)
      area_inter[c]  = #{mask & pred == c & target == c}  (mIoU.py:25,28)
      area_target[c] = #{target == c}  -- NOT masked      (mIoU.py:22,29)
      correct        = #{mask & pred == target}  on the raw values (Accuracy.py:19)
      valid          = #{mask}                            (Accuracy.py:20)
    """
    pred = np.asarray(pred)
    target = np.asarray(target)
    m = np.asarray(mask).astype(bool)
    assert pred.shape == target.shape == m.shape
    p = pred.astype(np.int64).ravel()
    t = target.astype(np.int64).ravel()
    m = m.ravel()
    nc = int(num_classes)
    p_ok = m & (p >= 0) & (p < nc)
    t_ok = (t >= 0) & (t < nc)
    area_pred = np.bincount(p[p_ok], minlength=nc).astype(np.int64)
    area_inter = np.bincount(p[p_ok & (p == t)], minlength=nc).astype(np.int64)
    area_target = np.bincount(t[t_ok], minlength=nc).astype(np.int64)
    correct = np.int64(np.count_nonzero(m & (p == t)))
    valid = np.int64(np.count_nonzero(m))
    return area_pred, area_inter, area_target, correct, valid


def iou_from_counts(area_pred, area_inter, area_target):
    """mIoU.py:30-35: union, BUG CHECK, nan-mean over classes (float64)."""
    area_union = area_pred + area_target - area_inter
    assert (area_inter <= area_union).all()
    with np.errstate(divide="ignore", invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmean(area_inter / area_union)


def accuracy_from_counts(correct, valid):
    """Accuracy.py:22-24: int64 / int64 -> float64 (0/0 -> NaN)."""
    assert correct <= valid
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.int64(correct) / np.int64(valid)


class MIoUOracle:
    """Stateful twin of the reference ``mIoU`` (mIoU.py:5-41) built on :func:`seg_counts`."""

    def __init__(self, num_classes):
        self.num_classes = num_classes
        self.ious = []

    def update(self, pred, target, mask):
        ap, ai, at, _, _ = seg_counts(pred, target, mask, self.num_classes)
        self.ious.append(iou_from_counts(ap, ai, at))

    def __call__(self):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            return np.nanmean(self.ious) * 100.0


class AccuracyOracle:
    """Stateful twin of the reference ``Accuracy`` (Accuracy.py:4-30)."""

    def __init__(self):
        self.accuracies = []

    def update(self, pred, target, mask):
        _, _, _, c, v = seg_counts(pred, target, mask, 1)
        self.accuracies.append(accuracy_from_counts(c, v))

    def __call__(self):
        return np.mean(self.accuracies) * 100.0


def argmax_first(logits: np.ndarray) -> np.ndarray:
    """``np.argmax(logits, axis=1)`` (benchmark.py:69) / ``t.argmax(.., dim=1)`` (train_or_resume.py:477):
    first index of the maximum along the class axis; a NaN counts as the maximum (first NaN wins)."""
    return np.argmax(np.asarray(logits), axis=1).astype(np.int64)


# --------------------------------------------------------------------------------------------------
# C restatement (single pass, plain C) -- used as the multi-threadable CPU baseline in bench.py
# --------------------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_C_SRC = os.path.join(_HERE, "seg_counts_ref.c")
_C_LIB = os.path.join(_HERE, "libseg_counts_ref.so")
_c_handle = None


def build_c(force: bool = False) -> str:
    """Compile ``oracle/seg_counts_ref.c`` with gcc (OpenMP if available)."""
    if force or not os.path.exists(_C_LIB) or os.path.getmtime(_C_LIB) < os.path.getmtime(_C_SRC):
        cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", _C_SRC, "-o", _C_LIB]
        subprocess.run(cmd, check=True)
    return _C_LIB


def seg_counts_c(pred, target, mask, num_classes: int, threads: int = 1):
    """Same contract as :func:`seg_counts` for pred int64 / target uint8 / mask bool, via the C restatement."""
    global _c_handle
    if _c_handle is None:
        _c_handle = ctypes.CDLL(build_c())
        _c_handle.seg_counts_ref.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        _c_handle.seg_counts_ref.restype = None
    pred = np.ascontiguousarray(pred, dtype=np.int64)
    target = np.ascontiguousarray(target, dtype=np.uint8)
    m = np.ascontiguousarray(mask).astype(np.uint8, copy=False)
    out = np.zeros(3 * num_classes + 2, dtype=np.int64)
    _c_handle.seg_counts_ref(pred.ctypes.data, target.ctypes.data, m.ctypes.data, pred.size,
                             int(num_classes), out.ctypes.data, int(threads))
    nc = num_classes
    return out[:nc], out[nc:2 * nc], out[2 * nc:3 * nc], out[3 * nc], out[3 * nc + 1]
