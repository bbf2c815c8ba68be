"""Feature-Affinity loss -- B200 drop-in for the reference's ``models.losses.FALoss``.

Mirrors ``models/losses/FALoss.py:5-34`` of the reference: same class name, base class
(``torch.nn.modules.loss._Loss``), constructor ``(subsample_factor=8, size_average=None, reduce=None,
reduction='mean')`` (the reference ignores ``size_average``/``reduce`` and so do we, FALoss.py:15), same
``forward(feature_map1, feature_map2)`` contract (4-D, equal shapes, FALoss.py:19-20) and the same result:
a 0-dim tensor for ``mean``/``sum``, ``(B, C, n*n)`` for ``none`` (n = (W // k)**2), differentiable w.r.t. both
inputs.  Used at ``command_handlers/train_or_resume.py:118,437,444`` as ``w2 * FALoss()(a, b)`` inside
``(CE + MSE + FA).backward()``.

All arithmetic runs in hand-written sm_100a kernels behind the C-ABI of ``libdsrl_b200.so``
(``include/dsrl_b200.h``); this file only owns tensors, the autograd node and error translation.  There is no
PyTorch/CPU fallback: CPU tensors raise.

Keyword-only extensions (defaults preserve the reference behaviour):
  affinity='reference' | 'position'   'position' = the paper's N x N position affinity (not in the reference; the two
                                      inputs may then differ in C; reduction 'mean' or 'sum')
  precision=None | 'tf32' | 'f16' | 'fp32'
                                      tensor-core arithmetic of 'position': 'tf32' (default) = one tcgen05 kind::tf32 pass
                                      with FP32 accumulation; 'f16' = FP16 operands (the normalised features are unit
                                      vectors, FP16 keeps the 11-bit significand of TF32), kind::f16, FP32 accumulation:
                                      the accuracy of 'tf32' at about half the time; 'fp32' = 3xTF32 split (hi*hi +
                                      hi*lo + lo*hi), about FP32 accuracy at ~2x the time of 'tf32'.  The reference
                                      semantics always runs in FP32 FMA.
  exact_signs=True | False            'position' only.  The gradient is a sum of sign(S1 - S2) terms and tensor-core operand
                                      rounding flips the ~1e-4 of them that lie within its error of zero (0.1-1 % relative
                                      -norm on the gradient of densely distributed inputs).  True (default) re-decides every
                                      entry within ~3.5 sigma of that error from the unrounded features -- FP32 with a
                                      rigorous error bound, FP64 below it -- in a second, memory-bound kernel over ~1e-3 of
                                      the entries: gradients then meet 1e-3 relative-norm against the float64 oracle on any
                                      input (about +25 % time at C = 256).  ``sign_stats()`` reports what it found.
"""
from __future__ import annotations

import ctypes
import os

import torch

from ... import _lib

_RED = {"none": _lib.REDUCE_NONE, "mean": _lib.REDUCE_MEAN, "sum": _lib.REDUCE_SUM}
_PREC = {None: _lib.PREC_TF32, "fp32": _lib.PREC_FP32, "tf32": _lib.PREC_TF32, "f16": _lib.PREC_F16}
_MODE = {"reference": _lib.FA_REFERENCE, "position": _lib.FA_POSITION}

_size_cache = {}


def _layout_hooks():
    """Test / tuning hooks of the library that change the workspace layout (part of every size-cache key)."""
    return tuple(os.environ.get(k) for k in ("DSRL_POS_JSPLIT", "DSRL_POS_AB", "DSRL_POS_ACHUNK", "DSRL_POS_PAIR"))


def _sizes(mode, precision, B, C1, C2, H, W, k):
    geom = (mode, precision, B, C1, C2, H, W, k)
    key = geom + _layout_hooks()
    v = _size_cache.get(key)
    if v is None:
        L = _lib.lib()
        v = (int(L.dsrl_fa_saved_bytes(*geom)), int(L.dsrl_fa_workspace_bytes(*geom)))
        _size_cache[key] = v
    return v


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _OnDevice:
    """``torch.cuda.device(dev)`` only when ``dev`` is not already current (the context manager costs ~10 us per call,
    a third of this wrapper's host time at the training shape)."""

    __slots__ = ("ctx",)

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


class _Scratch:
    """Per-module, per-geometry device buffers that only live for the duration of one call on one stream: the workspace, the
    `saved` blob when the call leaves nothing in it for backward, and a device 1.0.  Reused across calls (work on one stream
    is ordered; a module shared by concurrent streams needs one instance per stream, like any stateful cuDNN-style plan)."""

    __slots__ = ("ws", "saved", "one", "saved_bytes", "ws_bytes")

    def __init__(self, dev, saved_bytes, ws_bytes):
        self.saved_bytes, self.ws_bytes = saved_bytes, ws_bytes
        self.ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        self.saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev)
        self.one = torch.ones((), dtype=torch.float32, device=dev)


class _FAFunction(torch.autograd.Function):
    """Forward that needs gradients runs ``dsrl_fa_forward_backward`` with a unit upstream gradient: ONE pass that leaves the
    loss and dX1, dX2 (reference mode at the training shapes: one kernel launch; position mode without pooling: the gradient
    kernel writes dX itself).  ``backward`` then only applies the upstream gradient in place (``dsrl_scale_grads``, which
    returns at once on the device when it is 1).  ``reduction='none'`` keeps the two-call form (its upstream gradient is a
    tensor the pair pass needs)."""

    @staticmethod
    def forward(ctx, x1, x2, k, reduction, mode, precision, scratch_of):
        B, C1, H, W = x1.shape
        C2 = x2.shape[1]
        x1c, x2c = x1.contiguous(), x2.contiguous()
        saved_bytes, ws_bytes = _sizes(mode, precision, B, C1, C2, H, W, k)
        if saved_bytes == 0:
            raise _lib.DsrlError(_lib.ERR_UNSUPPORTED,
                                 f"FALoss: unsupported geometry B={B} C=({C1},{C2}) H={H} W={W} k={k}")
        dev = x1.device
        sc = scratch_of(dev, (mode, precision, B, C1, C2, H, W, k), saved_bytes, ws_bytes)
        need_grad = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        L = _lib.lib()
        ctx.fused = need_grad and reduction != _lib.REDUCE_NONE
        if ctx.fused:
            out = torch.empty((), dtype=torch.float32, device=dev)
            dx1, dx2 = torch.empty_like(x1c), torch.empty_like(x2c)
            with _OnDevice(dev), _lib.nvtx_range("dsrl.fa_forward_backward"):
                _lib.check(L.dsrl_fa_forward_backward(mode, precision, _ptr(x1c), _ptr(x2c), B, C1, C2, H, W, k, reduction,
                                                      _ptr(sc.one), _ptr(out), _ptr(dx1), _ptr(dx2), _ptr(sc.saved), saved_bytes,
                                                      _ptr(sc.ws), ws_bytes, stream))
            ctx.dx = (dx1, dx2)
            ctx.applied = None          # upstream gradient already multiplied into ctx.dx (None: 1)
            return out
        if reduction == _lib.REDUCE_NONE:
            n = (W // k) ** 2
            out = torch.empty((B, C1, n * n), dtype=torch.float32, device=dev)
            saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev) if need_grad else sc.saved
        else:
            out = torch.empty((), dtype=torch.float32, device=dev)
            saved = sc.saved
        with _OnDevice(dev), _lib.nvtx_range("dsrl.fa_forward"):
            _lib.check(L.dsrl_fa_forward(mode, precision, _ptr(x1c), _ptr(x2c), B, C1, C2, H, W, k, reduction,
                                         int(need_grad), _ptr(out), _ptr(saved), saved_bytes, _ptr(sc.ws), ws_bytes, stream))
        if need_grad:
            ctx.geom = (B, C1, C2, H, W, k, reduction, mode, precision, saved_bytes, ws_bytes)
            ctx.scratch = sc
            ctx.save_for_backward(saved)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        L = _lib.lib()
        if ctx.fused:
            dx1, dx2 = ctx.dx
            dev = dx1.device
            go = grad_out.to(torch.float32).contiguous()
            if ctx.applied is not None:
                # backward() again through a retained graph: the buffers already carry the first upstream gradient (and may
                # have become the inputs' .grad): rescale out of place
                r = go / ctx.applied
                return (dx1 * r if ctx.needs_input_grad[0] else None), (dx2 * r if ctx.needs_input_grad[1] else None), None, None, None, None, None
            ctx.applied = go
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            with _OnDevice(dev), _lib.nvtx_range("dsrl.fa_backward"):
                _lib.check(L.dsrl_scale_grads(_ptr(go), _ptr(dx1), dx1.numel(), _ptr(dx2), dx2.numel(), stream))
            return (dx1 if ctx.needs_input_grad[0] else None), (dx2 if ctx.needs_input_grad[1] else None), None, None, None, None, None
        B, C1, C2, H, W, k, reduction, mode, precision, saved_bytes, ws_bytes = ctx.geom
        saved, = ctx.saved_tensors
        dev = saved.device
        go = grad_out.to(torch.float32).contiguous()
        dx1 = torch.empty((B, C1, H, W), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        dx2 = torch.empty((B, C2, H, W), dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        sc = ctx.scratch
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with _OnDevice(dev), _lib.nvtx_range("dsrl.fa_backward"):
            _lib.check(L.dsrl_fa_backward(mode, precision, None, None, _ptr(saved), saved_bytes,
                                          _ptr(go), _ptr(dx1), _ptr(dx2), B, C1, C2, H, W, k, reduction,
                                          _ptr(sc.ws), ws_bytes, stream))
        return dx1, dx2, None, None, None, None, None


class FALoss(torch.nn.modules.loss._Loss):
    __constants__ = ['reduction']

    def __init__(self, subsample_factor: int = 8, size_average=None, reduce=None, reduction: str = 'mean', *,
                 affinity: str = 'reference', precision=None, exact_signs: bool = True) -> None:
        # the reference passes None for size_average/reduce whatever the caller gave (FALoss.py:15)
        super().__init__(size_average=None, reduce=None, reduction=reduction)
        if affinity not in _MODE:
            raise ValueError(f"affinity must be 'reference' or 'position', got {affinity!r}")
        if precision not in _PREC:
            raise ValueError(f"precision must be one of {sorted(p for p in _PREC if p)}, got {precision!r}")
        self.subsample_factor = subsample_factor
        self.affinity = affinity
        self.precision = precision
        self.exact_signs = bool(exact_signs)
        self._scratch = {}
        self._last = None

    def extra_repr(self) -> str:
        return f"subsample_factor={self.subsample_factor}, reduction={self.reduction!r}, affinity={self.affinity!r}"

    def _scratch_of(self, dev, geom, saved_bytes, ws_bytes):
        key = (dev, geom) + _layout_hooks()
        sc = self._scratch.get(key)
        if sc is None or sc.saved_bytes != saved_bytes or sc.ws_bytes != ws_bytes:
            if len(self._scratch) >= 4:                       # a loss module sees one or two geometries; do not hoard
                self._scratch.clear()
            sc = self._scratch[key] = _Scratch(dev, saved_bytes, ws_bytes)
        self._last = sc
        return sc

    def sign_stats(self):
        """``exact_signs=True``: what the last forward (one that needed gradients) found -- ``{'listed', 'corrected', 'dropped',
        'worst_ratio'}``: near ties re-decided in FP64, signs that changed, ties beyond a row's list capacity (they keep
        the tensor-core sign) and the largest |D_exact| / threshold among the corrected entries.  Synchronises."""
        if self._last is None or not (self.exact_signs and self.affinity == 'position'):
            return None
        out = (ctypes.c_uint64 * 4)()
        dev = self._last.saved.device
        with _OnDevice(dev):
            _lib.check(_lib.lib().dsrl_fa_sign_stats(_ptr(self._last.saved), out,
                                                     ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return {"listed": int(out[0]), "corrected": int(out[1]), "dropped": int(out[2]), "worst_ratio": out[3] * 1e-6}

    def forward(self, feature_map1: torch.Tensor, feature_map2: torch.Tensor) -> torch.Tensor:
        # same BUG CHECKs (and wording) as FALoss.py:19-20
        torch._assert(feature_map1.dim() == 4 and feature_map2.dim() == 4,
                      "BUG CHECK: Feature map inputs to FALoss.forward() must have 4 dimensions (B, C, H, W).")
        if self.affinity == 'reference':
            torch._assert(feature_map1.shape == feature_map2.shape,
                          "BUG CHECK: Feature map inputs to FALoss.forward() should be of same size.")
        else:
            torch._assert(feature_map1.shape[0] == feature_map2.shape[0] and feature_map1.shape[2:] == feature_map2.shape[2:],
                          "BUG CHECK: Feature map inputs to FALoss.forward() must agree in B, H and W.")
        if self.reduction not in _RED:
            raise ValueError(f"{self.reduction} is not a valid value for reduction")
        if self.affinity == 'position' and self.reduction == 'none':
            raise _lib.DsrlError(_lib.ERR_UNSUPPORTED, "FALoss(affinity='position'): reduction='none' would materialise "
                                                       "the N x N affinity; use 'mean' or 'sum'")
        if not (feature_map1.is_cuda and feature_map2.is_cuda):
            raise RuntimeError("FALoss (dsrl-b200) runs only on CUDA tensors: there is no CPU fallback on this path")
        if feature_map1.device != feature_map2.device:
            raise RuntimeError("FALoss inputs must live on the same device")
        if feature_map1.dtype != torch.float32 or feature_map2.dtype != torch.float32:
            # the reference itself rejects half/bfloat16 here ("Low precision dtypes not supported", linalg.norm)
            raise RuntimeError("FALoss (dsrl-b200) supports float32 feature maps only")
        k = int(self.subsample_factor)
        if feature_map1.shape[2] // k < 1 or feature_map1.shape[3] // k < 1:
            raise RuntimeError("FALoss: feature map smaller than the pooling window")
        prec = _PREC[self.precision]
        if self.exact_signs and self.affinity == 'position':
            prec |= _lib.PREC_EXACT_SIGNS
        return _FAFunction.apply(feature_map1, feature_map2, k, _RED[self.reduction], _MODE[self.affinity], prec,
                                 self._scratch_of)
