"""Low-level, allocation-free access to the FA kernels (what a CUDA-graph-captured training step would call).

`FAPlan` fixes the geometry once, owns every buffer (saved blob, workspace, loss, gradients) and issues exactly the
library's kernel launches per call -- no autograd bookkeeping.  `FALoss` (models/losses/FALoss.py) remains the
drop-in surface; this is the same C-ABI underneath.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_RED = {"none": _lib.REDUCE_NONE, "mean": _lib.REDUCE_MEAN, "sum": _lib.REDUCE_SUM}
_MODE = {"reference": _lib.FA_REFERENCE, "position": _lib.FA_POSITION}
_PREC = {None: _lib.PREC_TF32, "fp32": _lib.PREC_FP32, "tf32": _lib.PREC_TF32, "f16": _lib.PREC_F16}


class FAPlan:
    def __init__(self, shape1, shape2=None, subsample_factor=8, reduction="mean", affinity="reference", precision=None,
                 device=None):
        shape2 = tuple(shape2 or shape1)
        self.B, self.C1, self.H, self.W = (int(v) for v in shape1)
        self.C2 = int(shape2[1])
        self.k = int(subsample_factor)
        self.mode, self.red, self.prec = _MODE[affinity], _RED[reduction], _PREC[precision]
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        L = _lib.lib()
        geom = (self.mode, self.B, self.C1, self.C2, self.H, self.W, self.k)
        self.saved_bytes = int(L.dsrl_fa_saved_bytes(*geom))
        self.ws_bytes = int(L.dsrl_fa_workspace_bytes(*geom))
        if self.saved_bytes == 0:
            raise _lib.DsrlError(_lib.ERR_UNSUPPORTED, f"FAPlan: unsupported geometry {geom}")
        self.saved = torch.empty(self.saved_bytes, dtype=torch.uint8, device=dev)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        if self.red == _lib.REDUCE_NONE:
            n = (self.W // self.k) ** 2
            self.loss = torch.empty((self.B, self.C1, n * n), dtype=torch.float32, device=dev)
        else:
            self.loss = torch.empty((), dtype=torch.float32, device=dev)
        self.dx1 = torch.empty((self.B, self.C1, self.H, self.W), dtype=torch.float32, device=dev)
        self.dx2 = torch.empty((self.B, self.C2, self.H, self.W), dtype=torch.float32, device=dev)
        self._p = lambda t: ctypes.c_void_p(t.data_ptr())

    def forward(self, x1, x2, need_grad=True):
        st = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(_lib.lib().dsrl_fa_forward(self.mode, self.prec, self._p(x1), self._p(x2), self.B, self.C1, self.C2, self.H,
                                              self.W, self.k, self.red, int(need_grad), self._p(self.loss), self._p(self.saved),
                                              self.saved_bytes, self._p(self.ws), self.ws_bytes, st))
        return self.loss

    def backward(self, x1, x2, grad_out):
        st = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(_lib.lib().dsrl_fa_backward(self.mode, self.prec, self._p(x1), self._p(x2), self._p(self.saved), self.saved_bytes,
                                               self._p(grad_out), self._p(self.dx1), self._p(self.dx2), self.B, self.C1, self.C2,
                                               self.H, self.W, self.k, self.red, self._p(self.ws), self.ws_bytes, st))
        return self.dx1, self.dx2

    def forward_backward(self, x1, x2, grad_out):
        """x1, x2: contiguous fp32 CUDA tensors of the planned shapes; grad_out: fp32 CUDA tensor (1 element for
        mean/sum).  Returns (loss, dx1, dx2) -- the plan's own buffers, overwritten by the next call.  For mean/sum this
        is ``dsrl_fa_forward_backward`` (one kernel launch at the reference model's training shapes)."""
        if self.red == _lib.REDUCE_NONE:
            self.forward(x1, x2, True)
            self.backward(x1, x2, grad_out)
            return self.loss, self.dx1, self.dx2
        st = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(_lib.lib().dsrl_fa_forward_backward(self.mode, self.prec, self._p(x1), self._p(x2), self.B, self.C1, self.C2,
                                                       self.H, self.W, self.k, self.red, self._p(grad_out), self._p(self.loss),
                                                       self._p(self.dx1), self._p(self.dx2), self._p(self.saved), self.saved_bytes,
                                                       self._p(self.ws), self.ws_bytes, st))
        return self.loss, self.dx1, self.dx2
