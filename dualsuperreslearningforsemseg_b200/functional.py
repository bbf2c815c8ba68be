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
                 device=None, exact_signs=True):
        shape2 = tuple(shape2 or shape1)
        self.B, self.C1, self.H, self.W = (int(v) for v in shape1)
        self.C2 = int(shape2[1])
        self.k = int(subsample_factor)
        self.mode, self.red, self.prec = _MODE[affinity], _RED[reduction], _PREC[precision]
        if exact_signs and affinity == "position":
            self.prec |= _lib.PREC_EXACT_SIGNS
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        L = _lib.lib()
        geom = (self.mode, self.prec, self.B, self.C1, self.C2, self.H, self.W, self.k)
        self.saved_bytes = int(L.dsrl_fa_saved_bytes(*geom))
        self.ws_bytes = int(L.dsrl_fa_workspace_bytes(*geom))
        if self.saved_bytes == 0:
            raise _lib.DsrlError(_lib.ERR_UNSUPPORTED, f"FAPlan: unsupported geometry {geom}")
        self.saved = torch.empty(self.saved_bytes, dtype=torch.uint8, device=dev)
        self.ws = torch.empty(max(self.ws_bytes, 16), dtype=torch.uint8, device=dev)
        if self.red == _lib.REDUCE_NONE:
            n = (self.W // self.k) ** 2
            self.loss = torch.empty((self.B, self.C1, n * n), dtype=torch.float32, device=dev)
        else:
            self.loss = torch.empty((), dtype=torch.float32, device=dev)
        self.dx1 = torch.empty((self.B, self.C1, self.H, self.W), dtype=torch.float32, device=dev)
        self.dx2 = torch.empty((self.B, self.C2, self.H, self.W), dtype=torch.float32, device=dev)
        self._p = lambda t: ctypes.c_void_p(t.data_ptr())

    def sign_stats(self):
        """exact_signs: {'listed', 'corrected', 'dropped', 'worst_ratio'} of the last forward that produced gradients (synchronises)."""
        out = (ctypes.c_uint64 * 4)()
        st = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(_lib.lib().dsrl_fa_sign_stats(self._p(self.saved), out, st))
        return {"listed": int(out[0]), "corrected": int(out[1]), "dropped": int(out[2]), "worst_ratio": out[3] * 1e-6}

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

    def forward_backward(self, x1, x2, grad_out, out=None):
        """x1, x2: contiguous fp32 CUDA tensors of the planned shapes; grad_out: fp32 CUDA tensor (1 element for
        mean/sum).  Returns (loss, dx1, dx2) -- the plan's own buffers, overwritten by the next call, or the tensors
        given as ``out=(loss, dx1, dx2)`` (mean/sum only).  For mean/sum this is ``dsrl_fa_forward_backward`` (one
        kernel launch at the reference model's training shapes; position mode without pooling writes dX from the gradient
        kernel, so the plan's `saved` blob is scratch afterwards: do not follow this call with `backward()`)."""
        if self.red == _lib.REDUCE_NONE:
            self.forward(x1, x2, True)
            self.backward(x1, x2, grad_out)
            return self.loss, self.dx1, self.dx2
        loss, dx1, dx2 = out if out is not None else (self.loss, self.dx1, self.dx2)
        st = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(_lib.lib().dsrl_fa_forward_backward(self.mode, self.prec, self._p(x1), self._p(x2), self.B, self.C1, self.C2,
                                                       self.H, self.W, self.k, self.red, self._p(grad_out), self._p(loss),
                                                       self._p(dx1), self._p(dx2), self._p(self.saved), self.saved_bytes,
                                                       self._p(self.ws), self.ws_bytes, st))
        return loss, dx1, dx2


def chunk_bounds(batch, chunk, ramp=True):
    """[lo, hi) sample ranges of the host pipeline's chunks.  With ``ramp`` the first two chunks are single samples, so the
    kernels start after one sample's copy instead of ``chunk``; the last chunk may be shorter."""
    chunk = max(1, min(int(chunk), int(batch)))
    sizes, left = [], int(batch)
    while left > 0:
        n = 1 if (ramp and len(sizes) < 2) else chunk
        sizes.append(min(n, left))
        left -= sizes[-1]
    starts = [sum(sizes[:i]) for i in range(len(sizes))]
    return [(lo, lo + n) for lo, n in zip(starts, sizes)]


class FAHostPipeline:
    """FA loss forward+backward for feature maps that live in pinned HOST memory.

    Samples are independent units of the loss, so the batch is cut into chunks of ``chunk`` samples: chunk
    i+1 travels host -> device on a copy stream while the kernels work on chunk i on the caller's stream (two streams,
    one event per chunk, no host synchronisation; with ``ramp`` the first two chunks are single samples).  ``__call__(x1_host, x2_host, grad_out=1.0)`` returns
    ``(loss, dx1, dx2)`` as device tensors owned by the pipeline: the batch loss ('mean' or 'sum') and the gradients
    w.r.t. both inputs -- the same values ``FALoss`` + ``backward()`` give on the whole batch."""

    def __init__(self, shape1, shape2=None, subsample_factor=8, reduction="mean", affinity="reference", precision=None,
                 chunk=2, ramp=True, device=None, exact_signs=True):
        if reduction not in ("mean", "sum"):
            raise ValueError("FAHostPipeline: reduction must be 'mean' or 'sum'")
        shape2 = tuple(shape2 or shape1)
        self.B = int(shape1[0])
        self.reduction = reduction
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.bounds = chunk_bounds(self.B, chunk, ramp)
        self.plans = {}
        for lo, hi in self.bounds:                          # one plan per distinct chunk size (the last chunk may be shorter)
            n = hi - lo
            if n not in self.plans:
                self.plans[n] = FAPlan((n,) + tuple(shape1[1:]), (n,) + tuple(shape2[1:]), subsample_factor, reduction,
                                       affinity, precision, dev, exact_signs)
        self.x1 = torch.empty(tuple(shape1), dtype=torch.float32, device=dev)
        self.x2 = torch.empty(tuple(shape2), dtype=torch.float32, device=dev)
        self.dx1 = torch.empty_like(self.x1)
        self.dx2 = torch.empty_like(self.x2)
        self.losses = torch.zeros(len(self.bounds), dtype=torch.float32, device=dev)
        # chunk means -> batch mean: weight n_chunk / B on the loss and on the upstream gradient; sums add up unweighted
        w = [((hi - lo) / self.B if reduction == "mean" else 1.0) for lo, hi in self.bounds]
        self.weights = torch.tensor(w, dtype=torch.float32, device=dev)
        self.go = torch.empty(len(self.bounds), dtype=torch.float32, device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.events = [torch.cuda.Event() for _ in self.bounds]

    def __call__(self, x1_host, x2_host, grad_out=1.0):
        if x1_host.is_cuda or x2_host.is_cuda or not (x1_host.is_pinned() and x2_host.is_pinned()):
            raise ValueError("FAHostPipeline: inputs must be pinned host tensors (use FALoss / FAPlan for device tensors)")
        cur = torch.cuda.current_stream(self.dev)
        torch.mul(self.weights, float(grad_out), out=self.go)
        self.copy_stream.wait_stream(cur)                   # the previous call's kernels are done with the staging buffers
        with torch.cuda.stream(self.copy_stream):
            for i, (lo, hi) in enumerate(self.bounds):
                self.x1[lo:hi].copy_(x1_host[lo:hi], non_blocking=True)
                self.x2[lo:hi].copy_(x2_host[lo:hi], non_blocking=True)
                self.events[i].record(self.copy_stream)
        for i, (lo, hi) in enumerate(self.bounds):
            cur.wait_event(self.events[i])
            self.plans[hi - lo].forward_backward(self.x1[lo:hi], self.x2[lo:hi], self.go[i:i + 1],
                                                 out=(self.losses[i:i + 1], self.dx1[lo:hi], self.dx2[lo:hi]))
        return torch.dot(self.losses, self.weights), self.dx1, self.dx2

    # ---- launch-latency-bound shapes: the whole step as one CUDA graph -------------------------------------------
    def capture(self, x1_host, x2_host, grad_out=1.0):
        """Captures ``self(x1_host, x2_host, grad_out)`` plus the read-back of the loss into ONE CUDA graph: the host -> device
        copies (from THESE pinned tensors -- refill them in place between replays), the kernels on both streams and the
        device -> host copy of the loss.  At the reference's training shape the step is a handful of microsecond-sized nodes,
        so one graph launch replaces ~10 launches / stream operations."""
        self(x1_host, x2_host, grad_out)                  # eager once: one-time allocations inside the library happen here
        torch.cuda.synchronize(self.dev)
        self._loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            loss, _, _ = self(x1_host, x2_host, grad_out)
            self._loss_host.copy_(loss, non_blocking=True)
        return self

    def replay(self):
        """One launch of the captured step; returns the loss as a pinned host scalar (valid when this returns)."""
        self._graph.replay()
        torch.cuda.current_stream(self.dev).synchronize()
        return self._loss_host

