"""Drop-in for ``torch.nn.CrossEntropyLoss(ignore_index=...)`` as the reference's training step uses it
(command_handlers/train_or_resume.py:116 ``t.nn.CrossEntropyLoss(ignore_index=dataset['settings'].IGNORE_CLASS_LABEL)``,
applied at :435 ``loss_funcs[0](SSSR_output, target.long())``).

One pass over the (B,C,H,W) logits in forward and one in backward (``csrc/ce_loss.cu``) instead of torch's
log_softmax + nll_loss pair and their two backward kernels.  ``target`` may stay the uint8 map the dataset delivers
(the ``.long()`` copy is not needed) or be int32 / int64.  Only what the reference uses is supported: no class
weights, no label smoothing, class-index targets, reduction 'mean' or 'sum'; anything else raises.  No CPU fallback.
A target that is neither ``ignore_index`` nor a class index (torch: device assert) makes the loss NaN.
"""
from __future__ import annotations

import ctypes

import torch

from ... import _lib

_RED = {"mean": _lib.REDUCE_MEAN, "sum": _lib.REDUCE_SUM}
_DT = {torch.uint8: _lib.U8, torch.int32: _lib.I32, torch.int64: _lib.I64}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class _CEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index, reduction):
        x = logits.contiguous()
        tg = target.contiguous()
        B, C = int(x.shape[0]), int(x.shape[1])
        HW = int(x.numel() // max(B * C, 1))
        L = _lib.lib()
        saved_bytes = int(L.dsrl_ce_saved_bytes(B, HW))
        saved = torch.empty(saved_bytes, dtype=torch.uint8, device=x.device)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        st = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device), _lib.nvtx_range("dsrl.ce_forward"):
            _lib.check(L.dsrl_ce_forward(_ptr(x), _ptr(tg), _DT[tg.dtype], B, C, HW, ignore_index, reduction, _ptr(loss), _ptr(saved),
                                         saved_bytes, st))
        ctx.save_for_backward(x, tg, saved)
        ctx.geom = (B, C, HW, ignore_index, reduction, saved_bytes)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x, tg, saved = ctx.saved_tensors
        B, C, HW, ignore_index, reduction, saved_bytes = ctx.geom
        go = grad_out.to(dtype=torch.float32).contiguous()
        dx = torch.empty_like(x)
        st = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device), _lib.nvtx_range("dsrl.ce_backward"):
            _lib.check(_lib.lib().dsrl_ce_backward(_ptr(x), _ptr(tg), _DT[tg.dtype], B, C, HW, ignore_index, reduction, _ptr(saved),
                                                   saved_bytes, _ptr(go), _ptr(dx), st))
        return dx, None, None, None


class CrossEntropyLoss(torch.nn.modules.loss._WeightedLoss):
    __constants__ = ['ignore_index', 'reduction', 'label_smoothing']

    def __init__(self, weight=None, size_average=None, ignore_index: int = -100, reduce=None, reduction: str = 'mean',
                 label_smoothing: float = 0.0) -> None:
        super().__init__(weight, size_average, reduce, reduction)
        if weight is not None or label_smoothing != 0.0:
            raise NotImplementedError("CrossEntropyLoss (dsrl-b200): class weights / label smoothing are not used by the reference and not supported")
        self.ignore_index = int(ignore_index)
        self.label_smoothing = 0.0

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.reduction not in _RED:
            raise NotImplementedError("CrossEntropyLoss (dsrl-b200): reduction must be 'mean' or 'sum'")
        if input.dim() < 2 or target.dim() != input.dim() - 1 or input.shape[0] != target.shape[0] or tuple(input.shape[2:]) != tuple(target.shape[1:]):
            raise ValueError(f"CrossEntropyLoss: expected input (B,C,...) and class-index target (B,...), got {tuple(input.shape)} and {tuple(target.shape)}")
        if not input.is_cuda or not target.is_cuda:
            raise RuntimeError("CrossEntropyLoss (dsrl-b200) runs on CUDA tensors only: there is no CPU fallback")
        if input.dtype != torch.float32:
            raise TypeError("CrossEntropyLoss (dsrl-b200): logits must be float32")
        if target.dtype not in _DT:
            raise TypeError("CrossEntropyLoss (dsrl-b200): target must be uint8, int32 or int64 class indices")
        return _CEFunction.apply(input, target, self.ignore_index, _RED[self.reduction])
