"""The three losses of the reference's stage-3 training step in shared passes.

The reference computes (command_handlers/train_or_resume.py:435-438, models/DSRL.py:181,184)

    SSSR_t = SSSR_feature_transformer(SSSR_output)          # Conv2d(19 -> 1, k 1, stride 8, no bias) + BatchNorm2d(1) + ReLU
    SISR_t = SISR_feature_transformer(SISR_output)          # Conv2d( 3 -> 1, k 1, stride 8, no bias) + BatchNorm2d(1) + ReLU
    CE   = CrossEntropyLoss(ignore_index)(SSSR_output, target.long())
    MSE  = w1 * MSELoss()(SISR_output, input_org)
    FA   = w2 * FALoss()(SSSR_t, SISR_t)

as ~20 kernel launches forward and about as many backward.  ``Stage3Loss`` produces the same three numbers and the same
gradients (w.r.t. SSSR_output, SISR_output and the six transformer parameters) from

    forward   CE pass over the logits  (+ the strided 1x1 convolution of transformer 1 on the way)
              MSE pass over the image  (+ the convolution of transformer 2)
              one launch for both BatchNorms (batch statistics, running-statistics update)
              one FA launch (BatchNorm + ReLU folded into its pooling read; loss and gradient together)
    backward  one launch for both BatchNorm / ReLU backwards (dz, dgamma, dbeta)
              CE backward pass   (+ dz1 * w1[c] added at the strided pixels, dw1 collected)
              MSE backward pass  (+ dz2 * w2[c], dw2)

It takes the model's own two transformer modules, so parameters, running statistics and state-dict keys stay where the
reference keeps them (``DSRL.SSSR_feature_transformer`` / ``DSRL.SISR_feature_transformer``); the model then skips calling
them in its forward.  All arithmetic runs behind the C-ABI of libdsrl_b200.so; CUDA tensors only, no fallback.
"""
from __future__ import annotations

import ctypes

import torch

from ... import _lib

_DT = {torch.uint8: _lib.U8, torch.int32: _lib.I32, torch.int64: _lib.I64}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _parts(transformer):
    """(conv, bn) of a reference feature transformer (DSRL.py:86-95); checks it is what the kernels implement."""
    mods = list(transformer.children())
    if len(mods) != 3 or not isinstance(mods[0], torch.nn.Conv2d) or not isinstance(mods[1], torch.nn.BatchNorm2d) \
            or not isinstance(mods[2], torch.nn.ReLU):
        raise TypeError("feature transformer must be Sequential(Conv2d, BatchNorm2d, ReLU) as in models/DSRL.py:86-95")
    conv, bn = mods[0], mods[1]
    if conv.out_channels != 1 or conv.kernel_size != (1, 1) or conv.padding != (0, 0) or conv.bias is not None \
            or conv.stride[0] != conv.stride[1] or conv.groups != 1 or conv.dilation != (1, 1):
        raise TypeError("feature transformer convolution must be Conv2d(C, 1, kernel_size=1, stride=s, padding=0, bias=False)")
    if bn.num_features != 1 or not bn.affine or not bn.track_running_stats or bn.momentum is None:
        raise TypeError("feature transformer normalisation must be a plain BatchNorm2d(1) (affine, running statistics, momentum)")
    return conv, bn


class _Stage3Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sssr, sisr, image, w1, g1, b1, w2, g2, b2, target, rm1, rv1, rm2, rv2, cfg):
        ignore_index, stride, k, eps, momentum, training = cfg
        L = _lib.lib()
        dev = sssr.device
        sssr, sisr, image, target = sssr.contiguous(), sisr.contiguous(), image.contiguous(), target.contiguous()
        B, C1, H, W = sssr.shape
        C2 = sisr.shape[1]
        Hf, Wf = (H - 1) // stride + 1, (W - 1) // stride + 1
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        f32 = dict(dtype=torch.float32, device=dev)
        z1, z2 = torch.empty((B, 1, Hf, Wf), **f32), torch.empty((B, 1, Hf, Wf), **f32)
        losses = torch.empty(3, **f32)                           # CE, MSE, FA
        ce_saved_bytes = int(L.dsrl_ce_saved_bytes(B, H * W))
        ce_saved = torch.empty(ce_saved_bytes, dtype=torch.uint8, device=dev)
        mse_ws_bytes = int(L.dsrl_mse_workspace_bytes(B, C2, H, W))
        mse_ws = torch.empty(mse_ws_bytes, dtype=torch.uint8, device=dev)
        w1f, w2f = w1.reshape(-1).contiguous(), w2.reshape(-1).contiguous()
        with torch.cuda.device(dev), _lib.nvtx_range("dsrl.stage3_forward"):
            _lib.check(L.dsrl_ce_forward_tap(_ptr(sssr), _ptr(target), _DT[target.dtype], B, C1, H, W, ignore_index, _lib.REDUCE_MEAN,
                                             _ptr(losses[0:1]), _ptr(ce_saved), ce_saved_bytes, _ptr(w1f), _ptr(z1), stride, st))
            _lib.check(L.dsrl_mse_forward(_ptr(sisr), _ptr(image), B, C2, H, W, _ptr(losses[1:2]), _ptr(mse_ws), mse_ws_bytes,
                                          _ptr(w2f), _ptr(z2), stride, st))
            bn = torch.empty(8, **f32)
            _lib.check(L.dsrl_ft_bn_forward(_ptr(z1), _ptr(z2), B * Hf * Wf, _ptr(g1), _ptr(b1), _ptr(rm1), _ptr(rv1), _ptr(g2), _ptr(b2),
                                            _ptr(rm2), _ptr(rv2), eps, momentum, int(training), _ptr(bn), st))
            geom = (_lib.FA_REFERENCE, 0, B, 1, 1, Hf, Wf, k)
            fa_saved_bytes, fa_ws_bytes = int(L.dsrl_fa_saved_bytes(*geom)), int(L.dsrl_fa_workspace_bytes(*geom))
            if fa_saved_bytes == 0:
                raise _lib.DsrlError(_lib.ERR_UNSUPPORTED, f"Stage3Loss: unsupported transformer output geometry {(B, 1, Hf, Wf)} / k={k}")
            fa_saved = torch.empty(fa_saved_bytes, dtype=torch.uint8, device=dev)
            fa_ws = torch.empty(max(fa_ws_bytes, 16), dtype=torch.uint8, device=dev)
            one = torch.ones((), **f32)
            df1, df2 = torch.empty_like(z1), torch.empty_like(z2)
            _lib.check(L.dsrl_fa_forward_backward_transformed(_ptr(z1), _ptr(z2), _ptr(bn), B, Hf, Wf, k, _lib.REDUCE_MEAN, _ptr(one),
                                                              _ptr(losses[2:3]), _ptr(df1), _ptr(df2), _ptr(fa_saved), fa_saved_bytes,
                                                              _ptr(fa_ws), fa_ws_bytes, st))
        ctx.save_for_backward(sssr, sisr, image, target, w1f, w2f, z1, z2, df1, df2, bn, ce_saved)
        ctx.cfg = (ignore_index, stride, training, B, C1, C2, H, W, Hf, Wf, ce_saved_bytes)
        return losses[0], losses[1], losses[2]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_ce, g_mse, g_fa):
        sssr, sisr, image, target, w1f, w2f, z1, z2, df1, df2, bn, ce_saved = ctx.saved_tensors
        ignore_index, stride, training, B, C1, C2, H, W, Hf, Wf, ce_saved_bytes = ctx.cfg
        L = _lib.lib()
        dev = sssr.device
        f32 = dict(dtype=torch.float32, device=dev)
        zero = torch.zeros((), **f32)
        g_ce = (g_ce if g_ce is not None else zero).to(torch.float32).contiguous()
        g_mse = (g_mse if g_mse is not None else zero).to(torch.float32).contiguous()
        g_fa = (g_fa if g_fa is not None else zero).to(torch.float32).contiguous()
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        dz1, dz2 = torch.empty_like(z1), torch.empty_like(z2)
        dgb = torch.empty(4, **f32)                              # dgamma1, dbeta1, dgamma2, dbeta2
        nblk = int(L.dsrl_tap_dw_blocks(B, H, W, stride))
        dw1p, dw2p = torch.empty((nblk, C1), **f32), torch.empty((nblk, C2), **f32)
        dsssr, dsisr = torch.empty_like(sssr), torch.empty_like(sisr)
        used = ctypes.c_int64(0)
        with torch.cuda.device(dev), _lib.nvtx_range("dsrl.stage3_backward"):
            _lib.check(L.dsrl_ft_bn_backward(_ptr(z1), _ptr(df1), _ptr(dz1), _ptr(dgb[0:2]), _ptr(z2), _ptr(df2), _ptr(dz2), _ptr(dgb[2:4]),
                                             B * Hf * Wf, _ptr(bn), _ptr(g_fa), int(training), st))
            _lib.check(L.dsrl_ce_backward_tap(_ptr(sssr), _ptr(target), _DT[target.dtype], B, C1, H, W, ignore_index, _lib.REDUCE_MEAN,
                                              _ptr(ce_saved), ce_saved_bytes, _ptr(g_ce), _ptr(dsssr), _ptr(w1f), _ptr(dz1), _ptr(dw1p),
                                              stride, ctypes.byref(used), st))
            n1 = int(used.value)
            _lib.check(L.dsrl_mse_backward(_ptr(sisr), _ptr(image), B, C2, H, W, _ptr(g_mse), _ptr(dsisr), _ptr(w2f), _ptr(dz2), _ptr(dw2p),
                                           stride, st))
            n2 = ((H * W + 1023) // 1024) * B
        dw1 = dw1p[:n1].sum(dim=0).reshape(1, C1, 1, 1)
        dw2 = dw2p[:n2].sum(dim=0).reshape(1, C2, 1, 1)
        return (dsssr, dsisr, None, dw1, dgb[0:1], dgb[1:2], dw2, dgb[2:3], dgb[3:4], None, None, None, None, None, None)


class Stage3Loss(torch.nn.Module):
    """``ce, mse, fa = Stage3Loss(model.SSSR_feature_transformer, model.SISR_feature_transformer, ignore_index=255)(
    SSSR_output, SISR_output, target, input_org)`` -- the three stage-3 losses (unweighted; combine them as the reference does:
    ``ce + w1 * mse + w2 * fa``, train_or_resume.py:435-438).  ``target`` may stay uint8.  Follows ``self.training`` of the
    transformers' BatchNorms (batch statistics + running update, or running statistics)."""

    def __init__(self, sssr_transformer, sisr_transformer, ignore_index: int = -100, subsample_factor: int = 8):
        super().__init__()
        self.sssr_transformer, self.sisr_transformer = sssr_transformer, sisr_transformer
        (self._conv1, self._bn1), (self._conv2, self._bn2) = _parts(sssr_transformer), _parts(sisr_transformer)
        if self._conv1.stride != self._conv2.stride or self._bn1.eps != self._bn2.eps or self._bn1.momentum != self._bn2.momentum:
            raise TypeError("the two feature transformers must share stride, eps and momentum")
        self.ignore_index = int(ignore_index)
        self.subsample_factor = int(subsample_factor)

    def forward(self, sssr_output, sisr_output, target, image):
        for t in (sssr_output, sisr_output, target, image):
            if not t.is_cuda:
                raise RuntimeError("Stage3Loss (dsrl-b200) runs on CUDA tensors only: there is no CPU fallback")
        if sssr_output.dtype != torch.float32 or sisr_output.dtype != torch.float32 or image.dtype != torch.float32:
            raise TypeError("Stage3Loss: SSSR_output, SISR_output and the image must be float32")
        if target.dtype not in _DT:
            raise TypeError("Stage3Loss: target must be uint8, int32 or int64 class indices")
        if sssr_output.dim() != 4 or sisr_output.shape[2:] != sssr_output.shape[2:] or image.shape != sisr_output.shape \
                or tuple(target.shape) != (sssr_output.shape[0],) + tuple(sssr_output.shape[2:]) \
                or sssr_output.shape[1] != self._conv1.in_channels or sisr_output.shape[1] != self._conv2.in_channels:
            raise ValueError("Stage3Loss: expected SSSR_output (B,C1,H,W), SISR_output / image (B,C2,H,W), target (B,H,W)")
        bn1, bn2 = self._bn1, self._bn2
        training = bn1.training
        if training and bn1.num_batches_tracked is not None:
            bn1.num_batches_tracked += 1
            bn2.num_batches_tracked += 1
        cfg = (self.ignore_index, int(self._conv1.stride[0]), self.subsample_factor, float(bn1.eps), float(bn1.momentum), bool(training))
        return _Stage3Function.apply(sssr_output, sisr_output, image, self._conv1.weight, bn1.weight, bn1.bias, self._conv2.weight,
                                     bn2.weight, bn2.bias, target, bn1.running_mean, bn1.running_var, bn2.running_mean,
                                     bn2.running_var, cfg)
