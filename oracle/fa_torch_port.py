"""CPU timing port of the reference FA loss -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module.

The reference (models/losses/FALoss.py:8-34) is eager PyTorch with implicit autograd; it cannot travel to the
GPU box, so this is an independent PyTorch-CPU restatement with the same cost structure -- average pool,
spectral norm via SVD, batched Gram matmul, BOTH n^2 operands materialised, ``l1_loss``, autograd backward --
so that timing it on the box's host cores is a fair stand-in for timing the reference there
(``cpu_baseline.kind = "port"``).  ``tests/test_oracle_fa.py`` pins it against the reference's golden vectors.
"""
import torch
import torch.nn.functional as F


def _similarity(x, k):
    p = F.avg_pool2d(x, kernel_size=k)                                   # FALoss.py:23-24
    p = p / torch.linalg.matrix_norm(p, ord=2, keepdim=True)             # FALoss.py:10 (largest singular value)
    return torch.einsum("bchi,bchj->bcij", p, p)                         # FALoss.py:11 (A^T A over the height axis)


def fa_loss(x1, x2, k=8, reduction="mean"):
    s1 = _similarity(x1, k).flatten(2)
    s2 = _similarity(x2, k).flatten(2)
    B, C, n = s1.shape
    lhs = s1.unsqueeze(3).expand(B, C, n, n).reshape(B, C, n * n)        # element i*n+j = s1[i]  (FALoss.py:28)
    rhs = s2.unsqueeze(2).expand(B, C, n, n).reshape(B, C, n * n)        # element i*n+j = s2[j]  (FALoss.py:30)
    return F.l1_loss(lhs, rhs, reduction=reduction)                      # FALoss.py:32-34


def fwd_bwd(x1, x2, k=8, reduction="mean"):
    """One forward + backward; returns (loss, dx1, dx2) as tensors."""
    a = x1.detach().clone().requires_grad_(True)
    b = x2.detach().clone().requires_grad_(True)
    loss = fa_loss(a, b, k, reduction)
    loss.backward()
    return loss.detach(), a.grad, b.grad
