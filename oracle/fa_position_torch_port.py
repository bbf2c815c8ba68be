"""PyTorch-CPU timing port of the position-affinity FA loss -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module.

The reference has no position-affinity implementation ("parity unpinned by the reference"); this is what
a user would write with the reference's own tools -- ``F.normalize`` over channels, the ``transpose @ matmul`` Gram of
``FALoss.py:11`` on the (C, N) view, ``F.l1_loss`` and implicit autograd -- restricted to a block of affinity ROWS so
that a bounded sample of BASELINE configs[3] (32768 x 32768 per sample) fits in host memory and time.
``tests/test_oracle_fa.py`` pins it against the float64 oracle.
"""
import torch
import torch.nn.functional as F


def row_block_loss(x1, x2, k, r0, r1):
    """Sum over rows [r0, r1) and all columns of |S1 - S2| for sample 0 (diagonal excluded), as a differentiable tensor."""
    f1 = F.normalize(F.avg_pool2d(x1, k).flatten(2), dim=1, eps=1e-12)[0]      # (C1, N)
    f2 = F.normalize(F.avg_pool2d(x2, k).flatten(2), dim=1, eps=1e-12)[0]
    d = f1[:, r0:r1].t() @ f1 - f2[:, r0:r1].t() @ f2                          # (rows, N)
    idx = torch.arange(r0, r1)
    mask = torch.ones_like(d)
    mask[idx - r0, idx] = 0.0
    return F.l1_loss(d * mask, torch.zeros_like(d), reduction="sum")


def fwd_bwd_rows(x1, x2, k, r0, r1):
    a = x1.detach().clone().requires_grad_(True)
    b = x2.detach().clone().requires_grad_(True)
    loss = row_block_loss(a, b, k, r0, r1)
    loss.backward()
    return loss.detach(), a.grad, b.grad
