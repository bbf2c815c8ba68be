"""CPU oracle for the Feature-Affinity (FA) loss -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``dualsuperreslearningforsemseg_b200``) never does and fails loudly if the CUDA library is missing.

This is an independent float64 NumPy restatement of what the reference computes, written from the
maths, not from the reference's torch ops:

* ``reference`` semantics follows ``models/losses/FALoss.py:8-11`` (spectral normalisation + per-channel
  ``w x w`` Gram over the height axis), ``:23-24`` (non-overlapping ``k x k`` mean pooling, floor) and
  ``:27-34`` (all-pairs L1 between the two flattened Gram matrices, ``mean``/``sum``/``none``).
  The backward is the closed form of the autograd graph the reference builds implicitly
  (triggered at ``command_handlers/train_or_resume.py:444``).
* ``position`` semantics is the paper's N x N position affinity (Wang et al., CVPR 2020).  Its Gram
  sub-step is the same contraction as ``FALoss.py:11``; the loss/gradients are NOT in the reference:
  **parity unpinned by the reference** for this mode (pinned only against PyTorch autograd of the same
  formula in ``tests/golden/make_golden.py``).

Parity pin: ``tests/golden/fa_golden.npz`` holds inputs/outputs produced by importing the unmodified
reference ``FALoss`` (CPU, fp64 and fp32) in the build container; ``tests/test_oracle_fa.py`` checks this
restatement against every vector in it.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "avg_pool",
    "unpool_grad",
    "reference_gram",
    "fa_reference",
    "fa_position",
    "round_tf32",
    "fa_position_rows",
    "allpairs_l1_sorted",
]


# --------------------------------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------------------------------
def avg_pool(x: np.ndarray, k: int) -> np.ndarray:
    """Non-overlapping k x k mean, stride k, no padding, floor (FALoss.py:23-24: ``AvgPool2d(k)``)."""
    x = np.asarray(x, dtype=np.float64)
    B, C, H, W = x.shape
    h, w = H // k, W // k
    if h == 0 or w == 0:
        raise ValueError("feature map smaller than the pooling window")
    return x[:, :, : h * k, : w * k].reshape(B, C, h, k, w, k).mean(axis=(3, 5))


def unpool_grad(dP: np.ndarray, k: int, H: int, W: int) -> np.ndarray:
    """Adjoint of :func:`avg_pool`: every pooled cell's gradient is spread over its k x k window / k^2;
    rows/cols dropped by the floor get zero."""
    B, C, h, w = dP.shape
    dX = np.zeros((B, C, H, W), dtype=np.float64)
    dX[:, :, : h * k, : w * k] = np.repeat(np.repeat(dP, k, axis=2), k, axis=3) / float(k * k)
    return dX


def _top_singular(A: np.ndarray):
    """sigma_max and its singular pair for a batch of matrices (..., h, w)."""
    U, s, Vt = np.linalg.svd(A, full_matrices=False)
    return s[..., 0], U[..., :, 0], Vt[..., 0, :]


def reference_gram(P: np.ndarray):
    """FALoss.py:8-11.  ``S = (A/sigma)^T (A/sigma)`` per (b, c); sigma = matrix 2-norm of the h x w slice.

    An all-zero slice has sigma = 0 and the reference's 0/0 makes the whole slice NaN; reproduced here
    (loss NaN; the dead branch's gradient NaN, the other branch's gradient for that (b, c) exactly 0).
    Returns (S, sigma, u1, v1)."""
    sigma, u1, v1 = _top_singular(P)
    with np.errstate(divide="ignore", invalid="ignore"):
        Ah = P / sigma[..., None, None]
    S = np.swapaxes(Ah, -1, -2) @ Ah
    return S, sigma, u1, v1


def allpairs_l1_sorted(a: np.ndarray, b: np.ndarray):
    """Exact O(n log n) evaluation of the all-pairs terms for 1-D ``a``, ``b``:

    ``row_abs[i] = sum_j |a_i - b_j|``, ``row_sgn[i] = sum_j sign(a_i - b_j)``,
    ``col_sgn[j] = sum_i sign(a_i - b_j)``  (sign(0) = 0).
    Used when n^2 cannot be materialised (FALoss.py:27-30 would need n^2 floats per operand)."""
    n = b.size
    bs = np.sort(b)
    pre = np.concatenate(([0.0], np.cumsum(bs)))
    lt = np.searchsorted(bs, a, side="left")
    le = np.searchsorted(bs, a, side="right")
    gt = n - le
    row_abs = a * (lt - gt) - pre[lt] + (pre[n] - pre[le])
    row_sgn = (lt - gt).astype(np.float64)
    as_ = np.sort(a)
    a_lt = np.searchsorted(as_, b, side="left")          # #{a_i < b_j}
    a_gt = a.size - np.searchsorted(as_, b, side="right")  # #{a_i > b_j}
    col_sgn = (a_gt - a_lt).astype(np.float64)
    return row_abs, row_sgn, col_sgn


# --------------------------------------------------------------------------------------------------
# reference semantics (the graded parity target)
# --------------------------------------------------------------------------------------------------
def fa_reference(x1, x2, k: int = 8, reduction: str = "mean", grad_out=None, need_grad: bool = True,
                 materialise_limit: int = 1 << 22):
    """Reference-semantics FA loss and its gradients in float64.

    x1, x2 : (B, C, H, W).  ``reduction`` as ``torch.nn.functional.l1_loss`` (FALoss.py:32-34).
    grad_out : upstream gradient -- scalar for mean/sum (default 1), array (B, C, n^2) for ``none``
    (default ones).  Returns ``(loss, dX1, dX2)``; the gradients are None when ``need_grad`` is False.
    """
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    if x1.ndim != 4 or x2.ndim != 4:
        raise ValueError("FALoss inputs must be 4-D (B, C, H, W)")        # FALoss.py:19
    if x1.shape != x2.shape:
        raise ValueError("FALoss inputs must have the same shape")         # FALoss.py:20
    if reduction not in ("mean", "sum", "none"):
        raise ValueError(f"{reduction} is not a valid value for reduction")
    B, C, H, W = x1.shape
    P1, P2 = avg_pool(x1, k), avg_pool(x2, k)
    h, w = P1.shape[2:]
    n = w * w
    S1, sg1, u1, v1 = reference_gram(P1)
    S2, sg2, u2, v2 = reference_gram(P2)
    a = S1.reshape(B, C, n)
    b = S2.reshape(B, C, n)

    if reduction == "none" or n * n <= materialise_limit:
        # FALoss.py:27-30: a is repeat_interleave'd (index i*n+j -> a_i), b is tiled (-> b_j)
        D = a[..., :, None] - b[..., None, :]                  # (B, C, n, n)
        absD = np.abs(D)
        # torch's l1_loss backward uses sign() = (0 < d) - (d < 0): a NaN difference contributes 0
        sgn = (D > 0).astype(np.float64) - (D < 0).astype(np.float64)
        if reduction == "none":
            loss = absD.reshape(B, C, n * n)
            if grad_out is None:
                grad_out = np.ones_like(loss)
            Gw = np.asarray(grad_out, dtype=np.float64).reshape(B, C, n, n)
            g1 = (sgn * Gw).sum(axis=3)
            g2 = -(sgn * Gw).sum(axis=2)
        else:
            total = absD.sum()
            Z = float(B * C * n * n) if reduction == "mean" else 1.0
            loss = total / Z
            go = 1.0 if grad_out is None else float(grad_out)
            g1 = sgn.sum(axis=3) * (go / Z)
            g2 = -sgn.sum(axis=2) * (go / Z)
    else:
        total = 0.0
        g1 = np.empty((B, C, n))
        g2 = np.empty((B, C, n))
        for bi in range(B):
            for ci in range(C):
                ra, rs, cs = allpairs_l1_sorted(a[bi, ci], b[bi, ci])
                if not (np.isfinite(a[bi, ci]).all() and np.isfinite(b[bi, ci]).all()):
                    ra = np.full(n, np.nan); rs = np.zeros(n); cs = np.zeros(n)    # sign(NaN) -> 0, see above
                total += ra.sum()
                g1[bi, ci] = rs
                g2[bi, ci] = -cs
        Z = float(B * C * n * n) if reduction == "mean" else 1.0
        loss = total / Z
        go = 1.0 if grad_out is None else float(grad_out)
        g1 *= go / Z
        g2 *= go / Z

    if not need_grad:
        return loss, None, None

    def back(P, sigma, u, v, g):
        G = g.reshape(B, C, w, w)
        sig = sigma[..., None, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            Ah = P / sig
            Gh = Ah @ (G + np.swapaxes(G, -1, -2))             # dL/dA_hat
            inner = (Gh * P).sum(axis=(2, 3))[..., None, None]
            dP = Gh / sig - (inner / (sig * sig)) * (u[..., :, None] * v[..., None, :])
        return unpool_grad(dP, k, H, W)

    return loss, back(P1, sg1, u1, v1, g1), back(P2, sg2, u2, v2, g2)


# --------------------------------------------------------------------------------------------------
# position semantics (opt-in; parity unpinned by the reference)
# --------------------------------------------------------------------------------------------------
def _position_normalise(P: np.ndarray, eps: float = 1e-12):
    B, C, h, w = P.shape
    F = P.reshape(B, C, h * w)
    nrm = np.maximum(np.sqrt((F * F).sum(axis=1, keepdims=True)), eps)     # (B, 1, N)
    return F, F / nrm, nrm


def round_tf32(x: np.ndarray) -> np.ndarray:
    """float32 -> TF32 (10 explicit mantissa bits), round to nearest, ties away from zero: what PTX
    ``cvt.rna.tf32.f32`` does to the operands before the tensor-core contraction.  Returned as float64."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x1000)) & np.uint64(0xFFFFE000)
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def round_f16(x: np.ndarray) -> np.ndarray:
    """float32 -> IEEE binary16 (round to nearest even, what ``cvt.rn.f16.f32`` / ``__float2half_rn`` do): the same 11-bit
    significand as TF32 for |x| >= 2^-14; smaller values lose relative precision (absolute error <= 2^-25), which for
    unit-norm feature vectors is far below the ordinary rounding error.  Returned as float64."""
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float64)


def fa_position(x1, x2, k: int = 8, reduction: str = "mean", grad_out=None, need_grad: bool = True,
                chunk: int = 1024, operand_rounding: str | None = None):
    """Position-affinity FA loss: ``S = Fh^T Fh`` (N x N, Fh = channel-L2-normalised pooled features),
    ``L = reduce |S1 - S2|`` with the diagonal forced to zero.  x1 and x2 may differ in C.
    Evaluated in row chunks so N = 32768 never materialises N x N.

    ``operand_rounding='tf32'`` rounds the normalised features to TF32 before the (still float64) contractions:
    the exact value of a single-pass TF32 tensor-core evaluation.  The gradient contains sign(S1 - S2), so operand
    rounding flips the sign of the few entries with |S1 - S2| below the rounding error; against the unrounded
    oracle that is a 0.5-1 % relative-norm difference on random inputs whatever the kernel does, which is why
    the single-pass TF32 kernel is graded against this variant and the 3xTF32 kernel against the unrounded one."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    if x1.ndim != 4 or x2.ndim != 4:
        raise ValueError("FALoss inputs must be 4-D (B, C, H, W)")
    if x1.shape[0] != x2.shape[0] or x1.shape[2:] != x2.shape[2:]:
        raise ValueError("FALoss(position) inputs must agree in B, H, W")
    if reduction not in ("mean", "sum"):
        raise ValueError("position affinity supports reduction 'mean' or 'sum'")
    B, _, H, W = x1.shape
    P1, P2 = avg_pool(x1, k), avg_pool(x2, k)
    h, w = P1.shape[2:]
    N = h * w
    F1, Fh1, n1 = _position_normalise(P1)
    F2, Fh2, n2 = _position_normalise(P2)
    if operand_rounding == "tf32":
        Fh1, Fh2 = round_tf32(Fh1), round_tf32(Fh2)
    elif operand_rounding == "f16":
        Fh1, Fh2 = round_f16(Fh1), round_f16(Fh2)
    elif operand_rounding is not None:
        raise ValueError("operand_rounding must be None, 'tf32' or 'f16'")
    Z = float(B * N * N) if reduction == "mean" else 1.0
    go = 1.0 if grad_out is None else float(grad_out)
    total = 0.0
    Gh1 = np.zeros_like(Fh1)
    Gh2 = np.zeros_like(Fh2)
    for bi in range(B):
        for r0 in range(0, N, chunk):
            r1 = min(N, r0 + chunk)
            D = Fh1[bi][:, r0:r1].T @ Fh1[bi] - Fh2[bi][:, r0:r1].T @ Fh2[bi]    # (rows, N)
            D[np.arange(r1 - r0), np.arange(r0, r1)] = 0.0
            total += np.abs(D).sum()
            if need_grad:
                Sg = np.sign(D) * (go / Z)                                           # rows of Sigma
                # dL/dFh = 2 Fh Sigma (Sigma symmetric): column block r0:r1 of the result
                Gh1[bi][:, r0:r1] = 2.0 * (Fh1[bi] @ Sg.T)
                Gh2[bi][:, r0:r1] = -2.0 * (Fh2[bi] @ Sg.T)
    loss = total / Z
    if not need_grad:
        return loss, None, None

    def back(P, Fh, nrm, Gh, eps=1e-12):
        C = P.shape[1]
        raw = np.sqrt((P.reshape(B, C, N) ** 2).sum(axis=1, keepdims=True))
        proj = (Fh * Gh).sum(axis=1, keepdims=True)
        dF = np.where(raw > eps, (Gh - Fh * proj) / nrm, Gh / eps)
        return unpool_grad(dF.reshape(B, C, h, w), k, H, W)

    return loss, back(P1, Fh1, n1, Gh1), back(P2, Fh2, n2, Gh2)


def fa_position_rows(x1, x2, rows, k: int = 8, reduction: str = "mean", operand_rounding: str | None = None):
    """Sampled check for maps too large for :func:`fa_position` in seconds (N = 32768, C = 256): for sample 0 and the given
    position indices ``rows`` return ``(sum_j |D_ij| over those rows, dX1[0, :, rows], dX2[0, :, rows])`` -- the gradient of
    a position only needs that position's affinity row, so these columns are exact.  k must be 1 (no pooling)."""
    if k != 1:
        raise ValueError("fa_position_rows: only k = 1")
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    B, _, H, W = x1.shape
    N = H * W
    rows = np.asarray(rows, dtype=np.int64)
    F1, Fh1, n1 = _position_normalise(x1[:1])
    F2, Fh2, n2 = _position_normalise(x2[:1])
    if operand_rounding == "f16":
        Fh1, Fh2 = round_f16(Fh1), round_f16(Fh2)
    if operand_rounding == "tf32":
        Fh1, Fh2 = round_tf32(Fh1), round_tf32(Fh2)
    Z = float(B * N * N) if reduction == "mean" else 1.0
    D = Fh1[0][:, rows].T @ Fh1[0] - Fh2[0][:, rows].T @ Fh2[0]          # (len(rows), N)
    D[np.arange(rows.size), rows] = 0.0
    Sg = np.sign(D) / Z
    out = [np.abs(D).sum()]
    for Fh, nrm, sgn in ((Fh1, n1, 2.0), (Fh2, n2, -2.0)):
        Gh = sgn * (Fh[0] @ Sg.T)                                       # (C, len(rows)): dL/dFh at the sampled positions
        Fr = Fh[0][:, rows]
        proj = (Fr * Gh).sum(axis=0, keepdims=True)
        out.append((Gh - Fr * proj) / nrm[0][:, rows])
    return tuple(out)
