"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy, float64) of the cross-entropy the reference's training step applies to
the SSSR logits: ``t.nn.CrossEntropyLoss(ignore_index=IGNORE_CLASS_LABEL)`` (command_handlers/train_or_resume.py:116,435).

The arithmetic lives in a third-party dependency, not under /root/reference: torch (reference pin 1.7.0, requirements.txt;
2.11.0 here) -- ``log_softmax`` over the class axis followed by ``nll_loss`` with ``ignore_index``:
    loss_p = logsumexp_c(x[b,:,p]) - x[b, t_p, p]            for pixels with t_p != ignore_index
    mean   = sum_p loss_p / #{valid p}   (NaN when no pixel is valid);   sum = sum_p loss_p
    dloss/dx[b,c,p] = (softmax(x[b,:,p])_c - [c == t_p]) * g / (#valid or 1),   0 at ignored pixels
Pinned against torch.nn.functional.cross_entropy in float64 by tests/golden/make_golden.py -> tests/golden/ce_golden.npz
(tests/test_oracle_ce.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this."""
import numpy as np


def cross_entropy(logits, target, ignore_index=255, reduction="mean", grad_out=1.0):
    """logits (B,C,...) any float dtype; target (B,...) integer.  Returns (loss float64, dlogits float64)."""
    x = np.asarray(logits, dtype=np.float64)
    t = np.asarray(target).astype(np.int64)
    B, C = x.shape[0], x.shape[1]
    xf = x.reshape(B, C, -1)
    tf = t.reshape(B, -1)
    valid = tf != ignore_index
    m = xf.max(axis=1, keepdims=True)
    e = np.exp(xf - m)
    s = e.sum(axis=1, keepdims=True)
    lse = (m + np.log(s))[:, 0, :]
    tc = np.where(valid, tf, 0)
    xt = np.take_along_axis(xf, tc[:, None, :], axis=1)[:, 0, :]
    per = np.where(valid, lse - xt, 0.0)
    n = int(valid.sum())
    total = per.sum()
    with np.errstate(invalid="ignore", divide="ignore"):
        loss = total / n if reduction == "mean" else total
    scale = grad_out / n if (reduction == "mean" and n > 0) else (grad_out if reduction == "sum" else 0.0)
    p = e / s
    onehot = np.zeros_like(p)
    np.put_along_axis(onehot, tc[:, None, :], 1.0, axis=1)
    g = (p - onehot) * scale * valid[:, None, :]
    return float(loss), g.reshape(x.shape)
