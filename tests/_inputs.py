"""Seeded synthetic inputs shared by the oracle tests, the GPU parity tests and the golden generator.

`fa_inputs` / `seg_case` restate the generators in tests/golden/make_golden.py (which cannot be imported on
the GPU box because it imports the reference); tests/test_oracle_*.py prove they agree by reproducing the
golden outputs from these inputs.
"""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fa_inputs(shape, dist, seed):
    rng = np.random.default_rng(seed)
    x1 = rng.standard_normal(shape).astype(np.float32)
    x2 = rng.standard_normal(shape).astype(np.float32)
    if dist in ("relu", "dead"):
        x1 = np.maximum(x1, 0.0)
        x2 = np.maximum(x2, 0.0)
    if dist == "dead":
        x1[1, 0] = 0.0
    return x1, x2


def pos_inputs(s1, s2, seed):
    rng = np.random.default_rng(seed)
    x1 = np.maximum(rng.standard_normal(s1), 0).astype(np.float32)
    x2 = np.maximum(rng.standard_normal(s2), 0).astype(np.float32)
    return x1, x2


def pos_margin_inputs(B, C1, C2, H, W, seed):
    """Position-mode inputs whose affinity difference S1 - S2 stays away from zero everywhere off the diagonal.

    Each branch has two groups of positions (different groupings per branch) with a shared base vector per group plus
    noise; branch 1 is tight (cos ~0.95 inside a group, ~0.2 across), branch 2 loose (~0.75 / ~0.5), so S1 - S2 is
    about +0.2, +0.45, -0.55 or -0.3.  The gradient contains sign(S1 - S2): with a margin no rounding can flip a
    sign, and kernel-vs-oracle gradient differences measure the arithmetic rather than the discontinuity."""
    rng = np.random.default_rng(seed)
    N = H * W
    pos = np.arange(N)

    def branch(C, a, sigma, groups):
        base = np.full((2, C), a, dtype=np.float64)
        base[0, : C // 2] += 1.0
        base[1, C // 2:] += 1.0
        rms = np.sqrt((base ** 2).mean(axis=1))[groups]                       # (N,)
        x = base[groups].T[None] + sigma * rms[None, None, :] * rng.standard_normal((B, C, N))
        return x.reshape(B, C, H, W).astype(np.float32)

    return branch(C1, 0.1, 0.23, (pos // 3) % 2), branch(C2, 0.4, 0.58, (pos // 5) % 2)


def seg_case(kind, seed, shape=(2, 37, 53), nc=19, pred_dtype=np.int64, target_dtype=np.uint8):
    rng = np.random.default_rng(seed)
    target = rng.integers(0, nc, shape).astype(target_dtype)
    ign = rng.random(shape) < 0.1
    target[ign] = 255
    rnd = rng.integers(0, nc, shape)
    keep = rng.random(shape) < 0.7
    pred = np.where(keep, np.where(ign, 0, target), rnd).astype(pred_dtype)
    if kind == "all_ignored":
        target[...] = 255
    elif kind == "single_class":
        target[...] = 3
        pred[...] = 3
    elif kind == "oor_target":
        target[0, :5, :7] = 100
        pred[0, 2, :4] = 100
    elif kind == "oor_pred":
        pred[0, :3, :] = nc + 4
        if np.issubdtype(pred_dtype, np.signedinteger) and shape[0] > 1:
            pred[1, :2, :] = -1
    mask = target != 255
    if kind == "explicit_mask":
        mask = rng.random(shape) < 0.5
    return pred, target, mask


SEG_SEQS = [
    ("mixed19", 19, [("plain", 1, (2, 37, 53), "int64", "uint8"), ("plain", 2, (1, 64, 96), "int64", "uint8"),
                     ("oor_target", 3, (2, 37, 53), "int64", "uint8"), ("single_class", 4, (1, 16, 16), "int64", "uint8"),
                     ("oor_pred", 5, (2, 21, 35), "int64", "uint8"), ("explicit_mask", 6, (2, 37, 53), "int64", "uint8")]),
    ("with_all_ignored", 19, [("plain", 7, (1, 40, 40), "int64", "uint8"), ("all_ignored", 8, (1, 40, 40), "int64", "uint8"),
                              ("plain", 9, (3, 33, 31), "int64", "uint8")]),
    ("dtypes", 19, [("plain", 10, (2, 37, 53), "uint8", "uint8"), ("plain", 11, (2, 37, 53), "int32", "int64"),
                    ("oor_pred", 12, (2, 37, 53), "int32", "int32"), ("plain", 13, (2, 37, 53), "int64", "int64")]),
    ("nc6", 6, [("plain", 14, (2, 37, 53), "int64", "uint8")]),
    ("only_all_ignored", 19, [("all_ignored", 15, (1, 8, 8), "int64", "uint8")]),
]


def cfg3_maps(num_maps, seed=54321, shape=(1024, 2048), nc=19):
    """BASELINE config 3: uint8 targets with 10% ignore, int64 preds 70% correct."""
    rng = np.random.default_rng(seed)
    for _ in range(num_maps):
        target = rng.integers(0, nc, shape, dtype=np.uint8)
        ign = rng.random(shape, dtype=np.float32) < 0.1
        target[ign] = 255
        rnd = rng.integers(0, nc, shape, dtype=np.uint8)
        keep = rng.random(shape, dtype=np.float32) < 0.7
        pred = np.where(keep, np.where(ign, 0, target), rnd).astype(np.int64)
        yield pred[None], target[None], (target != 255)[None]


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def expand_pooled(gp, k, H, W):
    """Golden gradients are stored on the pooled grid (constant per k x k window, zero on the dropped border)."""
    B, C, h, w = gp.shape
    out = np.zeros((B, C, H, W), dtype=gp.dtype)
    out[:, :, : h * k, : w * k] = np.repeat(np.repeat(gp, k, 2), k, 3)
    return out


# ---- cross-entropy cases: name -> (shape (B,C,H,W), target dtype, ignore_index, reduction, logits scale, fraction ignored)
CE_CASES = {
    "plain_mean":  ((2, 19, 12, 20), "uint8", 255, "mean", 1.0, 0.1),
    "plain_sum":   ((2, 19, 12, 20), "uint8", 255, "sum", 1.0, 0.1),
    "long_m100":   ((3, 19, 8, 16), "int64", -100, "mean", 1.0, 0.2),
    "ragged":      ((3, 5, 7, 9), "uint8", 255, "mean", 1.0, 0.1),
    "int32_sum":   ((1, 3, 9, 11), "int32", 255, "sum", 1.0, 0.0),
    "one_class":   ((2, 1, 4, 8), "uint8", 255, "mean", 1.0, 0.3),
    "big_logits":  ((2, 19, 8, 8), "uint8", 255, "mean", 40.0, 0.1),
    "all_ignored": ((1, 19, 4, 8), "uint8", 255, "mean", 1.0, 1.0),
    "all_ign_sum": ((1, 19, 4, 8), "uint8", 255, "sum", 1.0, 1.0),
}


def ce_case(name, seed=54321):
    shape, tdt, ignore, red, scale, frac = CE_CASES[name]
    rng = np.random.default_rng([seed, sorted(CE_CASES).index(name)])
    B, C, H, W = shape
    x = (rng.standard_normal(shape) * scale).astype(np.float32)
    t = rng.integers(0, C, (B, H, W)).astype(np.int64)
    t[rng.random((B, H, W)) < frac] = ignore
    return x, t.astype(np.dtype(tdt)), ignore, red
