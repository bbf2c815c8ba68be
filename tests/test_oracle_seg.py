"""Pins oracle/seg_oracle.py (NumPy and C restatements) against golden vectors produced by the unmodified
reference metrices (tests/golden/make_golden.py), bit-exactly.  CPU only."""
import warnings

import numpy as np
import pytest

from _inputs import seg_case, SEG_SEQS, load_golden
from oracle import seg_oracle

G = load_golden("seg_golden.npz")


def bits(x):
    return np.asarray(x, dtype=np.float64).view(np.uint64)


def test_reference_fixture_scratchpad():
    # scratchpad.py:361-363 of the reference
    pred = np.array([[[0, 1, 3, 3, 4, 5], [2, 3, 1, 1, 3, 4]]], dtype=np.int64)
    target = np.array([[[0, 1, 2, 3, 4, 255], [2, 255, 1, 4, 255, 4]]], dtype=np.int64)
    ap, ai, at, c, v = seg_oracle.seg_counts(pred, target, target != 255, 6)
    assert ap.tolist() == [1, 3, 1, 2, 2, 0] and ai.tolist() == [1, 2, 1, 1, 2, 0] and at.tolist() == [1, 2, 2, 1, 3, 0]
    m, a = seg_oracle.MIoUOracle(6), seg_oracle.AccuracyOracle()
    m.update(pred, target, target != 255)
    a.update(pred, target, target != 255)
    assert m() == 66.66666666666666 == float(G["fixture/miou"])
    assert a() == 77.77777777777779 == float(G["fixture/acc"])


@pytest.mark.parametrize("name,nc,seq", SEG_SEQS, ids=[s[0] for s in SEG_SEQS])
def test_sequences_bit_exact(name, nc, seq):
    m, a = seg_oracle.MIoUOracle(nc), seg_oracle.AccuracyOracle()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for kind, seed, shape, pdt, tdt in seq:
            pred, target, mask = seg_case(kind, seed, shape, nc, np.dtype(pdt), np.dtype(tdt))
            m.update(pred, target, mask)
            a.update(pred, target, mask)
        miou, acc = m(), a()
    assert np.array_equal(bits(m.ious), bits(G[f"{name}/ious"]))
    assert np.array_equal(bits(a.accuracies), bits(G[f"{name}/accs"]))
    assert bits(miou) == bits(G[f"{name}/miou"])
    assert bits(acc) == bits(G[f"{name}/acc"])


def test_c_restatement_equals_numpy():
    for seed in range(4):
        pred, target, mask = seg_case("oor_target" if seed % 2 else "oor_pred", seed, (3, 65, 129), 19)
        ref = seg_oracle.seg_counts(pred, target, mask, 19)
        for threads in (1, 3):
            got = seg_oracle.seg_counts_c(pred, target, mask, 19, threads=threads)
            for r, g in zip(ref, got):
                assert np.array_equal(np.asarray(r), np.asarray(g))


def test_argmax_first_ties():
    x = np.zeros((1, 4, 2, 2), dtype=np.float32)
    x[0, 2, 0, 0] = 1.0
    x[0, 3, 0, 0] = 1.0
    assert seg_oracle.argmax_first(x)[0, 0, 0] == 2 and seg_oracle.argmax_first(x)[0, 1, 1] == 0
