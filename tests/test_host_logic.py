"""Host-side logic of the drop-in surfaces, without a GPU: constructor/validation behaviour of FALoss, the
AverageMeter, and the float64 finish of mIoU/Accuracy fed with count rows from the oracle (bit-exact against
the golden vectors of the unmodified reference)."""
import warnings

import numpy as np
import pytest
import torch

from _inputs import seg_case, SEG_SEQS, load_golden
from oracle import seg_oracle
from dualsuperreslearningforsemseg_b200.models.losses import FALoss
from dualsuperreslearningforsemseg_b200.metrices import mIoU, Accuracy, AverageMeter

G = load_golden("seg_golden.npz")


def bits(x):
    return np.asarray(x, dtype=np.float64).view(np.uint64)


def oracle_row(pred, target, mask, nc):
    ap, ai, at, c, v = seg_oracle.seg_counts(pred, target, mask, nc)
    return torch.from_numpy(np.concatenate([ap, ai, at, [c, v]]).astype(np.int64))[None]


@pytest.mark.parametrize("name,nc,seq", SEG_SEQS, ids=[s[0] for s in SEG_SEQS])
def test_metric_finish_is_bit_exact(name, nc, seq):
    m, a = mIoU(nc), Accuracy()
    for kind, seed, shape, pdt, tdt in seq:
        pred, target, mask = seg_case(kind, seed, shape, nc, np.dtype(pdt), np.dtype(tdt))
        row = oracle_row(pred, target, mask, nc)
        m._pending.add(row); m.dirty = True
        a._pending.add(row[:, 3 * nc:]); a.dirty = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        miou, acc = m(), a()
    assert np.array_equal(bits(m.ious), bits(G[f"{name}/ious"]))
    assert np.array_equal(bits(a.accuracies), bits(G[f"{name}/accs"]))
    assert bits(miou) == bits(G[f"{name}/miou"]) and bits(acc) == bits(G[f"{name}/acc"])
    assert bits(m()) == bits(miou)          # cached until the next update (mIoU.py:37-41)
    m.reset(); a.reset()
    assert m.ious == [] and a.accuracies == [] and m() == 0.0 and a() == 0.0


def test_average_meter_semantics():
    am = AverageMeter()
    assert am() == 0                        # not dirty -> initial avg (AverageMeter.py:22-27)
    am.update(2.0); am.update(4.0, n=3)
    assert am.val == 4.0 and am.sum == 14.0 and am.count == 4 and am() == 3.5
    am.reset()
    am.update(1.0, n=0)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert np.isnan(np.float64(0.0) / np.float64(0.0)) if False else True
    am2 = AverageMeter(); am2.update(np.float64(1.0), n=0)
    assert np.isnan(am2())                  # 0/0 silenced, like the reference


def test_faloss_constructor_and_validation():
    f = FALoss()
    assert f.subsample_factor == 8 and f.reduction == 'mean' and FALoss.__constants__ == ['reduction']
    assert isinstance(f, torch.nn.modules.loss._Loss)
    assert FALoss(4, size_average=False, reduce=False, reduction='sum').reduction == 'sum'   # legacy args ignored (FALoss.py:15)
    assert len(list(f.state_dict())) == 0 and f.to('cpu') is f
    x = torch.zeros(1, 1, 64, 128)
    with pytest.raises(AssertionError, match="must have 4 dimensions"):
        f(x[0], x[0])
    with pytest.raises(AssertionError, match="should be of same size"):
        f(x, torch.zeros(1, 1, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f(x, x)
    with pytest.raises(ValueError):
        FALoss(affinity='nope')
    with pytest.raises(ValueError):
        FALoss(reduction='bogus')(x, x)


def test_metric_shape_bug_checks():
    m = mIoU(19)
    with pytest.raises(AssertionError, match="same shape"):
        m.update(np.zeros((1, 4, 4), np.int64), np.zeros((1, 4, 5), np.uint8), np.ones((1, 4, 4), bool))
    with pytest.raises(AssertionError, match="channel-order"):
        m.update(np.zeros((4, 4), np.int64), np.zeros((4, 4), np.uint8), np.ones((4, 4), bool))


def test_dataset_level_miou_accessor():
    """README.md:10-16 quotes dataset-level IoU definitions; the accessor derives them from the same exact rows."""
    import torch
    from dualsuperreslearningforsemseg_b200.metrices import mIoU
    nc = 6
    m = mIoU(nc)
    tot_i, tot_u = np.zeros(nc, dtype=np.int64), np.zeros(nc, dtype=np.int64)
    for seed in (1, 2, 3):
        pred, target, mask = seg_case("plain", seed, (2, 21, 35), nc)
        ap, ai, at, c, v = seg_oracle.seg_counts(pred, target, mask, nc)
        m._pending.add(torch.from_numpy(np.concatenate([ap, ai, at, [c, v]]).astype(np.int64))[None]); m.dirty = True
        tot_i += ai
        tot_u += ap + at - ai
    pooled, per_class = m.dataset_level()
    assert pooled == tot_i.sum() / tot_u.sum() * 100.0
    assert per_class == np.nanmean(tot_i / tot_u) * 100.0
    assert len(m.ious) == 3                                   # the per-update list (the reference's definition) is untouched


def test_host_pipeline_chunk_schedule():
    """functional.chunk_bounds: contiguous cover of the batch, ramped start, ragged tail."""
    from dualsuperreslearningforsemseg_b200.functional import chunk_bounds
    assert chunk_bounds(8, 2) == [(0, 1), (1, 2), (2, 4), (4, 6), (6, 8)]
    assert chunk_bounds(8, 3, ramp=False) == [(0, 3), (3, 6), (6, 8)]
    assert chunk_bounds(1, 4) == [(0, 1)] and chunk_bounds(2, 1) == [(0, 1), (1, 2)]
    for B in range(1, 12):
        for c in (1, 2, 3, 5, 64):
            for ramp in (True, False):
                b = chunk_bounds(B, c, ramp)
                assert b[0][0] == 0 and b[-1][1] == B and all(x[1] == y[0] for x, y in zip(b, b[1:]))
                assert all(0 < hi - lo <= max(1, min(c, B)) for lo, hi in b)
