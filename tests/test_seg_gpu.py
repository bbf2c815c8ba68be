"""GPU parity of the mIoU / accuracy counts (K4) through the drop-in metrices -> C-ABI -> sm_100a kernel.
Everything here is bit-exact: integer counts, per-update float64 IoUs/accuracies and the final percentages."""
import warnings

import numpy as np
import pytest
import torch

from _inputs import seg_case, SEG_SEQS, load_golden, cfg3_maps
from oracle import seg_oracle

pytestmark = pytest.mark.gpu
G = load_golden("seg_golden.npz")


def bits(x):
    return np.asarray(x, dtype=np.float64).view(np.uint64)


def metrics():
    from dualsuperreslearningforsemseg_b200.metrices import mIoU, Accuracy
    return mIoU, Accuracy


def test_reference_fixture_scratchpad():
    mIoU, Accuracy = metrics()
    pred = np.array([[[0, 1, 3, 3, 4, 5], [2, 3, 1, 1, 3, 4]]], dtype=np.int64)
    target = np.array([[[0, 1, 2, 3, 4, 255], [2, 255, 1, 4, 255, 4]]], dtype=np.int64)
    m, a = mIoU(num_classes=6), Accuracy()
    m.update(pred, target, target != 255)
    a.update(pred, target, target != 255)
    assert m() == 66.66666666666666 and a() == 77.77777777777779


@pytest.mark.parametrize("as_cuda", [False, True], ids=["numpy_in", "cuda_in"])
@pytest.mark.parametrize("name,nc,seq", SEG_SEQS, ids=[s[0] for s in SEG_SEQS])
def test_sequences_bit_exact_vs_reference_golden(name, nc, seq, as_cuda):
    mIoU, Accuracy = metrics()
    m, a = mIoU(nc), Accuracy()
    for kind, seed, shape, pdt, tdt in seq:
        pred, target, mask = seg_case(kind, seed, shape, nc, np.dtype(pdt), np.dtype(tdt))
        if as_cuda:
            pred, target, mask = (torch.from_numpy(v).cuda() for v in (pred, target, mask))
        a.update(pred, target, mask)       # order of train_or_resume.py:480-481
        m.update(pred, target, mask)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        miou, acc = m(), a()
    assert np.array_equal(bits(m.ious), bits(G[f"{name}/ious"]))
    assert np.array_equal(bits(a.accuracies), bits(G[f"{name}/accs"]))
    assert bits(miou) == bits(G[f"{name}/miou"]) and bits(acc) == bits(G[f"{name}/acc"])


def raw_counts(pred, target, mask, nc, **kw):
    from dualsuperreslearningforsemseg_b200.metrices import _counts
    _counts._cache.update(key=None)
    rows = _counts.counts_for_update(pred, target, mask, nc, **kw)
    return rows.cpu().numpy()


def oracle_row(pred, target, mask, nc):
    ap, ai, at, c, v = seg_oracle.seg_counts(pred, target, mask, nc)
    return np.concatenate([ap, ai, at, [c, v]]).astype(np.int64)


@pytest.mark.parametrize("pdt", ["int64", "int32", "uint8"])
@pytest.mark.parametrize("tdt", ["uint8", "int32", "int64"])
@pytest.mark.parametrize("derive_mask", [False, True])
def test_counts_all_dtypes_and_ragged_sizes(pdt, tdt, derive_mask):
    for i, shape in enumerate([(1, 1, 1), (1, 7, 13), (2, 255, 257), (3, 301, 437)]):
        pred, target, mask = seg_case("oor_target" if i % 2 else "oor_pred", 40 + i, shape, 19, np.dtype(pdt), np.dtype(tdt))
        got = raw_counts(pred, target, None if derive_mask else mask, 19)
        assert np.array_equal(got[0], oracle_row(pred, target, target != 255 if derive_mask else mask, 19))


def test_empty_update_and_many_classes():
    pred, target, mask = seg_case("plain", 1, (2, 0, 5), 19)
    assert np.array_equal(raw_counts(pred, target, mask, 19)[0], np.zeros(59, np.int64))
    for nc in (1, 2, 37, 150, 254):
        pred, target, mask = seg_case("plain", nc, (2, 65, 63), nc)
        assert np.array_equal(raw_counts(pred, target, mask, nc)[0], oracle_row(pred, target, mask, nc))


def test_unaligned_views_take_the_scalar_path():
    pred, target, mask = seg_case("plain", 77, (1, 129, 515), 19)
    p = torch.from_numpy(np.concatenate([[0], pred.ravel()])).cuda()[1:].view(1, 129, 515)       # 8-byte offset
    t = torch.from_numpy(np.concatenate([[0], target.ravel()]).astype(np.uint8)).cuda()[1:].view(1, 129, 515)
    m = torch.from_numpy(np.concatenate([[0], mask.ravel()]).astype(np.uint8)).cuda()[1:].view(1, 129, 515)
    assert p.data_ptr() % 16 != 0 and t.data_ptr() % 2 != 0
    assert np.array_equal(raw_counts(p, t, m, 19)[0], oracle_row(pred, target, mask, 19))


def test_update_many_equals_sequential_updates_at_config3_size():
    """BASELINE config 3 geometry (1024 x 2048 maps, int64/uint8/bool) on a 6-map subset, one launch, checked
    against the C restatement; plus the size-independent invariants valid == #mask, sum(area_target) == #in-range."""
    mIoU, Accuracy = metrics()
    maps = list(cfg3_maps(6))
    pred = np.stack([p for p, _, _ in maps]); target = np.stack([t for _, t, _ in maps]); mask = np.stack([m for _, _, m in maps])
    rows = raw_counts(torch.from_numpy(pred).cuda(), torch.from_numpy(target).cuda(), torch.from_numpy(mask).cuda(), 19, updates_leading=True)
    for u in range(6):
        ap, ai, at, c, v = seg_oracle.seg_counts_c(pred[u], target[u], mask[u], 19, threads=4)
        assert np.array_equal(rows[u], np.concatenate([ap, ai, at, [c, v]]))
        assert rows[u, 58] == mask[u].sum() and rows[u, 38:57].sum() == (target[u] < 19).sum()
    m, a = mIoU(19), Accuracy()
    m.update_many(torch.from_numpy(pred).cuda(), torch.from_numpy(target).cuda(), torch.from_numpy(mask).cuda())
    a.update_many(torch.from_numpy(pred).cuda(), torch.from_numpy(target).cuda(), torch.from_numpy(mask).cuda())
    mo, ao = seg_oracle.MIoUOracle(19), seg_oracle.AccuracyOracle()
    for u in range(6):
        mo.update(pred[u], target[u], mask[u]); ao.update(pred[u], target[u], mask[u])
    assert np.array_equal(bits(m.ious), bits(mo.ious)) and bits(m()) == bits(mo())
    assert np.array_equal(bits(a.accuracies), bits(ao.accuracies)) and bits(a()) == bits(ao())


def test_coherent_label_maps():
    """Spatially coherent maps (whole warps see one class) -- the case that serialises atomic histograms."""
    nc = 19
    yy, xx = np.mgrid[0:512, 0:1024]
    target = ((yy // 64 + xx // 128) % nc).astype(np.uint8)[None]
    target[0, :32] = 255
    pred = np.roll(target, 5, axis=2).astype(np.int64)
    pred[target == 255] = 0
    mask = target != 255
    assert np.array_equal(raw_counts(pred, target, mask, nc)[0], oracle_row(pred, target, mask, nc))


@pytest.mark.parametrize("shape", [(2, 19, 64, 96), (1, 19, 33, 35), (3, 6, 50, 64)])
def test_fused_argmax_counts(shape):
    mIoU, _ = metrics()
    B, C, H, W = shape
    rng = np.random.default_rng(5)
    logits = rng.standard_normal(shape).astype(np.float32)
    logits = np.round(logits * 2) / 2                 # many exact ties -> first-max tie-break matters
    logits[0, 1, 0, :5] = np.nan                       # NaN counts as the maximum (numpy/torch argmax)
    _, target, mask = seg_case("plain", 3, (B, H, W), C)
    pred = seg_oracle.argmax_first(logits)
    assert np.array_equal(pred, torch.argmax(torch.from_numpy(logits), dim=1).numpy())
    m = mIoU(C)
    got_pred = m.update_from_logits(torch.from_numpy(logits).cuda(), torch.from_numpy(target).cuda(),
                                    torch.from_numpy(mask).cuda(), return_pred=True)
    assert np.array_equal(got_pred.cpu().numpy(), pred)
    mo = seg_oracle.MIoUOracle(C)
    mo.update(pred, target, mask)
    assert np.array_equal(bits(m.ious), bits(mo.ious))
    m2 = mIoU(C)
    m2.update_from_logits(torch.from_numpy(logits).cuda(), torch.from_numpy(target).cuda())      # mask derived in-kernel
    mo2 = seg_oracle.MIoUOracle(C); mo2.update(pred, target, target != 255)
    assert np.array_equal(bits(m2.ious), bits(mo2.ious))


def test_logits_many_updates_one_launch():
    """(U, B, NC, H, W) logits -> U rows in one launch, equal to U separate update_from_logits calls."""
    from dualsuperreslearningforsemseg_b200.metrices import _counts
    rng = np.random.default_rng(5)
    U, B, nc, H, W = 3, 2, 19, 40, 52
    logits = rng.standard_normal((U, B, nc, H, W)).astype(np.float32)
    target = rng.integers(0, nc, (U, B, H, W)).astype(np.uint8)
    target[rng.random(target.shape) < 0.1] = 255
    rows, _ = _counts.counts_from_logits(torch.from_numpy(logits).cuda(), torch.from_numpy(target).cuda(), None, nc, updates_leading=True)
    rows = rows.cpu().numpy()
    for u in range(U):
        pred = seg_oracle.argmax_first(logits[u])
        ap, ai, at, c, v = seg_oracle.seg_counts(pred, target[u], target[u] != 255, nc)
        assert np.array_equal(rows[u], np.concatenate([ap, ai, at, [c, v]]))
