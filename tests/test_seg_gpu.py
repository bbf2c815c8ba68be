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
    _counts.clear_shared_pass()
    rows, _ = _counts.counts_for_update(pred, target, mask, nc, **kw)
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


@pytest.mark.parametrize("as_cuda", [False, True], ids=["numpy_in", "cuda_in"])
def test_the_meter_pair_shares_one_pass(as_cuda):
    """The reference feeds the same arrays to both meters back to back, in either order (train_or_resume.py:480-481,
    benchmark.py:76-77): the second meter re-uses the first one's kernel pass -- one launch per batch, also for NumPy inputs
    (one host -> device copy) and whatever the number of classes, without any state shared between threads or instances."""
    from dualsuperreslearningforsemseg_b200 import _lib
    from dualsuperreslearningforsemseg_b200.metrices import _counts
    mIoU, Accuracy = metrics()
    probe = seg_case("plain", 20, (1, 8, 8), 19)
    n0 = _lib.launch_count()
    mIoU(19).update(*probe)
    one_pass = _lib.launch_count() - n0              # launches (kernel + memset node) of ONE counting pass
    assert one_pass >= 1
    for nc, acc_first in ((19, True), (19, False), (6, False)):
        pred, target, mask = seg_case("plain", 21, (2, 37, 53), nc)
        if as_cuda:
            pred_in, target_in, mask_in = (torch.from_numpy(v).cuda() for v in (pred, target, mask))
        else:
            pred_in, target_in, mask_in = pred, target, mask
        _counts.clear_shared_pass()
        m, a = mIoU(nc), Accuracy()
        n0 = _lib.launch_count()
        for meter in ((a, m) if acc_first else (m, a)):
            meter.update(pred_in, target_in, mask_in)
        assert _lib.launch_count() - n0 == one_pass, (nc, acc_first)
        mo, ao = seg_oracle.MIoUOracle(nc), seg_oracle.AccuracyOracle()
        mo.update(pred, target, mask); ao.update(pred, target, mask)
        assert bits(m()) == bits(mo()) and bits(a()) == bits(ao())
    # Accuracy first with another class count than the mIoU that follows: two passes, both right
    pred, target, mask = seg_case("plain", 22, (2, 37, 53), 6)
    _counts.clear_shared_pass()
    m, a = mIoU(6), Accuracy()
    n0 = _lib.launch_count()
    a.update(pred, target, mask); m.update(pred, target, mask)
    assert _lib.launch_count() - n0 == 2 * one_pass
    mo, ao = seg_oracle.MIoUOracle(6), seg_oracle.AccuracyOracle()
    mo.update(pred, target, mask); ao.update(pred, target, mask)
    assert bits(m()) == bits(mo()) and bits(a()) == bits(ao())
    # a modified tensor (new version) or a different array never matches the kept pass
    if as_cuda:
        p, t, k = (torch.from_numpy(v).cuda() for v in (pred, target, mask))
        m2 = mIoU(6)
        m2.update(p, t, k)
        p[0, 0, 0] = (p[0, 0, 0] + 1) % 6
        m2.update(p, t, k)
        mo2 = seg_oracle.MIoUOracle(6)
        mo2.update(pred, target, mask); mo2.update(p.cpu().numpy(), target, mask)
        assert np.array_equal(bits(m2.ious), bits(mo2.ious))


def test_dataset_level_iou_from_kernel_rows():
    """the README's dataset-level definitions (README.md:10-16; intersections and unions summed over ALL updates
    first) from the rows the counts kernel produced, against plain NumPy on the same maps -- including an all-ignored update,
    a class that never occurs and out-of-range labels."""
    mIoU, _ = metrics()
    nc = 19
    m = mIoU(nc)
    tot_i = np.zeros(nc, dtype=np.int64); tot_u = np.zeros(nc, dtype=np.int64)
    for kind, seed, shape in (("plain", 31, (2, 64, 96)), ("all_ignored", 32, (1, 40, 40)), ("oor_target", 33, (2, 37, 53)),
                              ("plain", 34, (3, 33, 31)), ("oor_pred", 35, (2, 21, 35))):
        pred, target, mask = seg_case(kind, seed, shape, nc)
        pred[pred == 7] = 8; target[target == 7] = 8                      # class 7 never occurs: its IoU is 0/0
        mask = mask & (target != 255)
        m.update(torch.from_numpy(pred).cuda(), torch.from_numpy(target).cuda(), torch.from_numpy(mask).cuda())
        for c in range(nc):
            p_c = mask & (pred == c); t_c = target == c
            inter = (p_c & (pred == target)).sum()
            tot_i[c] += inter
            tot_u[c] += p_c.sum() + t_c.sum() - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        pooled = np.float64(tot_i.sum()) / np.float64(tot_u.sum()) * 100.
        per_class = np.nanmean(tot_i / tot_u) * 100.
    got = m.dataset_level()
    assert bits(got[0]) == bits(pooled) and bits(got[1]) == bits(per_class), (got, pooled, per_class)
    assert tot_u[7] == 0 and 0. < got[0] < 100.


@pytest.mark.timeout(900)
def test_config3_all_500_updates_at_full_size():
    """BASELINE configs[2] at its stated size: 500 updates of one 1024 x 2048 map each (int64 pred, uint8 target, bool mask,
    10 % ignored, 70 % correct), fed as ten launches of 50 updates; every integer row, every per-update IoU / accuracy and
    the two final percentages bit for bit against the C restatement of the reference (oracle/seg_counts_ref.c) on the same
    maps; plus the reference's batch-4 grouping (125 updates) on the first 100 maps."""
    import os
    mIoU, Accuracy = metrics()
    nc, H, W = 19, 1024, 2048
    m, a = mIoU(nc), Accuracy()
    m4, a4 = mIoU(nc), Accuracy()
    mo, ao, mo4, ao4 = [], [], [], []
    threads = os.cpu_count()
    g = torch.Generator(device="cuda")
    for chunk in range(10):
        g.manual_seed(54321 + chunk)
        shape = (50, 1, H, W)
        target = torch.randint(0, nc, shape, device="cuda", generator=g, dtype=torch.uint8)
        ign = torch.rand(shape, device="cuda", generator=g) < 0.1
        target.masked_fill_(ign, 255)
        rnd = torch.randint(0, nc, shape, device="cuda", generator=g, dtype=torch.uint8)
        keep = torch.rand(shape, device="cuda", generator=g) < 0.7
        pred = torch.where(keep, torch.where(ign, torch.zeros_like(target), target), rnd).to(torch.int64)
        mask = target != 255
        del rnd, keep, ign
        m.update_many(pred, target, mask)
        a.update_many(pred, target, mask)
        hp, ht, hm = pred.cpu().numpy(), target.cpu().numpy(), mask.cpu().numpy()
        rows = raw_counts(pred, target, mask, nc, updates_leading=True)
        for u in range(50):
            ap, ai, at, c, v = seg_oracle.seg_counts_c(hp[u], ht[u], hm[u], nc, threads=threads)
            assert np.array_equal(rows[u], np.concatenate([ap, ai, at, [c, v]])), (chunk, u)
            mo.append(seg_oracle.iou_from_counts(ap, ai, at)); ao.append(seg_oracle.accuracy_from_counts(c, v))
        if chunk < 2:                                       # batch-4 grouping: (B, H, W) = (4, 1024, 2048) per update
            p4, t4, k4 = (v.view(-1, 4, H, W) for v in (pred[:48], target[:48], mask[:48]))
            m4.update_many(p4, t4, k4); a4.update_many(p4, t4, k4)
            for u in range(12):
                sl = slice(4 * u, 4 * u + 4)
                ap, ai, at, c, v = seg_oracle.seg_counts_c(hp[sl, 0], ht[sl, 0], hm[sl, 0], nc, threads=threads)
                mo4.append(seg_oracle.iou_from_counts(ap, ai, at)); ao4.append(seg_oracle.accuracy_from_counts(c, v))
        del pred, target, mask
    assert len(m.ious) == 500 and np.array_equal(bits(m.ious), bits(mo)) and np.array_equal(bits(a.accuracies), bits(ao))
    assert bits(m()) == bits(np.nanmean(mo) * 100.) and bits(a()) == bits(np.mean(ao) * 100.)
    assert np.array_equal(bits(m4.ious), bits(mo4)) and np.array_equal(bits(a4.accuracies), bits(ao4))
    assert bits(m4()) == bits(np.nanmean(mo4) * 100.) and bits(a4()) == bits(np.mean(ao4) * 100.)
