"""N>1 path on CPU: world_size-2 gloo.  The per-update integer rows of two ranks (each holding half of every
update's batch) all-reduce to exactly the rows of the unsharded run, so mIoU/accuracy are bit-identical; the
'gather' mode concatenates rank-major, the 'place' mode (one all-reduce per pass) returns the rows in global update order."""
import os
import socket
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _inputs import seg_case, load_golden
from oracle import seg_oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _row(pred, target, mask, nc):
    ap, ai, at, c, v = seg_oracle.seg_counts(pred, target, mask, nc)
    return torch.from_numpy(np.concatenate([ap, ai, at, [c, v]]).astype(np.int64))[None]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dualsuperreslearningforsemseg_b200.metrices import mIoU, Accuracy
    from dualsuperreslearningforsemseg_b200 import distributed as D
    nc = 19
    m, a = mIoU(nc), Accuracy()
    mg = mIoU(nc)
    for seed in (21, 22, 23):
        pred, target, mask = seg_case("plain", seed, (4, 33, 47), nc)
        sl = D.shard_slice(pred.shape[0], rank, world)
        row = _row(pred[sl], target[sl], mask[sl], nc)
        m._pending.add(row); m.dirty = True
        a._pending.add(row[:, 3 * nc:]); a.dirty = True
    for seed in (21, 22, 23, 24, 25):     # 'gather' mode: whole updates live on different ranks (3 on one, 2 on the other)
        if seed % world == rank:
            mg._pending.add(_row(*seg_case("plain", seed, (4, 33, 47), nc), nc)); mg.dirty = True
    # 'place' mode: the ranks know where their updates sit in the pass (rank 0: updates 0-2, rank 1: updates 3-4) -> one
    # all-reduce of the [5, row] table, rows in global order on both ranks; run twice (the table is kept between passes)
    mp_ = mIoU(nc)
    for _ in range(2):
        mp_.reset()
        seeds = (21, 22, 23) if rank == 0 else (24, 25)
        for seed in seeds:
            mp_._pending.add(_row(*seg_case("plain", seed, (4, 33, 47), nc), nc)); mp_.dirty = True
        mp_.sync(mode="place", offset=0 if rank == 0 else 3, total=5)
    m.sync(mode="sum"); a.sync(mode="sum"); mg.sync(mode="gather")
    loss = D.all_reduce_mean_loss(torch.tensor(float(rank + 1)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        q.put((rank, m(), a(), list(m.ious), sorted(mg.ious), float(loss), list(mp_.ious)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_metric_sync_is_bit_exact():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # single-process truth
    nc = 19
    m, a = seg_oracle.MIoUOracle(nc), seg_oracle.AccuracyOracle()
    for seed in (21, 22, 23):
        m.update(*seg_case("plain", seed, (4, 33, 47), nc))
        a.update(*seg_case("plain", seed, (4, 33, 47), nc))
    mg = seg_oracle.MIoUOracle(nc)
    for seed in (21, 22, 23, 24, 25):
        mg.update(*seg_case("plain", seed, (4, 33, 47), nc))
    for rank, miou, acc, ious, gious, loss, pious in res:
        assert miou == m() and acc == a() and ious == m.ious
        assert gious == sorted(mg.ious)
        assert pious == mg.ious                       # global update order, not rank-major
        assert loss == 1.5
