"""The product's multi-GPU API on NCCL with the real kernels, two ranks: batch-sharded FA loss with the scalar
loss all-reduce (`distributed.all_reduce_mean_loss`), and the per-update count rows of a validation pass exchanged ONCE
(`mIoU.sync` / `Accuracy.sync`, modes 'sum' and 'place') -- bit-exact against the single-process oracle.  Needs two GPUs;
skipped (not failed) on a one-GPU box.  Mirrors tests/test_distributed_cpu.py (gloo, synthetic rows)."""
import os
import socket
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import torch.distributed as dist
    from _inputs import seg_case, pos_inputs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from dualsuperreslearningforsemseg_b200.metrices import mIoU, Accuracy
    from dualsuperreslearningforsemseg_b200.models.losses import FALoss
    from dualsuperreslearningforsemseg_b200 import distributed as D
    nc = 19
    # (1) every rank holds half of each update's batch -> 'sum'
    m, a = mIoU(nc), Accuracy()
    for seed in (21, 22, 23):
        pred, target, mask = seg_case("plain", seed, (4, 33, 47), nc)
        sl = D.shard_slice(pred.shape[0], rank, world)
        p, t, k = (torch.from_numpy(np.ascontiguousarray(v[sl])).to(dev) for v in (pred, target, mask))
        a.update(p, t, k); m.update(p, t, k)
    m.sync(mode="sum"); a.sync(mode="sum")
    # (2) whole updates on different ranks, positions known -> 'place' (one all-reduce of the [5, 59] table), twice
    mp_ = mIoU(nc)
    for _ in range(2):
        mp_.reset()
        for seed in ((21, 22, 23) if rank == 0 else (24, 25)):
            pred, target, mask = seg_case("plain", seed, (4, 33, 47), nc)
            mp_.update(torch.from_numpy(pred).to(dev), torch.from_numpy(target).to(dev), torch.from_numpy(mask).to(dev))
        mp_.sync(mode="place", offset=0 if rank == 0 else 3, total=5)
    # (3) FA loss, batch of 4 sharded 2 + 2: local mean losses -> global mean; gradients are per-sample (no exchange)
    x1, x2 = pos_inputs((4, 64, 16, 32), (4, 64, 16, 32), 77)
    sl = D.shard_slice(4, rank, world)
    u = torch.from_numpy(x1[sl]).to(dev).requires_grad_(True)
    v = torch.from_numpy(x2[sl]).to(dev).requires_grad_(True)
    loss = FALoss(subsample_factor=1, affinity="position", precision="f16")(u, v)
    loss.backward()
    gl = D.all_reduce_mean_loss(loss)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        q.put((rank, m(), a(), list(m.ious), list(mp_.ious), float(gl), u.grad.cpu().numpy(), v.grad.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_nccl_fa_and_metric_sync():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from _inputs import seg_case, pos_inputs
    from oracle import seg_oracle, fa_oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nc = 19
    m, a = seg_oracle.MIoUOracle(nc), seg_oracle.AccuracyOracle()
    for seed in (21, 22, 23):
        m.update(*seg_case("plain", seed, (4, 33, 47), nc)); a.update(*seg_case("plain", seed, (4, 33, 47), nc))
    mg = seg_oracle.MIoUOracle(nc)
    for seed in (21, 22, 23, 24, 25):
        mg.update(*seg_case("plain", seed, (4, 33, 47), nc))
    x1, x2 = pos_inputs((4, 64, 16, 32), (4, 64, 16, 32), 77)
    ol, o1, o2 = fa_oracle.fa_position(x1, x2, 1, "mean")
    for rank, miou, acc, ious, pious, gl, g1, g2 in res:
        assert miou == m() and acc == a() and ious == m.ious and pious == mg.ious
        assert abs(gl - ol) <= 1e-4 * abs(ol), (gl, ol)
        # local-mean gradients are 2x the global-mean ones (half the batch per rank)
        sl = slice(2 * rank, 2 * rank + 2)
        for g, o in ((g1, o1[sl]), (g2, o2[sl])):
            assert np.linalg.norm(g - 2.0 * o) <= 1e-3 * np.linalg.norm(2.0 * o)
