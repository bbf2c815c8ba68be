"""Host side of the segmentation-count kernel (K4): tensor plumbing around ``dsrl_seg_counts`` and the
deferred, exact float64 finish shared by ``mIoU`` and ``Accuracy``."""
from __future__ import annotations

import ctypes
import threading

import numpy as np
import torch

from .. import _lib

_DT = {torch.uint8: _lib.U8, torch.int32: _lib.I32, torch.int64: _lib.I64}
IGNORE_DEFAULT = 255          # datasets/Cityscapes/settings.py IGNORE_CLASS_LABEL
NC_DEFAULT = 19               # datasets/Cityscapes/settings.py NUM_CLASSES (the row width Accuracy asks for when it runs first)


class _SharedPass(threading.local):
    """The last counting pass of THIS thread, kept so that the second meter of the reference's back-to-back pair

        mean_accuracy.update(pred, target, mask); miou.update(pred, target, mask)      (train_or_resume.py:480-481)
        miou.update(pred, target, mask); accuracy_mean.update(pred, target, mask)      (benchmark.py:76-77)

    re-uses the first one's kernel pass (and, for NumPy inputs, its host -> device copies) instead of repeating them.
    Contract: the entry matches only the very same Python objects (`is`; the entry holds references, so ids cannot be
    recycled), the same tensor versions, the same ignore label and -- for mIoU -- the same number of classes; callers must
    not modify a NumPy array in place between the two calls (the reference's loops do not).  Per thread, no module state
    shared between threads; rows consumed on another stream wait for the event recorded after the producing launch."""

    def __init__(self):
        self.entry = None


_shared = _SharedPass()


def clear_shared_pass():
    _shared.entry = None


def _versions(objs):
    return tuple(o._version if torch.is_tensor(o) else None for o in objs)


def row_len(nc: int) -> int:
    return 3 * nc + 2


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("dsrl-b200 metrics need a CUDA device: there is no CPU fallback on this path")
    return torch.device("cuda", torch.cuda.current_device())


def _as_cuda(x, dev, labels: bool):
    """numpy / CPU tensor / CUDA tensor -> contiguous CUDA tensor of a dtype the kernel reads natively."""
    if isinstance(x, np.ndarray):
        if x.dtype == np.bool_:
            x = x.view(np.uint8)
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif not torch.is_tensor(x):
        x = torch.as_tensor(np.asarray(x))
    if x.dtype == torch.bool:
        x = x.view(torch.uint8) if x.is_contiguous() else x.contiguous().view(torch.uint8)
    if labels and x.dtype not in _DT:
        if x.dtype in (torch.int8, torch.int16):
            x = x.to(torch.int32)
        else:
            raise TypeError(f"label maps must be integer typed, got {x.dtype}")
    if not labels and x.dtype != torch.uint8:
        x = (x != 0).view(torch.uint8)
    if not x.is_cuda:
        x = x.to(dev, non_blocking=True)
    return x.contiguous()


def counts_for_update(pred, target, mask, nc: int, updates_leading: bool = False, ignore_label: int = IGNORE_DEFAULT,
                      any_nc: bool = False):
    """Enqueue K4 for one ``update()`` (or U of them if ``updates_leading``) and return ``(rows, nc_of_rows)``: the device
    rows ``[U, 3*nc_of_rows+2]`` int64, without synchronising.  ``any_nc``: the caller only needs the last two columns
    (correct, valid), which do not depend on the number of classes, so a shared pass of any width will do."""
    shp = tuple(np.shape(pred)) if not torch.is_tensor(pred) else tuple(pred.shape)
    tshp = tuple(np.shape(target)) if not torch.is_tensor(target) else tuple(target.shape)
    want = 4 if updates_leading else 3
    # same BUG CHECKs as mIoU.py:16-17 / Accuracy.py:14-15
    assert shp == tshp, "BUG CHECK: 'pred' and 'target' must be of the same shape of (B, H, W)."
    assert len(shp) == want, "BUG CHECK: 'target' and 'pred' must be (B, H, W) channel-order dimensions."
    dev = pred.device if torch.is_tensor(pred) and pred.is_cuda else (
        target.device if torch.is_tensor(target) and target.is_cuda else _device())
    objs = (pred, target, mask)
    e = _shared.entry
    if (e is not None and all(a is b for a, b in zip(e["objs"], objs)) and e["versions"] == _versions(objs)
            and e["leading"] == updates_leading and e["ignore"] == (ignore_label if mask is None else None)
            and e["dev"] == dev and (any_nc or e["nc"] == nc)):
        cur = torch.cuda.current_stream(dev)
        if cur != e["stream"]:
            cur.wait_event(e["event"])
        return e["rows"], e["nc"]
    p = _as_cuda(pred, dev, True)
    t = _as_cuda(target, dev, True)
    m = _as_cuda(mask, dev, False) if mask is not None else None
    if m is not None and tuple(m.shape) != shp:
        m = m.expand(shp).contiguous()
    U = shp[0] if updates_leading else 1
    npix = int(np.prod(shp[1:] if updates_leading else shp, dtype=np.int64))
    if npix == 0 or U == 0:          # nothing to read: torch gives empty tensors a null data pointer
        return torch.zeros((U, row_len(nc)), dtype=torch.int64, device=dev), nc
    rows = torch.empty((U, row_len(nc)), dtype=torch.int64, device=dev)
    cur = torch.cuda.current_stream(dev)
    stream = ctypes.c_void_p(cur.cuda_stream)
    with torch.cuda.device(dev), _lib.nvtx_range("dsrl.seg_counts"):
        _lib.check(_lib.lib().dsrl_seg_counts(ctypes.c_void_p(p.data_ptr()), _DT[p.dtype], ctypes.c_void_p(t.data_ptr()),
                                              _DT[t.dtype], ctypes.c_void_p(m.data_ptr()) if m is not None else None,
                                              U, npix, nc, ignore_label, ctypes.c_void_p(rows.data_ptr()), stream))
    # keep the device inputs alive until the rows are consumed (the kernel is only enqueued) and remember the pass for
    # the other meter of the pair (see _SharedPass)
    rows._dsrl_keepalive = (p, t, m)
    event = torch.cuda.Event()
    event.record(cur)
    _shared.entry = {"objs": objs, "versions": _versions(objs), "leading": updates_leading,
                     "ignore": ignore_label if mask is None else None, "dev": dev, "nc": nc, "rows": rows, "stream": cur,
                     "event": event}
    return rows, nc


def counts_from_logits(logits, target, mask, nc: int, ignore_label: int = IGNORE_DEFAULT, want_pred: bool = False,
                       updates_leading: bool = False):
    """Fused argmax + counts: logits (B, NC, H, W) fp32 CUDA, target (B, H, W) -> ONE update; with ``updates_leading``
    logits (U, B, NC, H, W), target (U, B, H, W) -> U updates in one launch."""
    want = 5 if updates_leading else 4
    if not (torch.is_tensor(logits) and logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == want):
        raise TypeError(f"logits must be a {want}-D float32 CUDA tensor ({'U, ' if updates_leading else ''}B, NC, H, W)")
    U = logits.shape[0] if updates_leading else 1
    B, C, H, W = logits.shape[-4:]
    assert C == nc, "BUG CHECK: logits channel count must equal num_classes."
    assert tuple(target.shape) == tuple(logits.shape[:-3]) + (H, W), "BUG CHECK: 'pred' and 'target' must be of the same shape of (B, H, W)."
    dev = logits.device
    lg = logits.contiguous()
    t = _as_cuda(target, dev, True)
    m = _as_cuda(mask, dev, False) if mask is not None else None
    rows = torch.empty((U, row_len(nc)), dtype=torch.int64, device=dev)
    pred = torch.empty(tuple(target.shape), dtype=torch.int64, device=dev) if want_pred else None
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev), _lib.nvtx_range("dsrl.seg_counts_from_logits"):
        _lib.check(_lib.lib().dsrl_seg_counts_from_logits(
            ctypes.c_void_p(lg.data_ptr()), ctypes.c_void_p(t.data_ptr()), _DT[t.dtype],
            ctypes.c_void_p(m.data_ptr()) if m is not None else None, U, B, H * W, nc, ignore_label,
            ctypes.c_void_p(rows.data_ptr()), ctypes.c_void_p(pred.data_ptr()) if pred is not None else None, stream))
    rows._dsrl_keepalive = (lg, t, m)
    return rows, pred


class PendingRows:
    """Device count rows not yet brought to the host.  One D2H + sync when the owner needs numbers."""

    def __init__(self):
        self.chunks = []

    def add(self, rows):
        self.chunks.append(rows)

    def __len__(self):
        return sum(int(c.shape[0]) for c in self.chunks)

    def device_table(self):
        if not self.chunks:
            return None
        return self.chunks[0] if len(self.chunks) == 1 else torch.cat(self.chunks, dim=0)

    def replace(self, table):
        self.chunks = [table] if table is not None else []

    def drain(self) -> np.ndarray:
        table = self.device_table()
        self.chunks = []
        if table is None:
            return np.zeros((0, 0), dtype=np.int64)
        return table.cpu().numpy()


def gather_rows(table, group=None):
    """All-gather of per-update rows, rank-major; ranks may hold different numbers of updates (e.g. 500 updates over
    8 ranks): the row counts are exchanged first and the tables padded to the longest."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([table.shape[0]], dtype=torch.int64, device=table.device)
    counts = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    width = table.shape[1]
    padded = torch.zeros((max(counts), width), dtype=table.dtype, device=table.device)
    padded[: table.shape[0]] = table
    outs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(outs, padded, group=group)
    return torch.cat([o[:c] for o, c in zip(outs, counts)], dim=0)


_place_tables = {}


def sync_rows(pending: PendingRows, group=None, mode: str = "sum", offset: int = 0, total: int = 0):
    """Multi-GPU exchange for the per-update rows, once per validation pass.  ``sum``: every rank processed a
    shard of each update's pixels (same number of updates everywhere) -> element-wise int64 all-reduce, bit-exact at any
    world size.  ``place``: ranks processed different updates of a pass of ``total`` updates, this rank's starting at
    global index ``offset`` -> ONE all-reduce of the zero-initialised ``[total, row]`` table every rank writes its rows into
    (236 KB for the 500 updates of BASELINE configs[2]); rows come out in global update order on every rank; the table is
    kept between passes.  ``gather``: like ``place`` when the ranks do not know their offsets -> row counts exchanged
    first, padded all-gather, rank-major order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return
    table = pending.device_table()
    if mode == "place":
        if total <= 0:
            raise ValueError("mode 'place' needs the total number of updates of the pass")
        if table is None:
            raise ValueError("mode 'place': this rank has no rows; give it an empty update or use 'gather'")
        n, width = int(table.shape[0]), int(table.shape[1])
        if offset < 0 or offset + n > total:
            raise ValueError(f"mode 'place': rows [{offset}, {offset + n}) do not fit a pass of {total} updates")
        key = (table.device, total, width)
        full = _place_tables.get(key)
        if full is None:
            full = _place_tables[key] = torch.zeros((total, width), dtype=torch.int64, device=table.device)
        else:
            full.zero_()
        full[offset:offset + n].copy_(table)
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
        pending.replace(full.clone())
        return
    if table is None:
        return
    if mode == "sum":
        table = table.clone()
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
    elif mode == "gather":
        table = gather_rows(table, group)
    else:
        raise ValueError("mode must be 'sum', 'gather' or 'place'")
    pending.replace(table)
