"""Mean intersection-over-union -- B200 drop-in for the reference's ``metrices/mIoU.py:5-41``.

Same API: ``mIoU(num_classes)``, ``update(pred, target, valid_labels_mask)`` with ``(B, H, W)`` label maps,
``()`` -> percent, ``reset()``, attribute ``ious`` (one float64 per update).  The three ``np.histogram`` passes
(mIoU.py:27-29) run as ONE pass of the K4 CUDA kernel producing exact int64 counts; the float64 finish
below uses the reference's own expressions (mIoU.py:30-35,40) on those integers, so ``ious`` and the final
value are bit-identical to the reference.  Inputs may be NumPy arrays (as the reference's callers pass,
train_or_resume.py:477-481) or CUDA tensors (no D2H at all).  Counts stay on the device until a number is
asked for, so ``update()`` never synchronises.
"""
import warnings

import numpy as np

from . import _counts


class mIoU:
    def __init__(self, num_classes):
        self.num_classes = num_classes
        self.reset()

    def reset(self):
        self.dirty = False
        self.miou = 0.0
        self._ious = []
        self._pending = _counts.PendingRows()
        self._tot_inter = np.zeros(self.num_classes, dtype=np.int64)     # dataset-level sums (dataset_level())
        self._tot_union = np.zeros(self.num_classes, dtype=np.int64)

    # -- updates ------------------------------------------------------------------------------------------
    def update(self, pred, target, valid_labels_mask):
        self.dirty = True
        self._pending.add(_counts.counts_for_update(pred, target, valid_labels_mask, self.num_classes)[0])

    def update_many(self, pred, target, valid_labels_mask):
        """U updates in one launch: arrays shaped (U, B, H, W); equivalent to U consecutive ``update`` calls."""
        self.dirty = True
        self._pending.add(_counts.counts_for_update(pred, target, valid_labels_mask, self.num_classes, updates_leading=True)[0])

    def update_from_logits(self, logits, target, valid_labels_mask=None, return_pred=False):
        """Fused ``argmax(logits, dim=1)`` + update (replaces benchmark.py:61-77's D2H + host argmax)."""
        self.dirty = True
        rows, pred = _counts.counts_from_logits(logits, target, valid_labels_mask, self.num_classes, want_pred=return_pred)
        self._pending.add(rows)
        return pred

    def sync(self, group=None, mode="sum", offset=0, total=0):
        """All-reduce / all-gather the pending per-update rows across the process group, once per validation pass (see
        _counts.sync_rows: 'sum', 'place' -- this rank's updates start at global index ``offset`` of ``total`` -- or 'gather')."""
        _counts.sync_rows(self._pending, group, mode, offset, total)

    # -- results ------------------------------------------------------------------------------------------
    def _finish(self):
        if len(self._pending) == 0:
            return
        nc = self.num_classes
        for row in self._pending.drain():
            area_pred, area_inter, area_target = row[:nc], row[nc:2 * nc], row[2 * nc:3 * nc]
            area_union = area_pred + area_target - area_inter
            assert (area_inter <= area_union).all(), "BUG CHECK: Intersection area should always be less than or equal to union area."
            with np.errstate(divide='ignore', invalid='ignore'), warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                self._ious.append(np.nanmean(area_inter / area_union))
            self._tot_inter += area_inter
            self._tot_union += area_union

    @property
    def ious(self):
        self._finish()
        return self._ious

    def dataset_level(self):
        """The two dataset-level definitions the reference's README quotes next to its results (README.md:10-16) but
        does not implement in ``metrices``: intersections and unions summed over ALL updates first.  Returns percent
        ``(sum_c I_c / sum_c U_c, nanmean_c(I_c / U_c))``.  ``__call__`` keeps the reference's per-update mean."""
        self._finish()
        with np.errstate(divide='ignore', invalid='ignore'), warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            pooled = np.float64(self._tot_inter.sum()) / np.float64(self._tot_union.sum()) * 100.
            per_class = np.nanmean(self._tot_inter / self._tot_union) * 100.
        return pooled, per_class

    def __call__(self):
        if self.dirty:
            self.dirty = False
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                self.miou = (np.nanmean(self.ious) * 100.)
        return self.miou
