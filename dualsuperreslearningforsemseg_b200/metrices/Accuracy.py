"""Mean pixel accuracy -- B200 drop-in for the reference's ``metrices/Accuracy.py:4-30``.

Same API (``update(pred, target, valid_labels_mask)``, ``()`` -> percent, ``reset()``, ``accuracies``).  The two
reductions of Accuracy.py:19-20 come out of the same K4 pass that serves ``mIoU`` (when both meters are fed the
same arrays back to back, as at train_or_resume.py:480-481 and benchmark.py:76-77, the second one re-uses the first
one's pass whichever runs first: correct / valid do not depend on the number of classes, see _counts._SharedPass);
the division and the mean are the reference's own float64 NumPy expressions (Accuracy.py:24,29).
"""
import numpy as np

from . import _counts


class Accuracy:
    def __init__(self, num_classes: int = _counts.NC_DEFAULT):
        # the reference's Accuracy() takes no argument; the class count only sets the row width of a pass this meter starts
        self.num_classes = num_classes
        self.reset()

    def reset(self):
        self.dirty = False
        self.mean_accuracy = 0.0
        self._accuracies = []
        self._pending = _counts.PendingRows()

    def update(self, pred, target, valid_labels_mask):
        self.dirty = True
        rows, nc = _counts.counts_for_update(pred, target, valid_labels_mask, self.num_classes, any_nc=True)
        self._pending.add(rows[:, 3 * nc:3 * nc + 2])

    def update_many(self, pred, target, valid_labels_mask):
        self.dirty = True
        rows, nc = _counts.counts_for_update(pred, target, valid_labels_mask, self.num_classes, updates_leading=True, any_nc=True)
        self._pending.add(rows[:, 3 * nc:3 * nc + 2])

    def sync(self, group=None, mode="sum", offset=0, total=0):
        _counts.sync_rows(self._pending, group, mode, offset, total)

    def _finish(self):
        if len(self._pending) == 0:
            return
        for pixels_correct, total_pixels in self._pending.drain():
            assert pixels_correct <= total_pixels, "BUG CHECK: 'pixels_correct' cannot be be greater than 'total_pixels'."
            with np.errstate(divide='ignore', invalid='ignore'):
                self._accuracies.append(pixels_correct / total_pixels)      # np.int64 / np.int64 -> np.float64

    @property
    def accuracies(self):
        self._finish()
        return self._accuracies

    def __call__(self):
        if self.dirty:
            self.dirty = False
            self.mean_accuracy = (np.mean(self.accuracies) * 100.)
        return self.mean_accuracy
