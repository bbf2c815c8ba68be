"""Running weighted average -- behaviour-compatible with the reference's ``metrices/AverageMeter.py:4-27``
(``update(val, n=1)``, ``()`` -> sum/count with division-by-zero silenced, ``reset()``; attributes ``val``,
``avg``, ``sum``, ``count``, ``dirty``).  Pure host code; nothing here touches the GPU."""
import numpy as np


class AverageMeter:
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0
        self.dirty = False

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.dirty = True

    def __call__(self):
        if not self.dirty:
            return self.avg
        self.dirty = False
        with np.errstate(divide='ignore', invalid='ignore'):    # empty meter -> nan/inf, no warning (AverageMeter.py:25)
            self.avg = self.sum / self.count
        return self.avg
