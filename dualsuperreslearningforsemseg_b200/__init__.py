"""dsrl-b200: B200 (sm_100a) implementation of the DSRL Feature-Affinity loss and mIoU/accuracy counts.

Drop-in surfaces (same names / arguments as the reference):

    from dualsuperreslearningforsemseg_b200.models.losses import FALoss
    from dualsuperreslearningforsemseg_b200.metrices import mIoU, Accuracy, AverageMeter
"""
from . import _lib  # noqa: F401
from .models.losses import FALoss  # noqa: F401
from .metrices import mIoU, Accuracy, AverageMeter  # noqa: F401

__version__ = "0.1.0"
