"""Drop-in for the reference's ``metrices`` package (metrices/__init__.py:1-3): same three names."""
from .AverageMeter import AverageMeter  # noqa: F401
from .mIoU import mIoU  # noqa: F401
from .Accuracy import Accuracy  # noqa: F401
