"""One-call activation of the B200 hot path inside the reference code base.

The reference's scripts bind the hot path at import time::

    from models.losses import FALoss          # command_handlers/train_or_resume.py:14
    from metrices import *                    # command_handlers/train_or_resume.py:17, command_handlers/benchmark.py:10

``patch_reference()`` must therefore run before ``command_handlers`` is imported -- e.g. as the first statement of the
reference's ``main.py`` (with the reference tree on ``sys.path``)::

    import dualsuperreslearningforsemseg_b200.dropin as dropin; dropin.patch_reference()

It imports the reference's own ``models.losses`` and ``metrices`` packages and rebinds the four public names to the
classes of this package; nothing else of the reference is touched (same class names, constructors and methods, see
INTEGRATION.md).  ``unpatch_reference()`` restores the originals.
"""
from __future__ import annotations

import importlib

from .metrices import Accuracy, AverageMeter, mIoU
from .models.losses import FALoss

_ORIGINALS = {}
_TARGETS = (("models.losses", "FALoss", FALoss), ("models.losses.FALoss", "FALoss", FALoss),
            ("metrices", "mIoU", mIoU), ("metrices", "Accuracy", Accuracy), ("metrices", "AverageMeter", AverageMeter))


def patch_reference() -> list[str]:
    """Rebinds ``models.losses.FALoss`` and ``metrices.{mIoU,Accuracy,AverageMeter}`` of the reference (which must be
    importable) to this package's classes.  Returns the dotted names that were replaced."""
    done = []
    for mod_name, attr, cls in _TARGETS:
        mod = importlib.import_module(mod_name)
        key = (mod_name, attr)
        if key not in _ORIGINALS:
            _ORIGINALS[key] = getattr(mod, attr)
        setattr(mod, attr, cls)
        done.append(f"{mod_name}.{attr}")
    return done


def unpatch_reference() -> None:
    for (mod_name, attr), orig in _ORIGINALS.items():
        setattr(importlib.import_module(mod_name), attr, orig)
    _ORIGINALS.clear()
