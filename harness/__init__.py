"""Caller-side harness for BASELINE configs[4]: a stage-3 DSRL model and training step AROUND the hot path.
Plain PyTorch host code (cuDNN convolutions, torch DDP) -- not part of the product package; it exists so that the drop-in
FALoss / mIoU / Accuracy can be exercised and timed in the context the reference calls them from
(command_handlers/train_or_resume.py:404-481)."""
