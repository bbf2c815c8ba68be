"""BASELINE configs[4] in miniature: a stage-3 DSRL training step (harness/) with the drop-in FALoss against the same step
with a plain-PyTorch restatement of the reference FALoss (oracle/fa_torch_port.py) -- identical weights and inputs.
The FA term and the gradients it sends into the two feature transformers must agree; CE / MSE are bit-identical by
construction.  Also runs the metrics on the step's logits."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class TorchFALoss(torch.nn.Module):
    def forward(self, a, b):
        from oracle import fa_torch_port
        return fa_torch_port.fa_loss(a, b, 8, "mean")


def test_stage3_step_matches_pytorch_fa():
    from harness.train_step import Stage3Step, synthetic_batch
    from dualsuperreslearningforsemseg_b200.models.losses import FALoss
    from dualsuperreslearningforsemseg_b200.metrices import mIoU, Accuracy
    from oracle import seg_oracle
    dev = torch.device("cuda", 0)
    img, org, target = synthetic_batch(2, dev, 1)
    ours, ref = Stage3Step(FALoss(), dev), Stage3Step(TorchFALoss(), dev)
    ref.core.load_state_dict(ours.core.state_dict())
    # dropout makes two training-mode forwards differ: draw the same masks
    outs = []
    for step in (ours, ref):
        torch.manual_seed(7)
        ce, mse, fa, o = step.losses(img, org, target)
        (ce + mse + fa).backward()
        outs.append((ce, mse, fa, o))
    (ce1, mse1, fa1, o1), (ce2, mse2, fa2, _) = outs
    assert float(ce1) == float(ce2) and float(mse1) == float(mse2)
    assert o1[2].shape == (2, 1, 64, 128)                                     # the FA inputs of the real model
    assert abs(float(fa1) - float(fa2)) <= 1e-4 * abs(float(fa2)), (float(fa1), float(fa2))
    for name in ("SSSR_feature_transformer.0.weight", "SISR_feature_transformer.0.weight"):
        g1 = dict(ours.core.named_parameters())[name].grad
        g2 = dict(ref.core.named_parameters())[name].grad
        rel = float((g1 - g2).norm() / g2.norm())
        assert rel <= 1e-3, (name, rel)
    # a full optimiser step runs, and the metrics take the step's logits without leaving the device
    ours(img, org, target)
    m, a = mIoU(19), Accuracy()
    pred = torch.argmax(o1[0].detach(), dim=1)
    m.update(pred, target, target != 255)
    a.update(pred, target, target != 255)
    mo, ao = seg_oracle.MIoUOracle(19), seg_oracle.AccuracyOracle()
    p, t = pred.cpu().numpy(), target.cpu().numpy()
    mo.update(p, t, t != 255)
    ao.update(p, t, t != 255)
    assert m() == mo() and a() == ao()
