"""BASELINE configs[4] in miniature: a stage-3 DSRL training step (harness/) with the drop-in FALoss against the same step
with a plain-PyTorch restatement of the reference FALoss (oracle/fa_torch_port.py) applied to the same forward graph.  The FA term and the
gradients it sends into the feature transformers / SISR decoder must agree.  Also runs a full optimiser step and the
metrics on the step's logits."""
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
    step = Stage3Step(FALoss(), dev)
    # ONE forward graph (cuDNN may pick different algorithms for two forwards, so two model copies are not bit-comparable);
    # both FA implementations are applied to the same feature-transformer outputs and back-propagated through it
    ce, mse, fa_ours, o = step.losses(img, org, target)
    assert o[2].shape == (2, 1, 64, 128) and o[3].shape == (2, 1, 64, 128)    # the FA inputs of the real model
    fa_ref = TorchFALoss()(o[2], o[3])
    assert abs(float(fa_ours) - float(fa_ref)) <= 1e-4 * abs(float(fa_ref)), (float(fa_ours), float(fa_ref))
    params = dict(step.core.named_parameters())
    names = ("SSSR_feature_transformer.0.weight", "SISR_feature_transformer.0.weight", "SISR_decoder.0.weight")
    g_ours = torch.autograd.grad(fa_ours, [params[n] for n in names], retain_graph=True)
    g_ref = torch.autograd.grad(fa_ref, [params[n] for n in names], retain_graph=True)
    for n, g1, g2 in zip(names, g_ours, g_ref):
        rel = float((g1 - g2).norm() / g2.norm())
        assert rel <= 1e-3, (n, rel)
    # the whole objective back-propagates and a full optimiser step runs (train_or_resume.py:438-445)
    (ce + mse + fa_ours).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in step.core.parameters())
    losses = step(img, org, target)
    assert all(np.isfinite(float(x)) for x in losses)
    # the metrics take the step's logits without leaving the device (train_or_resume.py:476-481)
    m, a = mIoU(19), Accuracy()
    pred = torch.argmax(o[0].detach(), dim=1)
    m.update(pred, target, target != 255)
    a.update(pred, target, target != 255)
    mo, ao = seg_oracle.MIoUOracle(19), seg_oracle.AccuracyOracle()
    p, t = pred.cpu().numpy(), target.cpu().numpy()
    mo.update(p, t, t != 255)
    ao.update(p, t, t != 255)
    assert m() == mo() and a() == ao()


def test_stage3_step_with_fused_losses_matches_the_unfused_step():
    """The same model and batch through the unfused losses (torch CE / MSE, transformers in the model, drop-in FALoss) and through
    Stage3Loss: the three losses and the gradients that reach the transformers and both decoders agree."""
    from harness.train_step import Stage3Step, synthetic_batch
    from dualsuperreslearningforsemseg_b200.models.losses import FALoss
    dev = torch.device("cuda", 0)
    img, org, target = synthetic_batch(2, dev, 2)
    plain, fused = Stage3Step(FALoss(), dev), Stage3Step(FALoss(), dev, fused_losses=True)      # same seed: same weights
    fused.core.load_state_dict(plain.core.state_dict())
    # one forward graph for both loss paths (cuDNN algorithm choice): evaluate the decoders once, detach, feed both
    with torch.no_grad():
        sssr, sisr, _, _ = plain.core(img, apply_transformers=False)
    a1, b1 = sssr.clone().requires_grad_(True), sisr.clone().requires_grad_(True)
    a2, b2 = sssr.clone().requires_grad_(True), sisr.clone().requires_grad_(True)
    ce1 = plain.ce(a1, target.long()); mse1 = plain.mse(b1, org)
    fa1 = plain.fa(plain.core.SSSR_feature_transformer(a1), plain.core.SISR_feature_transformer(b1))
    (ce1 + 0.1 * mse1 + 1.0 * fa1).backward()
    ce2, mse2, fa2 = fused.stage3(a2, b2, target, org)
    (ce2 + 0.1 * mse2 + 1.0 * fa2).backward()
    for x, y, tol in ((ce1, ce2, 1e-5), (mse1, mse2, 1e-5), (fa1, fa2, 1e-4)):
        assert abs(float(x) - float(y)) <= tol * abs(float(x)), (float(x), float(y))
    for g1, g2 in ((a1.grad, a2.grad), (b1.grad, b2.grad)):
        assert float((g1 - g2).norm() / g1.norm()) <= 1e-3
    for n in ("SSSR_feature_transformer.0.weight", "SISR_feature_transformer.0.weight"):
        g1, g2 = dict(plain.core.named_parameters())[n].grad, dict(fused.core.named_parameters())[n].grad
        assert float((g1 - g2).norm() / g1.norm()) <= 1e-3, n
    # and a whole optimiser step runs through the fused path
    losses = fused(img, org, target)
    assert all(np.isfinite(float(x)) for x in losses)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in fused.core.parameters())
