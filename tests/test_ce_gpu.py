"""GPU parity of the one-pass cross-entropy (csrc/ce_loss.cu) through the drop-in CrossEntropyLoss -> ctypes -> C-ABI:
against the float64 oracle on the golden cases, and against torch's own CUDA implementation at the training shape
(train_or_resume.py:116,435: (B,19,512,1024) logits, uint8 target with 255 = ignore).  Tolerances: loss 1e-6 relative,
gradient 1e-6 relative-norm (fp32 exp/log per pixel, float64 accumulation of the sum)."""
import numpy as np
import pytest
import torch

from _inputs import CE_CASES, ce_case, load_golden
from oracle import ce_oracle

pytestmark = pytest.mark.gpu
G = load_golden("ce_golden.npz")
LOSS_RTOL = 1e-6
GRAD_RTOL = 1e-6


def relnorm(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", sorted(CE_CASES))
def test_ce_matches_oracle_and_golden(name):
    from dualsuperreslearningforsemseg_b200.models.losses import CrossEntropyLoss
    x, t, ignore, red = ce_case(name)
    go = float(G["grad_out"])
    a = torch.from_numpy(x).cuda().requires_grad_(True)
    loss = CrossEntropyLoss(ignore_index=ignore, reduction=red)(a, torch.from_numpy(t).cuda())
    (loss * go).backward()
    ol, og = ce_oracle.cross_entropy(x, t, ignore, red, grad_out=go)
    ref = float(G[f"{name}/loss64"])
    if np.isnan(ref):
        assert np.isnan(float(loss)) and np.isnan(ol)
        assert float(a.grad.abs().max()) == 0.0                 # torch: zero gradient when nothing is valid
    else:
        assert abs(float(loss) - ref) <= LOSS_RTOL * max(abs(ref), 1e-30), (float(loss), ref)
        if np.linalg.norm(og) > 0:
            assert relnorm(a.grad.cpu().numpy(), G[f"{name}/grad64"]) <= GRAD_RTOL
        else:
            assert float(a.grad.abs().max()) == 0.0


@pytest.mark.parametrize("tdt", [torch.uint8, torch.int64])
def test_ce_training_shape_vs_torch_cuda(tdt):
    """(2,19,512,1024): the reference's SSSR output shape at a smaller batch; torch's CUDA cross_entropy as the reference."""
    from dualsuperreslearningforsemseg_b200.models.losses import CrossEntropyLoss
    g = torch.Generator(device="cuda"); g.manual_seed(54321)
    x = torch.randn((2, 19, 512, 1024), device="cuda", generator=g) * 3
    t = torch.randint(0, 19, (2, 512, 1024), device="cuda", generator=g)
    t[torch.rand((2, 512, 1024), device="cuda", generator=g) < 0.1] = 255
    a = x.clone().requires_grad_(True)
    b = x.clone().double().requires_grad_(True)
    l1 = CrossEntropyLoss(ignore_index=255)(a, t.to(tdt))
    (l1 * 0.5).backward()
    l2 = torch.nn.CrossEntropyLoss(ignore_index=255)(b, t)
    (l2 * 0.5).backward()
    assert abs(float(l1) - float(l2)) <= LOSS_RTOL * abs(float(l2)), (float(l1), float(l2))
    assert float((a.grad.double() - b.grad).norm() / b.grad.norm()) <= GRAD_RTOL
    # deterministic: same bits on a second run
    a2 = x.clone().requires_grad_(True)
    l3 = CrossEntropyLoss(ignore_index=255)(a2, t.to(tdt))
    (l3 * 0.5).backward()
    assert float(l3) == float(l1) and torch.equal(a2.grad, a.grad)


def test_ce_out_of_range_target_is_loud_and_inf_logits():
    from dualsuperreslearningforsemseg_b200.models.losses import CrossEntropyLoss
    x = torch.randn((1, 5, 4, 8), device="cuda")
    t = torch.randint(0, 5, (1, 4, 8), device="cuda")
    t[0, 1, 1] = 0
    t2 = t.clone(); t2[0, 0, 0] = 77                               # torch would raise a device assert here
    t3 = t.clone(); t3[0, 0, 0] = 255
    f = CrossEntropyLoss(ignore_index=255)
    bad = x.clone().requires_grad_(True)
    lb = f(bad, t2)
    lb.backward()
    assert bool(torch.isnan(lb)) and bool(torch.isfinite(f(x, t3)))       # a label-mapping bug cannot train silently
    assert float(bad.grad[0, :, 0, 0].abs().sum()) == 0.0                  # ... and that pixel gets no gradient
    x[0, 2, 1, 1] = float("-inf")                                  # a -inf logit is a zero-probability class
    a = x.clone().requires_grad_(True)
    l = f(a, t3); l.backward()
    ref = torch.nn.functional.cross_entropy(x.double(), t3, ignore_index=255)
    assert abs(float(l) - float(ref)) <= 1e-6 * abs(float(ref)) and bool(torch.isfinite(a.grad).all())
