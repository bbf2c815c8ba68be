"""Stage3Loss (SURVEY 8f-2b + 8f-3): CE + MSE + both feature transformers + FA in shared passes, against the unfused
composition the reference runs (train_or_resume.py:435-438 with models/DSRL.py:86-95,181,184) evaluated in float64 PyTorch on
the same weights and inputs: the three losses, the gradients w.r.t. SSSR_output / SISR_output and the six transformer parameters,
and the BatchNorm running statistics.  Tolerances: CE / MSE 1e-6, FA loss 1e-4, gradients 1e-3 relative-norm (measured ~1e-6)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _transformer(cin, stride=8):
    return nn.Sequential(nn.Conv2d(cin, 1, 1, stride=stride, bias=False), nn.BatchNorm2d(1), nn.ReLU())


def _fa_reference_f64(a, b, k):
    """FALoss.py:8-34 in float64 (pool, spectral norm, Gram, all-pairs L1, mean)."""
    def sim(x):
        x = nn.functional.avg_pool2d(x, k)
        x = x / torch.linalg.matrix_norm(x, ord=2, dim=(2, 3), keepdim=True)
        return torch.matmul(x.transpose(2, 3), x)
    s1, s2 = sim(a).flatten(2), sim(b).flatten(2)
    n = s1.shape[2]
    return nn.functional.l1_loss(s1.repeat_interleave(n, dim=2), s2.repeat(1, 1, n), reduction="mean")


def relnorm(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("shape,k,training", [((2, 19, 64, 128), 4, True), ((3, 19, 128, 256), 8, True), ((2, 19, 128, 256), 8, False),
                                               ((6, 19, 512, 1024), 8, True)])
def test_stage3_loss_matches_the_unfused_composition(shape, k, training):
    from dualsuperreslearningforsemseg_b200.models.losses import Stage3Loss
    B, C1, H, W = shape
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(54321)
    torch.manual_seed(7)
    t1, t2 = _transformer(C1).to(dev), _transformer(3).to(dev)
    for t in (t1, t2):                                     # not the trivial affine / statistics
        t[1].weight.data.fill_(1.3); t[1].bias.data.fill_(0.2)
        t[1].running_mean.fill_(0.1); t[1].running_var.fill_(0.8)
    sssr = torch.randn(shape, device=dev, generator=g)
    sisr = torch.randn((B, 3, H, W), device=dev, generator=g)
    image = torch.randn((B, 3, H, W), device=dev, generator=g)
    target = torch.randint(0, C1, (B, H, W), device=dev, generator=g, dtype=torch.uint8)
    target.masked_fill_(torch.rand((B, H, W), device=dev, generator=g) < 0.1, 255)
    w1, w2 = 0.1, 1.0

    # float64 reference: the reference's own composition
    r1, r2 = copy.deepcopy(t1).double(), copy.deepcopy(t2).double()
    r1.train(training); r2.train(training)
    a = sssr.double().requires_grad_(True); b = sisr.double().requires_grad_(True)
    ce_r = nn.functional.cross_entropy(a, target.long(), ignore_index=255)
    mse_r = nn.functional.mse_loss(b, image.double())
    fa_r = _fa_reference_f64(r1(a), r2(b), k)
    (ce_r + w1 * mse_r + w2 * fa_r).backward()

    t1.train(training); t2.train(training)
    fn = Stage3Loss(t1, t2, ignore_index=255, subsample_factor=k)
    x = sssr.clone().requires_grad_(True); y = sisr.clone().requires_grad_(True)
    ce, mse, fa = fn(x, y, target, image)
    (ce + w1 * mse + w2 * fa).backward()
    torch.cuda.synchronize()

    assert abs(float(ce) - float(ce_r)) <= 1e-6 * abs(float(ce_r)), (float(ce), float(ce_r))
    assert abs(float(mse) - float(mse_r)) <= 1e-6 * abs(float(mse_r)), (float(mse), float(mse_r))
    assert abs(float(fa) - float(fa_r)) <= 1e-4 * abs(float(fa_r)), (float(fa), float(fa_r))
    assert relnorm(x.grad, a.grad) <= 1e-3 and relnorm(y.grad, b.grad) <= 1e-3, (relnorm(x.grad, a.grad), relnorm(y.grad, b.grad))
    # the transformer path alone (CE / MSE gradients removed): the part of dSSSR / dSISR that lives on the stride grid
    for got, ref, (name, mod, rmod) in ((x.grad, a.grad, ("sssr", t1, r1)), (y.grad, b.grad, ("sisr", t2, r2))):
        assert relnorm(mod[0].weight.grad, rmod[0].weight.grad) <= 1e-3, (name, "conv weight", relnorm(mod[0].weight.grad, rmod[0].weight.grad))
        # dgamma / dbeta are sums of ~B*Hf*Wf signed terms that cancel almost completely (FA is nearly invariant to the scale of
        # its inputs: they are divided by their spectral norm): held to 1e-3 of the LARGER of the two, i.e. of the un-cancelled scale
        scale = max(float(rmod[1].weight.grad.abs()), float(rmod[1].bias.grad.abs()))
        for what, gp, rp in (("gamma", mod[1].weight.grad, rmod[1].weight.grad), ("beta", mod[1].bias.grad, rmod[1].bias.grad)):
            assert abs(float(gp) - float(rp)) <= 1e-3 * scale, (name, what, float(gp), float(rp), scale)
        assert relnorm(mod[1].running_mean, rmod[1].running_mean) <= 1e-5 and relnorm(mod[1].running_var, rmod[1].running_var) <= 1e-5
        assert int(mod[1].num_batches_tracked) == int(rmod[1].num_batches_tracked)
    # the FA path's share of the input gradients, isolated by switching the other two losses off
    x2 = sssr.clone().requires_grad_(True); y2 = sisr.clone().requires_grad_(True)
    a2 = sssr.double().requires_grad_(True); b2 = sisr.double().requires_grad_(True)
    for t in (t1, t2, r1, r2):
        t.zero_grad()
    _, _, fa2 = fn(x2, y2, target, image)
    (2.5 * fa2).backward()
    (2.5 * _fa_reference_f64(r1(a2), r2(b2), k)).backward()
    assert relnorm(x2.grad, a2.grad) <= 1e-3 and relnorm(y2.grad, b2.grad) <= 1e-3, (relnorm(x2.grad, a2.grad), relnorm(y2.grad, b2.grad))


def test_stage3_loss_is_repeatable_and_launches_seven_kernels():
    from dualsuperreslearningforsemseg_b200.models.losses import Stage3Loss
    from dualsuperreslearningforsemseg_b200 import _lib
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    t1, t2 = _transformer(19).to(dev).eval(), _transformer(3).to(dev).eval()
    sssr = torch.randn((2, 19, 128, 256), device=dev); sisr = torch.randn((2, 3, 128, 256), device=dev)
    image = torch.randn((2, 3, 128, 256), device=dev)
    target = torch.randint(0, 19, (2, 128, 256), device=dev, dtype=torch.uint8)
    fn = Stage3Loss(t1, t2, ignore_index=255)
    outs = []
    for _ in range(2):
        x = sssr.clone().requires_grad_(True); y = sisr.clone().requires_grad_(True)
        for t in (t1, t2):
            t.zero_grad()
        n0 = _lib.launch_count()
        ce, mse, fa = fn(x, y, target, image)
        n_fwd = _lib.launch_count() - n0
        (ce + mse + fa).backward()
        n_all = _lib.launch_count() - n0
        outs.append((float(ce), float(mse), float(fa), x.grad.clone(), y.grad.clone(), t1[0].weight.grad.clone()))
    assert n_fwd == 4 and n_all == 7, (n_fwd, n_all)
    assert outs[0][:3] == outs[1][:3] and all(torch.equal(p, q) for p, q in zip(outs[0][3:], outs[1][3:]))
