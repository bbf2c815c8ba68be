"""GPU parity of the FA loss (reference semantics) through the drop-in FALoss -> C-ABI -> sm_100a kernels.

Tolerances are the north star's: loss <= 1e-4 relative, gradients <= 1e-3 relative-norm, both measured against
the FLOAT64 run of the unmodified reference (tests/golden/fa_golden.npz) and against the float64 oracle."""
import numpy as np
import pytest
import torch

from _inputs import fa_inputs, load_golden, expand_pooled
from oracle import fa_oracle

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_RTOL = 1e-3

G = load_golden("fa_golden.npz")
NAMES = [str(n) for n in G["names"]]


def relnorm(a, b):
    a = np.nan_to_num(np.asarray(a, dtype=np.float64))
    b = np.nan_to_num(np.asarray(b, dtype=np.float64))
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def run(x1, x2, k, red, go=None, **kw):
    from dualsuperreslearningforsemseg_b200.models.losses import FALoss
    a = torch.from_numpy(x1).cuda().requires_grad_(True)
    b = torch.from_numpy(x2).cuda().requires_grad_(True)
    loss = FALoss(subsample_factor=k, reduction=red, **kw)(a, b)
    if go is None:
        loss.backward()
    else:
        loss.backward(torch.from_numpy(go).cuda())
    torch.cuda.synchronize()
    return loss.detach().cpu().numpy(), a.grad.cpu().numpy(), b.grad.cpu().numpy()


@pytest.mark.parametrize("name", NAMES)
def test_matches_reference_golden(name):
    B, C, H, W, k, seed = (int(v) for v in G[f"{name}/meta"])
    red, dist = str(G[f"{name}/reduction"]), str(G[f"{name}/dist"])
    x1, x2 = fa_inputs((B, C, H, W), dist, seed)
    go = None
    if red == "none":
        n = (W // k) ** 2
        go = np.random.default_rng(seed + 1000).standard_normal((B, C, n * n)).astype(np.float32)
    loss, d1, d2 = run(x1, x2, k, red, go)
    g1 = expand_pooled(G[f"{name}/g1_64"], k, H, W)
    g2 = expand_pooled(G[f"{name}/g2_64"], k, H, W)
    if dist == "dead":
        # sigma = 0: the reference returns NaN loss, NaN gradient for the dead (b,c) of the dead branch and an
        # exactly-zero gradient for the same (b,c) of the other branch
        assert np.isnan(loss) and np.isnan(float(G[f"{name}/loss64"]))
        assert np.array_equal(np.isnan(d1), np.isnan(g1)) and np.array_equal(np.isnan(d2), np.isnan(g2))
        assert np.all(d2[1, 0] == 0.0)
    elif red == "none":
        ref = G[f"{name}/loss64"]
        assert loss.shape == ref.shape
        np.testing.assert_allclose(loss, ref, rtol=LOSS_RTOL, atol=1e-6)
    else:
        assert loss.shape == ()
        assert abs(float(loss) - float(G[f"{name}/loss64"])) <= LOSS_RTOL * abs(float(G[f"{name}/loss64"]))
    assert relnorm(d1, g1) <= GRAD_RTOL, relnorm(d1, g1)
    assert relnorm(d2, g2) <= GRAD_RTOL, relnorm(d2, g2)


@pytest.mark.parametrize("shape,k", [((1, 1, 512, 1024), 8), ((2, 1, 256, 1024), 8), ((1, 1, 1024, 2048), 8), ((1, 2, 515, 1030), 8)])
def test_large_maps_against_sorted_oracle(shape, k):
    """Beyond what the reference can materialise (w^4 floats): the O(n log n) oracle, validated against the
    reference at small n in tests/test_oracle_fa.py.  The last shape is the reference-mode counterpart of
    BASELINE config 4 (pooled 128 x 256, n = 65536, 4.3 G pairs per sample).  Maps of a megabyte and more are pooled by the
    grid-wide fa_ref_pool pass (515 x 1030: floor pooling with dropped rows / columns, unaligned scalar loads)."""
    x1, x2 = fa_inputs(shape, "relu", 99)
    loss, d1, d2 = run(x1, x2, k, "mean")
    ol, o1, o2 = fa_oracle.fa_reference(x1, x2, k, "mean", materialise_limit=0)
    assert abs(float(loss) - ol) <= LOSS_RTOL * abs(ol)
    assert relnorm(d1, o1) <= GRAD_RTOL and relnorm(d2, o2) <= GRAD_RTOL


@pytest.mark.parametrize("shape,k,dist", [((2, 1, 48, 80), 1, "randn"), ((1, 2, 80, 48), 1, "randn"), ((1, 1, 96, 160), 2, "relu"),
                                          ((1, 1, 100, 36), 1, "relu"), ((1, 1, 37, 300), 1, "relu")])
def test_mid_size_maps_both_solvers(shape, k, dist):
    """Short side 33..128: the squaring + power solver is tried first and the one-sided Jacobi takes over when the power steps
    do not converge.  Signed Gaussian maps have a spectral gap of a few per cent (Jacobi fallback), post-ReLU maps a dominant
    Perron vector (squarings converge); wide and tall maps.  Maps wider than 64 cells also take the row-chunked gradient kernel
    (fa_ref_grad_rows; 37 x 300: ragged last row chunk, two column blocks).  Against the float64 oracle."""
    x1, x2 = fa_inputs(shape, dist, 17)
    loss, d1, d2 = run(x1, x2, k, "mean")
    ol, o1, o2 = fa_oracle.fa_reference(x1, x2, k, "mean")
    assert abs(float(loss) - ol) <= LOSS_RTOL * abs(ol), (float(loss), ol)
    assert relnorm(d1, o1) <= GRAD_RTOL and relnorm(d2, o2) <= GRAD_RTOL, (relnorm(d1, o1), relnorm(d2, o2))


def test_wide_map_reduction_none():
    """reduction='none' on a map wider than 64 cells: the upstream-weighted g values reach the row-chunked gradient kernel as
    floats (not as integer counts)."""
    shape, k = (1, 2, 10, 72), 1
    x1, x2 = fa_inputs(shape, "relu", 23)
    n = 72 * 72
    go = np.random.default_rng(24).standard_normal((1, 2, n * n)).astype(np.float32)
    loss, d1, d2 = run(x1, x2, k, "none", go)
    ol, o1, o2 = fa_oracle.fa_reference(x1, x2, k, "none", grad_out=go)
    np.testing.assert_allclose(loss, ol, rtol=LOSS_RTOL, atol=1e-6)
    assert relnorm(d1, o1) <= GRAD_RTOL and relnorm(d2, o2) <= GRAD_RTOL, (relnorm(d1, o1), relnorm(d2, o2))


@pytest.mark.parametrize("shape,dist", [((3, 1, 64, 128), "relu"), ((5, 1, 64, 128), "randn"), ((7, 1, 64, 128), "relu"),
                                        ((8, 1, 64, 128), "relu"), ((3, 4, 64, 128), "relu"), ((2, 8, 64, 128), "randn"),
                                        ((2, 1, 128, 128), "relu"), ((1, 2, 256, 128), "randn"), ((1, 3, 96, 104), "relu")])
def test_fused_kernel_grid_and_solver_paths(shape, dist):
    """The single-launch kernel of the training shapes: every cluster size 1..8 of its loss reduction (B*C CTAs as one thread-block
    cluster, partials over distributed shared memory), the global-ticket form of larger grids (B*C = 12, 16), and the solver
    warp's two forms -- register-resident for a short side <= 8, shared-memory squarings for 9..16 (pooled 16 x 16, 32 x 16,
    12 x 13).  Against the float64 oracle; forward + backward through autograd and as the one-call form."""
    from dualsuperreslearningforsemseg_b200.functional import FAPlan
    x1, x2 = fa_inputs(shape, dist, 31)
    loss, d1, d2 = run(x1, x2, 8, "mean")
    ol, o1, o2 = fa_oracle.fa_reference(x1, x2, 8, "mean")
    assert abs(float(loss) - ol) <= LOSS_RTOL * abs(ol), (float(loss), ol)
    assert relnorm(d1, o1) <= GRAD_RTOL and relnorm(d2, o2) <= GRAD_RTOL, (relnorm(d1, o1), relnorm(d2, o2))
    if shape[2] % 8 == 0 and shape[3] % 8 == 0:
        plan = FAPlan(shape, subsample_factor=8)
        l2, e1, e2 = plan.forward_backward(torch.from_numpy(x1).cuda(), torch.from_numpy(x2).cuda(), torch.ones((), device="cuda"))
        torch.cuda.synchronize()
        assert float(l2) == float(loss) and np.array_equal(e1.cpu().numpy(), d1) and np.array_equal(e2.cpu().numpy(), d2)


@pytest.mark.parametrize("shape", [(2, 1, 256, 512), (2, 1, 1024, 1024), (2, 1, 512, 1152)])
def test_dead_channel_on_every_all_pairs_path(shape):
    """sigma = 0 (an all-zero pooled map) must give a NaN loss, NaN gradient for the dead (b, c) of the dead branch and an
    exactly-zero gradient for the same (b, c) of the other branch (torch's sign(NaN) = 0) -- on the n^2 stream (n = 4096),
    the single-chunk sorted path (n = 16384) and the chunked sorted path (n = 20736)."""
    x1, x2 = fa_inputs(shape, "relu", 5)
    x1[1, 0] = 0.0
    loss, d1, d2 = run(x1, x2, 8, "mean")
    assert np.isnan(float(loss))
    assert np.isnan(d1[1, 0]).all() and np.isfinite(d1[0, 0]).all()
    assert np.all(d2[1, 0] == 0.0) and np.isfinite(d2[0, 0]).all() and np.abs(d2[0, 0]).max() > 0
    # the live sample's gradients are those of the same inputs without the dead one, rescaled by the batch size
    _, e1, e2 = run(x1[:1], x2[:1], 8, "mean")
    assert relnorm(2.0 * d1[0], e1[0]) <= 1e-5 and relnorm(2.0 * d2[0], e2[0]) <= 1e-5


def test_size_independent_properties_at_training_shape():
    """BASELINE config 2 shape.  (i) the all-pairs L1 is symmetric in its arguments; (ii) spectral normalisation
    makes the loss invariant to a positive rescaling of either input, hence <dX, X> = 0; (iii) backward is
    linear in the upstream gradient (w2 * FA, train_or_resume.py:437)."""
    from dualsuperreslearningforsemseg_b200.models.losses import FALoss
    x1, x2 = fa_inputs((6, 1, 64, 128), "relu", 54321)
    l12, a1, a2 = run(x1, x2, 8, "mean")
    l21, b2, b1 = run(x2, x1, 8, "mean")
    assert abs(float(l12) - float(l21)) <= 1e-6 * abs(float(l12))
    assert relnorm(a1, b1) <= 1e-5 and relnorm(a2, b2) <= 1e-5
    ls, s1, _ = run(3.0 * x1, x2, 8, "mean")
    assert abs(float(ls) - float(l12)) <= 1e-5 * abs(float(l12))
    assert relnorm(3.0 * s1, a1) <= GRAD_RTOL
    assert abs(float((a1.astype(np.float64) * x1).sum())) <= 1e-4 * np.linalg.norm(a1) * np.linalg.norm(x1)
    a = torch.from_numpy(x1).cuda().requires_grad_(True)
    b = torch.from_numpy(x2).cuda().requires_grad_(True)
    (0.25 * FALoss()(a, b) + (a * 0).sum()).backward()
    np.testing.assert_allclose(a.grad.cpu().numpy(), 0.25 * a1, rtol=1e-6, atol=1e-12)


def test_only_one_input_requires_grad_and_no_grad():
    from dualsuperreslearningforsemseg_b200.models.losses import FALoss
    x1, x2 = fa_inputs((2, 1, 64, 128), "relu", 5)
    _, g1, _ = run(x1, x2, 8, "mean")
    a = torch.from_numpy(x1).cuda().requires_grad_(True)
    b = torch.from_numpy(x2).cuda()
    FALoss()(a, b).backward()
    np.testing.assert_array_equal(a.grad.cpu().numpy(), g1)
    with torch.no_grad():
        l = FALoss()(a, b)
    assert not l.requires_grad and l.shape == ()


def test_non_contiguous_and_repeatable():
    x1, x2 = fa_inputs((2, 2, 64, 128), "relu", 8)
    l0, a0, b0 = run(x1, x2, 8, "mean")
    l1, a1, b1 = run(x1, x2, 8, "mean")
    assert l0 == l1 and np.array_equal(a0, a1) and np.array_equal(b0, b1)      # deterministic
    from dualsuperreslearningforsemseg_b200.models.losses import FALoss
    big = torch.from_numpy(np.concatenate([x1, x1], axis=1)).cuda()
    a = big[:, ::2].detach().requires_grad_(True)           # non-contiguous view of channels 0 and 2
    ref = torch.from_numpy(np.ascontiguousarray(np.concatenate([x1, x1], axis=1)[:, ::2])).cuda()
    l_nc = FALoss()(a, torch.from_numpy(x2).cuda())
    l_c = FALoss()(ref, torch.from_numpy(x2).cuda())
    assert float(l_nc) == float(l_c)
