"""Pins oracle/fa_oracle.py against the golden vectors produced by the unmodified reference FALoss
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from _inputs import fa_inputs, pos_inputs, load_golden, expand_pooled
from oracle import fa_oracle

G = load_golden("fa_golden.npz")
NAMES = [str(n) for n in G["names"]]


def relerr(a, b):
    return np.linalg.norm(np.nan_to_num(a) - np.nan_to_num(b)) / max(np.linalg.norm(np.nan_to_num(b)), 1e-300)


@pytest.mark.parametrize("name", NAMES)
def test_reference_mode_matches_reference_fp64(name):
    B, C, H, W, k, seed = (int(v) for v in G[f"{name}/meta"])
    red = str(G[f"{name}/reduction"])
    dist = str(G[f"{name}/dist"])
    x1, x2 = fa_inputs((B, C, H, W), dist, seed)
    go = None
    if red == "none":
        n = (W // k) ** 2
        go = np.random.default_rng(seed + 1000).standard_normal((B, C, n * n)).astype(np.float32)
    loss, d1, d2 = fa_oracle.fa_reference(x1, x2, k, red, grad_out=go)
    gl = G[f"{name}/loss64"]
    g1 = expand_pooled(G[f"{name}/g1_64"], k, H, W)
    g2 = expand_pooled(G[f"{name}/g2_64"], k, H, W)
    if dist == "dead":
        assert np.isnan(loss) and np.isnan(gl)
        assert np.array_equal(np.isnan(d1), np.isnan(g1))
        assert np.array_equal(np.isnan(d2), np.isnan(g2))
    else:
        np.testing.assert_allclose(loss, gl, rtol=1e-12, atol=1e-14 if red == "none" else 0)
    assert relerr(d1, g1) < 1e-10
    assert relerr(d2, g2) < 1e-10


@pytest.mark.parametrize("name", ["cfg2_relu_6x1x64x128", "floor_relu_1x2x70x133", "tall_relu_1x1x128x64"])
def test_sorted_closed_form_equals_bruteforce(name):
    B, C, H, W, k, seed = (int(v) for v in G[f"{name}/meta"])
    x1, x2 = fa_inputs((B, C, H, W), str(G[f"{name}/dist"]), seed)
    a = fa_oracle.fa_reference(x1, x2, k, "mean")
    b = fa_oracle.fa_reference(x1, x2, k, "mean", materialise_limit=0)     # forces the O(n log n) path
    np.testing.assert_allclose(a[0], b[0], rtol=1e-12)
    assert relerr(a[1], b[1]) < 1e-12 and relerr(a[2], b[2]) < 1e-12


def test_fp32_reference_is_within_north_star_tolerance_of_fp64():
    """Context for the GPU tolerances: the reference's own fp32 run vs its fp64 run."""
    for name in NAMES:
        if str(G[f"{name}/dist"]) == "dead" or str(G[f"{name}/reduction"]) == "none":
            continue
        l64, l32 = float(G[f"{name}/loss64"]), float(G[f"{name}/loss32"])
        assert abs(l32 - l64) / abs(l64) < 1e-4
        assert relerr(G[f"{name}/g1_32"], G[f"{name}/g1_64"]) < 2e-2    # sign flips: see SURVEY section 7


def test_shape_errors():
    x = np.zeros((1, 1, 8, 8))
    with pytest.raises(ValueError):
        fa_oracle.fa_reference(x[0], x[0])
    with pytest.raises(ValueError):
        fa_oracle.fa_reference(x, np.zeros((1, 1, 8, 16)))


P = load_golden("fa_position_golden.npz")


@pytest.mark.parametrize("name", [str(n) for n in P["names"]])
def test_position_mode_matches_torch_autograd(name):
    B, C1, H, W, C2, k, seed = (int(v) for v in P[f"{name}/meta"])
    x1, x2 = pos_inputs((B, C1, H, W), (B, C2, H, W), seed)
    red = str(P[f"{name}/reduction"])
    loss, d1, d2 = fa_oracle.fa_position(x1, x2, k, red, chunk=100)
    np.testing.assert_allclose(loss, float(P[f"{name}/loss64"]), rtol=1e-12)
    assert relerr(d1, P[f"{name}/g1_64"]) < 1e-10
    assert relerr(d2, P[f"{name}/g2_64"]) < 1e-10


@pytest.mark.parametrize("name", ["cfg2_relu_6x1x64x128", "sum_relu_3x2x64x128", "floor_relu_1x2x70x133"])
def test_torch_timing_port_matches_reference(name):
    import torch
    from oracle import fa_torch_port
    B, C, H, W, k, seed = (int(v) for v in G[f"{name}/meta"])
    x1, x2 = fa_inputs((B, C, H, W), str(G[f"{name}/dist"]), seed)
    loss, d1, d2 = fa_torch_port.fwd_bwd(torch.from_numpy(x1).double(), torch.from_numpy(x2).double(), k, str(G[f"{name}/reduction"]))
    np.testing.assert_allclose(float(loss), float(G[f"{name}/loss64"]), rtol=1e-12)
    assert relerr(d1.numpy(), expand_pooled(G[f"{name}/g1_64"], k, H, W)) < 1e-10
    assert relerr(d2.numpy(), expand_pooled(G[f"{name}/g2_64"], k, H, W)) < 1e-10


def test_position_torch_port_matches_oracle():
    """oracle/fa_position_torch_port.py (the CPU timing baseline of the position stress): summing its row blocks gives
    the float64 oracle's loss and gradients."""
    import torch
    from oracle import fa_position_torch_port as tp
    x1, x2 = pos_inputs((1, 16, 16, 24), (1, 9, 16, 24), 3)
    k = 2
    N = (16 // k) * (24 // k)
    ol, o1, o2 = fa_oracle.fa_position(x1, x2, k, "sum")
    a, b = torch.from_numpy(x1).double(), torch.from_numpy(x2).double()
    tot, g1, g2 = 0.0, 0.0, 0.0
    for r0 in range(0, N, 40):
        l, d1, d2 = tp.fwd_bwd_rows(a, b, k, r0, min(N, r0 + 40))
        tot, g1, g2 = tot + float(l), g1 + d1.numpy(), g2 + d2.numpy()
    np.testing.assert_allclose(tot, ol, rtol=1e-12)
    assert relerr(g1, o1) < 1e-10 and relerr(g2, o2) < 1e-10


def test_tf32_operand_rounding_variant():
    """round_tf32 = cvt.rna.tf32.f32 (10 explicit mantissa bits, ties away from zero); the rounded-operand oracle stays
    within TF32 distance of the exact one on the loss while its gradient shows the sign-flip sensitivity."""
    x = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -10, -(1.0 + 2.0 ** -11), 3.1415927, 0.0], dtype=np.float32)
    r = fa_oracle.round_tf32(x)
    assert r[0] == 1.0 and r[1] == 1.0 + 2.0 ** -10 and r[2] == 1.0 + 2.0 ** -10 and r[3] == -(1.0 + 2.0 ** -10) and r[5] == 0.0
    assert abs(r[4] - 3.1415927) <= 2.0 ** -10 and (np.float32(r[4]).view(np.uint32) & 0x1FFF) == 0
    x1, x2 = pos_inputs((1, 32, 16, 16), (1, 32, 16, 16), 11)
    le, g1, _ = fa_oracle.fa_position(x1, x2, 1, "mean")
    lt, t1, _ = fa_oracle.fa_position(x1, x2, 1, "mean", operand_rounding="tf32")
    assert abs(lt - le) <= 1e-4 * abs(le)
    assert 1e-5 < relerr(t1, g1) < 5e-2


def test_position_sampled_rows_match_full_oracle():
    x1, x2 = pos_inputs((2, 12, 8, 16), (2, 7, 8, 16), 21)
    ol, o1, o2 = fa_oracle.fa_position(x1, x2, 1, "mean")
    rows = np.array([0, 5, 17, 100, 127])
    part, g1, g2 = fa_oracle.fa_position_rows(x1, x2, rows, 1, "mean")
    assert relerr(g1, o1[0].reshape(12, -1)[:, rows]) < 1e-12
    assert relerr(g2, o2[0].reshape(7, -1)[:, rows]) < 1e-12
    allrows, _, _ = fa_oracle.fa_position_rows(x1[:1], x2[:1], np.arange(128), 1, "sum")
    l0, _, _ = fa_oracle.fa_position(x1[:1], x2[:1], 1, "sum", need_grad=False)
    np.testing.assert_allclose(allrows, l0, rtol=1e-12)


def test_f16_operand_rounding_matches_tf32_width():
    """round_f16 is IEEE binary16 (what the FP16-operand kernel feeds the tensor cores): on unit-norm feature entries it
    has the 11-bit significand of TF32, so the two roundings differ by at most one unit in the last place."""
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(4096) / 16).astype(np.float32)
    h, t = fa_oracle.round_f16(x), fa_oracle.round_tf32(x)
    assert np.array_equal(h, x.astype(np.float16).astype(np.float64))
    big = np.abs(x) >= 2.0 ** -14
    assert np.all(np.abs(h - t)[big] <= np.abs(x[big]) * 2.0 ** -10)
    assert np.all(np.abs(h - x)[big] <= np.abs(x[big]) * 2.0 ** -11)
    assert np.all(np.abs(h - x)[~big] <= 2.0 ** -25)
    x1 = rng.standard_normal((1, 8, 8, 8)).astype(np.float32); x2 = rng.standard_normal((1, 8, 8, 8)).astype(np.float32)
    l0 = fa_oracle.fa_position(x1, x2, 1, "mean", need_grad=False)[0]
    lh = fa_oracle.fa_position(x1, x2, 1, "mean", need_grad=False, operand_rounding="f16")[0]
    assert abs(lh - l0) <= 1e-3 * abs(l0) and lh != l0
