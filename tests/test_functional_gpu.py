"""functional.FAPlan (allocation-free C-ABI access) against the autograd drop-in, eagerly and under CUDA-graph replay,
for both FA semantics (the position path launches TMA / cluster kernels: they must be capturable)."""
import pytest
import torch

from _inputs import fa_inputs, pos_margin_inputs

pytestmark = pytest.mark.gpu


def _autograd(x1, x2, k, go, **kw):
    from dualsuperreslearningforsemseg_b200.models.losses import FALoss
    a = x1.clone().requires_grad_(True)
    b = x2.clone().requires_grad_(True)
    loss = FALoss(subsample_factor=k, **kw)(a, b)
    (loss * go).backward()
    return loss.detach(), a.grad, b.grad


@pytest.mark.parametrize("mode", ["reference", "position_tf32", "position_fp32", "position_f16"])
def test_plan_matches_autograd_and_replays_in_a_graph(mode):
    from dualsuperreslearningforsemseg_b200.functional import FAPlan
    dev = torch.device("cuda", 0)
    if mode == "reference":
        x1, x2 = fa_inputs((6, 1, 64, 128), "relu", 54321)
        k, kw = 8, {}
    else:
        x1, x2 = pos_margin_inputs(2, 64, 64, 32, 32, 54321)
        k, kw = 1, {"affinity": "position", "precision": mode.split("_")[1]}
    x1, x2 = torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev)
    go = torch.full((), 0.5, device=dev)
    ref_loss, ref_d1, ref_d2 = _autograd(x1, x2, k, go, **kw)
    plan = FAPlan(tuple(x1.shape), tuple(x2.shape), subsample_factor=k, device=dev, **kw)
    loss, d1, d2 = plan.forward_backward(x1, x2, go)
    torch.cuda.synchronize()
    assert float(loss) == float(ref_loss) and torch.equal(d1, ref_d1) and torch.equal(d2, ref_d2)
    # capture the step once, then replay it on new inputs written into the captured buffers
    sx1, sx2 = x1.clone(), x2.clone()
    plan.forward_backward(sx1, sx2, go)                      # eager warm-up on the static buffers
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        plan.forward_backward(sx1, sx2, go)
    y1, y2 = x2.clone(), x1.clone()                          # swapped branches: loss identical, gradients swap and flip roles
    sx1.copy_(y1)
    sx2.copy_(y2)
    graph.replay()
    torch.cuda.synchronize()
    l2, e1, e2 = _autograd(y1, y2, k, go, **kw)
    assert float(plan.loss) == float(l2) and torch.equal(plan.dx1, e1) and torch.equal(plan.dx2, e2)


@pytest.mark.parametrize("mode,red,chunk", [("reference", "mean", 2), ("reference", "sum", 4), ("position_f16", "mean", 2),
                                             ("position_tf32", "sum", 3), ("position_f16", "mean", 8)])
def test_host_pipeline_matches_whole_batch(mode, red, chunk):
    """FAHostPipeline (chunked H2D on a copy stream overlapped with the kernels) == FALoss on the whole batch on the device:
    chunk means weighted n/B, per-sample gradients untouched by the chunking.  Called twice: the staging buffers are reused."""
    from dualsuperreslearningforsemseg_b200.functional import FAHostPipeline
    dev = torch.device("cuda", 0)
    if mode == "reference":
        x1, x2 = fa_inputs((6, 2, 64, 128), "relu", 54321)
        k, kw = 8, {}
    else:
        x1, x2 = pos_margin_inputs(5, 64, 64, 16, 32, 54321)          # 5 samples: ragged last chunk for chunk = 2, 3
        k, kw = 1, {"affinity": "position", "precision": mode.split("_")[1]}
    h1, h2 = torch.from_numpy(x1).pin_memory(), torch.from_numpy(x2).pin_memory()
    go = 0.25
    ref_loss, ref_d1, ref_d2 = _autograd(h1.to(dev), h2.to(dev), k, go, reduction=red, **kw)
    pipe = FAHostPipeline(tuple(h1.shape), tuple(h2.shape), subsample_factor=k, reduction=red, chunk=chunk, device=dev, **kw)
    for _ in range(2):
        loss, d1, d2 = pipe(h1, h2, go)
        torch.cuda.synchronize()
        assert abs(float(loss) - float(ref_loss)) <= 2e-6 * abs(float(ref_loss)), (float(loss), float(ref_loss))
        for d, r in ((d1, ref_d1), (d2, ref_d2)):
            assert float((d - r).norm() / r.norm()) <= 2e-6
    with pytest.raises(ValueError):
        pipe(h1.to(dev), h2.to(dev))                                   # device tensors belong to FALoss / FAPlan


@pytest.mark.parametrize("jsplit", ["1", "2"])
@pytest.mark.parametrize("prec", ["tf32", "f16"])
def test_fused_position_call_writes_dx_directly(jsplit, prec, monkeypatch):
    """Position mode without pooling: dsrl_fa_forward_backward lets the gradient kernel (or fa_pos_jacobian when the column
    range is split) write dX itself.  Ragged N (not a multiple of 128) and padded channel counts (40 and 33 -> 64 each):
    bit-identical to forward + backward through autograd, and nothing is written outside the real channels / positions."""
    from dualsuperreslearningforsemseg_b200.functional import FAPlan
    monkeypatch.setenv("DSRL_POS_JSPLIT", jsplit)
    dev = torch.device("cuda", 0)
    x1, x2 = pos_margin_inputs(2, 40, 33, 25, 30, 54321)              # N = 750 -> 6 row tiles, last one ragged
    x1, x2 = torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev)
    go = torch.full((), 0.3, device=dev)
    kw = {"affinity": "position", "precision": prec}
    ref_loss, ref_d1, ref_d2 = _autograd(x1, x2, 1, go, **kw)
    plan = FAPlan(tuple(x1.shape), tuple(x2.shape), subsample_factor=1, device=dev, **kw)
    plan.dx1.fill_(float("nan")); plan.dx2.fill_(float("nan"))
    loss, d1, d2 = plan.forward_backward(x1, x2, go)
    torch.cuda.synchronize()
    assert float(loss) == float(ref_loss) and torch.equal(d1, ref_d1) and torch.equal(d2, ref_d2)


def test_host_pipeline_graph_replay():
    """FAHostPipeline.capture(): copies, kernel and loss read-back as one CUDA graph; new inputs are written into the captured
    pinned buffers between replays."""
    from dualsuperreslearningforsemseg_b200.functional import FAHostPipeline
    dev = torch.device("cuda", 0)
    x1, x2 = fa_inputs((6, 1, 64, 128), "relu", 54321)
    h1, h2 = torch.from_numpy(x1).pin_memory(), torch.from_numpy(x2).pin_memory()
    pipe = FAHostPipeline(tuple(h1.shape), subsample_factor=8, chunk=6, ramp=False, device=dev).capture(h1, h2, 0.5)
    for seed in (54321, 7):
        y1, y2 = fa_inputs((6, 1, 64, 128), "relu", seed)
        h1.copy_(torch.from_numpy(y1)); h2.copy_(torch.from_numpy(y2))
        loss = float(pipe.replay())
        ref_loss, ref_d1, ref_d2 = _autograd(torch.from_numpy(y1).to(dev), torch.from_numpy(y2).to(dev), 8, 0.5)
        assert abs(loss - float(ref_loss)) <= 1e-6 * abs(float(ref_loss))
        assert float((pipe.dx1 - ref_d1).norm() / ref_d1.norm()) <= 1e-6 and float((pipe.dx2 - ref_d2).norm() / ref_d2.norm()) <= 1e-6


@pytest.mark.parametrize("case", [
    ((2, 64, 16, 32), (2, 64, 16, 32), 1, "f16", True),        # two-pass form, one channel group
    ((1, 200, 24, 40), (1, 200, 24, 40), 1, "f16", True),      # two groups, padded positions and channels
    ((1, 48, 32, 64), (1, 80, 32, 64), 2, "f16", True),        # pooling: dP + unpool
    ((1, 256, 16, 32), (1, 256, 16, 32), 1, "f16", False),     # tensor-core signs
    ((1, 96, 12, 32), (1, 96, 12, 32), 1, "f16", True),        # odd tile count: fused single-CTA kernel
    ((1, 64, 32, 32), (1, 64, 32, 32), 1, "tf32", True),       # fused pair kernel + resolve + finish
    ((6, 1, 64, 128), (6, 1, 64, 128), 8, "reference", True),  # reference semantics, training shape
])
def test_calls_stay_inside_their_buffers(case):
    """Guard bands instead of compute-sanitizer (closed on this pool): workspace, saved blob, gradients and loss are carved
    out of one allocation with 64 KB canary regions between them; after a forward + backward through the C-ABI at exactly
    the sizes the library asks for, every canary byte is intact (and the result equals the ordinary plan's)."""
    import ctypes
    from dualsuperreslearningforsemseg_b200 import _lib
    from dualsuperreslearningforsemseg_b200.functional import FAPlan
    from _inputs import pos_inputs
    s1, s2, k, prec, exact = case
    dev = torch.device("cuda", 0)
    if prec == "reference":
        a, b = fa_inputs(s1, "relu", 7)
        kw = {}
    else:
        a, b = pos_inputs(s1, s2, 7)
        kw = {"affinity": "position", "precision": prec, "exact_signs": exact}
    x1, x2 = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    plan = FAPlan(s1, s2, subsample_factor=k, device=dev, **kw)
    go = torch.full((), 0.75, device=dev)
    ref = [t.clone() for t in plan.forward_backward(x1, x2, go)]
    G = 65536
    sizes = [plan.ws_bytes, plan.saved_bytes, x1.numel() * 4, x2.numel() * 4, 4]
    offs, off = [], G
    for n in sizes:
        offs.append(off)
        off += (n + 255) // 256 * 256 + G
    arena = torch.full((off,), 0xA5, dtype=torch.uint8, device=dev)
    ws, saved, d1, d2, loss = (arena[o:o + n] for o, n in zip(offs, sizes))
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    for _ in range(2):
        _lib.check(_lib.lib().dsrl_fa_forward_backward(plan.mode, plan.prec, p(x1), p(x2), plan.B, plan.C1, plan.C2, plan.H, plan.W,
                                                       plan.k, plan.red, p(go), p(loss), p(d1), p(d2), p(saved), plan.saved_bytes,
                                                       p(ws), plan.ws_bytes, st))
    torch.cuda.synchronize()
    keep = torch.ones(off, dtype=torch.bool, device=dev)
    for o, n in zip(offs, sizes):
        keep[o:o + n] = False
    assert bool((arena[keep] == 0xA5).all()), "a kernel wrote outside the buffers it was given"
    assert torch.equal(loss.view(torch.float32).reshape(()), ref[0])
    assert torch.equal(d1.view(torch.float32).reshape(x1.shape), ref[1]) and torch.equal(d2.view(torch.float32).reshape(x2.shape), ref[2])
