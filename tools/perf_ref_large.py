"""Device time of FA(reference) fwd+bwd on large maps (general path: prepare / pairs / grad / unpool)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from dualsuperreslearningforsemseg_b200.functional import FAPlan
for shape, k in [((6, 1, 64, 128), 8), ((6, 1, 128, 256), 8), ((1, 1, 256, 512), 8), ((8, 1, 512, 1024), 8), ((8, 1, 1024, 2048), 8)]:
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    x1 = torch.relu(torch.randn(shape, device='cuda', generator=g)); x2 = torch.relu(torch.randn(shape, device='cuda', generator=g))
    plan = FAPlan(shape, subsample_factor=k)
    go = torch.ones((), device='cuda')
    for _ in range(2): plan.forward_backward(x1, x2, go)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n): plan.forward_backward(x1, x2, go)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    w = shape[3] // k
    pairs = shape[0] * shape[1] * w ** 4
    print(f"{shape} k={k}: w={w} n={w*w} pairs={pairs:.3e}  {ms:.3f} ms  {pairs / ms / 1e6:.1f} Gpairs/s  loss={float(plan.loss):.6e}", flush=True)
