"""Per-kernel device times of FA(reference) fwd+bwd on large maps (general path)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from dualsuperreslearningforsemseg_b200.functional import FAPlan
from torch.profiler import profile, ProfilerActivity
for shape, k in [((1, 1, 256, 512), 8), ((8, 1, 512, 1024), 8), ((8, 1, 1024, 2048), 8)]:
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    x1 = torch.relu(torch.randn(shape, device='cuda', generator=g)); x2 = torch.relu(torch.randn(shape, device='cuda', generator=g))
    plan = FAPlan(shape, subsample_factor=k)
    go = torch.ones((), device='cuda')
    for _ in range(2): plan.forward_backward(x1, x2, go)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): plan.forward_backward(x1, x2, go)
        torch.cuda.synchronize()
    print(shape)
    for e in prof.key_averages():
        if e.device_time > 0: print(f"   {e.key[:70]:70s} n={e.count} avg={e.device_time:.1f} us")
