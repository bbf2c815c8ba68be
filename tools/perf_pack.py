"""Times the three launches of one FA(position) step separately (pack / tiles / unpool) with CUDA events around each C-ABI call."""
import sys, torch
sys.path.insert(0, '/root/repo')
from dualsuperreslearningforsemseg_b200.functional import FAPlan
B, C, H, W = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "8,256,128,256").split(','))
g = torch.Generator(device='cuda'); g.manual_seed(1)
x1 = torch.relu(torch.randn((B, C, H, W), device='cuda', generator=g)); x2 = torch.relu(torch.randn((B, C, H, W), device='cuda', generator=g))
plan = FAPlan((B, C, H, W), subsample_factor=1, affinity='position')
go = torch.ones((), device='cuda')
for _ in range(2): plan.forward_backward(x1, x2, go)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): plan.forward_backward(x1, x2, go)
    torch.cuda.synchronize()
for e in prof.key_averages():
    print(f"{e.key[:60]:60s} n={e.count} avg={e.device_time / 1e3:.3f} ms")
