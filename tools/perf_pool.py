"""FA(position) with pooling (subsample_factor 8): per-kernel device times."""
import sys, torch
sys.path.insert(0, '/root/repo')
from dualsuperreslearningforsemseg_b200.functional import FAPlan
from torch.profiler import profile, ProfilerActivity
B, C, H, W, k = 2, 64, 1024, 2048, 8
g = torch.Generator(device='cuda'); g.manual_seed(1)
x1 = torch.relu(torch.randn((B, C, H, W), device='cuda', generator=g)); x2 = torch.relu(torch.randn((B, C, H, W), device='cuda', generator=g))
plan = FAPlan((B, C, H, W), subsample_factor=k, affinity='position', precision=(sys.argv[1] if len(sys.argv) > 1 else None))
go = torch.ones((), device='cuda')
for _ in range(2): plan.forward_backward(x1, x2, go)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): plan.forward_backward(x1, x2, go)
    torch.cuda.synchronize()
gb = 2 * B * C * H * W * 4 / 1e9
for e in prof.key_averages():
    if e.device_time > 0: print(f"   {e.key[:70]:70s} n={e.count} avg={e.device_time:.1f} us  ({gb / (e.device_time * 1e-6):.0f} GB/s if it moved the {gb:.2f} GB of inputs once)")
