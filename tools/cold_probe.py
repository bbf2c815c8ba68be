"""Dev probe: where the flushed-L2 time of the fused reference kernel goes -- replay after (a) an L2 flush, (b) a flush followed by a
read of the two input maps (inputs warm, everything else cold), (c) nothing (all warm)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from _inputs import fa_inputs
from dualsuperreslearningforsemseg_b200.functional import FAPlan
dev = torch.device('cuda', 0)
x1, x2 = fa_inputs((6, 1, 64, 128), "relu", 54321)
a, b = torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev)
plan = FAPlan((6, 1, 64, 128), subsample_factor=8, device=dev); go = torch.ones((), device=dev)
for _ in range(3): plan.forward_backward(a, b, go)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): plan.forward_backward(a, b, go)
buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(mode, n=200):
    tot = 0.0
    for _ in range(n):
        if mode != 'warm': buf.fill_(1)
        if mode == 'inputs': float(a.sum() + b.sum())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3
for mode in ('cold', 'inputs', 'warm', 'cold', 'inputs', 'warm'):
    print(mode, f"{run(mode):.2f} us")
