"""Dev probe: one FA(reference) forward + backward at a given shape (for ncu launch lists), e.g. `python tools/prof_ref.py 8,1,1024,2048 8`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualsuperreslearningforsemseg_b200.functional import FAPlan
shape = tuple(int(v) for v in sys.argv[1].split(",")); k = int(sys.argv[2])
g = torch.Generator(device='cuda'); g.manual_seed(1)
x1 = torch.relu(torch.randn(shape, device='cuda', generator=g)); x2 = torch.relu(torch.randn(shape, device='cuda', generator=g))
plan = FAPlan(shape, subsample_factor=k); go = torch.ones((), device='cuda')
for _ in range(3): plan.forward_backward(x1, x2, go)
torch.cuda.synchronize(); print(float(plan.loss))
