"""Dev probe: times FAPlan.forward_backward in position mode (CUDA events), e.g. `python tools/prof_pos.py --B 1 --C 256 --exact 1`."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualsuperreslearningforsemseg_b200.functional import FAPlan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=1)
ap.add_argument("--C", type=int, default=256)
ap.add_argument("--H", type=int, default=128)
ap.add_argument("--W", type=int, default=256)
ap.add_argument("--k", type=int, default=1)
ap.add_argument("--precision", default="f16")
ap.add_argument("--exact", type=int, default=0)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(54321)
x1 = torch.relu(torch.randn((a.B, a.C, a.H, a.W), device=dev, generator=g))
x2 = torch.relu(torch.randn((a.B, a.C, a.H, a.W), device=dev, generator=g))
plan = FAPlan(x1.shape, subsample_factor=a.k, affinity="position", precision=a.precision, device=dev, exact_signs=bool(a.exact))
go = torch.ones((), device=dev)
for _ in range(a.warmup):
    plan.forward_backward(x1, x2, go)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    plan.forward_backward(x1, x2, go)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
N = (a.H // a.k) * (a.W // a.k)
print(f"B={a.B} C={a.C} N={N} {a.precision} exact={a.exact}: {ms:.3f} ms/step  {a.B * N * N / ms / 1e6:.1f} Gpairs/s  loss={float(plan.loss):.6f}"
      + (f"  {plan.sign_stats()}" if a.exact else ""), flush=True)
