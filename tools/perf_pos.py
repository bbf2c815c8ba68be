"""Quick device-time probe of FA(position) fwd+bwd (not the bench; prints ms and algorithmic TFLOP/s = 12*B*C*N^2/t)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from dualsuperreslearningforsemseg_b200.functional import FAPlan
cases = [(2, 64, 128, 256), (2, 128, 128, 256), (2, 256, 128, 256)] if len(sys.argv) < 2 else [tuple(int(v) for v in sys.argv[1].split(','))]
prec = sys.argv[2] if len(sys.argv) > 2 else None
for (B, C, H, W) in cases:
    g = torch.Generator(device='cuda'); g.manual_seed(54321)
    x1 = torch.relu(torch.randn((B, C, H, W), device='cuda', generator=g))
    x2 = torch.relu(torch.randn((B, C, H, W), device='cuda', generator=g))
    plan = FAPlan((B, C, H, W), subsample_factor=1, affinity='position', precision=prec)
    go = torch.ones((), device='cuda')
    for need_grad in (True, False):
        def step():
            plan.forward(x1, x2, need_grad)
            if need_grad: plan.backward(x1, x2, go)
        step(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        N = H * W
        flops = (12 if need_grad else 4) * B * C * N * N
        print(f"B={B} C={C} N={N} grad={need_grad} prec={prec}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (algorithmic)  loss={float(plan.loss):.6f}", flush=True)
