"""FA(position) with FP16 operands against the TF32 kernels and the float64 oracle (development probe)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from dualsuperreslearningforsemseg_b200.models.losses import FALoss
from oracle import fa_oracle
from _inputs import pos_margin_inputs

def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

for name, (B, C1, C2, H, W, k) in {"c32": (1, 32, 32, 16, 32, 1), "c20_48": (2, 20, 48, 16, 16, 1), "c64": (2, 64, 64, 32, 32, 1),
                                   "c96_32": (1, 96, 32, 16, 32, 1), "c128": (1, 128, 128, 16, 64, 1), "c256": (1, 256, 256, 16, 32, 1)}.items():
    x1n, x2n = pos_margin_inputs(B, C1, C2, H, W, 7)
    out = {}
    for prec in ("tf32", "f16"):
        a = torch.tensor(x1n, device='cuda', requires_grad=True); b = torch.tensor(x2n, device='cuda', requires_grad=True)
        l = FALoss(k, affinity='position', precision=prec)(a, b); l.backward()
        out[prec] = (float(l), a.grad.cpu().numpy().astype(np.float64), b.grad.cpu().numpy().astype(np.float64))
    lo, g1, g2 = fa_oracle.fa_position(x1n, x2n, k=k)[:3]
    for prec in ("tf32", "f16"):
        l, a, b = out[prec]
        print(f"{name:8s} {prec:5s} loss rel {abs(l - lo) / lo:.2e}  g1 {rel(a, g1):.2e}  g2 {rel(b, g2):.2e}", flush=True)
