"""Dev probe: exact-sign position mode against the float64 oracle on relu(randn) inputs (prints errors and tie statistics)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _inputs import pos_inputs  # noqa: E402
from oracle import fa_oracle  # noqa: E402
from dualsuperreslearningforsemseg_b200.models.losses import FALoss  # noqa: E402

CASES = [
    ((2, 64, 32, 32), (2, 64, 32, 32), 1, "mean"),
    ((1, 128, 24, 40), (1, 128, 24, 40), 1, "mean"),
    ((1, 40, 50, 30), (1, 33, 50, 30), 2, "sum"),
    ((1, 256, 16, 32), (1, 256, 16, 32), 1, "mean"),
    ((2, 160, 16, 16), (2, 130, 16, 16), 1, "mean"),
    ((1, 32, 64, 64), (1, 32, 64, 64), 1, "mean"),
    ((1, 256, 16, 16), (1, 20, 16, 16), 1, "mean"),
    ((1, 96, 12, 32), (1, 96, 12, 32), 1, "sum"),
    ((1, 3, 32, 32), (1, 19, 32, 32), 1, "mean"),
    ((1, 64, 64, 128), (1, 64, 64, 128), 1, "mean"),
    ((1, 200, 24, 40), (1, 200, 24, 40), 1, "mean"),      # two channel groups, 64 padded positions
    ((2, 256, 32, 64), (2, 256, 32, 64), 1, "sum"),
]


def relnorm(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def main():
    for s1, s2, k, red in CASES:
        x1, x2 = pos_inputs(s1, s2, 54321)
        ol, o1, o2 = fa_oracle.fa_position(x1, x2, k, red)
        for prec in ("f16", "tf32", "fp32"):
            for exact in (False, True):
                a = torch.from_numpy(x1).cuda().requires_grad_(True)
                b = torch.from_numpy(x2).cuda().requires_grad_(True)
                fn = FALoss(subsample_factor=k, reduction=red, affinity="position", precision=prec, exact_signs=exact)
                loss = fn(a, b)
                loss.backward()
                torch.cuda.synchronize()
                st = fn.sign_stats()
                print(f"{s1}{s2} k={k} {prec:5s} exact={int(exact)} loss_rel={abs(float(loss) - ol) / abs(ol):.2e} "
                      f"g1={relnorm(a.grad.cpu().numpy(), o1):.2e} g2={relnorm(b.grad.cpu().numpy(), o2):.2e} {st}", flush=True)
    if "--big" in sys.argv:
        C = 256
        x1, x2 = pos_inputs((1, C, 128, 256), (1, C, 128, 256), 54321)
        rows = np.random.default_rng(3).choice(128 * 256, size=256, replace=False)
        _, o1, o2 = fa_oracle.fa_position_rows(x1, x2, rows, 1, "mean")
        for prec in ("f16", "tf32"):
            for exact in (False, True):
                a = torch.from_numpy(x1).cuda().requires_grad_(True)
                b = torch.from_numpy(x2).cuda().requires_grad_(True)
                fn = FALoss(subsample_factor=1, affinity="position", precision=prec, exact_signs=exact)
                for it in range(2):
                    a.grad = b.grad = None
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    loss = fn(a, b); loss.backward()
                    torch.cuda.synchronize(); dt = time.perf_counter() - t0
                g1 = a.grad[0].reshape(C, -1)[:, rows].cpu().numpy()
                g2 = b.grad[0].reshape(C, -1)[:, rows].cpu().numpy()
                print(f"N=32768 C=256 {prec} exact={int(exact)} {dt * 1e3:.2f} ms g1={relnorm(g1, o1):.2e} g2={relnorm(g2, o2):.2e} {fn.sign_stats()}", flush=True)


if __name__ == "__main__":
    main()
