"""Dev probe: first GPU check of the two-pass form against the float64 oracle and the fused kernels, then timings."""
import os, sys, subprocess
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _inputs import pos_inputs
from oracle import fa_oracle
from dualsuperreslearningforsemseg_b200.models.losses import FALoss

def rn(a, b): return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
for shape in ((1, 64, 16, 16), (1, 64, 32, 32), (1, 256, 16, 32), (2, 200, 24, 40)):
    x1, x2 = pos_inputs(shape, shape, 11)
    ol, o1, o2 = fa_oracle.fa_position(x1, x2, 1, "mean")
    for exact in (False, True):
        for ab in ("1", "0"):
            os.environ["DSRL_POS_AB"] = ab
            a = torch.from_numpy(x1).cuda().requires_grad_(True); b = torch.from_numpy(x2).cuda().requires_grad_(True)
            fn = FALoss(subsample_factor=1, affinity="position", precision="f16", exact_signs=exact)
            l = fn(a, b); l.backward(); torch.cuda.synchronize()
            print(shape, "exact", exact, "ab", ab, "loss rel %.2e" % (abs(float(l) - ol) / ol), "g1 %.2e g2 %.2e" % (rn(a.grad.cpu().numpy(), o1), rn(b.grad.cpu().numpy(), o2)),
                  fn.sign_stats() if exact else "", flush=True)
os.environ["DSRL_POS_AB"] = "1"
