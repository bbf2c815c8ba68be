"""Dev probe: gradient error (sampled rows, float64 oracle) and step time of the exact-sign path against the listing threshold
(DSRL_POS_KSIGMA, in sigmas of the operand-rounding error) at BASELINE configs[3]'s size, one sample."""
import os, sys, time
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from _inputs import pos_inputs
from oracle import fa_oracle
from dualsuperreslearningforsemseg_b200.functional import FAPlan
C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x1, x2 = pos_inputs((1, C, 128, 256), (1, C, 128, 256), 54321)
rows = np.random.default_rng(3).choice(128 * 256, size=192, replace=False)
_, o1, o2 = fa_oracle.fa_position_rows(x1, x2, rows, 1, "mean")
a = torch.from_numpy(x1).cuda(); b = torch.from_numpy(x2).cuda(); go = torch.ones((), device="cuda")
def rn(p, q): return float(np.linalg.norm(p.astype(np.float64) - q) / np.linalg.norm(q))
for ks in ("3.5", "3.0", "2.5", "2.0", "1.5", "1.0", "0.5"):
    os.environ["DSRL_POS_KSIGMA"] = ks
    plan = FAPlan(a.shape, subsample_factor=1, affinity="position", precision="f16", device=a.device, exact_signs=True)
    for _ in range(2): plan.forward_backward(a, b, go)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): plan.forward_backward(a, b, go)
    e1.record(); torch.cuda.synchronize()
    d1, d2 = plan.dx1.cpu().numpy(), plan.dx2.cpu().numpy()
    g1 = d1[0].reshape(C, -1)[:, rows]; g2 = d2[0].reshape(C, -1)[:, rows]
    print(f"C={C} ksigma {ks}: {e0.elapsed_time(e1)/5:.3f} ms  grad relnorm {rn(g1, o1):.2e} {rn(g2, o2):.2e}  {plan.sign_stats()}", flush=True)
