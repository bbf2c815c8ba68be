"""Dev probe for compute-sanitizer: small two-pass problems (one / two channel groups, padding, pooling, column chunks, column split)
through FALoss, exact signs on and off.  `compute-sanitizer --tool memcheck python tools/sanitize_pos.py`"""
import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from _inputs import pos_inputs
from dualsuperreslearningforsemseg_b200.models.losses import FALoss
cases = [((2, 64, 16, 32), 1, {}), ((1, 200, 24, 40), 1, {}), ((1, 48, 32, 64), 2, {}), ((1, 256, 16, 32), 1, {"DSRL_POS_ACHUNK": "2"}),
         ((1, 200, 32, 32), 1, {"DSRL_POS_JSPLIT": "2"}), ((1, 19, 32, 32), 1, {})]
for shape, k, env in cases:
    for key in ("DSRL_POS_ACHUNK", "DSRL_POS_JSPLIT"):
        os.environ.pop(key, None)
    os.environ.update(env)
    x1, x2 = pos_inputs(shape, shape, 3)
    for exact in (True, False):
        a = torch.from_numpy(x1).cuda().requires_grad_(True); b = torch.from_numpy(x2).cuda().requires_grad_(True)
        l = FALoss(subsample_factor=k, affinity="position", precision="f16", exact_signs=exact)(a, b)
        l.backward(); torch.cuda.synchronize()
        print(shape, k, env, exact, float(l.detach()), float(a.grad.abs().sum()), flush=True)
print("done")
