"""Small invocations of every kernel family for compute-sanitizer (memcheck): reference-mode FA (fused + general + none),
position-mode FA (all template variants), seg counts (+ logits)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from _inputs import fa_inputs, pos_inputs, seg_case
from dualsuperreslearningforsemseg_b200.models.losses import FALoss
from dualsuperreslearningforsemseg_b200.metrices import mIoU, Accuracy

def fa(shape, k, red, **kw):
    x1, x2 = fa_inputs(shape, "relu", 1) if kw.get("affinity", "reference") == "reference" else pos_inputs(shape, shape, 1)
    a = torch.from_numpy(x1).cuda().requires_grad_(True); b = torch.from_numpy(x2).cuda().requires_grad_(True)
    l = FALoss(subsample_factor=k, reduction=red, **kw)(a, b)
    (l.sum() if l.dim() else l).backward()
    torch.cuda.synchronize()
    return float(l.sum())

print("ref fused", fa((2, 1, 64, 128), 8, "mean"))
print("ref general", fa((1, 2, 70, 133), 2, "sum"))
print("ref none", fa((1, 1, 16, 24), 4, "none"))
for prec in ("tf32", "fp32"):
    print("pos resident", prec, fa((1, 40, 16, 24), 1, "mean", affinity="position", precision=prec))
    print("pos streamed", prec, fa((1, 160, 16, 16), 1, "mean", affinity="position", precision=prec))
    with torch.no_grad():
        x1, x2 = pos_inputs((1, 40, 16, 24), (1, 40, 16, 24), 1)
        print("pos fwd-only", prec, float(FALoss(subsample_factor=1, affinity="position", precision=prec)(torch.from_numpy(x1).cuda(), torch.from_numpy(x2).cuda())))
pred, target, mask = seg_case("plain", 1, (2, 37, 53), 19)
m, a = mIoU(19), Accuracy()
m.update(pred, target, mask); a.update(pred, target, mask)
logits = torch.randn((2, 19, 37, 53), device="cuda")
m.update_from_logits(logits, torch.from_numpy(target).cuda())
print("seg", m(), a())
