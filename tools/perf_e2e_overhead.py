"""Host-side cost of the drop-in FALoss call path at the training shape (what bench's e2e leg pays per step)."""
import sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from _inputs import fa_inputs
from dualsuperreslearningforsemseg_b200.models.losses import FALoss
from dualsuperreslearningforsemseg_b200.functional import FAPlan
dev = torch.device('cuda', 0)
x1h, x2h = fa_inputs((6, 1, 64, 128), "relu", 54321)
p1, p2 = torch.from_numpy(x1h).pin_memory(), torch.from_numpy(x2h).pin_memory()
a = p1.to(dev).requires_grad_(True); b = p2.to(dev).requires_grad_(True)
fn = FALoss()
plan = FAPlan((6, 1, 64, 128), device=dev)
go = torch.ones((), device=dev)
def t(f, n=2000):
    for _ in range(50): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
def fwd_bwd():
    a.grad = None; b.grad = None
    fn(a, b).backward()
def fwd_only():
    with torch.no_grad(): fn(a, b)
def full():
    u = p1.to(dev, non_blocking=True).requires_grad_(True); v = p2.to(dev, non_blocking=True).requires_grad_(True)
    l = fn(u, v); l.backward(); return l.item()
def h2d():
    p1.to(dev, non_blocking=True); p2.to(dev, non_blocking=True)
print(f"plan.forward_backward (2 ctypes calls)  {t(lambda: plan.forward_backward(a.detach(), b.detach(), go)):7.1f} us")
print(f"FALoss forward (no grad)                {t(fwd_only):7.1f} us")
print(f"FALoss forward + backward (autograd)    {t(fwd_bwd):7.1f} us")
print(f"2 x H2D of 196 KB from pinned memory    {t(h2d):7.1f} us")
print(f"full e2e step incl. .item()             {t(full):7.1f} us")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): fwd_bwd()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
