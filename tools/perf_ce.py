"""Device time of the two cross-entropy kernels through the C-ABI at the training shape (development probe)."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from dualsuperreslearningforsemseg_b200 import _lib
B, C, H, W = 6, 19, 512, 1024
dev = torch.device('cuda', 0)
x = torch.randn((B, C, H, W), device=dev) * 3
t = torch.randint(0, C, (B, H, W), device=dev).to(torch.uint8)
L = _lib.lib(); vp = lambda z: ctypes.c_void_p(z.data_ptr())
sb = int(L.dsrl_ce_saved_bytes(B, H * W)); saved = torch.empty(sb, dtype=torch.uint8, device=dev)
loss = torch.empty((), device=dev); go = torch.ones((), device=dev); dx = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
fwd = lambda: _lib.check(L.dsrl_ce_forward(vp(x), vp(t), _lib.U8, B, C, H * W, 255, 1, vp(loss), vp(saved), sb, st))
bwd = lambda: _lib.check(L.dsrl_ce_backward(vp(x), vp(t), _lib.U8, B, C, H * W, 255, 1, vp(saved), sb, vp(go), vp(dx), st))
for name, fn in (("forward", fwd), ("backward", bwd)):
    fn(); tot = 0.0
    for _ in range(20):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    print(f"{name}: {tot / 20 * 1e3:.1f} us")
