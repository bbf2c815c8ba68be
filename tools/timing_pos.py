import sys, torch, ctypes
sys.path.insert(0, '/root/repo')
from dualsuperreslearningforsemseg_b200.functional import FAPlan
B, C, H, W = (int(v) for v in sys.argv[1].split(','))
prec = sys.argv[2] if len(sys.argv) > 2 else None
g = torch.Generator(device='cuda'); g.manual_seed(1)
x1 = torch.relu(torch.randn((B, C, H, W), device='cuda', generator=g)); x2 = torch.relu(torch.randn((B, C, H, W), device='cuda', generator=g))
plan = FAPlan((B, C, H, W), subsample_factor=1, affinity='position', precision=prec)
for need_grad in (True, False):
    plan.forward(x1, x2, need_grad); plan.forward(x1, x2, need_grad); torch.cuda.synchronize()     # CTA (0,0,0) of the last launch leaves its clocks
    # partials offset inside ws: find via geometry -- the timing area is 1024 doubles past the partials start; scan for it
    ws = plan.ws.view(torch.int64)
    # PosWs layout: Fpm (2x for split sizing), Fcm, nrm, partials.  Recompute offsets like make_ws (split=1 sizing used by query, but kernel used split of this precision)
    split = 1 if prec == 'fp32' else 0
    N = H * W; Npad = (N + 127) // 128 * 128; Cp = (C + 31) // 32 * 32; Kc = 2 * Cp
    au = lambda x, a: (x + a - 1) // a * a
    off = au((1 + split) * B * Npad * Kc * 4, 1024); off = au(off + (B * Kc + 128) * Npad * 4, 1024)
    if prec == 'f16':                       # FP16 copies of both layouts sit between Fcm and the norms
        off = au(off + B * Npad * Kc * 2, 1024); off = au(off + (B * Kc + 128) * Npad * 2, 1024)
    off = au(off + B * 2 * Npad * 4, 256)
    tm = ws[off // 8 + 1024: off // 8 + 1032].cpu().tolist()
    T = Npad // 128
    print(f"C={C} grad={need_grad} prec={prec} tiles={T}: producer total {tm[0]} wait_empty {tm[1]} | mma total {tm[2]} wait_full {tm[3]} wait_p {tm[4]} wait_drain {tm[5]} | epi total {tm[6]} wait_d {tm[7]}  per-tile mma {tm[2] / T:.0f}")
