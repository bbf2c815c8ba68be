"""Phase clocks of fa_ref_fused_small (needs a -DDSRL_FUSED_TIMING build: DSRL_B200_LIB=tools/libdsrl_timing.so)."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from _inputs import fa_inputs
from dualsuperreslearningforsemseg_b200.functional import FAPlan
x1, x2 = fa_inputs((6, 1, 64, 128), "relu", 54321)
a, b = torch.from_numpy(x1).cuda(), torch.from_numpy(x2).cuda()
plan = FAPlan((6, 1, 64, 128), subsample_factor=8)
for _ in range(5):
    plan.forward(a, b, True); torch.cuda.synchronize()
    tm = plan.ws.view(torch.int64)[64:73].cpu().numpy()
    dbg = plan.ws.view(torch.int64)[73:81].cpu().numpy()
    print("  solver warp (m <= 8): M, squarings, start, power steps, final pair:", np.diff(dbg[:6]), " start after kernel begin:", dbg[0] - tm[0], " rank: setup", dbg[6] - tm[4], "searches", dbg[7] - dbg[6], "rest", tm[5] - dbg[7])
    print("phase cycles (pool, Gram+sort || sigma solve, -, scale+vote, rank, partial, grad, finish):", np.diff(tm), "total", tm[8] - tm[0])
