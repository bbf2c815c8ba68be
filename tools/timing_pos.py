"""Dev probe: role clocks of CTA (0,0,0) of the position-mode tile kernel.  Needs a -DDSRL_POS_TIMING build of the library
(`python tools/timing_pos.py --build` writes it to tools/_build and points DSRL_B200_LIB at it), e.g.
`python tools/timing_pos.py 8,256,128,256 f16`."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tools", "_build", "libdsrl_timing.so")      # git-ignored (*.so); travels with gpurun
if "--build" in sys.argv:
    from importlib import import_module
    sys.path.insert(0, ROOT)
    b = import_module("dualsuperreslearningforsemseg_b200.build")
    cmd = ["nvcc", *b.NVCC_FLAGS, "-DDSRL_POS_TIMING", "-I", b.INCLUDE, "-o", LIB, *b.sources()]
    subprocess.run(cmd, check=True)
    print(LIB)
    sys.exit(0)
os.environ["DSRL_B200_LIB"] = LIB
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dualsuperreslearningforsemseg_b200.functional import FAPlan  # noqa: E402

B, C, H, W = (int(v) for v in sys.argv[1].split(","))
prec = sys.argv[2] if len(sys.argv) > 2 else None
exact = len(sys.argv) > 3 and sys.argv[3] == "exact"
g = torch.Generator(device="cuda"); g.manual_seed(1)
x1 = torch.relu(torch.randn((B, C, H, W), device="cuda", generator=g)); x2 = torch.relu(torch.randn((B, C, H, W), device="cuda", generator=g))
plan = FAPlan((B, C, H, W), subsample_factor=1, affinity="position", precision=prec, exact_signs=exact)
plan.forward(x1, x2, True); plan.forward(x1, x2, True); torch.cuda.synchronize()
tm = plan.saved[64:256].view(torch.int64).cpu().tolist()

T = (H * W + 127) // 128
if prec == "f16" and os.environ.get("DSRL_POS_AB", "1") != "0" and T % 2 == 0:
    print(f"two-pass form, B={B} C={C} tiles={T}, CTA (0,0,0):\n"
          f"  pass A issuer: wait own rows {tm[0]} | loop {tm[1]} over {tm[4]} tiles = {tm[1] / max(tm[4], 1):.0f} clk/tile, wait_full {tm[2]}, wait_d_empty {tm[3]}\n"
          f"  pass A conversion warp: loop {tm[5]}, wait_d_full {tm[6]}\n"
          f"  pass B issuer: start to first MMA {tm[12]} | loop {tm[13]} over {tm[16]} tiles = {tm[13] / max(tm[16], 1):.0f} clk/tile, wait_full {tm[14]}, wait_p_full {tm[15]}\n"
          f"  pass B issuer of the last pair of row tiles (all column tiles below the diagonal): loop {tm[8]} = {tm[8] / max(tm[16], 1):.0f} clk/tile, wait_full {tm[9]}, wait_p_full {tm[10]}\n"
          f"  pass B conversion warp: loop {tm[17]}, wait_p_empty {tm[18]} | epilogue (Jacobian, dX) {tm[19]}: wait last MMAs {tm[20]}, projection pass {tm[21]}, output pass {tm[22]}")
    sys.exit(0)
print(f"B={B} C={C} prec={prec} exact={exact} tiles={T}: producer total {tm[0]} wait_empty {tm[1]} | "
      f"mma total {tm[2]} wait_full {tm[3]} wait_p(own) {tm[4]} wait_p_rem/drain {tm[5]} | epi total {tm[6]} wait_d {tm[7]} wait_xfull {tm[8]} wait_xempty {tm[9]}"
      f" conv {tm[10]} ship {tm[11]} | per column tile: mma {tm[2] / T:.0f} clk")
