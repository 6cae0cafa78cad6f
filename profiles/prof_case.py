"""A short, saturated solve for ncu: B instances of BASELINE config[1] resident in HBM, a few outer iterations.
   python profiles/prof_case.py [B] [max_outer] [arith] [gait]      (prints solves/s and cycles per inner iteration)"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic, SolverParams
from bunmpc_b200.solver import BatchSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
mo = int(sys.argv[2]) if len(sys.argv) > 2 else 6
arith = int(sys.argv[3]) if len(sys.argv) > 3 else 0
gait = sys.argv[4] if len(sys.argv) > 4 else "trot"
b = synthetic.perturbed(B, "solo12", gait, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=B)
dev = s.upload(b)
prm = SolverParams(max_outer=mo, slice_outer=-1)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = s.solve_resident(dev, params=prm, arith=arith)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
it = o["iters"].cpu().numpy()
cyc = o["cycles"].cpu().numpy()
inner = it[:, 1] + it[:, 2]
info = s.kernel_info()
per_sm = inner.sum() / info["num_sms"]
print(f"total inner iterations of one launch: {int(inner.sum())}")
print(f"B={B} n={b.n_col} max_outer={mo} arith={arith}: {dt*1e3:.2f} ms, {B/dt:.0f} solves/s, inner F/X mean {it[:,1].mean():.0f}/{it[:,2].mean():.0f}, "
      f"cycles per inner iteration per CTA {cyc.sum()/inner.sum():.0f}, SM cycles per inner iteration {dt*1.965e9/per_sm:.0f}, kernel {info}")
