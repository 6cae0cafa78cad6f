"""Small solves for compute-sanitizer (memcheck / racecheck), all through the C ABI: both problems, STRICT / FMA / MIXED
arithmetic, a forced line-search rejection (restores y_k and retries), a long horizon (more warps per CTA, generic
cross-warp totals), and -- with BUNMPC_MAX_CTAS=2 in the environment -- six instances on two resident CTAs, so that
instances are parked after one outer iteration and resumed by another CTA (time slicing, work queue).
    compute-sanitizer --tool memcheck  python profiles/sanitize_case.py
    compute-sanitizer --tool racecheck python profiles/sanitize_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, '.')
os.environ["BUNMPC_MAX_CTAS"] = "2"
from bunmpc_b200 import ARITH_FMA, ARITH_MIXED, SolverParams, synthetic
from bunmpc_b200.solver import BatchSolver

prm = SolverParams(max_outer=3, max_inner=10, slice_outer=1)
b = synthetic.perturbed(6, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=6)
r = s.solve(b, params=prm)
rf = s.solve(b, params=prm, arith=ARITH_FMA)
rm = s.solve(b, params=prm, arith=ARITH_MIXED)
b.L0 = np.array([[1.0, 40.0]])
r2 = s.solve(b, params=prm)
b3 = synthetic.perturbed(3, "solo12", "bound", seed=1, horizon_scale=2.0)
r3 = BatchSolver(b3.n_col, b3.n_eff, max_batch=3).solve(b3, params=prm)
dev = s.build_device(__import__("bunmpc_b200.motions", fromlist=["x"]).SOLO12,
                     __import__("bunmpc_b200.motions", fromlist=["x"]).solo12_trot,
                     np.array([[0.0, 0.0, 0.2]] * 6), np.zeros((6, 3)), np.zeros((6, 3)),
                     __import__("bunmpc_b200.motions", fromlist=["x"]).SOLO12.foot_pos[None].repeat(6, 0), np.zeros(6),
                     np.array([[0.2, 0.0, 0.0]] * 6), np.zeros(6))
r4 = s.solve_resident(dev, params=prm)
import torch
torch.cuda.synchronize()
print("ok", r.iters.tolist(), rf.iters[:, :3].tolist(), rm.iters[:, :3].tolist(), r2.iters.tolist(), r3.iters.tolist(),
      r4["iters"].cpu().numpy()[:, :3].tolist())
