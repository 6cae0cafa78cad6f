"""Ablation timing: cycles per inner iteration with pieces of the FISTA iteration removed (results are
garbage in ablated builds; only the timing is meaningful).  argv[1] = ablation mask."""
import sys, os, numpy as np
sys.path.insert(0, '.')
mask = sys.argv[1]
from bunmpc_b200 import _lib
_lib.LIB_PATH = os.path.join('profiles', 'ablate', f'lib_{mask}.so')
from bunmpc_b200 import synthetic, SolverParams
from bunmpc_b200.solver import BatchSolver
b = synthetic.config(1, B=148, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=1024)
res = []
for mi in (10, 40):
    prm = SolverParams(max_outer=10, max_inner=mi)
    s.solve(b, params=prm); sol = s.solve(b, params=prm)
    res.append((sol.iters[:, 1:3].sum(1).mean(), sol.cycles.mean()))
per_it = (res[1][1] - res[0][1]) / (res[1][0] - res[0][0])
print(f"mask {int(mask):3d}: {per_it:7.0f} cycles per inner iteration (avg of F and X), iters {res[0][0]:.0f}/{res[1][0]:.0f}")
