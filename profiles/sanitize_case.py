"""Small solve for compute-sanitizer (memcheck / racecheck): 4 instances, few iterations, both problems, incl. a
forced line-search rejection (replay path) and a long-horizon (combined roles) case."""
import sys, numpy as np
sys.path.insert(0, '.')
from bunmpc_b200 import synthetic, SolverParams
from bunmpc_b200.solver import BatchSolver
prm = SolverParams(max_outer=2, max_inner=12)
b = synthetic.perturbed(4, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=4)
r = s.solve(b, params=prm)
b.L0 = np.array([[1.0, 40.0]])
r2 = s.solve(b, params=prm)
b3 = synthetic.perturbed(2, "solo12", "bound", seed=1, horizon_scale=2.0)
r3 = BatchSolver(b3.n_col, b3.n_eff, max_batch=2).solve(b3, params=prm)
print("ok", r.iters.tolist(), r2.iters.tolist(), r3.iters.tolist())
