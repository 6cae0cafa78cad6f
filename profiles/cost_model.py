"""Least-squares cost model of the solve kernel from the per-instance cycle counters:
cycles = a * outer iterations + b * F iterations + c * X iterations.  GPU box only."""
import sys, numpy as np
sys.path.insert(0, '.')
from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver
for gait, scale in (("trot", 1.0), ("bound", 1.0), ("jump", 1.0), ("trot", 2.0)):
    b = synthetic.perturbed(148, "solo12", gait, seed=0, horizon_scale=scale)
    s = BatchSolver(b.n_col, b.n_eff, max_batch=148)
    s.solve(b); sol = s.solve(b)
    ok = sol.status != 2
    A = np.stack([sol.iters[ok, 0], sol.iters[ok, 1], sol.iters[ok, 2]], 1).astype(float)
    coef, res, *_ = np.linalg.lstsq(A, sol.cycles[ok].astype(float), rcond=None)
    tot = (A * coef).sum(0)
    print(f"{gait} n={b.n_col}: per outer iteration {coef[0]:.0f} cycles, per F iteration {coef[1]:.0f}, per X iteration {coef[2]:.0f}; "
          f"shares: outer {tot[0] / tot.sum():.1%}, F {tot[1] / tot.sum():.1%}, X {tot[2] / tot.sum():.1%}; "
          f"fit error {np.abs(A @ coef - sol.cycles[ok]).mean() / sol.cycles[ok].mean():.2%}")
