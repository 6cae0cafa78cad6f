"""Separates per-inner-iteration cost from per-outer-iteration overhead (set_data, matrix builds) using the
in-kernel cycle counter: solve with different max_inner caps and fit cycles = a*outer + b*inner."""
import sys, numpy as np
sys.path.insert(0, '.')
from bunmpc_b200 import synthetic, SolverParams
from bunmpc_b200.solver import BatchSolver
b = synthetic.config(1, B=148, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=1024)
rows = []
for mi in (2, 5, 10, 20, 40, 80, 150):
    prm = SolverParams(max_outer=30, max_inner=mi)
    s.solve(b, params=prm); sol = s.solve(b, params=prm)
    rows.append((sol.iters[:, 0].astype(float), sol.iters[:, 1].astype(float), sol.iters[:, 2].astype(float), sol.cycles.astype(float)))
    print(mi, 'outer', sol.iters[:, 0].mean(), 'F', sol.iters[:, 1].mean(), 'X', sol.iters[:, 2].mean(), 'cycles', sol.cycles.mean())
A = np.concatenate([np.stack([r[0], r[1], r[2]], 1) for r in rows]); y = np.concatenate([r[3] for r in rows])
coef, *_ = np.linalg.lstsq(A, y, rcond=None)
print('cycles per outer iteration (overhead): %.0f, per F inner iteration: %.0f, per X inner iteration: %.0f' % tuple(coef))
