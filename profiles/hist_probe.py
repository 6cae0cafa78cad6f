import sys, numpy as np
sys.path.insert(0, '.')
from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver
b = synthetic.config(1, B=1024, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=1024)
sol = s.solve(b, viol_hist=True)
np.savez_compressed('gpurun_out/hist_c1.npz', iters=sol.iters, viol_hist=sol.viol_hist, status=sol.status, cycles=sol.cycles)
print(sol.iters[:3], sol.viol_hist[:2, :10])
