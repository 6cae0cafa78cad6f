"""Cycles per inner iteration vs. resident CTAs (in-kernel clock64 per instance); run on the GPU box."""
import sys, numpy as np
sys.path.insert(0,'.')
from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver
b = synthetic.config(1, B=1024, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=8192)
print(s.kernel_info())
for B in (1, 148, 296, 1024):
    sub = b.select(np.arange(B))
    s.solve(sub); sol = s.solve(sub)
    it = sol.iters[:,1]+sol.iters[:,2]
    cpi = sol.cycles/it
    print(f"B={B:5d} cycles/inner-iter: mean {cpi.mean():7.0f} min {cpi.min():7.0f} max {cpi.max():7.0f}; outer mean {sol.iters[:,0].mean():.1f}; solve ms p50 {np.median(sol.cycles)/1.965e6:.2f}")
b8 = synthetic.config(1, B=8192, seed=0)
import time
s.solve(b8); t=time.perf_counter(); sol=s.solve(b8); dt=time.perf_counter()-t
print('B=8192 e2e solves/s', 8192/dt)
