import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic, SolverParams
from bunmpc_b200.solver import BatchSolver
b = synthetic.config(1, B=1024, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=1024)
dev = s.upload(b)
for sl in (-1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24):
    prm = SolverParams(slice_outer=sl)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve_resident(dev, params=prm); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(sl, [round(t, 2) for t in ts[1:]])
