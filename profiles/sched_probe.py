"""Sweep of the parked-instance scheduling heuristic (BUNMPC_LONG_INNER: threshold on the predicted remaining inner
iterations above which a parked instance goes to the queue that is served first).  Results never change; time does."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic, SolverParams
from bunmpc_b200.solver import BatchSolver
for B in (1024, 8192):
    b = synthetic.config(1, B=B, seed=0)
    s = BatchSolver(b.n_col, b.n_eff, max_batch=B)
    dev = s.upload(b)
    for sl in (6,):
        for li in (1e30, 1500, 2500, 3500, 5000, 7000):
            os.environ['BUNMPC_LONG_INNER'] = repr(li)
            prm = SolverParams(slice_outer=sl)
            ts = []
            for rep in range(4):
                torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); s.solve_resident(dev, params=prm); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(B, sl, li, round(min(ts[1:]), 2), round(B / min(ts[1:]) * 1e3), flush=True)
# other workloads at the default threshold against no prediction
del os.environ['BUNMPC_LONG_INNER']
for name, b in (("bound_n24", synthetic.perturbed(1024, "solo12", "bound", seed=1)), ("jump_n30", synthetic.perturbed(1024, "solo12", "jump", seed=2)),
                ("bayes_2048", synthetic.config(4, B=2048, seed=0))):
    s = BatchSolver(b.n_col, b.n_eff, max_batch=b.B)
    dev = s.upload(b)
    for li in (1e30, None):
        if li is None: os.environ.pop('BUNMPC_LONG_INNER', None)
        else: os.environ['BUNMPC_LONG_INNER'] = repr(li)
        ts = []
        for rep in range(3):
            torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.solve_resident(dev); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(name, li, round(min(ts[1:]), 2), flush=True)
