"""Tail of a launch on several batches of 1024 (the bench batch, the two shards of the 2-GPU global batch, other seeds)."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic, SolverParams
from bunmpc_b200.solver import BatchSolver
g2 = synthetic.config(1, B=2048, seed=0)
batches = [("seed0", synthetic.config(1, B=1024, seed=0)), ("g2048 shard0", g2.shard(0, 2)), ("g2048 shard1", g2.shard(1, 2)),
           ("seed5", synthetic.config(1, B=1024, seed=5)), ("seed9", synthetic.config(1, B=1024, seed=9))]
s = BatchSolver(20, 4, max_batch=1024)
for name, b in batches:
    dev = s.upload(b)
    res = []
    for li in (None, 1500, 4000):
        if li is None: os.environ.pop('BUNMPC_LONG_INNER', None)
        else: os.environ['BUNMPC_LONG_INNER'] = repr(li)
        ts = []
        for rep in range(3):
            torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = s.solve_resident(dev); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res.append(round(min(ts[1:]), 2))
    it = o["iters"][:, 1:3].sum().item()
    print(name, "inner total", it, "ms (default, 1500, 4000):", res, "bound ms @2.743us/kiter:", round(it * 2.743e-6, 2), flush=True)
