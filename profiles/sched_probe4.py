"""Companion of sched_probe3.py: the two shards of the 2-GPU global batch, a saturated batch, the bound / jump gaits at
their own horizons and the Bayes batch, at three threshold scales.  Launch time in ms.   (GPU box only)"""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver
g2 = synthetic.config(1, B=2048, seed=0)
batches = [("g2048 shard0", g2.shard(0, 2)), ("g2048 shard1", g2.shard(1, 2)),
           ("trot B=4096", synthetic.config(1, B=4096, seed=3)),
           ("bound n=24 B=2048", synthetic.perturbed(2048, "solo12", "bound", seed=1)),
           ("jump n=30 B=2048", synthetic.perturbed(2048, "solo12", "jump", seed=1))]
scales = (1000.0, 1250.0, 2500.0)
solvers = {}
for name, b in batches:
    s = solvers.setdefault(b.n_col, BatchSolver(b.n_col, 4, max_batch=4096))
    dev = s.upload(b)
    row = []
    for li in scales:
        os.environ['BUNMPC_LONG_INNER'] = repr(li)
        ts = []
        for rep in range(3):
            torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = s.solve_resident(dev); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        row.append(min(ts[1:]))
    print(f"{name:20s}", " ".join(f"{v:7.2f}" for v in row), "  (scales 1000 / 1250 / 2500)", flush=True)
