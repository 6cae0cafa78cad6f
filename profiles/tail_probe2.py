"""Why do some batches of 1024 end 5 % later than others with the same work?  Launch time against the slice length and the
queue thresholds for a slow shard, a fast shard and the bench batch.     (GPU box only)"""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic, SolverParams
from bunmpc_b200.solver import BatchSolver
g4 = synthetic.config(1, B=4096, seed=0)
batches = [("g4096 shard2 (slow)", g4.shard(2, 4)), ("g4096 shard0 (fast)", g4.shard(0, 4)), ("seed0 (bench)", synthetic.config(1, B=1024, seed=0))]
s = BatchSolver(20, 4, max_batch=1024)
for name, b in batches:
    dev = s.upload(b)
    for sl in (2, 3, 4, 5, 6, 7, 8):
        row = []
        for li in (1000, 2500, 5000):
            os.environ['BUNMPC_LONG_INNER'] = repr(li)
            ts = []
            for rep in range(3):
                torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); o = s.solve_resident(dev, params=SolverParams(slice_outer=sl)); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            row.append(round(min(ts[1:]), 2))
        print(name, "slice", sl, "ms at long_inner 1000 / 2500 / 5000:", row, flush=True)
    cyc = o["cycles"].cpu().numpy().astype(float); it = o["iters"].cpu().numpy()
    tot = it[:, 1] + it[:, 2]
    print(name, "cycles per inner iteration: mean", round(cyc.sum() / tot.sum(), 1), " busy time / (296 CTAs x launch):",
          round(cyc.sum() / 1.965e6 / 296 / min(ts[1:]), 4), flush=True)
