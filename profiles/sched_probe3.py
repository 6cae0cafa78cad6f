"""Queue thresholds of the parked-instance scheduler (BUNMPC_LONG_INNER scales them) over the batches the multi-GPU runs
see: the eight shards of the 8-GPU global batch, the four of the 4-GPU one, the bench batch and two more seeds.
Launch time in ms (min of 2 after a warm-up launch) per batch and scale, then mean / worst per scale.   (GPU box only)"""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver
g8, g4 = synthetic.config(1, B=8192, seed=0), synthetic.config(1, B=4096, seed=0)
batches = [(f"g8192 shard{r}", g8.shard(r, 8)) for r in range(8)] + [(f"g4096 shard{r}", g4.shard(r, 4)) for r in range(4)]
batches += [("seed0 (bench)", synthetic.config(1, B=1024, seed=0)), ("seed5", synthetic.config(1, B=1024, seed=5)),
            ("seed9", synthetic.config(1, B=1024, seed=9))]
scales = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "600,800,1000,1250,1500,2000,2500,3500".split(","))]
s = BatchSolver(20, 4, max_batch=1024)
tab = []
for name, b in batches:
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
    tab.append(row)
    print(f"{name:16s}", " ".join(f"{v:6.2f}" for v in row), flush=True)
tab = np.array(tab)
print(f"{'scale':16s}", " ".join(f"{v:6.0f}" for v in scales))
print(f"{'mean':16s}", " ".join(f"{v:6.2f}" for v in tab.mean(0)))
print(f"{'worst':16s}", " ".join(f"{v:6.2f}" for v in tab.max(0)))
print(f"{'worst of g8192':16s}", " ".join(f"{v:6.2f}" for v in tab[:8].max(0)))
