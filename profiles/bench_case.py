"""One launch of the bench workload for ncu: BASELINE config[1] (1024 perturbed Solo12 trot instances, horizon 20, default
solver parameters, time slicing on), resident in HBM -- the launch bench.py times.  Prints the inner-iteration total that
profiles/summarize.py --json needs.
    ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/prof python profiles/bench_case.py"""
import sys

sys.path.insert(0, '.')
import torch

from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver

arith = int(sys.argv[1]) if len(sys.argv) > 1 else 0
b = synthetic.config(1, B=1024, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=1024)
dev = s.upload(b)
o = s.solve_resident(dev, arith=arith)
torch.cuda.synchronize()
it = o["iters"].cpu().numpy()
print(f"inner iterations of the launch: {int(it[:, 1].sum() + it[:, 2].sum())}  kernel {s.kernel_info()}")
