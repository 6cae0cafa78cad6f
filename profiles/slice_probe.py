"""Time of one BASELINE config[1] launch (B instances resident in HBM) for several time-slice lengths.
   python profiles/slice_probe.py [B]"""
import sys, time
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic, SolverParams
from bunmpc_b200.solver import BatchSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = synthetic.config(1, B=B, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=B)
dev = s.upload(b)
for sl in (-1, 1, 2, 4, 8, 16):
    prm = SolverParams(slice_outer=sl)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.solve_resident(dev, params=prm); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"B={B} slice_outer={sl:3d}: {best*1e3:7.2f} ms  {B/best:8.0f} solves/s")
