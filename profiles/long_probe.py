"""Long-horizon kernels (combined warp roles): cycles per inner iteration and throughput; run on the GPU box."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
for gait, B in (("trot", 592), ("bound", 592), ("jump", 592)):
    b = synthetic.perturbed(B, "solo12", gait, seed=0, horizon_scale=scale)
    s = BatchSolver(b.n_col, b.n_eff, max_batch=B)
    s.solve(b); t = time.perf_counter(); sol = s.solve(b); dt = time.perf_counter() - t
    it = sol.iters[:, 1] + sol.iters[:, 2]
    cpi = sol.cycles / it
    print(f"{gait} n={b.n_col} {s.kernel_info()} cycles/inner-iter mean {cpi.mean():.0f} min {cpi.min():.0f}; "
          f"ls mean {sol.iters[:, 3:5].sum(1).mean():.2f}; solves/s {B / dt:.0f}")
