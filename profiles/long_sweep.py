"""Horizons around the CTA-size steps (n = 88 | 89..128 | 129..192 | 193..208): single-solve latency and batch throughput."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic
from bunmpc_b200.motions import GAITS
from bunmpc_b200.solver import BatchSolver
for gait, hs in (("trot", (8, 9, 10, 11, 12, 13, 16, 20)), ("bound", (15, 17, 19, 21, 23))):
    base = GAITS["solo12"][gait].gait_horizon
    for h in hs:
        one = synthetic.nominal("solo12", gait, v_des=(0.3, 0.0, 0.0), horizon_scale=h / base)
        s = BatchSolver(one.n_col, one.n_eff, max_batch=512)
        s.solve(one)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); sol = s.solve(one); ts.append(1e3 * (time.perf_counter() - t0))
        many = synthetic.perturbed(512, "solo12", gait, seed=h, horizon_scale=h / base)
        dev = s.upload(many)
        s.solve_resident(dev); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve_resident(dev); e1.record(); torch.cuda.synchronize()
        ki = s.kernel_info()
        it = int(sol.iters[0, 1] + sol.iters[0, 2])
        print(gait, one.n_col, ki["threads"], f"one {min(ts):.2f} ms", f"{sol.cycles[0] / it:.0f} cycles/iter", f"batch {512 / e0.elapsed_time(e1) * 1e3:.0f}/s", flush=True)
        s.close()
