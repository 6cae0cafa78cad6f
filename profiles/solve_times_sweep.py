"""The reference's own timing harness (examples/analysis/solve_times_test.py:55-73) on this path: for trot, jump and
bound it sweeps gait_horizon = 1..20 (bound: 1..32), i.e. horizons of 10..200 / 6..192 knots, solves ONE nominal problem
per horizon (t = 0, v_des = [0.3, 0, 0], w_des = 0) and records the dynamics solve time and the final dynamics
violation.  Here, per horizon:
  cpu_ms     the same single solve by the CPU restatement on one host core (what the reference's `dyn` stamp times)
  gpu_ms     the same single solve through the host API (BatchSolver.solve, B = 1: copies + one CTA on one SM)
  gpu_batch  solves/s of a batch of 512 perturbed instances of that gait and horizon (the GPU's own regime)
  viol       final ||A_f X - b_f|| (CPU == GPU bit for bit, asserted)
Run on the GPU box:  python profiles/solve_times_sweep.py > profiles/r02_solve_times_sweep.txt"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic
from bunmpc_b200.motions import GAITS
from bunmpc_b200.solver import BatchSolver
from oracle import oracle

step = int(sys.argv[1]) if len(sys.argv) > 1 else 1
print("gait gait_horizon n_col threads ctas_per_sm cpu_ms gpu_ms gpu_batch_solves_per_s outer inner viol status")
for gait, hmax in (("trot", 20), ("jump", 20), ("bound", 32)):
    base = GAITS["solo12"][gait].gait_horizon
    for h in range(1, hmax + 1, step):
        one = synthetic.nominal("solo12", gait, v_des=(0.3, 0.0, 0.0), horizon_scale=h / base)
        t0 = time.perf_counter(); ref = oracle.solve(one, n_threads=1); cpu_ms = 1e3 * (time.perf_counter() - t0)
        s = BatchSolver(one.n_col, one.n_eff, max_batch=512)
        sol = s.solve(one)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); sol = s.solve(one); ts.append(1e3 * (time.perf_counter() - t0))
        assert np.array_equal(sol.X, ref["X"], equal_nan=True) and np.array_equal(sol.iters, ref["iters"]), (gait, h)
        many = synthetic.perturbed(512, "solo12", gait, seed=h, horizon_scale=h / base)
        dev = s.upload(many)
        s.solve_resident(dev); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve_resident(dev); e1.record(); torch.cuda.synchronize()
        ki = s.kernel_info()
        print(gait, h, one.n_col, ki["threads"], ki["ctas_per_sm"], f"{cpu_ms:.1f}", f"{min(ts):.2f}",
              f"{512 / e0.elapsed_time(e1) * 1e3:.0f}", int(sol.iters[0, 0]), int(sol.iters[0, 1] + sol.iters[0, 2]),
              f"{sol.viol[0]:.3e}", int(sol.status[0]), flush=True)
        s.close()
