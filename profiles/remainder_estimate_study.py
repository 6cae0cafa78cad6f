"""How good is the remainder estimate that files parked instances into the queues (kernels.cuh, parking code)?  CPU study
with the oracle: one batch of 1024 solved as 100 chained solves of ONE outer iteration each (the state a park carries is
the state the reference carries across outer iterations, so the chain IS the solve), which gives the violation and the
inner-iteration counts after every outer iteration; then the kernel's estimate at every park point (slices of 8 outer
iterations) against what the instance really had left.      python profiles/remainder_estimate_study.py [shard 0..3]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from oracle import oracle
from bunmpc_b200 import synthetic

oracle.build()
shard = int(sys.argv[1]) if len(sys.argv) > 1 else 2
b = synthetic.config(1, B=4096, seed=0).shard(shard, 4)
B, n, e = b.B, b.n_col, b.n_eff
nx, nf = 9 * (n + 1), 3 * e * n
ex = oracle.expand(b)
X = np.tile(b.x_init, (1, n + 1)); F = np.zeros((B, nf)); P = np.zeros((B, nx)); L = np.broadcast_to(b.L0, (B, 2)).copy()
prm = oracle.default_params(max_outer=1)
exit_tol, max_outer, slice_outer = 1e-3, 100, 8
viol = np.full((B, max_outer), np.nan); inner = np.zeros((B, max_outer), dtype=np.int64)
alive = np.arange(B)
for k in range(max_outer):
    sel = lambda a: a[alive] if a.shape[0] == B else a
    r = oracle.solve_expanded(n, e, sel(np.broadcast_to(b.m, (B,))), sel(np.broadcast_to(b.rho, (B,))), b.x_init[alive], b.cnt_plan[alive],
                              b.dt[alive], ex["Qx"][alive], ex["qx"][alive], ex["Qf"][alive], ex["qf"][alive], ex["lbx"][alive],
                              ex["ubx"][alive], X[alive], F[alive], P[alive], L[alive], params=prm, n_threads=8)
    X[alive], F[alive], P[alive], L[alive] = r["X"], r["F"], r["P"], r["L"]
    viol[alive, k] = r["viol"]; inner[alive, k] = r["iters"][:, 1] + r["iters"][:, 2]
    alive = alive[~(r["viol"] < exit_tol) & ~np.isnan(r["viol"])]
    if len(alive) == 0:
        break
outer = (~np.isnan(viol)).sum(1)
total = inner.sum(1)
full = oracle.solve(b, n_threads=8)["iters"]
assert np.array_equal(outer, full[:, 0]) and np.array_equal(total, full[:, 1] + full[:, 2]), "the chain is the solve"
print(f"shard {shard}: {B} instances, outer mean {outer.mean():.1f}, capped {int((outer >= max_outer).sum())}, inner total {total.sum()}")

# the kernel's estimate at the end of every slice (k = slice_outer outer iterations in the slice, span 4)
rows = []
for i in range(B):
    for o in range(slice_outer, outer[i], slice_outer):          # parks after o outer iterations (o < outer[i] <= max_outer)
        if o >= max_outer:
            break
        v1, v0 = np.float32(viol[i, o - 1]), np.float32(viol[i, o - 1 - 4])
        rate = (np.log(v0) - np.log(v1)) / np.float32(4)
        rem = float(max_outer - o)
        if rate > 1e-3:
            rem = min(rem, max((np.log(v1) - np.log(np.float32(exit_tol))) / rate, 0.0))
        per_outer = inner[i, o - slice_outer:o].sum() / slice_outer
        rows.append((i, o, rem * per_outer, inner[i, o:].sum(), outer[i] >= max_outer))
rows = np.array(rows, dtype=float)
pred, true, capped = rows[:, 2], rows[:, 3], rows[:, 4] > 0
print(f"{len(rows)} park points; correlation of log(estimate + 1) and log(true remainder + 1): "
      f"{np.corrcoef(np.log1p(pred), np.log1p(true))[0, 1]:.3f}")
for lab, m in (("instances that converge", ~capped), ("instances that hit the cap", capped)):
    ratio = (pred[m] + 1) / (true[m] + 1)
    print(f"  {lab:28s} {int(m.sum()):5d} park points: estimate / true  median {np.median(ratio):.2f}, "
          f"10 % {np.percentile(ratio, 10):.2f}, 90 % {np.percentile(ratio, 90):.2f}; under-estimated by more than 2x: {100 * (ratio < 0.5).mean():.1f} %")
for o in (8, 16, 24, 32, 48, 64):
    m = capped & (rows[:, 1] == o)
    if m.any():
        print(f"  capped instances at the park after {o:3d} outer iterations: estimate / true  median {np.median((pred[m] + 1) / (true[m] + 1)):.2f}, "
              f"under 0.5: {100 * ((pred[m] + 1) / (true[m] + 1) < 0.5).mean():.0f} %")
