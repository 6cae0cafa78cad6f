"""Work / 296 CTAs against the simulated launch at the shipped setting, mean over the 15 batches.  CPU only."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'profiles')
import sched_sim
names = [f"g8192:{r}" for r in range(8)] + [f"g4096:{r}" for r in range(4)] + ["seed0", "seed5", "seed9"]
bs, ts, bz = [], [], []
for nm in names:
    v, f, x = sched_sim.per_outer_data(nm)
    run = ~np.isnan(v)
    w = ((17.2e3 * run + 1333.0 * f + 853.0 * x) * (1394.0 / 1120.0)).sum() / 1.965e6 / 296
    t, busy = sched_sim.simulate(v, f, x, 8, 1000.0)
    bs.append(w); ts.append(t); bz.append(busy)
print("work/296 mean", np.mean(bs).round(2), " simulated launch mean", np.mean(ts).round(2), " busy fraction mean", np.mean(bz).round(4), "min", np.min(bz).round(4))
