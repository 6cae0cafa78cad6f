"""Simulated launch time over the 15 batches of sched_probe3.py for a grid of (slice length, threshold scale) and for a true
priority order.  CPU only."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'profiles')
import sched_sim
names = [f"g8192:{r}" for r in range(8)] + [f"g4096:{r}" for r in range(4)] + ["seed0", "seed5", "seed9"]
data = {nm: sched_sim.per_outer_data(nm) for nm in names}
def ev(**kw):
    t = np.array([sched_sim.simulate(*data[nm], **kw)[0] for nm in names])
    return t.mean(), t.max(), t[:8].max(), t[12]
print("slice scale   mean  worst  worst-g8192  seed0")
for sl in (7, 8, 9, 10):
    for li in (700, 850, 1000, 1150, 1300, 1500):
        m, w, w8, s0 = ev(slice_outer=sl, long_inner=float(li))
        print(f"{sl:5d} {li:5d} {m:6.2f} {w:6.2f} {w8:6.2f} {s0:6.2f}", flush=True)
for sl in (6, 7, 8, 9, 10, 12):
    m, w, w8, s0 = ev(slice_outer=sl, long_inner=1000.0, order="priority")
    print(f"{sl:5d}  prio {m:6.2f} {w:6.2f} {w8:6.2f} {s0:6.2f}", flush=True)
