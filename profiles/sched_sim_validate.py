"""sched_sim.py against the measured launch times of profiles/r02_sched_probe3.txt (15 batches x 8 threshold scales).  CPU only."""
import sys, re, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'profiles')
import sched_sim
names = [f"g8192:{r}" for r in range(8)] + [f"g4096:{r}" for r in range(4)] + ["seed0", "seed5", "seed9"]
scales = [600, 800, 1000, 1250, 1500, 2000, 2500, 3500]
meas = []
for ln in open('profiles/r02_sched_probe3.txt'):
    m = re.match(r"^(g\d+ shard\d|seed\d)", ln)
    if m:
        meas.append([float(x) for x in ln.split()[-8:]])
meas = np.array(meas)
data = {nm: sched_sim.per_outer_data(nm) for nm in names}
sim = np.array([[sched_sim.simulate(*data[nm], 8, li)[0] for li in scales] for nm in names])
print("measured mean per scale ", meas.mean(0).round(2))
print("simulated mean per scale", sim.mean(0).round(2))
print("correlation of all 120 (batch, scale) pairs:", np.corrcoef(meas.ravel(), sim.ravel())[0, 1].round(3),
      " after removing each batch's mean:", np.corrcoef((meas - meas.mean(1, keepdims=True)).ravel(), (sim - sim.mean(1, keepdims=True)).ravel())[0, 1].round(3))
print("mean |sim - meas - offset|:", np.abs(sim - meas - (sim - meas).mean()).mean().round(3), "offset", (sim - meas).mean().round(3))
