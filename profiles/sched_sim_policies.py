"""Policy variants in the simulator (a longer first slice; a floor under the early estimates), 15 batches.  CPU only."""
import sys, numpy as np, heapq
sys.path.insert(0, '.'); sys.path.insert(0, 'profiles')
import sched_sim
names = [f"g8192:{r}" for r in range(8)] + [f"g4096:{r}" for r in range(4)] + ["seed0", "seed5", "seed9"]
data = {nm: sched_sim.per_outer_data(nm) for nm in names}

def simulate(viol, itf, itx, slice_outer=8, long_inner=1000.0, first_slice=None, floor_until=0, floor_frac=0.5, n_cta=296,
             end_slices=None, halving=0):
    """end_slices = ((unfinished instances <= a, slice), ...): shorter slices once fewer instances than that are unfinished"""
    B = viol.shape[0]
    outer_n = (~np.isnan(viol)).sum(1)
    cost = (17.2e3 + 1333.0 * itf + 853.0 * itx) * (1394.0 / 1120.0) / 1.965e6
    c = np.array([0.2, 0.4, 0.8, 1.2, 1.8, 2.6, 3.6]) * long_inner
    done = np.zeros(B, dtype=int); last_k = np.zeros(B, dtype=int); rem_est = np.full(B, 1e9)
    queues = [[] for _ in range(8)]
    fresh = 0; finished = 0
    events = [(0.0, k, -1) for k in range(n_cta)]; heapq.heapify(events)
    idle = []; t_end = 0.0
    def take(t, cta):
        nonlocal fresh
        i = -1
        if fresh < B: i = fresh; fresh += 1
        else:
            for q in range(7, -1, -1):
                if queues[q]: i = queues[q].pop(0); break
        if i < 0: return False
        o0 = done[i]
        sl = first_slice if (o0 == 0 and first_slice) else slice_outer
        if halving and rem_est[i] < 2 * slice_outer:      # the slice shrinks with the estimated remainder: ceil(rem / halving), at least 1
            sl = int(min(sl, max(1, np.ceil(rem_est[i] / halving))))
        for a, s_ in (end_slices or ()):
            if B - finished <= a: sl = min(sl, s_)
        o1 = min(o0 + sl, outer_n[i])
        last_k[i] = o1 - o0
        done[i] = o1
        heapq.heappush(events, (t + cost[i, o0:o1].sum(), cta, i))
        return True
    while events:
        t, cta, i = heapq.heappop(events)
        if i >= 0:
            o = done[i]
            if o >= outer_n[i]:
                finished += 1; t_end = max(t_end, t)
            else:
                k = last_k[i]; span = 4 if k > 4 else k - 1
                q = 0
                if span > 0:
                    v1, v0 = np.float32(viol[i, o - 1]), np.float32(viol[i, o - 1 - span])
                    rate = (np.log(v0) - np.log(v1)) / np.float32(span)
                    rem = float(100 - o)
                    if rate > 1e-3: rem = min(rem, max(float((np.log(v1) - np.log(np.float32(1e-3))) / rate), 0.0))
                    per = (itf[i, o - k:o].sum() + itx[i, o - k:o].sum()) / k
                    if o <= floor_until: rem = max(rem, floor_frac * (100 - o))
                    rem_est[i] = rem
                    q = int((rem * per > c).sum())
                queues[q].append(i)
                while idle and any(queues): take(t, idle.pop())
        if not take(t, cta):
            if finished < B: idle.append(cta)
    return t_end

def ev(**kw):
    t = np.array([simulate(*data[nm], **kw) for nm in names])
    return f"mean {t.mean():6.2f} worst {t.max():6.2f} worst-g8192 {t[:8].max():6.2f} seed0 {t[12]:6.2f}"
print("shipped (8, 1000)              ", ev())
for fs in (12, 16, 24):
    print(f"first slice {fs:2d}, then 8         ", ev(first_slice=fs), flush=True)
for fu, ff in ((16, 0.5), (16, 0.8), (24, 0.5), (24, 0.8), (32, 0.6)):
    print(f"floor until {fu} frac {ff}        ", ev(floor_until=fu, floor_frac=ff), flush=True)
for fs, fu, ff in ((16, 24, 0.5), (16, 32, 0.6)):
    print(f"first {fs}, floor until {fu} frac {ff}", ev(first_slice=fs, floor_until=fu, floor_frac=ff), flush=True)
# Shorter slices near the end, switched by the number of unfinished instances: no effect at all for thresholds up to 700 of
# 1024 -- longest-remainder-first makes nearly everything end together, so "few unfinished" comes too late to matter; a
# switch on the number of instances (1024, 2) is the short-slice case again (31 ms).
for es in (((296, 4), (148, 2)), ((700, 4), (450, 1)), ((1024, 2),)):
    print(f"end slices {str(es):40s}", ev(end_slices=es), flush=True)
# The slice of a parked instance shrinks with its estimated remaining outer iterations (ceil(rem / h) once rem < 16):
for h in (1.5, 2, 3):
    print(f"slice = ceil(estimated remainder / {h})          ", ev(halving=h), flush=True)
