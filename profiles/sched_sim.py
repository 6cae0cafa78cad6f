"""Simulator of the parked-instance scheduler of solve_kernel (kernels.cuh, work queue + parking code) on the CPU: 296
CTAs, fresh instances first, then eight threshold queues served from the longest predicted remainder down, first in first
out inside a queue; a slice costs what the cost model of DESIGN.md section 6 says (cycles per outer / force / state
iteration of a CTA that shares its SM).  Input: the per-outer-iteration data of one batch, produced with the oracle as a
chain of one-outer-iteration solves (which is the solve, bit for bit).
    python profiles/sched_sim.py [g4096:2 | g8192:r | seedN]          (CPU only; ~1 min per batch, cached in /tmp)"""
import heapq
import sys
import numpy as np
sys.path.insert(0, '.')


def batch_of(name):
    """'g8192:3' = shard 3 of the 8-GPU global batch (likewise g4096:r, g2048:r), 'seed5' = config[1] with that seed"""
    from bunmpc_b200 import synthetic
    if name.startswith("g"):
        tot, r = name[1:].split(":")
        return synthetic.config(1, B=int(tot), seed=0).shard(int(r), int(tot) // 1024)
    return synthetic.config(1, B=1024, seed=int(name[4:]))


def per_outer_data(name, cache="/tmp/bunmpc_sched_sim"):
    import os
    f = os.path.join(cache, name.replace(":", "_") + ".npz")
    if os.path.exists(f):
        z = np.load(f)
        return z["viol"], z["itf"], z["itx"]
    from oracle import oracle
    oracle.build()
    b = batch_of(name)
    B, n, e = b.B, b.n_col, b.n_eff
    nx, nf = 9 * (n + 1), 3 * e * n
    ex = oracle.expand(b)
    X = np.tile(b.x_init, (1, n + 1)); F = np.zeros((B, nf)); P = np.zeros((B, nx)); L = np.broadcast_to(b.L0, (B, 2)).copy()
    prm = oracle.default_params(max_outer=1)
    viol = np.full((B, 100), np.nan); itf = np.zeros((B, 100), dtype=np.int64); itx = np.zeros((B, 100), dtype=np.int64)
    alive = np.arange(B)
    m, rho = np.broadcast_to(b.m, (B,)), np.broadcast_to(b.rho, (B,))
    for k in range(100):
        r = oracle.solve_expanded(n, e, m[alive], rho[alive], b.x_init[alive], b.cnt_plan[alive], b.dt[alive], ex["Qx"][alive],
                                  ex["qx"][alive], ex["Qf"][alive], ex["qf"][alive], ex["lbx"][alive], ex["ubx"][alive],
                                  X[alive], F[alive], P[alive], L[alive], params=prm, n_threads=8)
        X[alive], F[alive], P[alive], L[alive] = r["X"], r["F"], r["P"], r["L"]
        viol[alive, k] = r["viol"]; itf[alive, k] = r["iters"][:, 1]; itx[alive, k] = r["iters"][:, 2]
        alive = alive[~(r["viol"] < 1e-3) & ~np.isnan(r["viol"])]
        if len(alive) == 0:
            break
    os.makedirs(cache, exist_ok=True)
    np.savez_compressed(f, viol=viol, itf=itf, itx=itx)
    return viol, itf, itx


def simulate(viol, itf, itx, slice_outer=8, long_inner=1000.0, n_cta=296, share=1394.0 / 1120.0, ghz=1.965,
             order="fifo"):
    """-> (launch time in ms, busy fraction).  Cost of one outer iteration: (17.2 k + 1333 it_f + 853 it_x) cycles alone on
    an SM (profiles/cost_model.py), x `share` with two busy CTAs per SM (1120 -> 1394 cycles per iteration, section 6)."""
    B = viol.shape[0]
    outer_n = (~np.isnan(viol)).sum(1)
    cost = (17.2e3 + 1333.0 * itf + 853.0 * itx) * share / (ghz * 1e6)          # ms per outer iteration
    c = np.array([0.2, 0.4, 0.8, 1.2, 1.8, 2.6, 3.6]) * long_inner
    done_outer = np.zeros(B, dtype=int)
    queues = [[] for _ in range(8)]                  # FIFO lists of (key, instance)
    fresh, finished, busy = 0, 0, 0.0
    events = [(0.0, k, -1) for k in range(n_cta)]    # (time a CTA becomes free, cta, instance it just ran)
    heapq.heapify(events)
    idle, t_end = [], 0.0

    def take(t, cta):
        nonlocal fresh, busy
        i = -1
        if fresh < B:
            i = fresh; fresh += 1
        else:
            for q in range(7, -1, -1):
                if queues[q]:
                    if order == "fifo":
                        i = queues[q].pop(0)[1]
                    else:                            # true priority inside the queue: longest predicted remainder first
                        j = max(range(len(queues[q])), key=lambda k: queues[q][k][0]); i = queues[q].pop(j)[1]
                    break
        if i < 0:
            return False
        o0 = done_outer[i]
        o1 = min(o0 + slice_outer, outer_n[i]) if outer_n[i] - o0 > slice_outer or True else outer_n[i]
        if o1 < outer_n[i] and o1 >= 100:
            o1 = outer_n[i]
        d = cost[i, o0:o1].sum()
        busy += d
        done_outer[i] = o1
        heapq.heappush(events, (t + d, cta, i))
        return True

    while events:
        t, cta, i = heapq.heappop(events)
        if i >= 0:
            o = done_outer[i]
            if o >= outer_n[i]:
                finished += 1; t_end = max(t_end, t)
            else:                                    # park: the kernel's estimate (span 4 of the slice just run)
                k = min(slice_outer, o)
                span = 4 if k > 4 else k - 1
                q = 0
                if span > 0:
                    v1, v0 = np.float32(viol[i, o - 1]), np.float32(viol[i, o - 1 - span])
                    rate = (np.log(v0) - np.log(v1)) / np.float32(span)
                    rem = float(100 - o)
                    if rate > 1e-3:
                        rem = min(rem, max(float((np.log(v1) - np.log(np.float32(1e-3))) / rate), 0.0))
                    work = rem * (itf[i, o - k:o].sum() + itx[i, o - k:o].sum()) / k
                    q = int((work > c).sum())
                else:
                    work = 0.0
                queues[q].append((work, i))
                while idle and any(queues):
                    take(t, idle.pop())
        if not take(t, cta):
            if finished < B:
                idle.append(cta)
    return t_end, busy / (n_cta * t_end)


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "g4096:2"
    viol, itf, itx = per_outer_data(name)
    cost_all = ((17.2e3 * ~np.isnan(viol) + 1333.0 * itf + 853.0 * itx) * (1394.0 / 1120.0)).sum() / 1.965e6
    print(f"{name}: work / 296 CTAs = {cost_all / 296:.2f} ms (every CTA sharing its SM all the time)")
    for sl in (2, 4, 6, 7, 8, 12):
        row = [simulate(viol, itf, itx, sl, li)[0] for li in (600, 1000, 1250, 2500, 5000)]
        print(f"  slice {sl:2d}: launch ms at threshold scale 600 / 1000 / 1250 / 2500 / 5000:", " ".join(f"{v:6.2f}" for v in row))
    for sl in (4, 8):
        print(f"  slice {sl:2d}, true priority order (longest estimated remainder first, no thresholds): "
              f"{simulate(viol, itf, itx, sl, 1000.0, order='priority')[0]:6.2f}")
