"""Per-slot timestamp trace of the FISTA fast path (profiling build profiles/libbunmpc_prof.so built with
-DBUNMPC_PHASE_PROF; clock reads are ordered after the data they follow by a resolved branch).
Traces instance 0, outer iteration 1: lane 0 of variable warp 0, row warp 0 and the scalar warp.  GPU box only."""
import sys, os, numpy as np
sys.path.insert(0, '.')
from bunmpc_b200 import _lib
_lib.LIB_PATH = os.path.join('profiles', 'libbunmpc_prof.so')
from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32   # >= 32: the trace needs 3072 doubles of the buffer
b = synthetic.config(1, B=max(B, 1), seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=1024)
sol = s.solve(b, viol_hist=True)       # profiling build: the viol_hist buffer of instance 0.. carries the trace
tr = sol.viol_hist.view(np.int64).reshape(-1)[: 2 * 3 * 64 * 8].reshape(2, 3, 64, 8)
for prob, pn in ((0, "F"), (1, "X")):
    t = tr[prob].astype(np.float64)
    base = t[0, 10, 0]
    print(f"== {pn} problem, slots 10..17 (cycles relative to slot 10 start of the variable warp) ==")
    for sl in range(10, 18):
        v, r, c = t[0, sl] - base, t[1, sl] - base, t[2, sl] - base
        print(f"slot {sl}: var start {v[0]:6.0f} grad {v[1]:6.0f} proj {v[2]:6.0f} sums {v[3]:6.0f} sts {v[4]:6.0f} bar {v[5]:6.0f} |"
              f" row start {r[0]:6.0f} done {r[6]:6.0f} bar {r[5]:6.0f} | scalar start {c[0]:6.0f} done {c[7]:6.0f} bar {c[5]:6.0f}")
    d = np.diff(t[0, 5:60, 0])
    print(f"   slot period: mean {d.mean():.0f} min {d.min():.0f} max {d.max():.0f}")
