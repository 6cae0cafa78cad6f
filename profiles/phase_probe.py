"""Per-phase cycle totals of the FISTA fast path (profiling build profiles/libbunmpc_prof.so built with
-DBUNMPC_PHASE_PROF; clock reads are ordered after the data they follow by a resolved branch).
Per instance and problem (F, X) the kernel accumulates, for lane 0 of variable warp 0, row warp 0 and the scalar
warp, the cycles spent in: 0 set-up, 1 gradient, 2 division+projection, 3 variable sums, 4 momentum+stores,
5 barrier wait, 6 row work, 7 scalar work, 8 tail.  GPU box only."""
import sys, os, numpy as np
sys.path.insert(0, '.')
from bunmpc_b200 import _lib
_lib.LIB_PATH = os.path.join('profiles', 'libbunmpc_prof.so')
from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
b = synthetic.config(1, B=B, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=1024)
sol = s.solve(b, viol_hist=True)       # profiling build: the viol_hist buffer of each instance carries 64 counters
tr = sol.viol_hist.view(np.int64).reshape(-1)[: B * 64].reshape(B, 64).astype(np.float64)
names = ["setup", "grad", "proj", "sums", "mom+sts", "barrier", "row", "scalar", "tail"]
for prob, pn, itc in ((0, "F", sol.iters[:, 1]), (1, "X", sol.iters[:, 2])):
    for role, rn in enumerate(("variable warp 0", "row warp 0", "scalar warp")):
        c = tr[:, 32 * prob + 9 * role: 32 * prob + 9 * role + 9]
        per = c.sum(0) / itc.sum()
        print(f"{pn} {rn:16s} cycles per inner iteration: " + "  ".join(f"{n} {v:6.0f}" for n, v in zip(names, per)) + f"   total {per.sum():6.0f}")
