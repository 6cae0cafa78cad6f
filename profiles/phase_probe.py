"""Per-phase cycle totals of one FISTA iteration (profiling build profiles/libbunmpc_prof.so, built by
`make -C bunmpc_b200/csrc prof` with -DBUNMPC_PHASE_PROF).  Per instance and problem (F, X) the kernel accumulates,
for thread 0, the cycles between: 0 start of the iteration -> end of the first phase (gradient, prox, sums, row sums of
y_k), 1 first barrier, 2 second phase (row sums of y_k_1, momentum store, warp reduction), 3 second barrier,
4 totals + line-search test.  GPU box only:   python profiles/phase_probe.py [B]"""
import os
import sys

import numpy as np

sys.path.insert(0, '.')
os.environ["BUNMPC_LIB"] = os.path.join('profiles', 'libbunmpc_prof.so')
from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
b = synthetic.config(1, B=B, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=B)
sol = s.solve(b, viol_hist=True)       # profiling build: the viol_hist buffer of each instance carries 32 counters
tr = sol.viol_hist.view(np.int64).reshape(-1)[: B * 32].reshape(B, 32).astype(np.float64)
names = ["leaves/stores/shift (X: after prox)", "barrier", "-", "-", "-", "totals+decision (F: +rows)", "gradient (X: +rows)", "warp tree", "prox (X: +leaves)"]
for prob, pn, itc in ((0, "F", sol.iters[:, 1]), (1, "X", sol.iters[:, 2])):
    per = tr[:, 16 * prob: 16 * prob + 9].sum(0) / itc.sum()
    print(f"B={B} {pn} cycles per inner iteration (thread 0): " + "  ".join(f"{n} {v:6.0f}" for n, v in zip(names, per))
          + f"   total {per.sum():6.0f}")
print("kernel", s.kernel_info(), "cycles per inner iteration per CTA", sol.cycles.sum() / (sol.iters[:, 1] + sol.iters[:, 2]).sum())
