"""Stage-level cycle counters of the FISTA fast path (profiling build, -DBUNMPC_PHASE_PROF; clock reads are
ordered after the data they follow by a resolved branch).  Run on the GPU box."""
import sys, os, numpy as np
sys.path.insert(0, '.')
from bunmpc_b200 import _lib
_lib.LIB_PATH = os.path.join('profiles', 'libbunmpc_prof.so')
from bunmpc_b200 import synthetic
from bunmpc_b200.solver import BatchSolver

b = synthetic.config(1, B=1024, seed=0)
s = BatchSolver(b.n_col, b.n_eff, max_batch=1024)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
sub = b.select(np.arange(B))
sol = s.solve(sub, viol_hist=True)       # the profiling build returns the counters in the viol_hist buffer
prof = sol.viol_hist.view(np.int64)[:, :64]
itf, itx, outer = (sol.iters[:, i].astype(float) for i in (1, 2, 0))
print(f"B={B}: total cycles/inner-iter {np.mean(sol.cycles/(itf+itx)):.0f}")
names = ["setup", "gradient", "div+project", "var_sums", "momentum+STS", "barrier", "row work", "stage2", "epilogue"]
for lab, off, it in (("F", 0, itf), ("X", 32, itx)):
    for role, rn in enumerate(("var warp 0", "row warp 0", "scalar warp")):
        c = prof[:, off + role * 9: off + role * 9 + 9]
        per_it = {names[i]: np.mean(c[:, i] / it) for i in range(1, 8)}
        per_call = {names[i]: np.mean(c[:, i] / outer) for i in (0, 8)}
        print(f"  {lab} {rn:12s} per iteration: " + " ".join(f"{k}={v:.0f}" for k, v in per_it.items() if v > 0.5)
              + " | per call: " + " ".join(f"{k}={v:.0f}" for k, v in per_call.items()))
