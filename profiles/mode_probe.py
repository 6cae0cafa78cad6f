"""Arithmetic modes x residency variants of the 128-thread kernel on the bench batch (and saturated): ms per launch."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from bunmpc_b200 import synthetic, ARITH_STRICT, ARITH_FMA, ARITH_MIXED
from bunmpc_b200.solver import BatchSolver
for B in (1024, 8192):
    b = synthetic.config(1, B=B, seed=0)
    for ctas in (None, "3"):
        if ctas: os.environ["BUNMPC_CTAS"] = ctas
        else: os.environ.pop("BUNMPC_CTAS", None)
        s = BatchSolver(b.n_col, b.n_eff, max_batch=B)
        dev = s.upload(b)
        for name, ar in (("strict", ARITH_STRICT), ("fma", ARITH_FMA), ("mixed", ARITH_MIXED)):
            ts = []
            for rep in range(3):
                torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); s.solve_resident(dev, arith=ar); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(B, "ctas/SM", ctas or 2, name, round(min(ts[1:]), 2), "ms", round(B / min(ts[1:]) * 1e3), "solves/s", s.kernel_info(), flush=True)
        s.close()
