"""GPU tests of kernel building blocks through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_division_identity():
    """div_fast (reciprocal hoisted out of the FISTA loop) == IEEE division, bit for bit, on 2^30 pairs."""
    from bunmpc_b200.solver import BatchSolver
    s = BatchSolver(20, 4, max_batch=1)
    for seed in (1, 2, 3, 4):
        assert s.selftest_division(1 << 28, seed) == 0


def test_fp64_peak_is_sane():
    from bunmpc_b200.solver import BatchSolver
    tf = BatchSolver(20, 4, max_batch=1).measure_fp64_peak()
    assert 20.0 < tf < 60.0, tf


def test_cycles_per_inner_iteration_budget():
    """Performance guard (a one-line change in the division helper once cost 35 % without failing any parity test, and
    a zero numerator in the state problem sent one lane through three real divisions in half of all iterations).
    With every CTA slot of the GPU busy (2 instances per SM at n = 20) an instance spends ~1600 SM cycles per FISTA
    iteration (in-kernel clock64), i.e. ~860 cycles per iteration and SM; alone on an SM it needs ~1240."""
    from bunmpc_b200 import SolverParams, synthetic
    from bunmpc_b200.solver import BatchSolver
    probe = BatchSolver(20, 4, max_batch=1)
    info = probe.kernel_info()
    slots = info["num_sms"] * info["ctas_per_sm"]
    prm = SolverParams(max_outer=6, slice_outer=-1)
    for B, budget in ((2 * slots, 2000), (info["num_sms"], 1550)):
        b = synthetic.config(1, B=B, seed=0)
        s = BatchSolver(b.n_col, b.n_eff, max_batch=B)
        s.solve(b, params=prm)
        sol = s.solve(b, params=prm)
        cpi = sol.cycles.sum() / (sol.iters[:, 1] + sol.iters[:, 2]).sum()
        assert cpi < budget, f"B={B}: {cpi:.0f} cycles per inner iteration and instance (budget {budget})"
    assert info["ctas_per_sm"] >= 2
