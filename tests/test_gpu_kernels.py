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
