"""Edge cases of the CUDA path against the oracle, bit for bit: shortest and longest horizons, iteration caps that
cut into the start-up and drain of the kernel's pipeline, plans without any contact, NaN inputs, warm starts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu



def assert_same(sol, ref, what=""):
    assert np.array_equal(sol.iters, ref["iters"]), f"{what}: iteration counters differ\n{sol.iters}\n{ref['iters']}"
    assert np.array_equal(sol.status, ref["status"]), f"{what}: status differs"
    for k in ("F", "X", "P", "L", "viol"):
        a, b = getattr(sol, k), ref[k]
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        assert same.all(), f"{what}: {k} differs in {np.count_nonzero(~same)} entries"


def _batch(n, B=5, seed=0, gait="trot"):
    from bunmpc_b200.motions import GAITS, ROBOTS
    from bunmpc_b200.plan_builder import build_batch
    rb, gp = ROBOTS["solo12"], GAITS["solo12"][gait]
    rng = np.random.default_rng(seed)
    com = np.array([0.0, 0.0, gp.nom_ht]) + rng.normal(0.0, 0.02, (B, 3))
    foot = np.broadcast_to(rb.foot_pos, (B, 4, 3)).copy()
    foot[:, :, :2] += rng.normal(0.0, 0.02, (B, 4, 2))
    v_des = np.zeros((B, 3)); v_des[:, 0] = rng.uniform(0.0, 0.3, B)
    return build_batch(rb, gp, com, rng.normal(0.0, 0.1, (B, 3)), rng.normal(0.0, 0.02, (B, 3)), foot,
                       rng.integers(0, 10, B) * gp.gait_dt, v_des, np.zeros(B), horizon=n)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 12, 16, 21, 28, 33, 41, 77])
def test_horizons_from_one_knot_to_the_largest(oracle, n):
    """Horizons that fall into every CTA size up to 384 threads (64, 96, 128, 160, 192, 256, 384); longer ones:
    tests/test_gpu_parity.py::test_horizons_of_the_reference_timing_sweep."""
    from bunmpc_b200.solver import BatchSolver
    from bunmpc_b200.problem import SolverParams
    b = _batch(n, B=4, seed=n)
    prm = SolverParams(max_outer=12 if n > 40 else 30)
    sol = BatchSolver(n, 4, max_batch=4).solve(b, prm)
    ref = oracle.solve(b, oracle.default_params(max_outer=prm.max_outer), n_threads=4)
    assert_same(sol, ref, f"n={n}")


@pytest.mark.parametrize("max_inner", [1, 2, 3, 4, 5, 7, 8, 9])
@pytest.mark.parametrize("n", [20, 24, 48])
def test_inner_iteration_caps_inside_the_pipeline_fill_and_drain(oracle, n, max_inner):
    """The pipeline resolves the decision of iteration j at slot j+3 and switches from checked to unrolled slots at
    slot 4: caps of 1..9 iterations end the inner solve in every one of those regimes."""
    from bunmpc_b200.solver import BatchSolver
    from bunmpc_b200.problem import SolverParams
    b = _batch(n, B=3, seed=100 + n, gait="trot" if n == 20 else "bound")
    prm = SolverParams(max_outer=6, max_inner=max_inner)
    sol = BatchSolver(n, 4, max_batch=3).solve(b, prm)
    ref = oracle.solve(b, oracle.default_params(max_outer=6, max_inner=max_inner), n_threads=3)
    assert (ref["iters"][:, 1] == 6 * max_inner).all()
    assert_same(sol, ref, f"n={n} max_inner={max_inner}")


def test_early_convergence_inside_the_first_slots(oracle):
    """Loose inner tolerances make FISTA stop after 1, 2, 3, ... iterations: exits in the checked start-up slots and
    in the first unrolled ones."""
    from bunmpc_b200.solver import BatchSolver
    from bunmpc_b200.problem import SolverParams
    b = _batch(20, B=6, seed=5)
    for tol in (30.0, 1.0, 0.3, 0.1, 0.03):
        prm = SolverParams(max_outer=10, tol=tol)
        sol = BatchSolver(20, 4, max_batch=6).solve(b, prm)
        ref = oracle.solve(b, oracle.default_params(max_outer=10, tol=tol), n_threads=6)
        assert ref["iters"][:, 1].min() < 10 * 15          # 1 to ~10 iterations per inner solve
        assert_same(sol, ref, f"tol={tol}")


def test_zero_outer_iterations_return_the_warm_start(oracle):
    from bunmpc_b200.solver import BatchSolver
    from bunmpc_b200.problem import SolverParams
    b = _batch(20, B=3, seed=8)
    rng = np.random.default_rng(1)
    b.X0, b.F0, b.P0 = rng.normal(size=(3, 189)), rng.normal(size=(3, 240)), rng.normal(size=(3, 189))
    sol = BatchSolver(20, 4, max_batch=3).solve(b, SolverParams(max_outer=0))
    assert np.array_equal(sol.X, b.X0) and np.array_equal(sol.F, b.F0) and np.array_equal(sol.P, b.P0)
    assert (sol.iters == 0).all() and (sol.status == 1).all()
    ref = oracle.solve(b, oracle.default_params(max_outer=0), n_threads=1)
    assert_same(sol, ref, "max_outer=0")


def test_warm_started_resolve_continues_bit_for_bit(oracle):
    """Second call warm-started with the first call's X, F, P and step sizes (what set_warm_start_vars plus the
    persistent FISTA objects do in the reference)."""
    from bunmpc_b200.solver import BatchSolver
    from bunmpc_b200.problem import SolverParams
    b = _batch(20, B=4, seed=12)
    s = BatchSolver(20, 4, max_batch=4)
    first = s.solve(b, SolverParams(max_outer=7))
    b.X0, b.F0, b.P0, b.L0 = first.X, first.F, first.P, first.L
    second = s.solve(b, SolverParams(max_outer=9))
    ref = oracle.solve(b, oracle.default_params(max_outer=9), n_threads=4)
    assert_same(second, ref, "warm start")


def test_plan_without_contacts_and_nan_inputs(oracle):
    from bunmpc_b200.solver import BatchSolver
    from bunmpc_b200.problem import SolverParams
    b = _batch(20, B=4, seed=3)
    cp = np.broadcast_to(b.cnt_plan, (4, 20, 4, 4)).copy()
    cp[0, :, :, 0] = 0.0                       # instance 0: flight phase only (no force can act, bounds stay +-inf)
    cp[1, 5:, :, 0] = 0.0                      # instance 1: loses all contacts after knot 5
    b.cnt_plan = cp
    x = np.broadcast_to(b.x_init, (4, 9)).copy()
    x[3, 4] = np.nan                           # instance 3: NaN state -> NaN everywhere, status 2 after one iteration
    b.x_init = x
    prm = SolverParams(max_outer=15)
    sol = BatchSolver(20, 4, max_batch=4).solve(b, prm)
    ref = oracle.solve(b, oracle.default_params(max_outer=15), n_threads=4)
    assert_same(sol, ref, "no contacts / NaN")
    assert sol.status[3] == 2 and sol.iters[3, 0] == 1 and np.isnan(sol.F[3]).any()
    assert (sol.F[0] == 0).all()


def test_step_size_overflows_to_infinity(oracle):
    """A diverging line search (Go2 mass, one inner iteration per solve) rejects ~1740 steps: L = L0 * 1.5^k overflows
    to +inf, after which gradient / L is exactly 0 in the reference.  Found by profiles/soak_parity.py."""
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    from bunmpc_b200.problem import SolverParams
    b = synthetic.perturbed(600, "go2", "jump", seed=756643367, vy_range=(-0.1, 0.1), w_range=(-0.1, 0.1)).select(np.arange(160, 176))
    b.L0 = np.array([[138.5017783022929, 341.3416382669748]])
    prm = SolverParams(max_outer=100, max_inner=1, tol=1e-3)
    sol = BatchSolver(b.n_col, 4, max_batch=16).solve(b, prm)
    ref = oracle.solve(b, oracle.default_params(max_outer=100, max_inner=1, tol=1e-3), n_threads=16)
    assert np.isinf(ref["L"]).any() and ref["iters"][:, 3].max() > 1700
    assert_same(sol, ref, "L overflow")


def test_randomised_soak_short():
    """Ten seconds of profiles/soak_parity.py: random gaits, robots, horizons 1..88, batch sizes, iteration caps,
    tolerances, beta/mu, initial step sizes, warm starts, arithmetic modes, slice lengths -- all bit-identical."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "profiles", "soak_parity.py"), "10", "123"], cwd=root,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 mismatching cases" in r.stdout
