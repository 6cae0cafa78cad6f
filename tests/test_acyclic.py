"""The acyclic caller of the path (`SoloAcyclicGen`, examples/mpc/abstract_acyclic_gen.py) against the reference's OWN
python: tests/golden/acyclic_cases.npz holds what `create_contact_plan` (:74-124) and `create_costs` (:126-190),
imported in place with the stand-ins of oracle/pinshim, handed to the solver on 62 cases -- the five motions of
examples/motions/acyclic that the reference's generator accepts, replans at the start, on the planning grid, on every
segment edge, off the grid, shifted motion starts (t0) and past the end of the motion.  Bit for bit; the solves then go
through the CUDA path with the reference's 50 outer iterations and are compared with the oracle."""
import numpy as np
import pytest

from tests.golden.make_acyclic_golden import load

CASES, MOTIONS = load()


def test_motion_records_equal_the_reference_files():
    from bunmpc_b200.acyclic import ACYCLIC_MOTIONS
    assert set(MOTIONS) == set(ACYCLIC_MOTIONS)
    for name, rec in MOTIONS.items():
        prm = ACYCLIC_MOTIONS[name]
        for k, v in rec.items():
            assert np.array_equal(np.asarray(getattr(prm, k), dtype=np.float64), v), (name, k)


@pytest.mark.parametrize("name,d", CASES, ids=[c[0] for c in CASES])
def test_host_builder_equals_reference_python(name, d):
    from bunmpc_b200.acyclic import ACYCLIC_F_MAX, ACYCLIC_MOTIONS, SoloAcyclicGen, build_batch
    prm = ACYCLIC_MOTIONS[d["motion"]]
    b = build_batch(prm, d["x_init"][None], d["t"][None], float(d["t0"]))
    n = prm.n_col
    assert b.n_col == n and float(b.m[0]) == float(d["mass"])
    for k in ("cnt_plan", "dt", "X_nom", "X_ter", "bounds"):
        assert np.array_equal(getattr(b, k)[0], d[k]), k
    assert np.array_equal(np.broadcast_to(b.W_X, (1, 9 * n))[0], d["W_X"])
    assert np.array_equal(np.broadcast_to(b.W_X_ter, (1, 9))[0], d["W_X_ter"])
    assert np.array_equal(np.broadcast_to(b.W_F, (1, 12 * n))[0], d["W_F"])
    assert float(np.ravel(b.rho)[0]) == float(d["rho"])
    assert np.array_equal(d["f_max"], 3 * [ACYCLIC_F_MAX])
    g = SoloAcyclicGen()
    g.update_motion_params(prm, None, float(d["t0"]))
    assert g.get_plan_freq(float(d["t"])) == float(d["plan_freq"])


def test_batched_builder_equals_case_by_case():
    """All cases of one motion in ONE build_batch call (per-instance t, shared t0) == the single-case builds."""
    from bunmpc_b200.acyclic import ACYCLIC_MOTIONS, build_batch
    for motion in MOTIONS:
        for t0 in (0.0, 0.1):
            ds = [d for _, d in CASES if d["motion"] == motion and float(d["t0"]) == t0]
            b = build_batch(ACYCLIC_MOTIONS[motion], np.stack([d["x_init"] for d in ds]), np.array([float(d["t"]) for d in ds]), t0)
            assert b.B == len(ds)
            for k in ("cnt_plan", "dt", "X_nom", "X_ter", "bounds"):
                assert np.array_equal(getattr(b, k), np.stack([d[k] for d in ds])), (motion, k)


def _perturbed(motion, B, seed):
    from bunmpc_b200 import synthetic
    return synthetic.acyclic_replans(B, motion, seed)


@pytest.mark.gpu
@pytest.mark.parametrize("motion", sorted(MOTIONS))
def test_acyclic_solves_gpu_equals_oracle(motion):
    """Perturbed replans of every motion over its whole duration (flight phases, three-legged and two-legged stances,
    +-inf boxes of rearing_jump), 50 outer iterations as abstract_acyclic_gen.py:319: CUDA == oracle, bit for bit."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from bunmpc_b200.acyclic import ACYCLIC_MAX_OUTER
    from bunmpc_b200.problem import SolverParams
    from bunmpc_b200.solver import BatchSolver
    from oracle import oracle
    b = _perturbed(motion, 24, seed=11)
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=24).solve(b, SolverParams(max_outer=ACYCLIC_MAX_OUTER))
    ref = oracle.solve(b, oracle.default_params(max_outer=ACYCLIC_MAX_OUTER), n_threads=16)
    assert np.array_equal(sol.iters, ref["iters"]) and np.array_equal(sol.status, ref["status"])
    for k in ("X", "F", "P", "L", "viol"):
        assert np.array_equal(getattr(sol, k), ref[k], equal_nan=True), k
    assert sol.iters[:, 0].max() <= ACYCLIC_MAX_OUTER


@pytest.mark.gpu
def test_acyclic_generator_replans_carry_step_sizes():
    """SoloAcyclicGen.optimize_centroidal: two replans of the jump; the second starts from the step sizes the first left
    (quirk Q3: the FISTA objects live as long as the KinoDynMP made in update_motion_params), f_int holds every knot's
    force for int(dt / 0.001) samples."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from bunmpc_b200.acyclic import ACYCLIC_MAX_OUTER, ACYCLIC_MOTIONS, SoloAcyclicGen, build_batch
    from oracle import oracle
    prm = ACYCLIC_MOTIONS["jump_fwd"]
    g = SoloAcyclicGen()
    g.update_motion_params(prm, None, 0.0)
    x0 = np.array([0.2, 0.0, 0.22, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
    L = None
    for t in (0.0, 0.3):
        f_int = g.optimize_centroidal(x0, t)
        b = build_batch(prm, x0[None], np.array([t]), 0.0, L0=L)
        ref = oracle.solve(b, oracle.default_params(max_outer=ACYCLIC_MAX_OUTER), n_threads=1)
        assert np.array_equal(g.last[1].F, ref["F"]) and np.array_equal(g.last[1].L, ref["L"])
        L = ref["L"]
        assert f_int.shape == (prm.n_col * int(prm.dt_arr[0] / 0.001), 12)
        assert np.array_equal(f_int[0], ref["F"][0, 0:12]) and np.array_equal(f_int[-1], ref["F"][0, -12:])


@pytest.mark.gpu
def test_device_builder_equals_reference_python():
    """build_acyclic_kernel (bunmpc_build_acyclic_device): all golden cases of a motion and motion start in one launch ==
    the reference's own python, bit for bit."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from bunmpc_b200.acyclic import ACYCLIC_MOTIONS
    from bunmpc_b200.solver import BatchSolver
    checked = 0
    for motion in sorted(MOTIONS):
        prm = ACYCLIC_MOTIONS[motion]
        s = BatchSolver(prm.n_col, 4, max_batch=32)
        for t0 in (0.0, 0.1):
            ds = [d for _, d in CASES if d["motion"] == motion and float(d["t0"]) == t0]
            dev = s.build_acyclic_device(prm, np.stack([d["x_init"] for d in ds]), np.array([float(d["t"]) for d in ds]), t0)
            f = {k: v.cpu().numpy() for k, v in dev.fields.items() if v is not None}
            for k in ("cnt_plan", "dt", "X_nom", "X_ter", "bounds"):
                want = np.stack([d[k].reshape(-1) for d in ds])
                assert np.array_equal(f[k], want), (motion, t0, k)
            checked += len(ds)
        s.close()
    assert checked == len(CASES)


@pytest.mark.gpu
def test_acyclic_device_path_equals_host_path():
    """SoloAcyclicGen.optimize_centroidal_batch(builder="device") == builder="host" == oracle (rearing_jump: +-inf boxes)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from bunmpc_b200.acyclic import ACYCLIC_MAX_OUTER, ACYCLIC_MOTIONS, SoloAcyclicGen, build_batch
    from oracle import oracle
    b = _perturbed("rearing_jump", 40, seed=5)
    prm = ACYCLIC_MOTIONS["rearing_jump"]
    rng = np.random.default_rng(5)
    t = np.round(rng.uniform(0, 1.4, 40) / 0.05) * 0.05
    g = SoloAcyclicGen()
    g.update_motion_params(prm, None, 0.0)
    host = g.optimize_centroidal_batch(b.x_init, t)
    devs = g.optimize_centroidal_batch(b.x_init, t, builder="device")
    ref = oracle.solve(build_batch(prm, b.x_init, t, 0.0), oracle.default_params(max_outer=ACYCLIC_MAX_OUTER), n_threads=16)
    for k in ("X", "F", "P", "L", "viol", "iters", "status"):
        assert np.array_equal(getattr(host, k), ref[k], equal_nan=True), k
        assert np.array_equal(getattr(devs, k), ref[k], equal_nan=True), k
