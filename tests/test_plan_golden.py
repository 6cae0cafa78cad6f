"""The batched problem builder against the reference's OWN python: tests/golden/plan_cases.npz holds what
examples/mpc/abstract_cyclic_gen.py (create_cnt_plan :159-414, create_costs :532-614, imported in place, pinocchio and
the pybind modules replaced by oracle/pinshim, phase lookup by the reference's own gait_planner.cpp) handed to its
solver on 36 cases: three gaits, both robots' masses, times on and off the planning grid and on phase edges, yaw over the
full circle, roll/pitch, turning (w_des != 0), overridden horizons.  Bit for bit."""
import dataclasses

import numpy as np
import pytest

from tests.golden.make_plan_golden import load

CASES, MOTIONS = load()


def _robot_and_params(d):
    from bunmpc_b200 import motions
    prm = motions.GAITS["solo12"][d["gait"]]
    robot = dataclasses.replace(motions.SOLO12, mass=float(d["mass"]), hip_offsets=d["hip_offsets"], I_zz=float(d["I_zz"]))
    return robot, prm


def test_motion_records_equal_the_reference_files():
    """bunmpc_b200/motions.py was transcribed from examples/motions/cyclic/solo12_{trot,bound,jump}.py; the golden file
    carries the records those files define when imported."""
    from bunmpc_b200 import motions
    for gait, rec in MOTIONS.items():
        prm = motions.GAITS["solo12"][gait]
        for k, v in rec.items():
            assert np.array_equal(np.asarray(getattr(prm, k), dtype=np.float64), v), (gait, k)
        assert prm.horizon() == int(np.round(rec["gait_horizon"] * rec["gait_period"] / rec["gait_dt"], 2))


@pytest.mark.parametrize("name,d", CASES, ids=[c[0] for c in CASES])
def test_host_builder_equals_reference_python(name, d):
    from bunmpc_b200 import plan_builder
    robot, prm = _robot_and_params(d)
    b = plan_builder.build_batch(robot, prm, d["com"][None], d["vcom"][None], d["amom"][None], d["foot_pos"][None],
                                 d["t"][None], d["v_des"][None], d["w_des"][None], yaw=d["yaw"][None],
                                 amom_des=d["amom_des"][None], horizon=int(d["horizon"]), hip_xy=d["hip_xy"][None])
    n = int(d["horizon"])
    assert b.n_col == n
    assert np.array_equal(b.cnt_plan[0], d["cnt_plan"]), "cnt_plan"
    assert np.array_equal(b.dt[0], d["dt"]), "dt"
    assert np.array_equal(b.x_init[0], d["x_init"]), "x_init"
    assert np.array_equal(b.X_nom[0], d["X_nom"]), "X_nom"
    assert np.array_equal(b.X_ter[0], d["X_ter"]), "X_ter"
    assert np.array_equal(np.broadcast_to(b.W_X, (1, 9 * n))[0], d["W_X"])
    assert np.array_equal(np.broadcast_to(b.W_X_ter, (1, 9))[0], d["W_X_ter"])
    assert np.array_equal(np.broadcast_to(b.W_F, (1, 12 * n))[0], d["W_F"])
    assert np.array_equal(np.broadcast_to(b.bounds, (1, n, 6))[0], d["bounds"])
    assert float(np.ravel(b.rho)[0]) == float(d["rho"])


@pytest.mark.gpu
def test_device_builder_equals_reference_python():
    """build_problem_kernel (one launch per group of cases with the same gait and horizon) == the reference's python."""
    from bunmpc_b200.solver import BatchSolver
    groups = {}
    for name, d in CASES:
        groups.setdefault((d["gait"], int(d["horizon"]), float(d["mass"])), []).append(d)
    checked = 0
    for (gait, n, mass), ds in groups.items():
        s = BatchSolver(n, 4, max_batch=len(ds))
        for d in ds:     # hip offsets / yaw inertia differ per case (they come from the injected robot): one launch each
            robot, prm = _robot_and_params(d)
            if prm.horizon() != n:
                prm = dataclasses.replace(prm, gait_horizon=n * prm.gait_dt / prm.gait_period)
                assert prm.horizon() == n
            dev = s.build_device(robot, prm, d["com"][None], d["vcom"][None], d["amom"][None], d["foot_pos"][None],
                                 d["t"][None], d["v_des"][None], d["w_des"][None], yaw=d["yaw"][None],
                                 amom_des=d["amom_des"][None], hip_xy=d["hip_xy"][None])
            f = {k: v.cpu().numpy() for k, v in dev.fields.items() if v is not None}
            assert np.array_equal(f["cnt_plan"][0], d["cnt_plan"].reshape(-1)), (gait, "cnt_plan")
            assert np.array_equal(f["dt"][0], d["dt"]) and np.array_equal(f["x_init"][0], d["x_init"])
            if prm.horizon() == motions_horizon(gait):      # X_ter uses gait_horizon itself (:593)
                assert np.array_equal(f["X_ter"][0], d["X_ter"]), (gait, "X_ter")
            assert np.array_equal(f["X_nom"][0], d["X_nom"]), (gait, "X_nom")
            checked += 1
    assert checked == len(CASES)


def motions_horizon(gait):
    from bunmpc_b200 import motions
    return motions.GAITS["solo12"][gait].horizon()


# ---- the reference's other cyclic generator: AbstractGaitGen, examples/mpc/abstract_cyclic_gen1.py ----
def _gen1_robot(d):
    from bunmpc_b200 import motions
    g1 = d["gen1"]
    return dataclasses.replace(motions.SOLO12, mass=float(d["mass"]), hip_offsets=g1["hip_offsets"], I_zz=float(d["I_zz"]))


@pytest.mark.parametrize("name,d", CASES, ids=[c[0] for c in CASES])
def test_host_builder_equals_abstract_gait_gen(name, d):
    """Same states through abstract_cyclic_gen1.py: its own hip offsets (round(foot - com, 3)), its swing rule (no Raibert
    step in the second half of a swing, :211-215); costs and bounds are those of SoloMpcGaitGen."""
    from bunmpc_b200 import motions, plan_builder
    g1 = d["gen1"]
    prm = motions.GAITS["solo12"][d["gait"]]
    n = int(g1["horizon"])
    b = plan_builder.build_batch(_gen1_robot(d), prm, d["com"][None], d["vcom"][None], d["amom"][None], d["foot_pos"][None],
                                 d["t"][None], d["v_des"][None], d["w_des"][None], yaw=d["yaw"][None],
                                 amom_des=d["amom_des"][None], hip_xy=g1["hip_xy"][None],
                                 swing_rule=plan_builder.SWING_RULE_ABSTRACT)
    assert b.n_col == n
    for k in ("cnt_plan", "dt", "x_init", "X_nom", "X_ter"):
        assert np.array_equal(getattr(b, k)[0], g1[k]), k
    assert np.array_equal(np.broadcast_to(b.W_X, (1, 9 * n))[0], g1["W_X"])
    assert np.array_equal(np.broadcast_to(b.W_F, (1, 12 * n))[0], g1["W_F"])
    assert np.array_equal(np.broadcast_to(b.bounds, (1, n, 6))[0], g1["bounds"])
    # the rule matters: with SoloMpcGaitGen's rule the plan differs whenever a foot is late in its swing inside the horizon
    b0 = plan_builder.build_batch(_gen1_robot(d), prm, d["com"][None], d["vcom"][None], d["amom"][None], d["foot_pos"][None],
                                  d["t"][None], d["v_des"][None], d["w_des"][None], yaw=d["yaw"][None],
                                  amom_des=d["amom_des"][None], hip_xy=g1["hip_xy"][None])
    assert not np.array_equal(b0.cnt_plan[0], g1["cnt_plan"])


def test_abstract_gait_gen_class_uses_its_rule():
    from bunmpc_b200 import AbstractGaitGen, SoloMpcGaitGen, plan_builder
    assert AbstractGaitGen.swing_rule == plan_builder.SWING_RULE_ABSTRACT
    assert SoloMpcGaitGen.swing_rule == plan_builder.SWING_RULE_SOLO_MPC
    assert issubclass(AbstractGaitGen, SoloMpcGaitGen)


@pytest.mark.gpu
def test_device_builder_equals_abstract_gait_gen():
    """build_problem_kernel with bunmpc_gait.swing_rule = 1 == abstract_cyclic_gen1.py, bit for bit."""
    from bunmpc_b200 import motions, plan_builder
    from bunmpc_b200.solver import BatchSolver
    solvers = {}
    for name, d in CASES:
        g1 = d["gen1"]
        prm = motions.GAITS["solo12"][d["gait"]]
        n = int(g1["horizon"])
        s = solvers.setdefault(n, BatchSolver(n, 4, max_batch=1))
        dev = s.build_device(_gen1_robot(d), prm, d["com"][None], d["vcom"][None], d["amom"][None], d["foot_pos"][None],
                             d["t"][None], d["v_des"][None], d["w_des"][None], yaw=d["yaw"][None],
                             amom_des=d["amom_des"][None], hip_xy=g1["hip_xy"][None],
                             swing_rule=plan_builder.SWING_RULE_ABSTRACT)
        f = {k: v.cpu().numpy() for k, v in dev.fields.items() if v is not None}
        assert np.array_equal(f["cnt_plan"][0], g1["cnt_plan"].reshape(-1)), (name, "cnt_plan")
        for k in ("dt", "x_init", "X_nom", "X_ter"):
            assert np.array_equal(f[k][0], g1[k]), (name, k)


# ---- the generator's own entry point: SoloMpcGaitGen.optimize(q, v, t, v_des, w_des), abstract_cyclic_gen.py:629-698 ----
def _fake_robot(d):
    """The robot wrapper of oracle/pinshim (test infrastructure) with the case's kinematics injected: what pinocchio
    would return for (q, v) is whatever the golden generator injected when the reference's python ran."""
    import os
    import sys
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pinshim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import pinocchio as pin
    robot = pin.FakeRobot(float(d["mass"]), nv=18)
    robot.inject(com=d["com"], foot_pos=d["foot_pos"], hg=d["hg"])
    return robot


def _shim_gen(d):
    from bunmpc_b200 import SoloMpcGaitGen
    rc, prm = _robot_and_params(d)
    gen = SoloMpcGaitGen(_fake_robot(d), "solo12.urdf", np.zeros(37), 0.05, None, robot_constants=rc)
    gen.update_gait_params(prm, float(d["t"]), horizon=int(d["horizon"]))
    return gen


@pytest.mark.parametrize("name,d", CASES, ids=[c[0] for c in CASES])
def test_optimize_q_v_front_end_equals_reference_python(name, d):
    """(q, v) -> builder inputs (gait_gen.centroidal_inputs: origin reset, body-frame velocity goal, yaw, yaw-rotated hip
    offsets, orientation-correction momentum) == what the reference's optimize / create_cnt_plan / create_costs derived
    from the same (q, v) when the golden file was made.  Bit for bit."""
    gen = _shim_gen(d)
    q = d["q"].copy()
    q[0:2] = [0.3, -0.2]                      # optimize resets the origin itself (:633)
    c = gen.centroidal_inputs(q, d["qv"].copy(), d["v_in"].copy(), float(d["w_des"]))
    assert q[0] == 0 and q[1] == 0
    for k in ("com", "vcom", "amom", "foot_pos", "v_des", "yaw", "amom_des", "hip_xy"):
        assert np.array_equal(np.asarray(c[k]), d[k]), k


@pytest.mark.gpu
def test_optimize_q_v_equals_reference_plan_and_oracle_solution(oracle):
    """SoloMpcGaitGen.optimize(q, v, ...) end to end on the GPU: the problem it hands to the solver is the reference
    python's, bit for bit; the solution is the oracle's for that problem; f_int / com_int / mom_int are the 1 kHz
    interpolation of :677-692."""
    for name, d in CASES[:12]:
        gen = _shim_gen(d)
        xs, us, f_int = gen.optimize(d["q"].copy(), d["qv"].copy(), float(d["t"]), d["v_in"].copy(), float(d["w_des"]))
        assert xs is None and us is None                     # the IK is out of scope
        batch, sol = gen.last
        n = int(d["horizon"])
        assert batch.n_col == n
        for k in ("cnt_plan", "dt", "x_init", "X_nom", "X_ter"):
            assert np.array_equal(getattr(batch, k)[0], d[k]), (name, k)
        assert np.array_equal(np.broadcast_to(batch.bounds, (1, n, 6))[0], d["bounds"])
        ref = oracle.solve(batch)
        for k in ("X", "F", "iters", "status"):
            assert np.array_equal(getattr(sol, k), ref[k].reshape(getattr(sol, k).shape), equal_nan=True), (name, k)
        F = sol.F[0].reshape(n, 12)
        dt = batch.dt[0]
        want = np.vstack([np.linspace(F[i], F[i + 1], int(dt[i] / 0.001)) for i in range(gen.size)])
        assert np.array_equal(f_int, want, equal_nan=True) and np.array_equal(gen.f_int, want, equal_nan=True), name   # (Go2 mass: NaN)
        assert gen.com_int.shape == (want.shape[0], 3) and gen.mom_int.shape == (want.shape[0], 6)
