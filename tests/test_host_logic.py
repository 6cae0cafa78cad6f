"""CPU tests of the host-side mirror of the reference interface (no GPU, no solve)."""
import numpy as np
import pytest


def test_gait_planner_phases():
    from bunmpc_b200.gait_planner import GaitPlanner
    gp = GaitPlanner(0.5, np.array([0.6] * 4), np.array([0.0, 0.5, 0.5, 0.0]), 0.075)
    assert gp.get_phase(0.0, 0) == 1 and gp.get_phase(0.0, 1) == 1       # phi = 0 and 0.25 <= 0.3
    assert gp.get_phase(0.31, 0) == 0 and gp.get_phase(0.3, 0) == 1     # |phi - stance| < 1e-4 rule
    assert gp.get_phase(0.1, 1) == 0                                     # phi = 0.35 > 0.3
    t = np.array([0.0, 0.1, 0.31, 0.45])
    assert np.array_equal(gp.get_phase(t, 0), [1, 1, 0, 0])
    assert np.isclose(gp.get_percent_in_phase(0.15, 0), 0.5)
    assert np.isclose(gp.get_percent_in_phase(0.4, 0), 0.5)


def test_contact_plan_rules():
    from bunmpc_b200 import synthetic
    b = synthetic.nominal()
    assert b.n_col == 20 and b.n_eff == 4 and b.B == 1
    c = b.cnt_plan[0, :, :, 0]
    # trot: diagonal pairs alternate, every knot has at least two feet down
    assert np.array_equal(c[:, 0], c[:, 3]) and np.array_equal(c[:, 1], c[:, 2])
    assert (c.sum(1) >= 2).all()
    # stance feet keep their position (abstract_cyclic_gen.py:269-271)
    pos = b.cnt_plan[0, :, :, 1:4]
    for j in range(4):
        for i in range(1, 20):
            if c[i, j] == 1 and c[i - 1, j] == 1:
                assert np.array_equal(pos[i, j], pos[i - 1, j])
    # knot 0 uses the current foot position rounded to 3 dp (:213)
    assert np.allclose(pos[0, 0], [0.195, 0.147, 0.018])
    assert np.allclose(b.dt, 0.05)
    # dt[0] rule (:385-388): gait_dt - round(t mod gait_dt, 2), replaced by gait_dt when 0
    b2 = synthetic.nominal(t=0.02)
    assert np.isclose(b2.dt[0, 0], 0.03) and np.allclose(b2.dt[0, 1:], 0.05)


def test_costs_follow_reference_rules():
    from bunmpc_b200 import synthetic
    b = synthetic.nominal(v_des=(0.2, 0.0, 0.0))
    Xn = b.X_nom[0].reshape(20, 9)
    assert np.allclose(Xn[:, 0], 0.2 * 0.05 * np.arange(20))           # x advances by v_des*dt (:574-575)
    assert (Xn[:, 1] == 0).all() and np.allclose(Xn[:, 2], 0.2) and np.allclose(Xn[:, 3], 0.2)
    assert np.allclose(b.X_ter[0, :3], [2.0 * 0.5 * 0.2, 0.0, 0.2])    # :593
    assert np.allclose(b.bounds[0, 0], [-0.45, -0.45, 0, 0.45, 0.45, 0.45])
    assert b.W_X.shape == (1, 180) and b.W_F.shape == (1, 240)


def test_batch_container_select_and_shard():
    from bunmpc_b200 import synthetic
    b = synthetic.perturbed(10, seed=0)
    assert b.B == 10 and b.m.shape == (1,) and b.x_init.shape == (10, 9)
    s = b.shard(1, 4)
    assert s.B == 3 and np.array_equal(s.x_init, b.x_init[[1, 5, 9]])
    assert s.W_X.shape[0] == 1                                           # shared fields stay shared
    assert b.input_bytes() > 0
    with pytest.raises(ValueError):
        type(b)(b.n_col, b.n_eff, m=b.m, rho=b.rho, x_init=b.x_init, cnt_plan=b.cnt_plan[:5], dt=b.dt, W_X=b.W_X,
                W_X_ter=b.W_X_ter, X_nom=b.X_nom, X_ter=b.X_ter, W_F=b.W_F, bounds=b.bounds)


def test_biconvexmp_shell_matches_oracle_builders(oracle):
    """create_cost_X / create_cost_F / create_bound_constraints of the python class (host bookkeeping, as in
    the reference) against the oracle's restatement of biconvex.cpp:27-78."""
    from bunmpc_b200 import BiconvexMP, synthetic
    b = synthetic.perturbed(1, seed=4)
    mp = BiconvexMP(2.5, b.n_col, b.n_eff)
    assert (mp.lb_x == 0).all() and (mp.ub_x == 0).all() and mp.rho_ == 1e5 and mp.L_f == 506.25
    for i in range(b.n_col):
        mp.set_contact_plan(b.cnt_plan[0, i], b.dt[0, i])
    with pytest.raises(IndexError):
        mp.set_contact_plan(b.cnt_plan[0, 0], 0.05)
    mp.create_bound_constraints(b.bounds[0], 15.0, 15.0, 15.0)
    mp.create_cost_X(b.W_X[0], b.W_X_ter[0], b.X_ter[0], b.X_nom[0])
    mp.create_cost_F(b.W_F[0])
    ex = oracle.expand(b)
    assert np.array_equal(mp.Q_x, ex["Qx"][0]) and np.array_equal(mp.q_x, ex["qx"][0])
    assert np.array_equal(mp.Q_f, ex["Qf"][0]) and np.array_equal(mp.q_f, ex["qf"][0])
    assert np.array_equal(mp.lb_x, ex["lbx"][0]) and np.array_equal(mp.ub_x, ex["ubx"][0])
    # set_cost_x accepts scipy.sparse like the pybind signature and rejects non-diagonal matrices
    import scipy.sparse as sp
    mp.set_cost_x(sp.diags(ex["Qx"][0]).tocsc(), ex["qx"][0])
    assert np.array_equal(mp.Q_x, ex["Qx"][0])
    bad = sp.lil_matrix((mp.nx, mp.nx)); bad[0, 1] = 1.0
    with pytest.raises(NotImplementedError):
        mp.set_cost_x(bad.tocsc(), ex["qx"][0])
    assert mp.return_opt_com().shape == (b.n_col + 1, 3) and mp.return_opt_mom().shape == (b.n_col + 1, 6)


def test_aliases_and_api_surface():
    import bunmpc_b200 as pkg
    assert pkg.BiConvexMP is pkg.BiconvexMP
    methods = ["set_contact_plan", "set_rotation_matrix_f", "return_A_x", "return_b_x", "return_A_f", "return_b_f",
               "set_cost_x", "create_cost_X", "set_cost_f", "create_cost_F", "set_bounds_x", "set_bounds_f",
               "create_bound_constraints", "set_rho", "return_opt_x", "return_opt_f", "return_opt_p",
               "return_opt_com", "return_opt_mom", "set_warm_start_vars", "optimize", "return_dyn_viol_hist",
               "collect_statistics"]          # srcpy/motion_planner/biconvex.cpp:21-44
    for m in methods:
        assert callable(getattr(pkg.BiconvexMP, m)), m


def test_posterior_update_helpers():
    from bunmpc_b200 import dist
    axes = (np.linspace(0, 0.3, 10), np.linspace(-0.1, 0.1, 10), np.linspace(-0.1, 0.1, 10))
    prior = np.full((10, 10, 10), 1e-3)
    like = dist.gaussian_likelihood_grid(axes, (0.2, 0.0, 0.0), sigma=0.1)
    post = dist.posterior_update(prior, like)
    assert np.isclose(post.sum(), 1.0) and np.unravel_index(post.argmax(), post.shape)[0] in (6, 7)
    st = dist.goal_sufficient_stats(np.ones((5, 3)), np.arange(5.0))
    assert st.shape == (17,) and st[0] == 5 and st[13] == 10.0


def test_gait_gen_shim_sizes_and_interpolation():
    """update_gait_params sizes (abstract_cyclic_gen.py:125-153) and the 1 kHz interpolation (:677-692)."""
    import bunmpc_b200 as pkg
    from bunmpc_b200.motions import solo12_trot
    assert pkg.SoloMpcGaitGen is pkg.CyclicQuadrupedGaitGen and issubclass(pkg.AbstractGaitGen, pkg.CyclicQuadrupedGaitGen)
    gg = pkg.CyclicQuadrupedGaitGen(None, None, None, planning_time=0.05)
    gg.update_gait_params(solo12_trot, 0.0)
    assert gg.horizon == 20 and gg.ik_horizon == 10 and gg.size == 3      # min(10, int(0.05/0.05)+2) = 3
    knots = np.arange(12.0).reshape(4, 3)
    out = gg.interpolate(knots, np.array([0.03, 0.05, 0.05, 0.05]), 3)
    assert out.shape == (30 + 50 + 50, 3)
    assert np.array_equal(out[0], knots[0]) and np.array_equal(out[29], knots[1]) and np.array_equal(out[30], knots[1])
    with pytest.raises(ImportError):
        gg.optimize(np.zeros(19), np.zeros(18), 0.0, np.array([0.2, 0, 0]), 0.0)


def test_bench_reference_arm_line():
    """`bench.py --impl reference`: the CPU arm of the bench contract (the oracle on the host cores, a bounded sample of
    the same workload) prints one JSON line with the keys the driver reads; no GPU involved."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "centroidal_mpc_solves_per_sec" and d["unit"] == "solves/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert "BASELINE config[1]" in d["config"]["workload"] and d["config"]["n_col"] == 20
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
