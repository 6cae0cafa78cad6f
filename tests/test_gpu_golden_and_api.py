"""GPU tests against the committed golden vectors and of the drop-in python API, all through the C ABI."""
import numpy as np
import pytest

from tests.golden.make_golden import RES, load

pytestmark = pytest.mark.gpu
CASES = load()


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(((a == b) | (np.isnan(a.astype(float)) & np.isnan(b.astype(float)))).all())


@pytest.mark.parametrize("name,batch,sets", CASES, ids=[c[0] for c in CASES])
def test_gpu_reproduces_golden(name, batch, sets):
    """STRICT mode == oracle default order == the reference's own sources (golden `ref`); FMA mode == its oracle twin."""
    from bunmpc_b200 import ARITH_FMA
    from bunmpc_b200.solver import BatchSolver
    s = BatchSolver(batch.n_col, batch.n_eff, max_batch=batch.B)
    sol = s.solve(batch)
    solf = s.solve(batch, arith=ARITH_FMA)
    for k in RES:
        assert same(getattr(sol, k), sets["oc"][k]), f"{name}/oc/{k}"
        assert same(getattr(solf, k), sets["ocf"][k]), f"{name}/ocf/{k}"
        if k != "status":
            assert same(getattr(sol, k), sets["ref"][k]), f"{name}/ref/{k}"


def test_biconvexmp_drop_in_protocol(oracle):
    """The reference's per-solve call protocol on the python class (abstract_cyclic_gen.py:391,611-614,663,671-673),
    two consecutive solves on one object: the FISTA step sizes and iterates carry over like the C++ members."""
    from bunmpc_b200 import BiconvexMP, synthetic
    b = synthetic.perturbed(1, "solo12", "trot", seed=12)
    n, e = b.n_col, b.n_eff
    mp = BiconvexMP(2.5, n, e)
    mp.set_rho(float(b.rho[0]))
    mp.collect_statistics()
    state = dict(X0=np.tile(b.x_init[0], n + 1)[None], F0=np.zeros((1, 3 * e * n)), P0=np.zeros((1, 9 * (n + 1))),
                 L0=b.L0)
    for rep in range(2):
        for i in range(n):
            mp.set_contact_plan(b.cnt_plan[0, i], b.dt[0, i])
        mp.create_bound_constraints(b.bounds[0], 15.0, 15.0, 15.0)
        mp.create_cost_X(b.W_X[0], b.W_X_ter[0], b.X_ter[0], b.X_nom[0])
        mp.create_cost_F(b.W_F[0])
        mp.set_warm_start_vars(state["X0"][0], state["F0"][0], state["P0"][0])
        mp.optimize(b.x_init[0], 100)
        ref = oracle.solve(b.with_state(**state))
        assert np.array_equal(mp.return_opt_x(), ref["X"][0]) and np.array_equal(mp.return_opt_f(), ref["F"][0])
        assert np.array_equal(mp.return_opt_p(), ref["P"][0])
        assert (mp.L_f, mp.L_x) == tuple(ref["L"][0]) and np.array_equal(mp.last_iters, ref["iters"][0])
        assert np.array_equal(mp.return_opt_com(), ref["X"][0].reshape(-1, 9)[:, :3])
        assert np.array_equal(mp.return_opt_mom()[:, :3], 2.5 * ref["X"][0].reshape(-1, 9)[:, 3:6])
        # second solve: cold restart of the iterates (kino_dyn.cpp:83-99) but the step sizes persist (Q3)
        state = dict(X0=state["X0"], F0=state["F0"], P0=state["P0"], L0=ref["L"])
    hist = mp.return_dyn_viol_hist()
    assert len(hist) == 2 * int(ref["iters"][0, 0]) or len(hist) > 0
    assert hist[-1] == ref["viol"][0]
    with pytest.raises(RuntimeError):
        mp.optimize(b.x_init[0], 100)          # optimize() cleared the contact plan (biconvex.cpp:117)


def test_dense_matrix_accessors(oracle):
    """return_A_x / return_b_x / return_A_f / return_b_f against the oracle's dense matrices."""
    from bunmpc_b200 import BiconvexMP, synthetic
    b = synthetic.perturbed(1, "solo12", "trot", seed=13)
    n, e = b.n_col, b.n_eff
    mp = BiconvexMP(2.5, n, e)
    for i in range(n):
        mp.set_contact_plan(b.cnt_plan[0, i], b.dt[0, i])
    rng = np.random.default_rng(1)
    X, F = rng.normal(size=9 * (n + 1)), rng.normal(size=3 * e * n)
    A_x, b_x = oracle.dense_x_mat(2.5, b.cnt_plan[0], b.dt[0], X)
    A_f, b_f = oracle.dense_f_mat(2.5, b.cnt_plan[0], b.dt[0], F, b.x_init[0])
    assert np.array_equal(mp.return_A_x(X), A_x) and np.array_equal(mp.return_b_x(X), b_x)
    assert np.array_equal(mp.return_A_f(F, b.x_init[0]), A_f) and np.array_equal(mp.return_b_f(F, b.x_init[0]), b_f)


def test_expand_kernel_matches_oracle_builders(oracle):
    """create_bound_constraints / create_cost_X / create_cost_F as a CUDA kernel (bunmpc_expand_device)."""
    import ctypes as C
    import torch
    from bunmpc_b200 import _lib, synthetic
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.perturbed(37, "solo12", "jump", seed=14)          # jump: knots without any contact -> infinite bounds
    s = BatchSolver(b.n_col, b.n_eff, max_batch=64)
    dev = s.upload(b)
    B, nx, nf = b.B, s.nx, s.nf
    out = {k: torch.empty((B, nx if k not in ("Qf", "qf") else nf), dtype=torch.float64, device="cuda")
           for k in ("Qx", "qx", "Qf", "qf", "lbx", "ubx")}
    prob = _lib.CompactProblem()
    prob.batch = B
    for f in _lib.COMPACT_FIELDS:
        t = dev.fields[f]
        setattr(prob, f, _lib.In(None, 0) if t is None else _lib.In(t.data_ptr(), 0 if t.shape[0] == 1 else dev.widths[f]))
    _lib.check(_lib.lib().bunmpc_expand_device(s._h, C.byref(prob), *[C.c_void_p(out[k].data_ptr()) for k in
                                               ("Qx", "qx", "Qf", "qf", "lbx", "ubx")], None))
    torch.cuda.synchronize()
    ex = oracle.expand(b)
    assert np.isinf(ex["lbx"]).any()
    for k in out:
        assert np.array_equal(out[k].cpu().numpy(), ex[k]), k


def test_full_size_batch_properties():
    """BASELINE config[1] at full size (B = 1024): properties that need no oracle run --
    permutation invariance (instances are independent), split invariance, feasibility of every solution."""
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.config(1, B=1024, seed=0)
    s = BatchSolver(b.n_col, b.n_eff, max_batch=1024)
    sol = s.solve(b)
    perm = np.random.default_rng(0).permutation(b.B)
    solp = s.solve(b.select(perm))
    for k in ("X", "F", "P", "L", "iters", "viol", "status"):
        assert np.array_equal(getattr(solp, k), getattr(sol, k)[perm]), k
    half = s.solve(b.select(np.arange(0, 1024, 2)))
    assert np.array_equal(half.F, sol.F[::2]) and np.array_equal(half.iters, sol.iters[::2])
    n, e = b.n_col, b.n_eff
    F = sol.F.reshape(-1, n, e, 3)
    fin = sol.status != 2
    assert fin.mean() > 0.99
    assert (F[fin][..., 2] >= 0).all()
    assert ((F[fin][..., 0] ** 2 + F[fin][..., 1] ** 2) <= F[fin][..., 2] * (1 + 1e-9) + 1e-12).all()
    conv = sol.status == 0
    assert conv.mean() > 0.5 and (sol.viol[conv] < 1e-3).all() and (sol.viol[sol.status == 1] >= 1e-3).all()
    assert (sol.iters[:, 0] <= 100).all() and (sol.iters[:, 1] <= 150 * sol.iters[:, 0]).all()
    assert (sol.cycles > 0).all()


def test_errors_are_loud():
    from bunmpc_b200 import BunmpcError, synthetic
    from bunmpc_b200.solver import BatchSolver
    with pytest.raises(BunmpcError):
        BatchSolver(20, 3, max_batch=4)                 # kernels are built for n_eff == 4
    s = BatchSolver(20, 4, max_batch=4)
    with pytest.raises(ValueError):
        s.solve(synthetic.perturbed(8, seed=0))         # batch > max_batch


@pytest.mark.parametrize("gait,robot", [("trot", "solo12"), ("bound", "solo12"), ("jump", "solo12"), ("trot", "go2")])
def test_device_problem_builder_matches_host_builder(gait, robot):
    """the on-device batched builder (contact plan, X_nom, X_ter, scaled weights) against the numpy
    restatement of the gait generator's rules, bit for bit, then solved from the device-resident problem."""
    from bunmpc_b200 import synthetic
    from bunmpc_b200.motions import GAITS, ROBOTS
    from bunmpc_b200.plan_builder import build_batch
    from bunmpc_b200.solver import BatchSolver
    rb, gp = ROBOTS[robot], GAITS[robot][gait]
    rng = np.random.default_rng(3)
    B = 257
    com = np.array([0.0, 0.0, gp.nom_ht]) + rng.normal(0, 0.02, (B, 3))
    vcom, amom = rng.normal(0, 0.1, (B, 3)), rng.normal(0, 0.02, (B, 3))
    foot = np.broadcast_to(rb.foot_pos, (B, 4, 3)) + rng.normal(0, 0.02, (B, 4, 3))
    t = rng.uniform(0, gp.gait_period, B).round(3)
    v_des = np.stack([rng.uniform(0, 0.3, B), rng.uniform(-0.1, 0.1, B), np.zeros(B)], 1)
    w_des = np.where(rng.uniform(size=B) < 0.5, 0.0, rng.uniform(-0.1, 0.1, B))
    yaw = rng.uniform(-0.3, 0.3, B)
    amom_des = rng.normal(0, 0.05, (B, 3))
    sc = np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 3)))
    host = build_batch(rb, gp, com, vcom, amom, foot, t, v_des, w_des, yaw=yaw, amom_des=amom_des,
                       scale_W_X=sc[:, 0], scale_W_F=sc[:, 1], scale_rho=sc[:, 2])
    s = BatchSolver(host.n_col, host.n_eff, max_batch=B)
    dev = s.build_device(rb, gp, com, vcom, amom, foot, t, v_des, w_des, yaw=yaw, amom_des=amom_des, scales=sc)
    for f in ("x_init", "cnt_plan", "dt", "X_nom", "X_ter", "W_X", "W_X_ter", "W_F", "rho"):
        got = dev.fields[f].cpu().numpy().reshape(getattr(host, f).shape)
        assert np.array_equal(got, getattr(host, f)), f
    out = s.solve_resident(dev)
    import torch
    torch.cuda.synchronize()
    ref = s.solve(host)
    assert np.array_equal(out["iters"].cpu().numpy(), ref.iters)
    Fd = out["F"].cpu().numpy()
    assert ((Fd == ref.F) | (np.isnan(Fd) & np.isnan(ref.F))).all()


def test_gait_generator_shim_replans(oracle):
    """CyclicQuadrupedGaitGen (alias SoloMpcGaitGen): two consecutive replans from centroidal states; the FISTA step
    sizes persist across replans like the KinoDynMP's solver objects (simulation.py:408)."""
    from bunmpc_b200 import CyclicQuadrupedGaitGen
    from bunmpc_b200.motions import SOLO12, solo12_trot
    gg = CyclicQuadrupedGaitGen(None, None, None, planning_time=0.05)
    gg.update_gait_params(solo12_trot, 0.0)
    com, foot = np.array([0.01, -0.02, 0.21]), SOLO12.foot_pos
    L_prev = None
    for t in (0.0, 0.05):
        com_int, mom_int, f_int = gg.optimize_centroidal(com, [0.05, 0.0, 0.0], np.zeros(3), foot, t, np.array([0.2, 0.0, 0.0]), 0.0)
        batch, sol = gg.last
        ref = oracle.solve(batch)
        assert np.array_equal(sol.F, ref["F"]) and np.array_equal(sol.iters, ref["iters"])
        if L_prev is not None:
            assert np.array_equal(batch.L0, L_prev)
        L_prev = sol.L
        steps = sum(int(d / 0.001) for d in batch.dt[0][: gg.size])
        assert com_int.shape == (steps, 3) and mom_int.shape == (steps, 6) and f_int.shape == (steps, 12)
        assert np.array_equal(f_int[0], sol.F[0][:12])
