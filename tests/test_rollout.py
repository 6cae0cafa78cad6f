"""Lock-step batched rollouts (f-3): the batched loop must produce what B sequential episode loops produce,
including the FISTA step sizes carried from replan to replan and the handling of failed episodes."""
import numpy as np
import pytest


def _oracle_solver(oracle, n_threads=8):
    from bunmpc_b200.problem import BatchSolution

    def solve(batch):
        r = oracle.solve(batch, n_threads=n_threads)
        return BatchSolution(m=batch.m, **r)
    return solve


def _initial_states(B, seed=3):
    from bunmpc_b200.motions import GAITS, ROBOTS
    from bunmpc_b200.rollout import EpisodeState
    rb, gp = ROBOTS["solo12"], GAITS["solo12"]["trot"]
    rng = np.random.default_rng(seed)
    com = np.array([0.0, 0.0, gp.nom_ht]) + rng.normal(0.0, 0.01, (B, 3))
    foot = np.broadcast_to(rb.foot_pos, (B, 4, 3)).copy()
    st = EpisodeState(com, rng.normal(0.0, 0.05, (B, 3)), np.zeros((B, 3)), foot,
                      rng.integers(0, 10, B) * gp.gait_dt, np.zeros(B))
    v_des = np.zeros((B, 3)); v_des[:, 0] = rng.uniform(0.0, 0.3, B)
    return rb, gp, st, v_des


def _records_equal(a, b, idx=None):
    for name in ("com", "vcom", "F0", "iters"):
        for x, y in zip(getattr(a, name), getattr(b, name)):
            if idx is not None:
                x = x[idx]
            assert ((x == y) | (np.isnan(x) & np.isnan(y))).all(), name


def test_lockstep_equals_sequential_episodes_cpu(oracle):
    from bunmpc_b200.rollout import LockstepRollouts, TrackingPlant
    rb, gp, st, v_des = _initial_states(3)
    solve = _oracle_solver(oracle)
    both = LockstepRollouts(rb, gp, plant=TrackingPlant(), solve_fn=solve).run(st, v_des, 0.0, n_ticks=3)
    assert len(both.com) == 3 and (both.failed_at == -1).all()
    for i in range(3):       # the reference's way: one episode at a time
        one = LockstepRollouts(rb, gp, plant=TrackingPlant(), solve_fn=solve).run(st.select(np.array([i])), v_des[i:i + 1], 0.0, n_ticks=3)
        _records_equal(both, one, idx=np.array([i]))
    # step sizes were carried: the second replan of an episode differs from a cold one only through L0, which shows
    # up in the line-search counters of tick 0 (cold: rejections from L0 = 506.25/2.25e6 upward) vs later ticks
    assert both.iters[0][:, 3:5].sum() >= both.iters[1][:, 3:5].sum()


def test_failed_episodes_leave_the_batch(oracle):
    from bunmpc_b200.problem import BatchSolution
    from bunmpc_b200.rollout import LockstepRollouts
    rb, gp, st, v_des = _initial_states(4)
    base = _oracle_solver(oracle)
    sizes = []

    def solve(batch):
        sizes.append(batch.B)
        sol = base(batch)
        if len(sizes) == 1:
            sol.F[1, 5] = np.nan                   # episode 1 diverges at the first replan (simulation.py:513-516)
        return sol
    rec = LockstepRollouts(rb, gp, solve_fn=solve).run(st, v_des, 0.0, n_ticks=3)
    assert sizes == [4, 3, 3]
    assert rec.failed_at.tolist() == [-1, 0, -1, -1]
    assert np.isnan(rec.com[1][1]).all() and np.isfinite(rec.com[2][0]).all()
    err = rec.tracking_error(v_des)
    assert np.isinf(err[1]) and np.isfinite(err[[0, 2, 3]]).all()


def test_tracking_plant_follows_the_plan():
    from bunmpc_b200 import synthetic
    from bunmpc_b200.problem import BatchSolution
    from bunmpc_b200.rollout import EpisodeState, TrackingPlant
    b = synthetic.perturbed(2, "solo12", "trot", seed=1)
    n = b.n_col
    X = np.arange(2 * 9 * (n + 1), dtype=np.float64).reshape(2, -1)
    sol = BatchSolution(X=X, F=np.zeros((2, 12 * n)), P=X * 0, L=np.ones((2, 2)), iters=np.zeros((2, 5), int),
                        viol=np.zeros(2), status=np.zeros(2, int), m=b.m)
    st = EpisodeState(np.zeros((2, 3)), np.zeros((2, 3)), np.zeros((2, 3)), np.zeros((2, 4, 3)), np.zeros(2), np.zeros(2))
    dt0 = np.broadcast_to(b.dt, (2, n))[:, 0]
    out = TrackingPlant()(st, b, sol, 0.5 * dt0[0])
    Xr = X.reshape(2, n + 1, 9)
    a = 0.5 * dt0[0] / dt0
    assert np.allclose(out.com, (1 - a)[:, None] * Xr[:, 0, 0:3] + a[:, None] * Xr[:, 1, 0:3])
    assert np.allclose(out.t, 0.5 * dt0[0])


def test_goal_posterior_matches_single_updates():
    from bunmpc_b200 import dist
    from bunmpc_b200.rollout import GoalPosterior
    gp = GoalPosterior(n=20)
    g1, g2 = np.array([0.1, 0.02, -0.03]), np.array([0.25, -0.05, 0.0])
    ref = dist.posterior_update(gp.p, dist.gaussian_likelihood_grid(gp.axes, g1, 0.1))
    assert np.allclose(gp.update(g1), ref, rtol=1e-12)
    ref = dist.posterior_update(ref, dist.gaussian_likelihood_grid(gp.axes, g2, 0.1))
    both = GoalPosterior(n=20)
    assert np.allclose(both.update_batch(np.stack([g1, g2])), ref, rtol=1e-10)
    s = both.sample(1000, np.random.default_rng(0))
    assert s.shape == (1000, 3) and (s[:, 0] >= 0).all() and (s[:, 0] <= 0.3).all()
    assert abs(np.average(gp.axes[0], weights=both.p.sum((1, 2))) - s[:, 0].mean()) < 0.02


def test_goal_posterior_update_as_the_reference_calls_it():
    """locosafedagger_modified.py:611 passes (v_des[0], v_des[1], w_des, error) into (observed_goal, vx_obs, vy_obs, w_obs):
    the triple loop of :374-384 with those bindings, restated here on a small grid, is what update_as_called multiplies
    in; update() is the documented (unshifted) behaviour and differs."""
    from bunmpc_b200.rollout import GoalPosterior
    gp, doc = GoalPosterior(n=9), GoalPosterior(n=9)
    v_des, w_des, error, sigma = np.array([0.22, -0.04, 0.0]), 0.06, 0.031, 0.1
    vx_obs, vy_obs, w_obs = v_des[1], w_des, error                  # the call's bindings
    lik = np.zeros((9, 9, 9))
    for i, vx in enumerate(gp.axes[0]):
        for j, vy in enumerate(gp.axes[1]):
            for k, w in enumerate(gp.axes[2]):
                d = np.array([vx - vx_obs, vy - vy_obs, w - w_obs])
                lik[i, j, k] = np.exp(-np.sum(d ** 2) / (2 * sigma ** 2))
    lik /= lik.sum()
    post = np.full((9, 9, 9), 1.0 / 9 ** 3) * lik
    post /= post.sum()
    assert np.allclose(gp.update_as_called(v_des, w_des, error), post, rtol=1e-12)
    assert not np.allclose(doc.update([v_des[0], v_des[1], w_des]), post, rtol=1e-3)


@pytest.mark.gpu
def test_lockstep_rollouts_gpu_equals_oracle(oracle):
    import torch
    assert torch.cuda.is_available()
    from bunmpc_b200.rollout import GoalPosterior, LockstepRollouts, TrackingPlant
    rb, gp, st, v_des = _initial_states(24, seed=9)
    gpu = LockstepRollouts(rb, gp, plant=TrackingPlant(0.003, 0.02, 0.0, seed=4))
    rec = gpu.run(st, v_des, 0.0, n_ticks=4)
    ref = LockstepRollouts(rb, gp, plant=TrackingPlant(0.003, 0.02, 0.0, seed=4), solve_fn=_oracle_solver(oracle, 16)).run(st, v_des, 0.0, n_ticks=4)
    assert gpu.launches == 4
    _records_equal(rec, ref)
    # the posterior grid on the device agrees with the host one
    d, h = GoalPosterior(n=40, device="cuda:0"), GoalPosterior(n=40)
    w = np.exp(-rec.tracking_error(v_des))
    d.update_batch(v_des, w); h.update_batch(v_des, w)
    assert np.allclose(d.p.cpu().numpy(), h.p, rtol=1e-9, atol=1e-300)


@pytest.mark.gpu
def test_lockstep_rollouts_device_builder_equals_host_builder():
    """f-3 on the device path: states up, build_problem_kernel + solve where the problem lies, plan down.  Same records,
    bit for bit, as the host-builder loop (disturbances included), with ~20x fewer bytes towards the GPU."""
    import torch
    assert torch.cuda.is_available()
    from bunmpc_b200.rollout import GoalPosterior, LockstepRollouts, TrackingPlant
    rb, gp, st, v_des = _initial_states(40, seed=11)
    host = LockstepRollouts(rb, gp, plant=TrackingPlant(0.003, 0.02, 0.0, seed=4))
    devr = LockstepRollouts(rb, gp, plant=TrackingPlant(0.003, 0.02, 0.0, seed=4), builder="device")
    rh = host.run(st, v_des, 0.0, n_ticks=5)
    rd = devr.run(st, v_des, 0.0, n_ticks=5)
    assert devr.launches == 5
    _records_equal(rd, rh)
    assert devr.h2d_bytes * 10 < host.h2d_bytes
    # goals drawn on the device come from the grid and follow the posterior
    post = GoalPosterior(n=40, device="cuda:0")
    post.update_batch(v_des, np.exp(-rd.tracking_error(v_des)))
    g = post.sample_device(4000, generator=torch.Generator(device="cuda:0").manual_seed(1))
    assert g.shape == (4000, 3) and g.is_cuda
    mean_vx = float((post.p.sum((1, 2)) * post._ax[0]).sum())
    assert abs(float(g[:, 0].mean()) - mean_vx) < 0.01
