"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle, bit for bit
(integer counters and IEEE binary64 values identical), on the same seeded inputs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _require_gpu():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"


def assert_same(sol, ref, what=""):
    assert np.array_equal(sol.iters, ref["iters"]), f"{what}: iteration counters differ\n{sol.iters}\n{ref['iters']}"
    assert np.array_equal(sol.status, ref["status"]), f"{what}: status differs"
    for k in ("F", "X", "P", "L", "viol"):
        a, b = getattr(sol, k), ref[k]
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        assert same.all(), f"{what}: {k} differs in {np.count_nonzero(~same)} entries, max abs {np.nanmax(np.abs(a - b))}"


def test_nominal_single_solve(oracle):
    """BASELINE config 1: Solo12 trot, one solve."""
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.nominal()
    s = BatchSolver(b.n_col, b.n_eff, max_batch=1)
    sol = s.solve(b)
    ref = oracle.solve(b)
    assert_same(sol, ref, "nominal")
    assert sol.status[0] == 0 and sol.viol[0] < 1e-3
    assert s.launch_count() == 2          # expand + solve


@pytest.mark.parametrize("gait,B,seed", [("trot", 64, 0), ("bound", 16, 1), ("jump", 16, 2)])
def test_perturbed_batches(oracle, gait, B, seed):
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.perturbed(B, "solo12", gait, seed=seed)
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=B).solve(b)
    ref = oracle.solve(b, n_threads=8)
    assert_same(sol, ref, gait)


def test_fma_mode_matches_fma_oracle(oracle):
    _require_gpu()
    from bunmpc_b200 import synthetic, ARITH_FMA
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.perturbed(32, "solo12", "trot", seed=5)
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=32).solve(b, arith=ARITH_FMA)
    ref = oracle.solve(b, params=oracle.default_params(use_fma=1), n_threads=8)
    assert_same(sol, ref, "fma")


def test_line_search_rejections_replay(oracle):
    """Tiny initial step sizes force many line-search rejections (L *= 1.5), which sends the kernel's
    speculative pipeline through its sequential-replay path; counters and iterates must still match."""
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.perturbed(16, "solo12", "trot", seed=7)
    b.L0 = np.array([[1.0, 40.0]])
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=16).solve(b)
    ref = oracle.solve(b, n_threads=8)
    assert ref["iters"][:, 3].min() > 5 and ref["iters"][:, 4].min() > 5     # rejections did happen
    assert_same(sol, ref, "rejections")


def test_go2_mass_hits_cap_and_backtracks(oracle):
    """Solo-tuned weights with the Go2 mass: iteration cap and occasional L_x backtracks."""
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.perturbed(16, "go2", "trot", seed=11)
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=16).solve(b)
    ref = oracle.solve(b, n_threads=8)
    assert_same(sol, ref, "go2")


@pytest.mark.parametrize("gait,scale,n_expected", [("trot", 2.0, 40), ("bound", 2.0, 48), ("jump", 2.0, 60)])
def test_longer_horizons(oracle, gait, scale, n_expected):
    """BASELINE config 4: longer horizons (gait_horizon x2 as in analysis/solve_times_test.py:60-66).
    n = 40 still fits the split warp roles; n = 48 and 60 run with combined roles."""
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.perturbed(6, "solo12", gait, seed=17, horizon_scale=scale)
    assert b.n_col == n_expected
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=8).solve(b)
    ref = oracle.solve(b, n_threads=8)
    assert_same(sol, ref, f"{gait} n={n_expected}")


@pytest.mark.parametrize("gait,n_expected", [("bound", 48), ("jump", 60)])
def test_longer_horizons_with_rejections(oracle, gait, n_expected):
    """Combined-role kernels (n = 48, 60) with tiny initial step sizes: the line-search test needs the
    |A y + b|^2 partials of the row work in the one-barrier pipeline as well as in the sequential replay."""
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.perturbed(6, "solo12", gait, seed=23, horizon_scale=2.0)
    assert b.n_col == n_expected
    b.L0 = np.array([[1.0, 40.0]])
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=8).solve(b)
    ref = oracle.solve(b, n_threads=8)
    assert ref["iters"][:, 3].min() > 5 and ref["iters"][:, 4].min() > 5
    assert_same(sol, ref, f"{gait} n={n_expected} with rejections")


def test_time_sliced_instances(oracle):
    """More instances than resident CTAs: the kernel parks an instance every few outer iterations and resumes it
    later, possibly on another SM.  Results, counters and the violation history must not notice."""
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    probe = BatchSolver(20, 4, max_batch=1)
    B = probe.kernel_info()["num_sms"] * probe.kernel_info()["ctas_per_sm"] + 141      # more than fit at once
    b = synthetic.perturbed(B, "solo12", "trot", seed=5)
    s = BatchSolver(b.n_col, b.n_eff, max_batch=B)
    sol = s.solve(b, viol_hist=True)
    ref = oracle.solve(b, n_threads=16)
    assert_same(sol, ref, "time-sliced")
    assert sol.iters[:, 0].max() > 16                       # some instances were parked at least twice
    for i in (0, 100, B - 1):                               # history: one entry per outer iteration, NaN after the exit
        k = sol.iters[i, 0]
        assert np.isfinite(sol.viol_hist[i, :k]).all() and np.isnan(sol.viol_hist[i, k:]).all()
        assert sol.viol_hist[i, k - 1] == sol.viol[i]
    assert (sol.cycles > 0).all()


@pytest.mark.parametrize("slice_outer", [1, 3, -1])
def test_slice_length_never_changes_results(oracle, slice_outer):
    """Parking after every single outer iteration (up to 11 parks per instance), every third, or never."""
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.problem import SolverParams
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.perturbed(300, "solo12", "trot", seed=31)
    prm = SolverParams(max_outer=12, slice_outer=slice_outer)
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=300).solve(b, prm, viol_hist=True)
    ref = oracle.solve(b, oracle.default_params(max_outer=12), n_threads=16)
    assert_same(sol, ref, f"slice_outer={slice_outer}")
    assert np.isfinite(sol.viol_hist[np.arange(300), sol.iters[:, 0] - 1]).all()


@pytest.mark.parametrize("long_inner", ["0", "40", "700", "1e30"])
def test_parked_queue_thresholds_never_change_results(oracle, monkeypatch, long_inner):
    """Parked instances wait in eight queues by predicted remaining work (longest first, kernels.cuh parking code).  The
    threshold only decides who runs next: everything in the last queue (0), spread over all of them (40, 700: the
    instances of this batch have 100-3000 inner iterations left when they park), everything in the first (1e30), with
    few resident CTAs (BUNMPC_MAX_CTAS) so that every queue is drained by CTAs other than the ones that filled it."""
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.problem import SolverParams
    from bunmpc_b200.solver import BatchSolver
    monkeypatch.setenv("BUNMPC_LONG_INNER", long_inner)
    monkeypatch.setenv("BUNMPC_MAX_CTAS", "7")
    b = synthetic.perturbed(60, "solo12", "trot", seed=77)
    prm = SolverParams(max_outer=30, slice_outer=2)
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=64).solve(b, prm)
    ref = oracle.solve(b, oracle.default_params(max_outer=30), n_threads=16)
    assert sol.iters[:, 0].max() > 6                        # parked at least three times
    assert_same(sol, ref, f"long_inner={long_inner}")


def test_baseline_config1_full_batch_bit_for_bit(oracle):
    """BASELINE config[1] at its full size: 1024 perturbed Solo12 trot states on one B200, every instance compared
    with the oracle -- values, step sizes, iteration counters, status (the oracle needs a few seconds on the host cores)."""
    _require_gpu()
    import os
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.config(1, B=1024, seed=0)
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=1024).solve(b)
    ref = oracle.solve(b, n_threads=max(1, len(os.sched_getaffinity(0))))
    assert_same(sol, ref, "config1 B=1024")
    assert (sol.status == 0).mean() > 0.8


def test_bayesian_samples_batch(oracle):
    """BASELINE config 5 shape: per-instance goals (vx, vy, w) and log-uniform scalings of W_X, W_F and rho."""
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.config(4, B=96, seed=5)
    assert b.rho.shape == (96,) and b.W_F.shape[0] == 96
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=96).solve(b)
    ref = oracle.solve(b, n_threads=8)
    assert_same(sol, ref, "bayes")


def test_mixed_mode_matches_its_oracle_and_the_tolerance(oracle):
    """BUNMPC_ARITH_MIXED (ATA_ / A_ entries stored in binary32, binary64 arithmetic -- the path's FP32 mode): bit-exact
    against the oracle with storage = 1 on BASELINE config[1] (B = 1024).  Against the binary64 solve: within the north
    star's 1e-3 relative on every instance that converges with the same iteration counters in both modes; the shares
    that keep the counters / converge are printed (instances that run into the 100-iteration cap without converging
    are chaotic in either mode and are excluded from the tolerance claim)."""
    _require_gpu()
    from bunmpc_b200 import synthetic, ARITH_MIXED
    from bunmpc_b200.solver import BatchSolver
    b = synthetic.config(1, B=1024, seed=0)
    s = BatchSolver(b.n_col, b.n_eff, max_batch=1024)
    sol = s.solve(b, arith=ARITH_MIXED)
    ref = oracle.solve(b, params=oracle.default_params(storage=1), n_threads=16)
    assert_same(sol, ref, "mixed")
    f64 = oracle.solve(b, n_threads=16)
    same = (sol.iters == f64["iters"]).all(axis=1)
    conv = (sol.status == 0) & (f64["status"] == 0)

    def rel(mask):
        r = np.zeros(int(mask.sum()))
        for k in ("F", "X"):
            a, c = getattr(sol, k)[mask], f64[k][mask]
            r = np.maximum(r, np.abs(a - c).max(axis=1) / np.abs(c).max(axis=1))
        return r
    r_sc, r_c = rel(same & conv), rel(conv)
    print(f"mixed mode vs binary64 on 1024 instances: {same.mean() * 100:.1f} % keep the iteration counters, "
          f"{conv.mean() * 100:.1f} % converge in both; same counters & converged: median {np.median(r_sc):.1e}, "
          f"max {r_sc.max():.1e}; all converged: 99th percentile {np.percentile(r_c, 99):.1e}, max {r_c.max():.1e}")
    assert same.mean() > 0.8
    assert r_sc.max() <= 1e-3       # north star: within 1e-3 relative in FP32 mode
    assert np.median(r_c) <= 1e-6


@pytest.mark.parametrize("cfg,per_gpu", [(2, 2048), (3, 512), (4, 8192)])
def test_per_gpu_shard_sizes_sampled(oracle, cfg, per_gpu):
    """BASELINE configs 3-5 at the size ONE of eight GPUs sees (16 384 / 4 096 / 65 536 instances sharded interleaved):
    many more instances than CTA slots, so every instance is parked and resumed several times and the work queue is under
    pressure.  The GPU solves the whole shard; the oracle re-solves 256 randomly drawn instances, bit for bit."""
    _require_gpu()
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    full = synthetic.config(cfg, B=8 * per_gpu, seed=5)
    b = full.shard(3, 8)
    assert b.B == per_gpu
    sol = BatchSolver(b.n_col, b.n_eff, max_batch=per_gpu).solve(b)
    idx = np.sort(np.random.default_rng(cfg).choice(per_gpu, 256, replace=False))
    ref = oracle.solve(b.select(idx), n_threads=os.cpu_count() or 8)
    assert np.array_equal(sol.iters[idx], ref["iters"]), f"config {cfg}: iteration counters differ"
    assert np.array_equal(sol.status[idx], ref["status"])
    for k in ("F", "X", "P", "L", "viol"):
        a, r = getattr(sol, k)[idx], ref[k]
        same = (a == r) | (np.isnan(a) & np.isnan(r))
        assert same.all(), f"config {cfg}: {k} differs in {np.count_nonzero(~same)} entries"


@pytest.mark.parametrize("n", [64, 93, 94, 120, 128, 129, 150, 160, 161, 192, 200, 206])
def test_horizons_of_the_reference_timing_sweep(oracle, n):
    """analysis/solve_times_test.py:60-66 sweeps gait_horizon up to 20 periods (trot: n = 200) and 32 (bound: n = 192).
    Up to n = 160 an instance runs the pipelined loops -- from n = 94 on (CTAs of 512 and 640 threads) without the bank
    padding and the row records (kernels.cuh: big_cta); beyond (768 and 1024 threads, seq_cta) the shared memory of one
    SM only holds the sequential loops' buffers, up to n = 206; beyond that the solver refuses loudly.  The horizons
    here sit on both sides of every CTA-size step."""
    _require_gpu()
    from bunmpc_b200.motions import GAITS, ROBOTS
    from bunmpc_b200.plan_builder import build_batch
    from bunmpc_b200.problem import SolverParams
    from bunmpc_b200.solver import BatchSolver
    rb, gp = ROBOTS["solo12"], GAITS["solo12"]["trot"]
    B = 3
    rng = np.random.default_rng(n)
    com = np.array([0.0, 0.0, gp.nom_ht]) + rng.normal(0.0, 0.02, (B, 3))
    foot = np.broadcast_to(rb.foot_pos, (B, 4, 3)).copy()
    foot[:, :, :2] += rng.normal(0.0, 0.02, (B, 4, 2))
    v_des = np.zeros((B, 3)); v_des[:, 0] = rng.uniform(0.0, 0.3, B)
    b = build_batch(rb, gp, com, rng.normal(0.0, 0.1, (B, 3)), rng.normal(0.0, 0.02, (B, 3)), foot,
                    rng.integers(0, 10, B) * gp.gait_dt, v_des, np.zeros(B), horizon=n)
    assert b.n_col == n
    prm = SolverParams(max_outer=8)
    s = BatchSolver(n, 4, max_batch=B)
    assert s.kernel_info()["threads"] == {64: 256, 93: 384, 94: 512, 120: 512, 128: 512, 129: 640, 150: 640, 160: 640,
                                          161: 768, 192: 768, 200: 1024, 206: 1024}[n]
    sol = s.solve(b, prm)
    ref = oracle.solve(b, oracle.default_params(max_outer=8), n_threads=B)
    assert_same(sol, ref, f"n={n}")
    b.L0 = np.array([[1.0, 40.0]])            # tiny step sizes: rejections, i.e. the replay with the sequential loops
    sol = s.solve(b, prm)
    ref = oracle.solve(b, oracle.default_params(max_outer=8), n_threads=B)
    assert ref["iters"][:, 3:5].min() > 3
    assert_same(sol, ref, f"n={n} with rejections")
    with pytest.raises(RuntimeError, match="shared memory"):
        BatchSolver(207, 4, max_batch=1)
    with pytest.raises(RuntimeError, match="too large"):
        BatchSolver(249, 4, max_batch=1)


def test_three_warp_ctas_sharing_an_sm(oracle):
    """n = 16: CTAs of three warps (two workers + the service warp), two resident per SM, which start at different
    schedulers -- the case in which logical_warp() permutes the warp roles by scheduler placement.  400 instances keep
    every SM's two slots busy; results must not depend on which hardware warp took which share."""
    _require_gpu()
    from bunmpc_b200.motions import GAITS, ROBOTS
    from bunmpc_b200.plan_builder import build_batch
    from bunmpc_b200.problem import SolverParams
    from bunmpc_b200.solver import BatchSolver
    rb, gp = ROBOTS["solo12"], GAITS["solo12"]["trot"]
    B, n = 400, 16
    rng = np.random.default_rng(16)
    com = np.array([0.0, 0.0, gp.nom_ht]) + rng.normal(0.0, 0.02, (B, 3))
    foot = np.broadcast_to(rb.foot_pos, (B, 4, 3)).copy()
    foot[:, :, :2] += rng.normal(0.0, 0.02, (B, 4, 2))
    v_des = np.zeros((B, 3)); v_des[:, 0] = rng.uniform(0.0, 0.3, B)
    b = build_batch(rb, gp, com, rng.normal(0.0, 0.1, (B, 3)), rng.normal(0.0, 0.02, (B, 3)), foot,
                    rng.integers(0, 10, B) * gp.gait_dt, v_des, np.zeros(B), horizon=n)
    prm = SolverParams(max_outer=10)
    s = BatchSolver(n, 4, max_batch=B)
    assert s.kernel_info()["threads"] == 96
    sol = s.solve(b, prm)
    ref = oracle.solve(b, oracle.default_params(max_outer=10), n_threads=os.cpu_count() or 8)
    assert_same(sol, ref, "n=16, two CTAs of three warps per SM")
