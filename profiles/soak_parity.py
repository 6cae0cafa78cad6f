"""Randomised GPU-vs-oracle soak: random gaits, robots, horizons, batch sizes, iteration caps, tolerances, initial step
sizes (forcing line-search rejections), arithmetic modes and slice lengths; every result must be bit-identical.
    python profiles/soak_parity.py [seconds] [seed]          (GPU box only; the oracle is the checker)"""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import oracle
from bunmpc_b200 import synthetic
from bunmpc_b200._lib import ARITH_FMA, ARITH_STRICT
from bunmpc_b200.problem import SolverParams
from bunmpc_b200.solver import BatchSolver



def truncate_or_tile(b, n):
    """The same states with a horizon of n knots: contact plan, dt, weights and references cut or repeated."""
    from bunmpc_b200.problem import CentroidalBatch
    n0 = b.n_col
    idx = np.arange(n) % n0

    def knots(a, per):           # [Bf, n0*per] -> [Bf, n*per]
        return np.ascontiguousarray(a.reshape(a.shape[0], n0, per)[:, idx].reshape(a.shape[0], n * per))
    return CentroidalBatch(n_col=n, n_eff=b.n_eff, m=b.m, rho=b.rho, x_init=b.x_init,
                           cnt_plan=np.ascontiguousarray(b.cnt_plan[:, idx]), dt=np.ascontiguousarray(b.dt[:, idx]),
                           W_X=knots(b.W_X, 9), W_X_ter=b.W_X_ter, X_nom=knots(b.X_nom, 9), X_ter=b.X_ter,
                           W_F=knots(b.W_F, 3 * b.n_eff), bounds=np.ascontiguousarray(b.bounds[:, idx]), L0=b.L0)


oracle.build()
BATCHES = [int(x) for x in os.environ.get("SOAK_BATCHES", "1,7,33,150,300,600").split(",")]
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t_end = time.time() + budget
n_cases = n_inst = n_bad = 0
solvers = {}
while time.time() < t_end:
    gait = rng.choice(["trot", "bound", "jump"])
    robot = rng.choice(["solo12", "solo12", "go2"])
    scale = float(rng.choice([0.5, 1.0, 1.0, 1.5, 2.0]))
    B = int(rng.choice(BATCHES))
    bseed = int(rng.integers(1 << 30))
    b = synthetic.perturbed(B, robot, gait, seed=bseed, horizon_scale=scale,
                            vy_range=(-0.1, 0.1), w_range=(-0.1, 0.1),
                            weight_scale_range=(0.5, 2.0) if rng.random() < 0.3 else None)
    if rng.random() < 0.35:                                   # any horizon 1..SOAK_NMAX (every CTA size; default 88)
        n = int(rng.integers(1, int(os.environ.get("SOAK_NMAX", "88")) + 1))
        B = min(B, 150 if n <= 88 else 12)
        b = b.select(np.arange(B))
        b = truncate_or_tile(b, n)
    prm = SolverParams(max_outer=int(rng.choice([3, 10, 25, 100])), max_inner=int(rng.choice([1, 2, 3, 4, 5, 9, 40, 150])),
                       tol=float(rng.choice([1e-5, 1e-3, 1e-1])), slice_outer=int(rng.choice([0, 0, 1, 3, -1])),
                       exit_tol=float(rng.choice([1e-3, 1e-3, 1e-2, 1e-5])), beta=float(rng.choice([1.5, 1.5, 2.0, 1.1])),
                       mu=float(rng.choice([1.0, 1.0, 0.6, 1.3])))
    if prm.slice_outer > 0 and prm.max_outer > 16 * prm.slice_outer:
        prm.slice_outer = 0
    if rng.random() < 0.3:
        b.L0 = np.array([[float(10 ** rng.uniform(-1, 2.7)), float(10 ** rng.uniform(1, 6.3))]])
    if rng.random() < 0.2:
        B = b.B
        nx, nf = 9 * (b.n_col + 1), 12 * b.n_col
        b.X0, b.F0, b.P0 = rng.normal(0, 0.1, (B, nx)), rng.normal(0, 1.0, (B, nf)), rng.normal(0, 1e-3, (B, nx))
    fma = rng.random() < 0.25
    key = (b.n_col, max(BATCHES))
    if key not in solvers:
        solvers[key] = BatchSolver(b.n_col, 4, max_batch=max(BATCHES))
    sol = solvers[key].solve(b, prm, arith=ARITH_FMA if fma else ARITH_STRICT)
    ref = oracle.solve(b, oracle.default_params(max_outer=prm.max_outer, max_inner=prm.max_inner, tol=prm.tol,
                                                exit_tol=prm.exit_tol, beta=prm.beta, mu=prm.mu,
                                                use_fma=1 if fma else 0), n_threads=32)
    ok = np.array_equal(sol.iters, ref["iters"]) and np.array_equal(sol.status, ref["status"])
    detail = []
    if not ok:
        bad_i = np.flatnonzero((sol.iters != ref["iters"]).any(1) | (sol.status != ref["status"]))
        detail.append(f"iters/status differ in {len(bad_i)} instances, first {bad_i[:3]}: gpu {sol.iters[bad_i[:2]].tolist()} {sol.status[bad_i[:2]].tolist()} ref {ref['iters'][bad_i[:2]].tolist()} {ref['status'][bad_i[:2]].tolist()}")
    for k in ("F", "X", "P", "L", "viol"):
        a, r = getattr(sol, k), ref[k]
        same = (a == r) | (np.isnan(a) & np.isnan(r))
        if not same.all():
            ok = False
            rows = np.flatnonzero(~same.reshape(same.shape[0], -1).all(1))
            detail.append(f"{k}: {len(rows)} instances, first {rows[:3]}, gpu nan {np.isnan(a[rows[0]]).any()} ref nan {np.isnan(r[rows[0]]).any()}")
    n_cases += 1; n_inst += B
    if not ok:
        n_bad += 1
        print("MISMATCH", gait, robot, scale, B, "n", b.n_col, "seed", bseed, prm, "fma" if fma else "strict", "L0", b.L0.tolist(), "warm", b.X0 is not None, "|", " ; ".join(detail), flush=True)
print(f"soak: {n_cases} cases, {n_inst} instances, {n_bad} mismatching cases")
sys.exit(1 if n_bad else 0)
