"""Randomised CPU soak of the parity chain's first links (runs where /root/reference is mounted and oracle/_ref is built):
  (1) the reference's own sources compiled against the Eigen stand-in  ==  the oracle (canonical order, what the GPU
      computes),
  (2) the oracle with a plain tree of 32-leaf blocks (a rounding variant of the dense sums)  ==  (1),
bit for bit, on random gaits, robots, horizons, initial step sizes, warm starts and iteration caps.
    python oracle/soak_ref.py [seconds] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
from bunmpc_b200 import synthetic

oracle.build()
if not oracle.ref_available():
    print("oracle/_ref not built (needs /root/reference)"); sys.exit(0)
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t_end = time.time() + budget


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(((a == b) | (np.isnan(a.astype(float)) & np.isnan(b.astype(float)))).all())


n_cases = n_bad32 = n_bad30 = 0
while time.time() < t_end:
    gait, robot = str(rng.choice(["trot", "bound", "jump"])), str(rng.choice(["solo12", "solo12", "go2"]))
    scale = float(rng.choice([0.5, 1.0, 1.0, 1.5]))
    b = synthetic.perturbed(3, robot, gait, seed=int(rng.integers(1 << 30)), horizon_scale=scale,
                            vy_range=(-0.1, 0.1), w_range=(-0.1, 0.1),
                            weight_scale_range=(0.5, 2.0) if rng.random() < 0.3 else None)
    if rng.random() < 0.4:
        b.L0 = np.array([[float(10 ** rng.uniform(-1, 2.7)), float(10 ** rng.uniform(1, 6.3))]])
    if rng.random() < 0.25:
        nx, nf = 9 * (b.n_col + 1), 12 * b.n_col
        b.X0, b.F0, b.P0 = rng.normal(0, 0.1, (3, nx)), rng.normal(0, 1.0, (3, nf)), rng.normal(0, 1e-3, (3, nx))
    mo = int(rng.choice([4, 15, 100]))
    ref = oracle.ref_solve(b, max_outer=mo)
    o30 = oracle.solve(b, params=oracle.default_params(max_outer=mo, reduction=32), n_threads=3)
    o32 = oracle.solve(b, params=oracle.default_params(max_outer=mo), n_threads=3)
    ok32 = all(same(ref[k], o32[k]) for k in ("X", "F", "P", "L", "iters", "viol"))
    ok30 = all(same(o32[k], o30[k]) for k in ("X", "F", "P", "L", "iters", "status"))   # `viol` is itself such a sum
    n_cases += 1
    if not ok32:
        n_bad32 += 1
        print("REF != ORACLE", gait, robot, scale, mo, b.L0.tolist(), flush=True)
    if not ok30:
        n_bad30 += 1
        print("ORACLE(plain 32) != ORACLE", gait, robot, scale, mo, b.L0.tolist(), flush=True)
print(f"reference soak: {n_cases} cases x 3 instances; reference-sources vs oracle: {n_bad32} mismatches; "
      f"oracle(plain 32-leaf tree) vs oracle: {n_bad30} mismatches")
