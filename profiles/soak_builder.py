"""Randomised soak of the on-device problem builder (contact plan, references, scaled weights) against the numpy
builder, bit for bit: random gaits, robots, horizons, gait times (incl. phase boundaries and many periods ahead),
velocities of either sign, turning rates, yaw over the full circle, scattered feet.
    python profiles/soak_builder.py [seconds] [seed]          (GPU box only)"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from bunmpc_b200.motions import GAITS, ROBOTS
from bunmpc_b200.plan_builder import build_batch
from bunmpc_b200.solver import BatchSolver

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t_end = time.time() + budget
n_cases = n_bad = 0
solvers = {}
while time.time() < t_end:
    robot = str(rng.choice(["solo12", "go2"]))
    gait = str(rng.choice(["trot", "bound", "jump"]))
    rb, gp = ROBOTS[robot], GAITS[robot][gait].scaled(float(rng.choice([0.5, 1.0, 1.5, 2.0])))
    B = int(rng.choice([1, 31, 257]))
    com = np.array([0.0, 0.0, gp.nom_ht]) + rng.normal(0, 0.05, (B, 3)) + np.concatenate([rng.uniform(-3, 3, (B, 2)), np.zeros((B, 1))], 1)
    vcom, amom = rng.normal(0, 0.3, (B, 3)), rng.normal(0, 0.05, (B, 3))
    foot = com[:, None, :] * np.array([1.0, 1.0, 0.0]) + rb.foot_pos + rng.normal(0, 0.03, (B, 4, 3))
    kind = rng.integers(0, 3)
    if kind == 0:
        t = rng.integers(0, 400, B) * gp.gait_dt                       # replanning grid, many periods ahead
    elif kind == 1:
        t = rng.uniform(0, 20 * gp.gait_period, B).round(3)
    else:
        t = rng.integers(0, 40, B) * gp.gait_period * np.array(gp.stance_percent)[rng.integers(0, 4, B)]   # phase edges
    v_des = np.stack([rng.uniform(-0.5, 0.5, B), rng.uniform(-0.2, 0.2, B), np.zeros(B)], 1)
    w_des = np.where(rng.uniform(size=B) < 0.4, 0.0, rng.uniform(-0.5, 0.5, B))
    yaw = rng.uniform(-np.pi, np.pi, B)
    amom_des = rng.normal(0, 0.05, (B, 3)) if rng.random() < 0.5 else None
    sc = np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 3))) if rng.random() < 0.5 else None
    kw = {} if sc is None else dict(scale_W_X=sc[:, 0], scale_W_F=sc[:, 1], scale_rho=sc[:, 2])
    host = build_batch(rb, gp, com, vcom, amom, foot, t, v_des, w_des, yaw=yaw, amom_des=amom_des, **kw)
    key = host.n_col
    if key not in solvers:
        solvers[key] = BatchSolver(host.n_col, host.n_eff, max_batch=257)
    dev = solvers[key].build_device(rb, gp, com, vcom, amom, foot, t, v_des, w_des, yaw=yaw, amom_des=amom_des, scales=sc)
    bad = []
    for f in ("x_init", "cnt_plan", "dt", "X_nom", "X_ter", "W_X", "W_X_ter", "W_F", "rho"):
        got = dev.fields[f].cpu().numpy().reshape(-1)
        want = np.broadcast_to(getattr(host, f), (B,) + getattr(host, f).shape[1:]).reshape(-1) if got.size != getattr(host, f).size else getattr(host, f).reshape(-1)
        if got.size != want.size or not (((got == want) | (np.isnan(got) & np.isnan(want))).all()):
            bad.append(f)
            d = np.flatnonzero(~((got == want) | (np.isnan(got) & np.isnan(want))))
            per = got.size // B
            i0 = d[0] // per
            print("  ", f, "differs in", len(d), "entries; instance", i0, "offsets", (d[:6] % per).tolist(), "gpu", got[d[:4]].tolist(), "host", want[d[:4]].tolist(),
                  "| t", float(np.atleast_1d(t)[i0]), "com", com[i0].tolist(), "v_des", v_des[i0].tolist(), "w_des", float(w_des[i0]), "yaw", float(yaw[i0]), flush=True)
    n_cases += 1
    if bad:
        n_bad += 1
        print("MISMATCH", robot, gait, gp.gait_horizon, B, "kind", kind, bad, flush=True)
print(f"builder soak: {n_cases} cases, {n_bad} mismatching cases")
sys.exit(1 if n_bad else 0)
