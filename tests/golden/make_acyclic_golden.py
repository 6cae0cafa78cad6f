"""Generates tests/golden/acyclic_cases.npz: what the REFERENCE'S OWN acyclic gait generator hands to the solver --
examples/mpc/abstract_acyclic_gen.py (`SoloAcyclicGen.create_contact_plan` :74-124, the dynamics part of
`create_costs` :126-190) imported in place from /root/reference and run on the reference's own motion files
(examples/motions/acyclic/{plan_jump,rearing,plan_hifive,plan_cartwheel,rearing_jump}.py).  pinocchio, the pybind
modules and matplotlib are the stand-ins of oracle/pinshim (the centre of mass and the centroidal momentum the
generator reads from pinocchio are injected; the solver and IK objects are recorders).

Run in the build container:   python tests/golden/make_acyclic_golden.py

Stored: per motion, the centroidal part of the reference's motion record (pins bunmpc_b200/acyclic.py's table); per
case, the inputs of bunmpc_b200.acyclic.build_batch (x_init, t, t0) and what the reference passed to
set_contact_plan / create_bound_constraints / create_cost_X / create_cost_F / set_rho.

`stand.py` is not covered: it still has the three-column bounds of an older record format and no `plan_freq`, and the
reference's own update_motion_params / create_costs raise on it."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/iterative_supervised_learning/examples"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "acyclic_cases.npz")
MOTION_FILES = {"jump_fwd": "plan_jump", "rearing": "rearing", "hifive": "plan_hifive",
                "cartwheel": "plan_cartwheel", "rearing_jump": "rearing_jump"}
OUT_KEYS = ("cnt_plan", "dt", "X_nom", "X_ter", "W_X", "W_X_ter", "W_F", "bounds", "rho", "f_max")


def main():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "pinshim"), os.path.join(REF, "mpc"), REF, ROOT]
    import pinocchio as pin                                     # the stand-in
    gen = importlib.import_module("abstract_acyclic_gen")      # the reference's own file
    rng = np.random.default_rng(7)
    out, names = {}, []
    for key, fname in MOTION_FILES.items():
        plan = importlib.import_module(f"motions.acyclic.{fname}").plan
        for k in ("n_col", "rho"):
            out[f"motion/{key}/{k}"] = np.float64(getattr(plan, k))
        for k in ("dt_arr", "cnt_plan", "X_nom", "X_ter", "bounds", "W_X", "W_X_ter", "W_F", "plan_freq"):
            out[f"motion/{key}/{k}"] = np.asarray(getattr(plan, k), dtype=np.float64)
        T_end = float(np.asarray(plan.cnt_plan)[-1][0][5])
        dt0 = float(plan.dt_arr[0])
        # times into the plan: the start, replanning instants, segment edges, off the grid, past the end of the motion
        edges = sorted({float(s[0][4]) for s in plan.cnt_plan} | {float(s[0][5]) for s in plan.cnt_plan})
        times = [0.0, dt0, 3 * dt0, 0.013, 0.3, 0.5] + edges[1:] + [edges[1] - dt0, T_end - 2 * dt0, T_end + 0.37]
        for ci, t in enumerate(times):
            t0 = 0.0 if ci % 3 else (0.0 if ci == 0 else 0.1)
            t = float(t) + t0
            robot = pin.FakeRobot(2.5, nv=18)
            g = gen.SoloAcyclicGen(robot, "solo12.urdf")
            q0 = np.array([0.2, 0.0, 0.25, 0.0, 0.0, 0.0, 1.0] + 12 * [0.0])
            g.update_motion_params(plan, q0, t0)
            com = np.array([0.2, 0.0, 0.22]) + rng.normal(0, 0.02, 3)
            hg = np.concatenate([2.5 * rng.normal(0, 0.1, 3), rng.normal(0, 0.02, 3)])
            robot.inject(com=com, hg=hg)
            q, v = q0.copy(), np.zeros(18)
            g.create_contact_plan(q, v, t)
            g.create_costs(q, v, t)
            mp = g.mp
            name = f"{key}_{ci:02d}"
            names.append(name)
            vals = dict(x_init=np.concatenate([com, hg[0:3] / 2.5, hg[3:6]]), t=t, t0=t0, mass=2.5,
                        cnt_plan=np.array(mp.cnt_plan), dt=np.array(mp.dt), X_nom=mp.X_nom, X_ter=mp.X_ter, W_X=mp.W_X,
                        W_X_ter=mp.W_X_ter, W_F=mp.W_F, bounds=mp.bounds, rho=mp.rho, f_max=np.array(mp.f_max),
                        plan_freq=g.get_plan_freq(t))
            out[f"{name}/motion"] = key
            for k, v_ in vals.items():
                out[f"{name}/{k}"] = np.asarray(v_, dtype=np.float64)
    out["names"] = np.array(names)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(names), "cases")


def load(path=OUT):
    z = np.load(path)
    cases = []
    for name in z["names"]:
        d = {k: z[f"{name}/{k}"] for k in OUT_KEYS + ("x_init", "t", "t0", "mass", "plan_freq")}
        d["motion"] = str(z[f"{name}/motion"])
        cases.append((str(name), d))
    motions = {}
    for key in z.files:
        if key.startswith("motion/"):
            _, g, k = key.split("/")
            motions.setdefault(g, {})[k] = z[key]
    return cases, motions


if __name__ == "__main__":
    main()
