"""Generates tests/golden/plan_cases.npz: the contact plans and cost references that the REFERENCE'S OWN python
produces -- examples/mpc/abstract_cyclic_gen.py (`SoloMpcGaitGen.create_cnt_plan` :159-414, `create_costs` :532-614)
imported in place from /root/reference, with its gait parameters from the reference's own motion files
(examples/motions/cyclic/solo12_{trot,bound,jump}.py) and its phase lookup from the reference's own gait_planner.cpp
(oracle/_ref/libgait_ref.so).  pinocchio, the pybind modules and matplotlib are the stand-ins of oracle/pinshim
(kinematic quantities injected, see its README).

Run in the build container:   make -C oracle/refshim && python tests/golden/make_plan_golden.py

Stored per case: what bunmpc_b200.plan_builder / the CUDA build_problem_kernel take as inputs (centroidal state, foot
positions, time in the gait, desired velocities, yaw, orientation-correction momentum, hip offsets, yaw inertia) and
what the reference handed to its solver (cnt_plan, dt, X_nom, X_ter, W_X, W_X_ter, W_F, bounds, rho).

Every case is also run through the reference's OTHER cyclic generator, `AbstractGaitGen` of
examples/mpc/abstract_cyclic_gen1.py (create_cnt_plan :136-237, create_costs :239-322), on the same state, with the hip
offsets that class derives itself (round(foot - com, 3) at q0) and its own default horizon: keys `<case>/gen1/...`."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/iterative_supervised_learning/examples"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "plan_cases.npz")
IN_KEYS = ("mass", "com", "vcom", "amom", "foot_pos", "t", "v_des", "w_des", "yaw", "amom_des", "hip_offsets", "hip_xy",
           "I_zz", "horizon")
OUT_KEYS = ("cnt_plan", "dt", "X_nom", "X_ter", "W_X", "W_X_ter", "W_F", "bounds", "rho", "x_init")


def quat_rpy(r, p, y):
    """(x, y, z, w) of Rz(y) Ry(p) Rx(r)"""
    cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy])


def main():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "pinshim"), os.path.join(REF, "mpc"), REF, ROOT]
    import pinocchio as pin                                     # the stand-in
    gen = importlib.import_module("abstract_cyclic_gen")       # the reference's own files
    gen1 = importlib.import_module("abstract_cyclic_gen1")
    motions = {g: getattr(importlib.import_module(f"motions.cyclic.solo12_{g}"), g) for g in ("trot", "bound", "jump")}
    rng = np.random.default_rng(0)
    out, names = {}, []
    hip0 = np.array([[0.1946, 0.14695, 0.0], [0.1946, -0.14695, 0.0], [-0.1946, 0.14695, 0.0], [-0.1946, -0.14695, 0.0]])
    for ci in range(36):
        gait = ("trot", "bound", "jump")[ci % 3]
        prm = motions[gait]
        robot = pin.FakeRobot(2.5 if ci % 5 else 15.099, nv=18)
        # ---- the "robot": nominal configuration q0 (level base), then the state the plan is made for ----
        q0 = np.zeros(19); q0[2] = 0.25; q0[6] = 1.0
        com0 = np.array([0.0, 0.0, 0.2]) + rng.normal(0, 0.003, 3)
        I_comp = np.diag([0.03, 0.06, 0.0885]) + rng.normal(0, 1e-3, (3, 3))
        hip_pos0 = hip0 + np.array([0, 0, 0.2]) + rng.normal(0, 1e-3, (4, 3))
        robot.inject(com=com0, hip_pos=hip_pos0, I_composite=I_comp)
        gg = gen.SoloMpcGaitGen(robot, "solo12.urdf", np.zeros(37), 0.05, q0)
        # time in the gait: on the planning grid, on phase edges, off the grid, several periods ahead
        t = float([0.0, 0.05, 0.3, 0.15, 0.25, 0.1, 0.013, 0.262, 1.2, 0.5, 0.45, 2.05][ci % 12])
        horizon = None if ci % 4 else int(prm.gait_horizon * prm.gait_period / prm.gait_dt) + 3
        gg.update_gait_params(prm, t, horizon=horizon)
        turning = ci % 3 == 1 or ci % 7 == 0
        w_des = float(rng.uniform(-0.3, 0.3)) if turning else 0.0
        v_in = np.array([rng.uniform(-0.1, 0.4), rng.uniform(-0.15, 0.15), 0.0])
        rpy = np.array([rng.normal(0, 0.05), rng.normal(0, 0.05), rng.uniform(-np.pi, np.pi) if ci % 2 else rng.normal(0, 0.2)])
        q = np.zeros(19); q[0:3] = [rng.normal(0, 0.3), rng.normal(0, 0.3), 0.22]; q[3:7] = quat_rpy(*rpy)
        v = rng.normal(0, 0.2, 18)
        com = np.array([rng.normal(0, 0.02), rng.normal(0, 0.02), 0.2 + rng.normal(0, 0.02)])
        hg = np.concatenate([robot.model.mass * rng.normal(0, 0.1, 3), rng.normal(0, 0.02, 3)])
        feet = hip0 * [1, 1, 0] + [0, 0, 0.018] + rng.normal(0, 0.02, (4, 3)) * [1, 1, 0.1]
        robot.inject(com=com, foot_pos=feet, hg=hg)
        # ---- SoloMpcGaitGen.optimize up to the solver call, abstract_cyclic_gen.py:633-659 ----
        q[0:2] = 0
        ori_des = q[3:7] if w_des != 0 else [0, 0, 0, 1]
        R = pin.Quaternion(np.array(q[3:7])).toRotationMatrix()
        v_des = np.matmul(R, v_in)
        gg.create_cnt_plan(q, v, t, v_des, w_des)
        gg.create_costs(q, v, v_des, w_des, ori_des)
        mp = gg.mp
        # ---- the inputs of our builder, derived the way the reference derives them ----
        rpyv = pin.rpy.matrixToRpy(R)
        des_quat = pin.Quaternion(pin.rpy.rpyToMatrix(np.array([0.0, 0.0, pin.rpy.matrixToRpy(
            pin.Quaternion(np.array(ori_des)).toRotationMatrix())[2]])))
        amom_des = gg.compute_ori_correction(q, des_quat.coeffs())
        R_yaw = pin.rpy.rpyToMatrix(np.array([0.0, 0.0, rpyv[2]]))                       # :172-177
        hip_xy = np.array([np.matmul(R_yaw, gg.offsets[j])[0:2] for j in range(4)])     # :279,347
        name = f"{ci:02d}_{gait}"
        names.append(name)
        vals = dict(mass=robot.model.mass, com=com, vcom=hg[0:3] / robot.model.mass, amom=hg[3:6], foot_pos=feet, t=t,
                    v_des=v_des, w_des=w_des, yaw=rpyv[2], amom_des=amom_des, hip_offsets=gg.offsets.copy(), hip_xy=hip_xy,
                    I_zz=I_comp[2, 2], horizon=gg.horizon,
                    cnt_plan=np.array(mp.cnt_plan), dt=np.array(mp.dt), X_nom=mp.X_nom, X_ter=mp.X_ter, W_X=mp.W_X,
                    W_X_ter=mp.W_X_ter, W_F=mp.W_F, bounds=mp.bounds, rho=mp.rho, x_init=gg.X_init.copy())
        out[f"{name}/gait"] = gait
        # what SoloMpcGaitGen.optimize itself is called with (no further random draws): tests run the shim's
        # optimize(q, v, t, v_in, w_des) on a robot wrapper with the same injected kinematics
        vals.update(q=q.copy(), qv=v.copy(), v_in=v_in, hg=hg)
        for k, v_ in vals.items():
            out[f"{name}/{k}"] = np.asarray(v_, dtype=np.float64)
        # ---- the same state through AbstractGaitGen (abstract_cyclic_gen1.py); no further random draws ----
        robot1 = pin.FakeRobot(robot.model.mass, nv=18)
        robot1.inject(com=com0, hip_pos=hip_pos0, foot_pos=hip0 * [1, 1, 0] + [0, 0, 0.018], I_composite=I_comp)
        pin.register_urdf("solo12.urdf", robot1)
        g1 = gen1.AbstractGaitGen("solo12.urdf", list(pin.FakeRobot.FRAMES[0:4]), list(pin.FakeRobot.FRAMES[4:8]),
                                  np.zeros(37), 0.05, q0)
        g1.update_gait_params(prm, t)
        robot1.inject(com=com, foot_pos=feet, hg=hg)
        g1.create_cnt_plan(q, v, t, v_des, w_des)
        g1.create_costs(q, v, v_des, w_des, ori_des)
        mp1 = g1.mp
        vals1 = dict(hip_offsets=g1.offsets.copy(), hip_xy=np.array([np.matmul(R_yaw, g1.offsets[j])[0:2] for j in range(4)]),
                     horizon=g1.horizon, cnt_plan=np.array(mp1.cnt_plan), dt=np.array(mp1.dt), X_nom=mp1.X_nom, X_ter=mp1.X_ter,
                     W_X=mp1.W_X, W_X_ter=mp1.W_X_ter, W_F=mp1.W_F, bounds=mp1.bounds, rho=mp1.rho, x_init=g1.X_init.copy())
        for k, v_ in vals1.items():
            out[f"{name}/gen1/{k}"] = np.asarray(v_, dtype=np.float64)
        # the reference's own gait record, to pin bunmpc_b200/motions.py against it
        for k in ("gait_period", "gait_dt", "gait_horizon", "nom_ht", "rho", "step_ht"):
            out[f"motion/{gait}/{k}"] = np.float64(getattr(prm, k))
        for k in ("stance_percent", "phase_offset", "W_X", "W_X_ter", "W_F", "ori_correction"):
            out[f"motion/{gait}/{k}"] = np.asarray(getattr(prm, k), dtype=np.float64)
    out["names"] = np.array(names)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(names), "cases")


def load(path=OUT):
    z = np.load(path)
    cases = []
    for name in z["names"]:
        d = {k: z[f"{name}/{k}"] for k in IN_KEYS + OUT_KEYS}
        d["gait"] = str(z[f"{name}/gait"])
        d.update({k: z[f"{name}/{k}"] for k in ("q", "qv", "v_in", "hg")})
        d["gen1"] = {k: z[f"{name}/gen1/{k}"] for k in OUT_KEYS + ("hip_offsets", "hip_xy", "horizon")}
        cases.append((str(name), d))
    motions = {}
    for key in z.files:
        if key.startswith("motion/"):
            _, g, k = key.split("/")
            motions.setdefault(g, {})[k] = z[key]
    return cases, motions


if __name__ == "__main__":
    main()
