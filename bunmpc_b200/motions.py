"""Gait parameter records and robot constants used to build centroidal MPC problems.

Parameter values are those of the reference's motion files (examples/motions/cyclic/solo12_trot.py:16-40,
solo12_bound.py:16-40, solo12_jump.py:17-41); only the fields the centroidal solve reads are kept
(the IK weights belong to the whole-body IK, which is out of scope).  The record type mirrors
examples/motions/weight_abstract.py:7-43.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np


@dataclass
class BiconvexMotionParams:
    robot_name: str
    motion_name: str
    gait_period: float = None
    stance_percent: list = None
    gait_dt: float = 0.05
    phase_offset: list = None
    step_ht: float = None
    W_X: np.ndarray = None
    W_X_ter: np.ndarray = None
    W_F: np.ndarray = None
    nom_ht: float = None
    rho: float = None
    ori_correction: list = None
    gait_horizon: float = None

    def horizon(self) -> int:
        """abstract_cyclic_gen.py:125"""
        return int(np.round(self.gait_horizon * self.gait_period / self.gait_dt, 2))

    def scaled(self, gait_horizon_scale: float) -> "BiconvexMotionParams":
        """Longer horizons as in examples/analysis/solve_times_test.py:60-66 (gait_horizon sweep)."""
        return replace(self, gait_horizon=self.gait_horizon * gait_horizon_scale)


@dataclass
class RobotConstants:
    """What the centroidal problem needs from the robot description (no pinocchio here)."""
    name: str
    mass: float
    foot_pos: np.ndarray          # [4,3] nominal foot positions (FL, FR, HL, HR) relative to the base xy, z on ground
    hip_offsets: np.ndarray       # [4,3] round(hip - com, 3) with the +-0.04 y shift, abstract_cyclic_gen.py:51-69
    foot_size: float = 0.018      # abstract_cyclic_gen.py:31
    I_zz: float = 0.0             # composite inertia about yaw, used only when w_des != 0.  AN INPUT: the reference reads it
                                  # from pinocchio (crba, abstract_cyclic_gen.py:46-47); the defaults below are assumed values
    bx: float = 0.45              # abstract_cyclic_gen.py:92-97
    by: float = 0.45
    bz: float = 0.45
    f_max: tuple = (15.0, 15.0, 15.0)
    eff_names: tuple = ("FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT")


def _quad(x, y, z):
    return np.array([[x, y, z], [x, -y, z], [-x, y, z], [-x, -y, z]], dtype=np.float64)


# Solo12: mass = sum of <mass> in robots/solo12/urdf/solo12.urdf (2.5 kg); feet below the joint chain
# HAA(0.1946, 0.0875) + HFE(0.014) + KFE(0.03745) + ANKLE(0.008), solo12.urdf:49,89,134,176.
SOLO12 = RobotConstants(
    name="solo12", mass=2.5,
    foot_pos=_quad(0.1946, 0.14695, 0.018),
    hip_offsets=_quad(0.195, 0.102 + 0.04, 0.0),
    I_zz=0.0885,
)

# Go2: constants from robot_properties_go2 (xacro/const.xacro:21-32,70-119; config.py:162-165).
# The reference never wires Go2 to the MPC; Solo12 gait timing/weights are reused.
GO2 = RobotConstants(
    name="go2", mass=0.001 + 6.921 + 0.001 + 4 * (0.678 + 1.152 + 0.154 + 0.06),
    foot_pos=_quad(0.1934, 0.142, 0.02),
    hip_offsets=_quad(0.193, 0.142, 0.0),
    foot_size=0.02,
    I_zz=0.25,
)

solo12_trot = BiconvexMotionParams(
    "solo12", "Trot", gait_period=0.5, stance_percent=[0.6] * 4, gait_dt=0.05,
    phase_offset=[0.0, 0.5, 0.5, 0.0], step_ht=0.075, nom_ht=0.2,
    W_X=np.array([1e-5, 1e-5, 1e+5, 1e+1, 1e+1, 2e+2, 1e+4, 1e+4, 1e4]),
    W_X_ter=10 * np.array([1e+5, 1e-5, 1e+5, 1e+1, 1e+1, 2e+2, 1e+5, 1e+5, 1e+5]),
    W_F=np.array(4 * [1e+1, 1e+1, 1e+1]), rho=5e+4, ori_correction=[0.3, 0.5, 0.4], gait_horizon=2.0)

solo12_bound = BiconvexMotionParams(
    "solo12", "Bound", gait_period=0.3, stance_percent=[0.5] * 4, gait_dt=0.05,
    phase_offset=[0.0, 0.0, 0.5, 0.5], step_ht=0.07, nom_ht=0.25,
    W_X=np.array([1e-5, 1e-5, 5e+4, 1e1, 1e1, 1e+3, 5e+3, 1e+4, 5e+3]),
    W_X_ter=10 * np.array([1e-5, 1e-5, 5e+4, 1e1, 1e1, 1e+3, 1e+4, 1e+4, 1e+4]),
    W_F=np.array(4 * [1e1, 1e+1, 1.5e+1]), rho=5e+4, ori_correction=[0.2, 0.8, 0.8], gait_horizon=4.0)

solo12_jump = BiconvexMotionParams(
    "solo12", "Jump", gait_period=0.5, stance_percent=[0.3] * 4, gait_dt=0.05,
    phase_offset=[0.7] * 4, step_ht=0.05, nom_ht=0.25,
    W_X=np.array([1e-5, 1e-5, 1e+5, 1e+1, 1e+1, 2e+2, 1e+4, 1e+4, 1e4]),
    W_X_ter=10 * np.array([1e+5, 1e-5, 1e+5, 1e+1, 1e+1, 2e+2, 1e+5, 1e+5, 1e+5]),
    W_F=np.array(4 * [1e+1, 1e+1, 1.5e+1]), rho=5e+4, ori_correction=[0.2, 0.5, 0.4], gait_horizon=3.0)

# Go2 runs the Solo12 records with its own nominal height.
go2_trot = replace(solo12_trot, robot_name="go2", nom_ht=0.30)
go2_bound = replace(solo12_bound, robot_name="go2", nom_ht=0.30)
go2_jump = replace(solo12_jump, robot_name="go2", nom_ht=0.30)

# NON-REFERENCE retuned Go2 set (SURVEY 8(d) config 3 allows one, "clearly labelled non-reference").  With the Solo12
# records the reference algorithm itself does not converge for Go2's mass (the cone step of fista.cpp:59-67
# compares a SQUARED tangential force with mu * f_z and diverges once forces exceed a few newtons; the penalty rho is
# too weak against the cost weights for a 6x heavier robot; the exit test is absolute).  The retune keeps every rule of
# the algorithm and changes three numbers, found with the CPU oracle: cost weights W_X, W_X_ter, W_F x 0.01 (rho / W
# x 100), friction coefficient mu = 200 (the squared-norm cone then only cuts in above ~90 N of tangential force per
# foot), exit tolerance 6e-3 = 1e-3 x the mass ratio.  Trot: 97 % of perturbed instances converge (38 outer iterations
# on average), bound 88 %; the jump gait still does not.
GO2_RETUNE_WEIGHT_SCALE = 0.01
GO2_RETUNE_SOLVER = dict(mu=200.0, exit_tol=6e-3)


def _retuned(p):
    return replace(p, robot_name="go2_retuned", W_X=p.W_X * GO2_RETUNE_WEIGHT_SCALE,
                   W_X_ter=p.W_X_ter * GO2_RETUNE_WEIGHT_SCALE, W_F=p.W_F * GO2_RETUNE_WEIGHT_SCALE)


GAITS = {"solo12": {"trot": solo12_trot, "bound": solo12_bound, "jump": solo12_jump},
         "go2": {"trot": go2_trot, "bound": go2_bound, "jump": go2_jump},
         "go2_retuned": {"trot": _retuned(go2_trot), "bound": _retuned(go2_bound), "jump": _retuned(go2_jump)}}
ROBOTS = {"solo12": SOLO12, "go2": GO2, "go2_retuned": GO2}
