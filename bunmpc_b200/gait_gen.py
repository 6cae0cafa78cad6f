"""Cyclic gait generator shim: the reference's `SoloMpcGaitGen` / `AbstractGaitGen` call shape
(examples/mpc/abstract_cyclic_gen.py:15-157,629-698; upstream BiConMP calls the class CyclicQuadrupedGaitGen) for the
part of it that is the hot path: contact plan + costs + centroidal solve + the 1 kHz interpolation of the plan.

What is NOT here: the whole-body IK (crocoddyl DDP, `KinoDynMP.optimize` beyond `dyn.optimize`) and pinocchio
kinematics -- both out of scope and absent from this image.  `optimize(q, v, ...)` therefore
needs pinocchio to turn (q, v) into the centroidal state and raises a clear error without it; the entry points that
start from the centroidal state (`optimize_centroidal`, `optimize_centroidal_batch`) are complete and run on the GPU.
"""
from __future__ import annotations

import numpy as np

from .motions import SOLO12, BiconvexMotionParams, RobotConstants
from .plan_builder import SWING_RULE_ABSTRACT, SWING_RULE_SOLO_MPC, build_batch
from .problem import BatchSolution
from .solver import get_solver


class CyclicQuadrupedGaitGen:
    swing_rule = SWING_RULE_SOLO_MPC       # the one contact-plan rule in which the reference's two cyclic generators differ

    def __init__(self, robot=None, r_urdf=None, x_reg=None, planning_time=0.05, q0=None, height_map=None,
                 robot_constants: RobotConstants = SOLO12, device: int = 0):
        self.robot, self.r_urdf, self.x_reg, self.q0 = robot, r_urdf, x_reg, q0
        self.planning_time = planning_time                       # abstract_cyclic_gen.py:35
        self.height_map = height_map
        self.rc = robot_constants
        self.device = device
        self.eff_names = list(robot_constants.eff_names)
        self.n_eff = 4
        self.m = robot_constants.mass
        self.foot_size = robot_constants.foot_size
        self.params = None
        self.L = None                                            # FISTA step sizes carried between replans (Q3)
        self.last = None

    def update_gait_params(self, weight_abstract: BiconvexMotionParams, t, ik_hor_ratio=0.5, horizon=None):
        """abstract_cyclic_gen.py:108-157"""
        self.params = weight_abstract
        self.gait_horizon = self.params.gait_horizon
        self.horizon = int(horizon) if horizon is not None else self.params.horizon()                 # :125
        self.ik_horizon = int(np.round(ik_hor_ratio * self.params.gait_horizon * self.params.gait_period
                                       / self.params.gait_dt, 2))                                      # :128
        self.size = min(self.ik_horizon, int(self.planning_time / self.params.gait_dt) + 2)           # :151-153
        if self.planning_time > self.params.gait_dt:
            self.size -= 1
        self.L = None                                            # a new KinoDynMP is created here in the reference (:133)

    # ---- the hot path, from centroidal states ----
    def _solve(self, com, vcom, amom, foot_pos, t, v_des, w_des, yaw, amom_des):
        batch = build_batch(self.rc, self.params, com, vcom, amom, foot_pos, t, v_des, w_des, yaw=yaw,
                            amom_des=amom_des, horizon=self.horizon, L0=self.L, swing_rule=self.swing_rule)
        sol = get_solver(batch.n_col, batch.n_eff, batch.B, self.device).solve(batch)
        self.L = sol.L.copy()                                    # the FISTA objects keep their L_ across replans
        self.last = (batch, sol)
        return batch, sol

    @staticmethod
    def interpolate(knots: np.ndarray, dt: np.ndarray, size: int) -> np.ndarray:
        """abstract_cyclic_gen.py:677-692: np.linspace between consecutive knots with int(dt[i]/0.001) samples for
        the first `size` knots, stacked.  knots [n_knots, d] -> [sum_i int(dt[i]/0.001), d]."""
        segs = [np.linspace(knots[i], knots[i + 1], int(dt[i] / 0.001)) for i in range(size)]
        return np.vstack(segs)

    def optimize_centroidal(self, com, vcom, amom, foot_pos, t, v_des, w_des, yaw=0.0, amom_des=None):
        """One replan from the centroidal state.  Returns (com_int, mom_int, f_int) at 1 kHz like :677-692, plus the
        raw solution in self.last.  f_int rows are the 3*n_eff stacked contact forces."""
        batch, sol = self._solve(np.atleast_2d(com), vcom, amom, np.asarray(foot_pos)[None], t, np.asarray(v_des)[None],
                                 w_des, yaw, amom_des)
        n, e = batch.n_col, batch.n_eff
        dt = batch.dt[0]
        F = sol.F[0].reshape(n, 3 * e)
        com_opt, mom_opt = sol.com()[0], sol.mom()[0]
        return (self.interpolate(com_opt, dt, self.size), self.interpolate(mom_opt, dt, self.size),
                self.interpolate(F, dt, self.size))

    def optimize_centroidal_batch(self, com, vcom, amom, foot_pos, t, v_des, w_des, yaw=0.0, amom_des=None) -> BatchSolution:
        """B independent replans in one launch (perturbed restarts of data_collection.py:181-277, goal samples of
        locosafedagger_modified.py:449-614).  Returns the BatchSolution; use interpolate() per instance as needed."""
        _, sol = self._solve(com, vcom, amom, foot_pos, t, v_des, w_des, yaw, amom_des)
        return sol

    def optimize(self, q, v, t, v_des, w_des, X_wm=None, F_wm=None, P_wm=None, noise_std=None, mcts_x_y_cnt_loc=None,
                 v_feet_des=None, ee_pos=None, z_height=None):
        """abstract_cyclic_gen.py:629-698.  Needs pinocchio (centroidal state and foot kinematics from q, v) and the
        reference's IK module for xs/us; neither is part of the hot path nor available in this image."""
        try:
            import pinocchio as pin                                # noqa: F401
        except ImportError as e:
            raise ImportError("CyclicQuadrupedGaitGen.optimize(q, v, ...) needs pinocchio to compute the centroidal "
                              "state; use optimize_centroidal(...) with com / momentum / foot positions") from e
        if self.robot is None:
            raise ValueError("optimize(q, v, ...) needs the pinocchio robot wrapper passed to the constructor")
        rmodel, rdata = self.robot.model, self.robot.data
        q = np.asarray(q, dtype=np.float64).copy()
        q[0:2] = 0                                                  # :633
        R = pin.Quaternion(np.array(q[3:7])).toRotationMatrix()
        v_des = np.matmul(R, v_des)                                 # :642-643
        pin.forwardKinematics(rmodel, rdata, q, v)
        pin.updateFramePlacements(rmodel, rdata)
        com = pin.centerOfMass(rmodel, rdata, q, v)
        pin.computeCentroidalMomentum(rmodel, rdata)
        hg = np.array(rdata.hg)
        foot = np.stack([rdata.oMf[rmodel.getFrameId(nm)].translation for nm in self.eff_names])
        yaw = pin.rpy.matrixToRpy(R)[2]
        com_int, mom_int, f_int = self.optimize_centroidal(com, hg[0:3] / self.m, hg[3:6], foot, t, v_des, w_des, yaw=yaw)
        self.com_int, self.mom_int, self.f_int = com_int, mom_int, f_int
        return None, None, f_int                                    # xs_int, us_int come from the IK (out of scope)


# the names the reference and upstream BiConMP use for this class
SoloMpcGaitGen = CyclicQuadrupedGaitGen


class AbstractGaitGen(CyclicQuadrupedGaitGen):
    """examples/mpc/abstract_cyclic_gen1.py: the robot-agnostic variant (end-effector and hip names are constructor
    arguments, hip offsets = round(foot - com, 3) without the +-0.04 y shift -- pass them in `robot_constants`).  Its
    contact plan leaves the Raibert step out of the second half of a swing (:211-215); costs, bounds, horizon, dt rule and
    interpolation are those of SoloMpcGaitGen (tests/golden/plan_cases.npz holds its outputs next to SoloMpcGaitGen's).
    `update_gait_params` has no `horizon` override there (:97)."""
    swing_rule = SWING_RULE_ABSTRACT
