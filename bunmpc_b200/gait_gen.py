"""Cyclic gait generator shim: the reference's `SoloMpcGaitGen` / `AbstractGaitGen` call shape
(examples/mpc/abstract_cyclic_gen.py:15-157,629-698; upstream BiConMP calls the class CyclicQuadrupedGaitGen) for the
part of it that is the hot path: contact plan + costs + centroidal solve + the 1 kHz interpolation of the plan.

What is NOT here: the whole-body IK (crocoddyl DDP, `KinoDynMP.optimize` beyond `dyn.optimize`) and pinocchio
kinematics -- both out of scope and absent from this image.  `optimize(q, v, ...)` uses the caller's
pinocchio robot wrapper to turn (q, v) into the centroidal problem exactly as the reference does (`centroidal_inputs`;
pinned bit for bit against the reference's own python with the injected-kinematics wrapper of oracle/pinshim,
tests/test_plan_golden.py), solves it on the GPU and returns (None, None, f_int): xs / us are the IK's.  Without
pinocchio it raises a clear error; the entry points that start from the centroidal state (`optimize_centroidal`,
`optimize_centroidal_batch`) need nothing but numpy and the GPU.
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
    def _solve(self, com, vcom, amom, foot_pos, t, v_des, w_des, yaw, amom_des, hip_xy=None):
        batch = build_batch(self.rc, self.params, com, vcom, amom, foot_pos, t, v_des, w_des, yaw=yaw,
                            amom_des=amom_des, horizon=self.horizon, L0=self.L, hip_xy=hip_xy,
                            swing_rule=self.swing_rule)
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

    def optimize_centroidal(self, com, vcom, amom, foot_pos, t, v_des, w_des, yaw=0.0, amom_des=None, hip_xy=None):
        """One replan from the centroidal state.  Returns (com_int, mom_int, f_int) at 1 kHz like :677-692, plus the
        raw solution in self.last.  f_int rows are the 3*n_eff stacked contact forces.  hip_xy [n_eff, 2]: the yaw-rotated
        hip offsets if the caller has them (optimize(q, v, ...) passes the reference's own product)."""
        batch, sol = self._solve(np.atleast_2d(com), vcom, amom, np.asarray(foot_pos)[None], t, np.asarray(v_des)[None],
                                 w_des, yaw, amom_des, None if hip_xy is None else np.asarray(hip_xy)[None])
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

    def centroidal_inputs(self, q, v, v_des, w_des):
        """What abstract_cyclic_gen.py:633-643 + create_cnt_plan (:159-177) + create_costs (:532-560) read from (q, v)
        through pinocchio: the centroidal state, the feet, the yaw, the body-frame velocity goal, the orientation-correction
        momentum and the yaw-rotated hip offsets -- the inputs of the batched problem builder.  q is modified in place
        like the reference does (q[0:2] = 0).  Needs pinocchio and the robot wrapper passed to the constructor."""
        try:
            import pinocchio as pin
        except ImportError as e:
            raise ImportError("CyclicQuadrupedGaitGen.optimize(q, v, ...) needs pinocchio to compute the centroidal "
                              "state; use optimize_centroidal(...) with com / momentum / foot positions") from e
        if self.robot is None:
            raise ValueError("optimize(q, v, ...) needs the pinocchio robot wrapper passed to the constructor")
        rmodel, rdata = self.robot.model, self.robot.data
        q[0:2] = 0                                                  # :633
        ori_des = q[3:7] if w_des != 0 else [0, 0, 0, 1]            # :636-639
        R = pin.Quaternion(np.array(q[3:7])).toRotationMatrix()
        v_des = np.matmul(R, v_des)                                 # :642-643
        pin.forwardKinematics(rmodel, rdata, q, v)                  # :161-166
        pin.updateFramePlacements(rmodel, rdata)
        com = np.array(pin.centerOfMass(rmodel, rdata, q, v))
        pin.computeCentroidalMomentum(rmodel, rdata)                # :536-540
        hg = np.array(rdata.hg)
        foot = np.stack([np.array(rdata.oMf[rmodel.getFrameId(nm)].translation) for nm in self.eff_names])
        yaw = pin.rpy.matrixToRpy(R)[2]                             # :172-177
        R_yaw = pin.rpy.rpyToMatrix(np.array([0.0, 0.0, yaw]))
        hip_xy = np.array([np.matmul(R_yaw, np.asarray(self.rc.hip_offsets)[j])[0:2] for j in range(self.n_eff)])   # :279,347
        # the desired orientation keeps only its yaw (:547-549); the momentum that corrects the difference (:616-627)
        yaw_des = pin.rpy.matrixToRpy(pin.Quaternion(np.array(ori_des, dtype=np.float64)).toRotationMatrix())[2]
        des_quat = pin.Quaternion(pin.rpy.rpyToMatrix(np.array([0.0, 0.0, yaw_des])))
        amom_des = pin.log3((des_quat * (pin.Quaternion(np.array(q[3:7])).inverse())).toRotationMatrix())
        return dict(com=com, vcom=hg[0:3] / self.m, amom=hg[3:6], foot_pos=foot, v_des=v_des, yaw=yaw,
                    amom_des=np.asarray(amom_des, dtype=np.float64), hip_xy=hip_xy)

    def optimize(self, q, v, t, v_des, w_des, X_wm=None, F_wm=None, P_wm=None, noise_std=None, mcts_x_y_cnt_loc=None,
                 v_feet_des=None, ee_pos=None, z_height=None):
        """abstract_cyclic_gen.py:629-698 without the IK: (q, v) -> centroidal problem (centroidal_inputs) -> solve ->
        1 kHz interpolation.  Returns (None, None, f_int): xs_int / us_int come from the reference's IK module, which is
        out of scope; com_int / mom_int / f_int are kept on the object like the reference does."""
        c = self.centroidal_inputs(q, v, v_des, w_des)
        com_int, mom_int, f_int = self.optimize_centroidal(c["com"], c["vcom"], c["amom"], c["foot_pos"], t, c["v_des"],
                                                           w_des, yaw=c["yaw"], amom_des=c["amom_des"], hip_xy=c["hip_xy"])
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
