"""Batched host-side builder of centroidal MPC problems (contact plan, nominal trajectory, bounds).

Vectorised (numpy, leading batch dimension) restatement of what `SoloMpcGaitGen` does before it calls
the solver: `create_cnt_plan` (examples/mpc/abstract_cyclic_gen.py:159-414) and the dynamics part of
`create_costs` (:564-614).  Pinocchio is not available, so the builder starts from the centroidal state
(com, vcom, angular momentum), the current foot positions and a yaw angle instead of (q, v); everything
after that point follows the reference rule by rule.
"""
from __future__ import annotations

import numpy as np

from .gait_planner import GaitPlanner
from .motions import BiconvexMotionParams, RobotConstants
from .problem import CentroidalBatch, L0_F, L0_X

GRAVITY = 9.81   # abstract_cyclic_gen.py:49

# The reference has two cyclic generators that differ in ONE rule of the contact plan (found by running both through
# oracle/pinshim on the same inputs): where a foot is planned to be in the second half of its swing.
SWING_RULE_SOLO_MPC = 0      # SoloMpcGaitGen, abstract_cyclic_gen.py:351-355: hip + yaw step + Raibert step
SWING_RULE_ABSTRACT = 1      # AbstractGaitGen, abstract_cyclic_gen1.py:211-215: hip + yaw step (the Raibert step it
                             # computes there is not added)


def _b(a, B, tail):
    a = np.asarray(a, dtype=np.float64)
    return np.broadcast_to(a, (B,) + tuple(tail)).copy()


def rotated_hip_offsets(robot: RobotConstants, yaw):
    """(R hip_offset_j)[0:2] of abstract_cyclic_gen.py:279,347 with R = rpyToMatrix(0, 0, yaw), formed with numpy.matmul
    exactly as the reference forms it (its rounding is the BLAS's, typically fused); [B,e,2]."""
    yaw = np.atleast_1d(np.asarray(yaw, dtype=np.float64))
    out = np.zeros((yaw.shape[0], len(robot.eff_names), 2))
    for b, y in enumerate(yaw):
        cy, sy = np.cos(y), np.sin(y)
        R = np.array([[cy, -sy, 0.0], [sy, cy, 0.0], [0.0, 0.0, 1.0]])
        for j in range(out.shape[1]):
            out[b, j] = np.matmul(R, np.asarray(robot.hip_offsets[j], dtype=np.float64))[0:2]
    return out


def build_contact_plan(robot: RobotConstants, params: BiconvexMotionParams, com, foot_pos, t, v_des, w_des,
                       yaw=0.0, horizon=None, hip_xy=None, swing_rule=SWING_RULE_SOLO_MPC):
    """create_cnt_plan, abstract_cyclic_gen.py:159-414 (height_map=None, noise_std=None, mcts=None).

    com [B,3], foot_pos [B,e,3] (current end-effector positions; rounded to 3 dp as :211/:240 unless they
    are given as ee_pos), t [B], v_des [B,3] (already in the local frame, :642-643), w_des [B], yaw [B].
    hip_xy [B,e,2]: the rotated hip offsets (R hip_offset_j)[0:2]; None = cos/sin(yaw) products, unfused (what the CUDA
    builder computes when it is not handed the products either).
    Returns cnt_plan [B,n,e,4], dt [B,n]."""
    com = np.atleast_2d(np.asarray(com, dtype=np.float64))
    B = com.shape[0]
    e = len(robot.eff_names)
    n = params.horizon() if horizon is None else int(horizon)
    foot_pos = _b(foot_pos, B, (e, 3))
    t = _b(t, B, ())
    v_des = _b(v_des, B, (3,))
    w_des = _b(w_des, B, ())
    yaw = _b(yaw, B, ())
    gp = GaitPlanner(params.gait_period, np.array(params.stance_percent), np.array(params.phase_offset),
                     params.step_ht)
    gait_dt = params.gait_dt

    com_xy = np.round(com[:, 0:2], 3)                                  # :164
    z_height = com[:, 2]                                               # :165
    cy, sy = np.cos(yaw), np.sin(yaw)                                  # R = rpyToMatrix(0, 0, yaw), :172-177
    vtrack = v_des[:, 0:2]                                             # :179
    # ang_step = cross(0.5*sqrt(z/g)*vtrack, [0, 0, w_des])[0:2], :286-287
    a = 0.5 * np.sqrt(z_height / GRAVITY)[:, None] * vtrack
    ang_step = np.stack([a[:, 1] * w_des, -a[:, 0] * w_des], axis=1)

    cnt_plan = np.zeros((B, n, e, 4))
    dt = np.zeros((B, n))
    for i in range(n):
        for j in range(e):
            if i == 0:
                cnt_plan[:, 0, j, 0] = gp.get_phase(t, j)              # :206-208,236
                cnt_plan[:, 0, j, 1:4] = np.round(foot_pos[:, j], 3)   # :213,240
                continue
            ft = np.round(t + i * gait_dt, 3)                          # :260
            stance = gp.get_phase(ft, j) == 1                          # :263
            prev_stance = cnt_plan[:, i - 1, j, 0] == 1                # :269
            off = robot.hip_offsets[j]
            rot_off = (np.stack([cy * off[0] - sy * off[1], sy * off[0] + cy * off[1]], axis=1) if hip_xy is None
                       else _b(hip_xy, B, (e, 2))[:, j])
            hip_loc = com_xy + rot_off + i * gait_dt * vtrack          # :279,347
            raibert = 0.5 * vtrack * params.gait_period * params.stance_percent[j] \
                - 0.05 * (vtrack - v_des[:, 0:2])                      # :282
            per_ph = np.round(gp.get_percent_in_phase(ft, j), 3)       # :346
            touchdown_xy = raibert + hip_loc + ang_step                # :289
            swing_xy = np.where((per_ph < 0.5)[:, None] | (swing_rule == SWING_RULE_ABSTRACT), hip_loc + ang_step,
                                hip_loc + ang_step + raibert)          # :351-355; abstract_cyclic_gen1.py:211-215
            xy = np.where(stance[:, None],
                          np.where(prev_stance[:, None], cnt_plan[:, i - 1, j, 1:3], touchdown_xy), swing_xy)
            z = np.where(stance & prev_stance, cnt_plan[:, i - 1, j, 3], robot.foot_size)   # :271,337,374
            cnt_plan[:, i, j, 0] = stance
            cnt_plan[:, i, j, 1:3] = xy
            cnt_plan[:, i, j, 3] = z
        if i == 0:
            d0 = gait_dt - np.round(np.remainder(t, gait_dt), 2)       # :385-388
            dt[:, 0] = np.where(d0 == 0, gait_dt, d0)
        else:
            dt[:, i] = gait_dt
    return cnt_plan, dt


def build_costs(robot: RobotConstants, params: BiconvexMotionParams, x_init, dt, v_des, w_des, amom_des=None):
    """Dynamics part of create_costs, abstract_cyclic_gen.py:564-614.
    x_init [B,9] = [com, hg_lin/m, hg_ang] (:567-571); amom_des [B,3] = log3(R_des R_q^T) (:616-627),
    zero when the base is level.  Returns W_X, W_X_ter, X_nom, X_ter, W_F, bounds."""
    x_init = np.atleast_2d(np.asarray(x_init, dtype=np.float64))
    B = x_init.shape[0]
    n = dt.shape[1]
    v_des = _b(v_des, B, (3,))
    w_des = _b(w_des, B, ())
    amom = _b(0.0 if amom_des is None else amom_des, B, (3,))
    X_nom = np.zeros((B, n, 9))
    X_nom[:, :, 0] = x_init[:, 0:1]                                    # :573
    for i in range(1, n):                                              # :574-576 (y of knot 0 stays 0)
        X_nom[:, i, 0] = X_nom[:, i - 1, 0] + v_des[:, 0] * dt[:, i]
        X_nom[:, i, 1] = X_nom[:, i - 1, 1] + v_des[:, 1] * dt[:, i]
    X_nom[:, :, 2] = params.nom_ht                                     # :578
    X_nom[:, :, 3:6] = v_des[:, None, :]                               # :579-581
    X_ter = np.zeros((B, 9))
    X_ter[:, 0:2] = x_init[:, 0:2] + (params.gait_horizon * params.gait_period * v_des)[:, 0:2]   # :593
    X_ter[:, 2] = params.nom_ht
    X_ter[:, 3:6] = v_des
    X_ter[:, 6:9] = amom
    X_nom[:, :, 6] = (amom[:, 0] * params.ori_correction[0])[:, None]  # :598-599
    X_nom[:, :, 7] = (amom[:, 1] * params.ori_correction[1])[:, None]
    yaw_momentum = robot.I_zz * w_des                                  # :604 (I_composite_b @ [0,0,w])[2]
    turning = w_des != 0
    X_nom[:, :, 8] = np.where(turning, yaw_momentum, amom[:, 2] * params.ori_correction[2])[:, None]
    X_ter[:, 8] = np.where(turning, yaw_momentum, X_ter[:, 8])
    bounds = np.tile([-robot.bx, -robot.by, 0, robot.bx, robot.by, robot.bz], (n, 1))[None]    # :611
    W_X = np.tile(params.W_X, n)[None]                                 # :613
    W_X_ter = np.asarray(params.W_X_ter, dtype=np.float64)[None]
    W_F = np.tile(params.W_F, n)[None]                                 # :614
    return W_X, W_X_ter, X_nom.reshape(B, 9 * n), X_ter, W_F, bounds


def build_batch(robot: RobotConstants, params: BiconvexMotionParams, com, vcom, amom, foot_pos, t, v_des, w_des,
                yaw=0.0, amom_des=None, horizon=None, L0=None, scale_W_X=None, scale_W_F=None,
                scale_rho=None, hip_xy=None, swing_rule=SWING_RULE_SOLO_MPC) -> CentroidalBatch:
    """One CentroidalBatch from centroidal states: contact plan + costs + bounds.  The optional per-instance
    scalings multiply W_X / W_X_ter, W_F and rho (BASELINE config 5's cost-weight samples)."""
    com = np.atleast_2d(np.asarray(com, dtype=np.float64))
    B = com.shape[0]
    x_init = np.concatenate([com, _b(vcom, B, (3,)), _b(amom, B, (3,))], axis=1)
    v_des = _b(v_des, B, (3,))
    cnt_plan, dt = build_contact_plan(robot, params, com, foot_pos, t, v_des, w_des, yaw=yaw, horizon=horizon,
                                      hip_xy=hip_xy, swing_rule=swing_rule)
    W_X, W_X_ter, X_nom, X_ter, W_F, bounds = build_costs(robot, params, x_init, dt, v_des, w_des, amom_des)
    rho = np.array([params.rho], dtype=np.float64)
    if scale_W_X is not None:
        s = _b(scale_W_X, B, ())[:, None]
        W_X, W_X_ter = W_X * s, W_X_ter * s
    if scale_W_F is not None:
        W_F = W_F * _b(scale_W_F, B, ())[:, None]
    if scale_rho is not None:
        rho = params.rho * _b(scale_rho, B, ())
    return CentroidalBatch(
        n_col=dt.shape[1], n_eff=len(robot.eff_names), m=np.array([robot.mass]), rho=rho, x_init=x_init,
        cnt_plan=cnt_plan, dt=dt, W_X=W_X, W_X_ter=W_X_ter, X_nom=X_nom, X_ter=X_ter, W_F=W_F, bounds=bounds,
        L0=np.array([[L0_F, L0_X]]) if L0 is None else L0)
