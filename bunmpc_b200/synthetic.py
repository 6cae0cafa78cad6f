"""Synthetic batches for the BASELINE.json configurations.

The reference perturbs generalized coordinates and maps them through pinocchio
(iterative_algorithm/data_collection.py:232-252, cfgs/data_collection_config.yaml:19-50); pinocchio is not
available, so the perturbation is applied at the centroidal level: CoM, CoM velocity, angular momentum,
stance-foot positions, replanning phase (i_replan * plan_freq, data_collection.py:185) and desired
velocity (data_collection_config.yaml:11-16).  Everything downstream is the reference's rule set
(plan_builder.py).
"""
from __future__ import annotations

import numpy as np

from .motions import GAITS, ROBOTS
from .plan_builder import build_batch
from .problem import CentroidalBatch


def nominal(robot="solo12", gait="trot", v_des=(0.2, 0.0, 0.0), w_des=0.0, t=0.0, horizon_scale=1.0) -> CentroidalBatch:
    """BASELINE config 1: one solve from the nominal standing state (test_mpc.py:53-55 goal)."""
    rb, gp = ROBOTS[robot], GAITS[robot][gait].scaled(horizon_scale)
    com = np.array([[0.0, 0.0, gp.nom_ht]])
    return build_batch(rb, gp, com, np.zeros(3), np.zeros(3), rb.foot_pos[None], t, np.asarray(v_des), w_des)


def perturbed(B, robot="solo12", gait="trot", seed=0, horizon_scale=1.0, vx_range=(0.0, 0.3),
              vy_range=(0.0, 0.0), w_range=(0.0, 0.0), sigma_com=0.02, sigma_vcom=0.1, sigma_amom=0.02,
              sigma_foot=0.02, weight_scale_range=None) -> CentroidalBatch:
    """BASELINE configs 2-5: B perturbed initial states / goals, numpy default_rng(seed).

    weight_scale_range=(lo, hi): additionally draw log-uniform scalings of W_X, W_F and rho per instance
    (config 5's cost-weight samples)."""
    rb, gp = ROBOTS[robot], GAITS[robot][gait].scaled(horizon_scale)
    rng = np.random.default_rng(seed)
    com = np.array([0.0, 0.0, gp.nom_ht]) + rng.normal(0.0, sigma_com, (B, 3))
    vcom = rng.normal(0.0, sigma_vcom, (B, 3))
    amom = rng.normal(0.0, sigma_amom, (B, 3))
    foot = np.broadcast_to(rb.foot_pos, (B, 4, 3)).copy()
    foot[:, :, 0:2] += rng.normal(0.0, sigma_foot, (B, 4, 2))
    n_phase = int(round(gp.gait_period / gp.gait_dt))
    t0 = rng.integers(0, n_phase, B) * gp.gait_dt                       # i_replan * plan_freq
    v_des = np.zeros((B, 3))
    v_des[:, 0] = rng.uniform(vx_range[0], vx_range[1], B)
    v_des[:, 1] = rng.uniform(vy_range[0], vy_range[1], B)
    w_des = rng.uniform(w_range[0], w_range[1], B)
    kw = {}
    if weight_scale_range is not None:
        lo, hi = np.log(weight_scale_range[0]), np.log(weight_scale_range[1])
        kw = dict(scale_W_X=np.exp(rng.uniform(lo, hi, B)), scale_W_F=np.exp(rng.uniform(lo, hi, B)),
                  scale_rho=np.exp(rng.uniform(lo, hi, B)))
    return build_batch(rb, gp, com, vcom, amom, foot, t0, v_des, w_des, **kw)


def config(idx: int, B=None, seed=0) -> CentroidalBatch:
    """The five BASELINE.json configs by index (0-based)."""
    if idx == 0:
        return nominal("solo12", "trot")
    if idx == 1:
        return perturbed(1024 if B is None else B, "solo12", "trot", seed=seed)
    if idx == 2:
        return perturbed(16384 if B is None else B, "go2", "trot", seed=seed)
    if idx == 3:
        return perturbed(4096 if B is None else B, "go2", "bound", seed=seed, horizon_scale=2.0)
    if idx == 4:
        return perturbed(65536 if B is None else B, "solo12", "trot", seed=seed, vx_range=(0.0, 0.3),
                         vy_range=(-0.1, 0.1), w_range=(-0.1, 0.1), weight_scale_range=(0.5, 2.0))
    raise ValueError(idx)


def acyclic_replans(B, motion="jump_fwd", seed=0, sigma_com=0.02, sigma_vcom=0.1, sigma_amom=0.02) -> CentroidalBatch:
    """B replans of an acyclic motion (abstract_acyclic_gen.py): replanning instants on a 50 ms grid over the whole
    motion, the state perturbed around the motion's first nominal state.  Solve with max_outer = ACYCLIC_MAX_OUTER."""
    from .acyclic import ACYCLIC_MOTIONS, build_batch
    prm = ACYCLIC_MOTIONS[motion]
    rng = np.random.default_rng(seed)
    x0 = np.asarray(prm.X_nom[0][0:9]) + np.concatenate([rng.normal(0, sigma_com, (B, 3)), rng.normal(0, sigma_vcom, (B, 3)),
                                                         rng.normal(0, sigma_amom, (B, 3))], axis=1)
    t = np.round(rng.uniform(0, prm.cnt_plan[-1][0][5], B) / 0.05) * 0.05
    return build_batch(prm, x0, t, 0.0)
