"""Phase/stance lookup of the cyclic gait, vectorised over a batch of times.

Host-side restatement of gait_planner::QuadrupedGait (src/gait_planner/gait_planner.cpp:4-58,104-121,
bound as `GaitPlanner` in srcpy/gait_planner/py_gait_planner.cpp).  Scalars in, scalars out keeps the
reference's call shape; arrays in, arrays out is what the batched contact-plan builder uses.
"""
from __future__ import annotations

import numpy as np


class QuadrupedGait:
    def __init__(self, gait_period, stance_percent, phase_offset, step_height):
        self.gait_period_ = float(gait_period)
        self.stance_percent_ = np.asarray(stance_percent, dtype=np.float64).copy()
        self.stance_time_ = self.gait_period_ * self.stance_percent_          # gait_planner.cpp:12
        self.swing_percent_ = 1.0 - self.stance_percent_
        self.swing_time_ = self.gait_period_ - self.stance_time_
        self.phase_offset_ = np.asarray(phase_offset, dtype=np.float64).copy()
        self.step_height_ = float(step_height)
        self.n_eff = self.stance_percent_.size

    def get_phi(self, time_in, foot_ID):
        """gait_planner.cpp:41-44: fmod(t + offset*T, T)"""
        return np.fmod(np.asarray(time_in, dtype=np.float64) + self.phase_offset_[foot_ID] * self.gait_period_,
                       self.gait_period_)

    def get_phase(self, time_in, foot_ID):
        """gait_planner.cpp:46-58: 1 = stance iff phi <= stance_time or |phi - stance_time| < 1e-4"""
        phi = self.get_phi(time_in, foot_ID)
        st = self.stance_time_[foot_ID]
        ph = ((phi <= st) | (np.abs(phi - st) < 1e-4)).astype(np.int64)
        return int(ph) if ph.ndim == 0 else ph

    def get_percent_in_phase(self, time_in, foot_ID):
        """gait_planner.cpp:104-121"""
        phi = self.get_phi(time_in, foot_ID)
        st = self.stance_time_[foot_ID]
        out = np.where(phi <= st, phi / st, (phi - st) / (self.gait_period_ - st))
        return float(out) if out.ndim == 0 else out

    def set_step_height(self, step_height):
        self.step_height_ = float(step_height)


GaitPlanner = QuadrupedGait
