"""Recorder stand-ins of the pybind module biconvex_mpc_cpp (srcpy/motion_planner/biconvex.cpp:15-63): they keep
what the reference's gait generator hands to the solver.  TEST INFRASTRUCTURE ONLY (oracle/pinshim/README.md)."""
import numpy as np

from inverse_kinematics_cpp import InverseKinematics


class BiconvexMP:
    def __init__(self, m, n_col, n_eff):
        self.m, self.n_col, self.n_eff = m, n_col, n_eff
        self.reset()

    def reset(self):
        self.cnt_plan, self.dt = [], []
        self.bounds = self.W_X = self.W_X_ter = self.X_ter = self.X_nom = self.W_F = self.rho = None

    def set_rho(self, rho):
        self.rho = float(rho)

    def set_contact_plan(self, plan, dt):
        self.cnt_plan.append(np.array(plan, dtype=np.float64))
        self.dt.append(float(dt))

    def create_bound_constraints(self, b, fx, fy, fz):
        self.bounds = np.array(b, dtype=np.float64)
        self.f_max = (fx, fy, fz)

    def create_cost_X(self, W_X, W_X_ter, X_ter, X_nom):
        self.W_X, self.W_X_ter = np.array(W_X, dtype=np.float64), np.array(W_X_ter, dtype=np.float64)
        self.X_ter, self.X_nom = np.array(X_ter, dtype=np.float64), np.array(X_nom, dtype=np.float64)

    def create_cost_F(self, W_F):
        self.W_F = np.array(W_F, dtype=np.float64)


class KinoDynMP:
    def __init__(self, urdf, m, n_eff, dyn_col, ik_col):
        self._dyn = BiconvexMP(m, dyn_col, n_eff)
        self._ik = InverseKinematics()

    def return_dyn(self):
        return self._dyn

    def return_ik(self):
        return self._ik

    def set_com_tracking_weight(self, w):
        pass

    def set_mom_tracking_weight(self, w):
        pass
