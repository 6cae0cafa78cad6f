"""`BiconvexMP`: the reference's python class, same constructor and the same 23 methods
(pybind11 module `biconvex_mpc_cpp`, iterative_supervised_learning/srcpy/motion_planner/biconvex.cpp:19-44),
with `optimize()` running on the GPU through libbunmpc.so.

The object is a thin stateful shell (exactly the state the C++ object keeps between calls: contact arrays,
costs, bounds, rho, warm-start iterates, the two FISTA step sizes and the violation history) around a
batch-of-one call of the batched solver.  Setters are host bookkeeping as in the reference; there is no
CPU solve path.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .problem import L0_F, L0_X, SolverParams
from .solver import get_solver


def _diag_of(Q, n):
    """Diagonal of a scipy.sparse / dense matrix; the kernels take diagonal Q (what create_cost_X/F build)."""
    if hasattr(Q, "diagonal") and hasattr(Q, "nnz"):          # scipy.sparse
        d = np.asarray(Q.diagonal(), dtype=np.float64)
        off = Q.copy()
        off.setdiag(0)
        if hasattr(off, "eliminate_zeros"):
            off.eliminate_zeros()
        if off.nnz:
            raise NotImplementedError("only diagonal cost matrices are supported (create_cost_X/F shape)")
    else:
        Qd = np.asarray(Q, dtype=np.float64)
        if Qd.ndim == 1:
            d = Qd
        else:
            d = np.diag(Qd).copy()
            if np.any(Qd - np.diag(d)):
                raise NotImplementedError("only diagonal cost matrices are supported (create_cost_X/F shape)")
    if d.shape != (n,):
        raise ValueError(f"cost matrix must be {n}x{n}")
    return d


class BiconvexMP:
    def __init__(self, m: float, n_col: int, n_eff: int, device: int = 0):
        # BiConvexMP::BiConvexMP, biconvex.cpp:6-25; ProblemData::ProblemData, problem.cpp:11-29
        self.m_, self.n_col_, self.n_eff_ = float(m), int(n_col), int(n_eff)
        self._m_dyn = float(m)                            # CentroidalDynamics::m_ is const (centroidal.hpp:50)
        n, e = self.n_col_, self.n_eff_
        self.nx, self.nf = 9 * (n + 1), 3 * e * n
        self.device = device
        self.rho_ = 1e5                                   # biconvex.hpp:148
        self.params = SolverParams()                      # maxit 150, tol 1e-5, exit_tol 1e-3 (biconvex.hpp:154-160)
        self.arith = _lib.ARITH_STRICT
        self.cnt_arr_ = np.zeros((n, e))                  # centroidal.cpp:32-33
        self.dt_ = np.zeros(n)                            # centroidal.cpp:9
        self.r_ = []                                      # std::vector<MatrixXd>, cleared by optimize()
        self.Q_x, self.q_x = np.zeros(self.nx), np.zeros(self.nx)
        self.Q_f, self.q_f = np.zeros(self.nf), np.zeros(self.nf)
        self.lb_x, self.ub_x = np.zeros(self.nx), np.zeros(self.nx)     # problem.cpp:24-25 (setZero)
        self.lb_f, self.ub_f = np.zeros(self.nf), np.zeros(self.nf)
        self.X_k, self.F_k, self.P_k_ = np.zeros(self.nx), np.zeros(self.nf), np.zeros(self.nx)
        self.L_f, self.L_x = L0_F, L0_X                   # biconvex.cpp:20-21; never reset (fista.hpp:28)
        self.log_statistics = False
        self.dyn_violation_hist_ = []
        self.rotation_matrices = []
        self.last_iters = None
        self.last_status = None
        self._solver = None

    # ---- contact plan ----
    def set_contact_plan(self, cnt_plan, dt):
        """centroidal.cpp:39-49"""
        cnt_plan = np.asarray(cnt_plan, dtype=np.float64)
        if cnt_plan.shape != (self.n_eff_, 4):
            raise ValueError(f"cnt_plan must be ({self.n_eff_}, 4)")
        i = len(self.r_)
        if i >= self.n_col_:
            raise IndexError("set_contact_plan called more than n_col times before optimize()")
        self.r_.append(cnt_plan[:, 1:4].copy())
        self.dt_[i] = float(dt)
        self.cnt_arr_[i, :] = cnt_plan[:, 0]

    def _cnt_plan(self):
        if len(self.r_) != self.n_col_:
            raise RuntimeError(f"contact plan holds {len(self.r_)} of {self.n_col_} knots "
                               "(set_contact_plan must be called n_col times; optimize() clears it)")
        return np.concatenate([self.cnt_arr_[:, :, None], np.stack(self.r_)], axis=2)

    def set_rotation_matrix_f(self, rot_matrix):
        """biconvex.hpp:85-89 (stored, never used by the projection: fista.cpp:57-58)"""
        self.rotation_matrices.append(np.asarray(rot_matrix, dtype=np.float64).copy())

    # ---- dense debug accessors, biconvex.hpp:30-51 ----
    def _mats(self, **kw):
        return self._get_solver().centroidal_mats(self._m_dyn, self._cnt_plan(), self.dt_, **kw)

    def return_A_x(self, X):
        return self._mats(X=X)["A_x"]

    def return_b_x(self, X):
        return self._mats(X=X)["b_x"]

    def return_A_f(self, F, x_init):
        return self._mats(F=F, x_init=x_init)["A_f"]

    def return_b_f(self, F, x_init):
        return self._mats(F=F, x_init=x_init)["b_f"]

    # ---- costs and bounds ----
    def set_cost_x(self, Q_x, q_x):
        self.Q_x = _diag_of(Q_x, self.nx)
        self.q_x = np.asarray(q_x, dtype=np.float64).reshape(self.nx).copy()

    def set_cost_f(self, Q_f, q_f):
        self.Q_f = _diag_of(Q_f, self.nf)
        self.q_f = np.asarray(q_f, dtype=np.float64).reshape(self.nf).copy()

    def create_cost_X(self, W_X, W_X_ter, X_ter, X_nom):
        """biconvex.cpp:60-72"""
        W_X, X_nom = np.asarray(W_X, dtype=np.float64), np.asarray(X_nom, dtype=np.float64)
        W_X_ter, X_ter = np.asarray(W_X_ter, dtype=np.float64), np.asarray(X_ter, dtype=np.float64)
        n9 = self.nx - 9
        self.Q_x[:n9] = W_X[:n9]
        self.Q_x[n9:] = W_X_ter
        self.q_x[:n9] = -2 * (X_nom * W_X)
        self.q_x[n9:] = -2 * (X_ter * W_X_ter)

    def create_cost_F(self, W_F):
        """biconvex.cpp:74-78 (q_f is left as it is: zero unless set_cost_f was called)"""
        self.Q_f[:] = np.asarray(W_F, dtype=np.float64)[: self.nf]

    def set_bounds_x(self, lb, ub):
        self.lb_x = np.asarray(lb, dtype=np.float64).reshape(self.nx).copy()
        self.ub_x = np.asarray(ub, dtype=np.float64).reshape(self.nx).copy()

    def set_bounds_f(self, lb, ub):
        """Stored only: the force box is dead code in the reference (fista.cpp:9-14)."""
        self.lb_f = np.asarray(lb, dtype=np.float64).reshape(self.nf).copy()
        self.ub_f = np.asarray(ub, dtype=np.float64).reshape(self.nf).copy()

    def create_bound_constraints(self, b, fx_max, fy_max, fz_max):
        """biconvex.cpp:27-58"""
        b = np.asarray(b, dtype=np.float64)
        if b.shape[1] != 6:
            print("bound constraints wrong size. Expected 6 ...")      # biconvex.cpp:33-35
        cnt = self._cnt_plan()
        n, e = self.n_col_, self.n_eff_
        self.lb_x = -np.inf * np.ones(self.nx)
        self.ub_x = np.inf * np.ones(self.nx)
        f = np.array([fx_max, fy_max, fz_max], dtype=np.float64)
        self.lb_f = np.tile(np.array([-f[0], -f[1], 0.0]), n * e)
        self.ub_f = np.tile(f, n * e)
        for i in range(n):
            if self.cnt_arr_[i].sum() > 0:
                r = cnt[i, :, 1:4]
                self.lb_x[9 * i: 9 * i + 3] = r.max(axis=0) + b[i, 0:3]
                self.ub_x[9 * i: 9 * i + 3] = r.min(axis=0) + b[i, 3:6]

    def set_rho(self, rho):
        self.rho_ = float(rho)

    def set_warm_start_vars(self, x_wm, f_wm, P_wm):
        """biconvex.hpp:66-70"""
        self.X_k = np.asarray(x_wm, dtype=np.float64).reshape(self.nx).copy()
        self.F_k = np.asarray(f_wm, dtype=np.float64).reshape(self.nf).copy()
        self.P_k_ = np.asarray(P_wm, dtype=np.float64).reshape(self.nx).copy()

    # ---- the solve ----
    def _get_solver(self):
        # looked up on every use: the cache may have replaced (and closed) the handle when a larger batch arrived
        return get_solver(self.n_col_, self.n_eff_, 1, self.device)

    def optimize(self, x_init, num_iters):
        """BiConvexMP::optimize, biconvex.cpp:80-120.  Runs on the GPU; iterates, dual and FISTA step sizes
        carry over to the next call exactly as the C++ members do."""
        x_init = np.asarray(x_init, dtype=np.float64).reshape(9)
        cnt = self._cnt_plan()
        prm = SolverParams(max_outer=int(num_iters), max_inner=self.params.max_inner, tol=self.params.tol,
                           exit_tol=self.params.exit_tol, beta=self.params.beta, mu=self.params.mu)
        sol = self._get_solver().solve_expanded(
            m=[self._m_dyn], rho=[self.rho_], x_init=x_init, cnt_plan=cnt, dt=self.dt_, Qx=self.Q_x, qx=self.q_x,
            Qf=self.Q_f, qf=self.q_f, lbx=self.lb_x, ubx=self.ub_x, X0=self.X_k, F0=self.F_k, P0=self.P_k_,
            L0=[self.L_f, self.L_x], params=prm, arith=self.arith, viol_hist=self.log_statistics)
        self.X_k, self.F_k, self.P_k_ = sol.X[0].copy(), sol.F[0].copy(), sol.P[0].copy()
        self.L_f, self.L_x = float(sol.L[0, 0]), float(sol.L[0, 1])
        self.last_iters, self.last_status = sol.iters[0].copy(), int(sol.status[0])
        if self.log_statistics:
            self.dyn_violation_hist_.extend(float(v) for v in sol.viol_hist[0, : int(sol.iters[0, 0])])
        if self.last_status == _lib.NAN:
            print("ERROR: solver diverged, Dyn violation is NaN")      # biconvex.cpp:107
        self.r_ = []                                                    # biconvex.cpp:117

    def return_opt_x(self):
        return self.X_k.copy()

    def return_opt_f(self):
        return self.F_k.copy()

    def return_opt_p(self):
        return self.P_k_.copy()

    def return_opt_com(self):
        """biconvex.cpp:122-130"""
        return self.X_k.reshape(-1, 9)[:, 0:3].copy()

    def return_opt_mom(self):
        """biconvex.cpp:132-142"""
        Xr = self.X_k.reshape(-1, 9)
        return np.concatenate([self.m_ * Xr[:, 3:6], Xr[:, 6:9]], axis=1)

    def return_dyn_viol_hist(self):
        return list(self.dyn_violation_hist_)

    def collect_statistics(self):
        self.log_statistics = True

    def set_robot_mass(self, m):
        """biconvex.hpp:135-137: changes only the mass used by return_opt_mom (not bound in python)."""
        self.m_ = float(m)


# upstream BiConMP spells the class both ways
BiConvexMP = BiconvexMP


class CentroidalDynamics:
    """Constructor-only binding in the reference (srcpy/motion_planner/biconvex.cpp:51-52)."""

    def __init__(self, m, n_col, n_eff):
        self.m_, self.n_col_, self.n_eff_ = float(m), int(n_col), int(n_eff)
