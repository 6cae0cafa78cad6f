"""ctypes wrapper of the CPU oracle (oracle/bicon_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under bunmpc_b200/ imports this module.
PARITY UNPINNED against real Eigen (see bicon_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbicon_oracle.so")
_lib = None


class Params(C.Structure):
    _fields_ = [("max_outer", C.c_int), ("max_inner", C.c_int), ("tol", C.c_double),
                ("exit_tol", C.c_double), ("beta", C.c_double), ("mu", C.c_double),
                ("use_fma", C.c_int), ("reduction", C.c_int), ("storage", C.c_int)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bicon_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _lib.bicon_default_params.argtypes = [C.POINTER(Params)]
        _lib.bicon_solve_batch.restype = C.c_int
        _lib.bicon_solve_batch.argtypes = ([C.c_int] * 3 + [dp] * 15 + [C.POINTER(Params), C.c_int]
                                           + [dp] * 4 + [ip, dp, ip])
        _lib.bicon_create_bound_constraints.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp]
        _lib.bicon_create_cost_X.argtypes = [C.c_int] + [dp] * 6
        _lib.bicon_dense_x_mat.argtypes = [C.c_int, C.c_int, C.c_double] + [dp] * 5
        _lib.bicon_dense_f_mat.argtypes = [C.c_int, C.c_int, C.c_double] + [dp] * 6
    return _lib


def default_params(**kw) -> Params:
    p = Params()
    lib().bicon_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _d(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
    return a


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def create_bound_constraints(cnt_plan, bounds):
    """[n,e,4], [n,6] -> lbx, ubx  (biconvex.cpp:27-58)"""
    cnt_plan = _d(cnt_plan)
    n, e = cnt_plan.shape[0], cnt_plan.shape[1]
    bounds = _d(bounds, (n, 6))
    nx = 9 * (n + 1)
    lb, ub = np.empty(nx), np.empty(nx)
    lib().bicon_create_bound_constraints(n, e, _p(cnt_plan), _p(bounds), _p(lb), _p(ub))
    return lb, ub


def create_cost_X(W_X, W_X_ter, X_ter, X_nom):
    """-> Qx (diag), qx  (biconvex.cpp:60-72)"""
    W_X, X_nom = _d(W_X), _d(X_nom)
    n = W_X.shape[0] // 9
    Qx, qx = np.empty(9 * (n + 1)), np.empty(9 * (n + 1))
    lib().bicon_create_cost_X(n, _p(W_X), _p(_d(W_X_ter)), _p(_d(X_ter)), _p(X_nom), _p(Qx), _p(qx))
    return Qx, qx


def dense_x_mat(m, cnt_plan, dt, X):
    cnt_plan = _d(cnt_plan)
    n, e = cnt_plan.shape[0], cnt_plan.shape[1]
    nx, nf = 9 * (n + 1), 3 * e * n
    A, b = np.empty((nx, nf)), np.empty(nx)
    lib().bicon_dense_x_mat(n, e, float(m), _p(cnt_plan), _p(_d(dt)), _p(_d(X)), _p(A), _p(b))
    return A, b


def dense_f_mat(m, cnt_plan, dt, F, x_init):
    cnt_plan = _d(cnt_plan)
    n, e = cnt_plan.shape[0], cnt_plan.shape[1]
    nx = 9 * (n + 1)
    A, b = np.empty((nx, nx)), np.empty(nx)
    lib().bicon_dense_f_mat(n, e, float(m), _p(cnt_plan), _p(_d(dt)), _p(_d(F)), _p(_d(x_init)), _p(A), _p(b))
    return A, b


def expand(batch):
    """Compact gait-generator inputs -> the expanded per-instance arrays the solver consumes,
    through the oracle's own restatement of create_bound_constraints / create_cost_X / create_cost_F.
    `batch` is any object with the attributes of bunmpc_b200.problem.CentroidalBatch."""
    B, n, e = batch.B, batch.n_col, batch.n_eff
    nx, nf = 9 * (n + 1), 3 * e * n
    cnt = _d(batch.cnt_plan, (B, n, e, 4))
    bounds = _d(batch.bounds, (B, n, 6))
    W_X, W_X_ter = _d(batch.W_X, (B, 9 * n)), _d(batch.W_X_ter, (B, 9))
    X_nom, X_ter = _d(batch.X_nom, (B, 9 * n)), _d(batch.X_ter, (B, 9))
    out = dict(Qx=np.empty((B, nx)), qx=np.empty((B, nx)), lbx=np.empty((B, nx)), ubx=np.empty((B, nx)))
    for b in range(B):
        out["lbx"][b], out["ubx"][b] = create_bound_constraints(cnt[b], bounds[b])
        out["Qx"][b], out["qx"][b] = create_cost_X(W_X[b], W_X_ter[b], X_ter[b], X_nom[b])
    out["Qf"] = _d(batch.W_F, (B, nf)).copy()          # create_cost_F, biconvex.cpp:74-78
    out["qf"] = np.zeros((B, nf))                      # q_ is zero-initialised, problem.cpp:27-28
    return out


def solve_expanded(n_col, n_eff, m, rho, x_init, cnt_plan, dt, Qx, qx, Qf, qf, lbx, ubx,
                   X0, F0, P0, L0, params=None, n_threads=1):
    """B instances in struct-of-arrays form (leading batch dimension) -> dict of results."""
    n, e = int(n_col), int(n_eff)
    nx, nf = 9 * (n + 1), 3 * e * n
    x_init = _d(x_init)
    if x_init.ndim == 1:
        x_init = x_init[None]
    B = x_init.shape[0]
    a = dict(m=_d(m, (B,)), rho=_d(rho, (B,)), x_init=_d(x_init, (B, 9)), cnt_plan=_d(cnt_plan, (B, n, e, 4)),
             dt=_d(dt, (B, n)), Qx=_d(Qx, (B, nx)), qx=_d(qx, (B, nx)), Qf=_d(Qf, (B, nf)), qf=_d(qf, (B, nf)),
             lbx=_d(lbx, (B, nx)), ubx=_d(ubx, (B, nx)), X0=_d(X0, (B, nx)), F0=_d(F0, (B, nf)),
             P0=_d(P0, (B, nx)), L0=_d(L0, (B, 2)))
    prm = params if params is not None else default_params()
    X, F, P = np.empty((B, nx)), np.empty((B, nf)), np.empty((B, nx))
    L, viol = np.empty((B, 2)), np.empty(B)
    iters, status = np.empty((B, 5), dtype=np.int32), np.empty(B, dtype=np.int32)
    rc = lib().bicon_solve_batch(
        B, n, e, _p(a["m"]), _p(a["rho"]), _p(a["x_init"]), _p(a["cnt_plan"]), _p(a["dt"]),
        _p(a["Qx"]), _p(a["qx"]), _p(a["Qf"]), _p(a["qf"]), _p(a["lbx"]), _p(a["ubx"]),
        _p(a["X0"]), _p(a["F0"]), _p(a["P0"]), _p(a["L0"]), C.byref(prm), int(n_threads),
        _p(X), _p(F), _p(P), _p(L), iters.ctypes.data_as(C.POINTER(C.c_int)), _p(viol),
        status.ctypes.data_as(C.POINTER(C.c_int)))
    if rc != 0:
        raise RuntimeError(f"bicon_solve_batch failed: {rc}")
    return dict(X=X, F=F, P=P, L=L, iters=iters, viol=viol, status=status)


def solve(batch, params=None, n_threads=1):
    """Solve a compact CentroidalBatch: expand (a7/a8), cold/warm start (kino_dyn.cpp:83-99), solve."""
    B, n, e = batch.B, batch.n_col, batch.n_eff
    nx, nf = 9 * (n + 1), 3 * e * n
    ex = expand(batch)
    x_init = _d(batch.x_init, (B, 9))
    X0 = np.tile(x_init, (1, n + 1)) if batch.X0 is None else batch.X0      # X_wm = tile(x0)
    F0 = np.zeros((B, nf)) if batch.F0 is None else batch.F0                # F_wm = 0
    P0 = np.zeros((B, nx)) if batch.P0 is None else batch.P0                # P_wm = 0
    return solve_expanded(n, e, batch.m, batch.rho, x_init, batch.cnt_plan, batch.dt,
                          ex["Qx"], ex["qx"], ex["Qf"], ex["qf"], ex["lbx"], ex["ubx"],
                          X0, F0, P0, batch.L0, params=params, n_threads=n_threads)


# ---------------------------------------------------------------------------------------------------
# oracle/_ref: the reference's own sources compiled against the Eigen stand-in (oracle/refshim)
# ---------------------------------------------------------------------------------------------------
_REF_SO = os.path.join(_HERE, "_ref", "libbicon_ref.so")
_ref = None


def ref_available() -> bool:
    return os.path.exists(_REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF_SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _ref.ref_create.restype = C.c_void_p
        _ref.ref_create.argtypes = [C.c_double, C.c_int, C.c_int]
        _ref.ref_destroy.argtypes = [C.c_void_p]
        _ref.ref_solve.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_double] + [dp] * 12 + [C.c_int] + [dp] * 4
                                   + [ip, dp])
    return _ref


def ref_solve(batch, max_outer=100):
    """Every instance of a CentroidalBatch through the reference's own BiConvexMP (fresh object per instance,
    its own create_* setters and optimize()).  Same result dict as solve()."""
    R = ref_lib()
    B, n, e = batch.B, batch.n_col, batch.n_eff
    nx, nf = 9 * (n + 1), 3 * e * n
    g = lambda name, shape: _d(getattr(batch, name), (B,) + shape)
    m, rho, x_init = g("m", ()), g("rho", ()), g("x_init", (9,))
    cnt, dt, bounds = g("cnt_plan", (n, e, 4)), g("dt", (n,)), g("bounds", (n, 6))
    W_X, W_X_ter, X_ter, X_nom, W_F = g("W_X", (9 * n,)), g("W_X_ter", (9,)), g("X_ter", (9,)), g("X_nom", (9 * n,)), g("W_F", (nf,))
    L0 = g("L0", (2,))
    X0 = np.tile(x_init, (1, n + 1)) if batch.X0 is None else _d(batch.X0, (B, nx))
    F0 = np.zeros((B, nf)) if batch.F0 is None else _d(batch.F0, (B, nf))
    P0 = np.zeros((B, nx)) if batch.P0 is None else _d(batch.P0, (B, nx))
    X, F, P = np.empty((B, nx)), np.empty((B, nf)), np.empty((B, nx))
    L, viol = L0.copy(), np.empty(B)
    iters, status = np.empty((B, 5), dtype=np.int32), np.empty(B, dtype=np.int32)
    for b in range(B):
        h = R.ref_create(float(m[b]), n, e)
        it = (C.c_int * 5)()
        v = C.c_double()
        R.ref_solve(h, n, e, float(rho[b]), _p(x_init[b]), _p(cnt[b]), _p(dt[b]), _p(bounds[b]), _p(W_X[b]),
                    _p(W_X_ter[b]), _p(X_ter[b]), _p(X_nom[b]), _p(W_F[b]), _p(X0[b]), _p(F0[b]), _p(P0[b]),
                    int(max_outer), _p(L[b]), _p(X[b]), _p(F[b]), _p(P[b]), it, C.byref(v))
        R.ref_destroy(h)
        iters[b] = list(it)
        viol[b] = v.value
        status[b] = 2 if np.isnan(v.value) else (0 if v.value < 1e-3 else 1)
    return dict(X=X, F=F, P=P, L=L, iters=iters, viol=viol, status=status)
