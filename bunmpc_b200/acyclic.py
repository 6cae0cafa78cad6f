"""Acyclic motions: the second caller of the centroidal solve (`SoloAcyclicGen`, examples/mpc/abstract_acyclic_gen.py).

An acyclic motion (jump, rearing, hi-five, cartwheel) is a TIME TABLE instead of a gait: contact segments
`[c, x, y, z, t_start, t_end]` per foot, nominal-state segments `[9 values, t_start, t_end]`, box segments
`[6 values, t_start, t_end]` (examples/motions/weight_abstract.py:46-83).  For a replan at time `t` the generator looks
every knot of the horizon up in those tables (`create_contact_plan` :74-124, the dynamics part of `create_costs`
:126-190) and solves with 50 outer iterations (`self.kd.optimize(q, v, 50, 1)` :319).

This module holds
* `ACyclicMotionParams` + `ACYCLIC_MOTIONS`: the centroidal part of the reference's motion records
  (examples/motions/acyclic/plan_jump.py, rearing.py, plan_hifive.py, plan_cartwheel.py, rearing_jump.py; `stand.py` is
  a stale record the reference's own generator raises on),
* `build_batch`: the table look-ups for B replans at once (numpy, sequential over the knots because the reference
  accumulates and rounds the knot time knot by knot, vectorised over the batch) -> `CentroidalBatch`,
* `SoloAcyclicGen`: the reference's call shape around it (update_motion_params / optimize / get_plan_freq), from the
  centroidal state; IK (xs, us) is out of scope as in `gait_gen.py`.

Pinned bit for bit against the reference's own python run in place (tests/golden/make_acyclic_golden.py ->
tests/golden/acyclic_cases.npz, tests/test_acyclic.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .problem import BatchSolution, CentroidalBatch, L0_F, L0_X, SolverParams

ACYCLIC_MAX_OUTER = 50        # abstract_acyclic_gen.py:319 (cyclic gaits: 100)
ACYCLIC_F_MAX = 25.0          # :34-36 (dead in the solve: quirk Q4)


@dataclass
class ACyclicMotionParams:
    """examples/motions/weight_abstract.py:46-83, the fields the centroidal solve reads."""
    robot_name: str
    motion_name: str
    n_col: int = None
    dt_arr: list = None
    plan_freq: list = None        # [[period, t_start, t_end]]
    cnt_plan: list = None         # [segment][foot] = [c, x, y, z, t_start, t_end]
    W_X: np.ndarray = None
    W_X_ter: np.ndarray = None
    W_F: np.ndarray = None
    X_nom: list = None            # [segment] = [9 values, t_start, t_end]
    X_ter: list = None
    rho: float = None
    bounds: list = None           # [segment] = [6 values, t_start, t_end]
    mass: float = 2.5             # pin.computeTotalMass of Solo12 (abstract_acyclic_gen.py:24)
    eff_names: tuple = ("FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT")


# ---- the reference's records.  Feet of the standing Solo12 with the base at x = 0.2: hips at 0.2 +- 0.1946 ----
_FRONT, _HIND, _SIDE = 0.3946, 0.0054, 0.14695


def _feet(contacts, t0, t1, front=_FRONT, hind=_HIND):
    """one contact segment: FL, FR, HL, HR on the ground (1) or not (0) between t0 and t1"""
    xs, ys = (front, front, hind, hind), (_SIDE, -_SIDE, _SIDE, -_SIDE)
    return [[float(c), x, y, 0.0, t0, t1] for c, x, y in zip(contacts, xs, ys)]


def _nom(com, amom_y, t0, t1):
    return [com[0], com[1], com[2], 0.0, 0.0, 0.0, 0.0, amom_y, 0.0, t0, t1]


def _box(z_lo, z_hi, t0, t1, xy=0.25):
    return [-xy, -xy, z_lo, xy, xy, z_hi, t0, t1]


def _jump_fwd():
    st, flight, T = 0.4, 0.3, 1.2                                  # plan_jump.py:25-27
    p = ACyclicMotionParams("solo12", "jump_fwd", n_col=25, dt_arr=25 * [T / 30], rho=7e4)
    p.plan_freq = [[0.3, 0, st + flight], [0.5, st + flight, T]]
    p.cnt_plan = [_feet((1, 1, 1, 1), 0.0, st), _feet((0, 0, 0, 0), st, st + flight), _feet((1, 1, 1, 1), st + flight, T)]
    p.W_X = np.array([1e-5, 1e-5, 1e5, 1e-4, 1e-4, 1e-4, 3e4, 3e4, 3e4])
    p.W_X_ter = 10 * np.array([1e-5, 1e-5, 1e5, 1e2, 1e1, 1e2, 1e5, 1e5, 1e5])
    p.W_F = np.array(4 * [5.0, 5.0, 7.0])
    c = (0.2, 0.0, 0.22)
    p.X_nom = [_nom(c, 0.0, 0.0, st), _nom(c, 0.0, st, st + flight), _nom(c, 0.0, st + flight, T)]
    p.X_ter = [0.2, 0, 0.2, 0, 0, 0, 0, 0.0, 0.0]
    p.bounds = [_box(0.1, 0.25, 0, st), _box(0.1, 0.3, st, T)]
    return p


def _rearing_family(name):
    """rearing.py, plan_hifive.py, rearing_jump.py share their first two phases: stand, then rear up on the hind legs"""
    st, rear = 0.5, 0.4
    common_W_X_ter = 10 * np.array([1e3, 1e1, 1e5, 1e-1, 1e-1, 1e-1, 1e2, 1e4, 1e2])
    stand, up = (0.2, 0.0, 0.22), (0.18, 0.0, 0.28)
    if name == "rearing":                                          # rearing.py:22-63
        T = 1.2
        p = ACyclicMotionParams("solo12", name, n_col=20, dt_arr=20 * [5e-2], rho=5e4)
        p.plan_freq = [[0.4, 0, st + rear], [0.4, st + rear, T]]
        p.cnt_plan = [_feet((1, 1, 1, 1), 0.0, st), _feet((0, 0, 1, 1), st, st + rear),
                      _feet((1, 1, 1, 1), st + rear, T, front=0.41)]
        p.W_X = np.array([1e3, 1e1, 1e5, 1e-4, 1e-4, 1e-4, 1e2, 5e3, 1e2])
        p.W_F = np.array(4 * [1e1, 1e1, 1e0])
        p.X_nom = [_nom(stand, -0.05, 0.0, st), _nom(up, -0.45, st, st + rear), _nom(stand, 0.0, st + rear, T)]
        p.bounds = [_box(0.1, 0.25, 0, st), _box(0.1, 0.4, st, st + rear), _box(0.1, 0.25, st + rear, T)]
    elif name == "hifive":                                         # plan_hifive.py:22-70
        five, T = 0.1, 1.4
        p = ACyclicMotionParams("solo12", name, n_col=25, dt_arr=25 * [5e-2], rho=5e4)
        p.plan_freq = [[1.4, 0, st], [1.4, st, st + rear + five], [0.05, st + rear + five, T]]
        p.cnt_plan = [_feet((1, 1, 1, 1), 0.0, st), _feet((0, 0, 1, 1), st, st + rear),
                      _feet((0, 0, 0, 0), st + rear, st + rear + five),
                      _feet((1, 1, 1, 1), st + rear + five, T, front=0.41, hind=-0.0054)]
        p.W_X = np.array([1e3, 1e1, 1e2, 1e-4, 1e-4, 1e-4, 1e2, 5e3, 1e2])
        p.W_F = np.array(4 * [1e1, 1e1, 5e-1])
        p.X_nom = [_nom(stand, -0.05, 0.0, st), _nom(up, -0.45, st, st + rear),
                   _nom((0.18, 0.0, 0.32), 0.0, st + rear, st + rear + five), _nom(stand, 0.0, st + rear + five, T)]
        p.bounds = [_box(0.1, 0.25, 0, st), _box(0.1, 0.4, st, st + rear), _box(0.1, 0.25, st + rear, T)]
    else:                                                          # rearing_jump.py:22-71
        jump, T = 0.4, 1.4
        p = ACyclicMotionParams("solo12", name, n_col=20, dt_arr=20 * [5e-2], rho=5e4)
        p.plan_freq = [[0.4, 0, st + rear], [0.4, st + rear, T]]
        p.cnt_plan = [_feet((1, 1, 1, 1), 0.0, st), _feet((0, 0, 1, 1), st, st + rear),
                      _feet((0, 0, 0, 0), st + rear, st + rear + jump),
                      _feet((1, 1, 1, 1), st + rear + jump, T, front=0.41)]
        p.W_X = np.array([1e3, 1e1, 1e5, 1e-4, 1e-4, 1e-4, 1e2, 5e3, 1e2])
        p.W_F = np.array(4 * [1e1, 1e1, 1e0])
        p.X_nom = [_nom(stand, -0.05, 0.0, st), _nom(up, -0.45, st, st + rear),
                   _nom((0.23, 0.0, 0.3), 0.0, st + rear, st + rear + jump), _nom((0.23, 0.0, 0.22), 0.0, st + rear + jump, T)]
        # the third segment overlaps the second (both start at st): the first match wins, as in the reference's loop
        p.bounds = [_box(0.1, 0.25, 0, st), _box(0.1, 0.4, st, st + rear),
                    [-np.inf, -np.inf, 0.0, np.inf, np.inf, 0.7, st, st + rear + jump], _box(0.1, 0.25, st + rear + jump, T)]
    p.W_X_ter = common_W_X_ter
    p.X_ter = [0.2, 0, 0.22, 0, 0, 0, 0, 0.0, 0.0]
    return p


def _cartwheel():
    st, flip, T = 0.4, 0.5, 1.2                                    # plan_cartwheel.py:22-60 (24 knots, 25 step lengths)
    p = ACyclicMotionParams("solo12", "cartwheel", n_col=24, dt_arr=25 * [5e-2], rho=5e4)
    p.plan_freq = [[0.6, 0, T], [1.0, T, T + 1.5]]
    p.cnt_plan = [_feet((1, 1, 1, 1), 0.0, st), _feet((1, 1, 0, 0), st, st + flip),
                  _feet((1, 1, 1, 1), st + flip, T, hind=0.8054)]
    p.W_X = np.array([1e-2, 1e-2, 1e5, 1e-2, 1e-2, 1e-4, 1e3, 1e3, 1e4])
    p.W_X_ter = 10 * np.array([1e-2, 1e-2, 1e5, 1e-2, 1e-2, 1e-4, 1e3, 1e4, 1e4])
    p.W_F = np.array(4 * [1e1, 1e1, 2e0])
    p.X_nom = [_nom((0.2, 0.0, 0.2), 0.1, 0.0, st), _nom((0.4, 0.0, 0.3), 0.6, st, st + flip),
               _nom((0.6, 0.0, 0.2), 0.0, st + flip, T)]
    p.X_ter = [0.2, 0, 0.2, 0, 0, 0, 0, 0.0, 0.0]
    p.bounds = [_box(0.0, 0.3, 0, st, xy=0.45), _box(0.0, 0.45, st, T, xy=0.45)]
    return p


ACYCLIC_MOTIONS = {"jump_fwd": _jump_fwd(), "rearing": _rearing_family("rearing"), "hifive": _rearing_family("hifive"),
                   "cartwheel": _cartwheel(), "rearing_jump": _rearing_family("rearing_jump")}


# ---- the table look-ups, for B replans at once ----
def _lookup(ft, starts, ends, values, beyond):
    """values[k] of the FIRST segment k with starts[k] <= ft < ends[k] while ft < ends[-1] (zeros if none matches, as the
    reference's loops leave them), `beyond` otherwise.  ft [B]; values [K, ...]; returns [B, ...] and the mask ft < ends[-1]."""
    inside = ft < ends[-1]
    out = np.zeros((ft.shape[0],) + values.shape[1:])
    taken = np.zeros(ft.shape[0], dtype=bool)
    for k in range(len(starts)):
        hit = inside & ~taken & (ft >= starts[k]) & (ft < ends[k])
        out[hit] = values[k]
        taken |= hit
    out[~inside] = beyond
    return out, inside


def build_batch(params: ACyclicMotionParams, x_init, t, t0=0.0, L0=None) -> CentroidalBatch:
    """`create_contact_plan` (abstract_acyclic_gen.py:74-124) and the dynamics part of `create_costs` (:126-190) for B
    replans: x_init [B,9] = [com, hg_lin / m, hg_ang] (:136-140), t [B] the time of each replan, t0 the time the motion
    started (`update_motion_params`, :42-54).  make_cyclic = False and the use_current_* switches off, as everywhere in
    the reference."""
    x_init = np.atleast_2d(np.asarray(x_init, dtype=np.float64))
    B = x_init.shape[0]
    t = np.broadcast_to(np.asarray(t, dtype=np.float64), (B,)).copy()
    t0 = np.float64(t0)
    n, e = int(params.n_col), len(params.eff_names)
    dt_arr = np.asarray(params.dt_arr, dtype=np.float64)

    seg = np.asarray(params.cnt_plan, dtype=np.float64)                 # [K, e, 6]
    cnt_plan, dt = np.zeros((B, n, e, 4)), np.zeros((B, n))
    ft = np.round(t - dt_arr[0] - t0, 3)                                # :83
    for i in range(n):
        ft = ft + np.round(dt_arr[i], 3)                                # :88 (the sum is not rounded again here)
        cnt_plan[:, i], _ = _lookup(ft, seg[:, 0, 4], seg[:, 0, 5], seg[:, :, 0:4], seg[-1, :, 0:4])    # :90-110
        dt[:, i] = dt_arr[i]
    first = dt_arr[0] - np.round(np.remainder(t, dt_arr[0]), 2)         # :116-119: the first step ends on the time grid
    dt[:, 0] = np.where(first == 0, dt_arr[0], first)

    nom = np.asarray(params.X_nom, dtype=np.float64)                    # [K, 11]
    box = np.asarray(params.bounds, dtype=np.float64)                   # [K, 8]
    X_ter_rec = np.asarray(params.X_ter, dtype=np.float64)
    X_nom, bounds = np.zeros((B, n, 9)), np.zeros((B, n, 6))
    ft = t - dt_arr[0] - t0                                             # :142, :166 (not rounded at the start here)
    for i in range(n):
        ft = np.round(ft + dt_arr[i], 3)                                # :145-146
        X_nom[:, i], inside = _lookup(ft, nom[:, 9], nom[:, 10], nom[:, 0:9], X_ter_rec)        # :147-157
        bounds[:, i], _ = _lookup(ft, box[:, -2], box[:, -1], box[:, 0:6], box[-1, 0:6])        # :170-178
    X_ter = X_nom[:, n - 1].copy()                                      # :152-153, :157-158: the last knot's nominal state
    X_nom[:, 0] = x_init                                                # :183
    if n == 1:
        X_ter = np.where(inside[:, None], x_init, X_ter)                # X_ter is a VIEW of X_nom[-9:] in the reference
    return CentroidalBatch(
        n_col=n, n_eff=e, m=np.array([params.mass]), rho=np.array([params.rho], dtype=np.float64), x_init=x_init,
        cnt_plan=cnt_plan, dt=dt, W_X=np.tile(params.W_X, n)[None], W_X_ter=np.asarray(params.W_X_ter, dtype=np.float64)[None],
        X_nom=X_nom.reshape(B, 9 * n), X_ter=X_ter, W_F=np.tile(params.W_F, n)[None], bounds=bounds,
        L0=np.array([[L0_F, L0_X]]) if L0 is None else L0)


def _table(rows, x, value=lambda r: r[0]):
    """get_plan_freq / get_gains, abstract_acyclic_gen.py:349-369: the row whose [t_start, t_end) holds x, the last row
    from its end on; None in a gap (the reference falls off its loop there)."""
    for r in rows:
        if x < rows[-1][-1]:
            if r[-2] <= x < r[-1]:
                return value(r)
        else:
            return value(rows[-1])
    return None


class SoloAcyclicGen:
    """abstract_acyclic_gen.py:13-369 without pinocchio and the IK: replans start from the centroidal state."""

    def __init__(self, robot=None, r_urdf=None, device: int = 0):
        self.robot, self.r_urdf, self.device = robot, r_urdf, device
        self.eff_names = ["FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT"]
        self.n_eff = 4
        self.fx_max = self.fy_max = self.fz_max = ACYCLIC_F_MAX
        self.params = None
        self.L = None
        self.last = None

    def update_motion_params(self, weight_abstract: ACyclicMotionParams, q0=None, t0=0.0):
        """:42-72"""
        self.params, self.q0, self.t0 = weight_abstract, q0, t0
        self.m = weight_abstract.mass
        self.freq = self.params.plan_freq[0][0]
        self.horizon = self.ik_horizon = self.params.n_col
        self.size = min(self.ik_horizon, int(self.freq / self.params.dt_arr[0]) + 2)      # :65-68
        if self.freq > self.params.dt_arr[0]:
            self.size += 1
        self.L = None                                               # a new KinoDynMP (fresh FISTA objects) is made here, :56

    def optimize_centroidal_batch(self, x_init, t, params: SolverParams = None, builder: str = "host") -> BatchSolution:
        """B replans in one launch; 50 outer iterations (:319) unless params says otherwise.  builder = "device": only
        the states and replanning instants go up, build_acyclic_kernel looks the knots up where the problem is solved."""
        from .solver import get_solver
        prm = params if params is not None else SolverParams(max_outer=ACYCLIC_MAX_OUTER)
        if builder == "device":
            x_init = np.atleast_2d(np.asarray(x_init, dtype=np.float64))
            s = get_solver(int(self.params.n_col), self.n_eff, x_init.shape[0], self.device)
            dev = s.build_acyclic_device(self.params, x_init, t, self.t0, L0=self.L)
            o = s.solve_resident(dev, params=prm)
            sol = BatchSolution(X=o["X"].cpu().numpy(), F=o["F"].cpu().numpy(), P=o["P"].cpu().numpy(), L=o["L"].cpu().numpy(),
                                iters=o["iters"].cpu().numpy(), viol=o["viol"].cpu().numpy(), status=o["status"].cpu().numpy(),
                                m=np.array([self.params.mass]), cycles=o["cycles"].cpu().numpy())
            self.last = (dev, sol)
            return sol
        batch = build_batch(self.params, x_init, t, self.t0, L0=self.L)
        sol = get_solver(batch.n_col, batch.n_eff, batch.B, self.device).solve(batch, prm)
        self.last = (batch, sol)
        return sol

    def optimize_centroidal(self, x_init, t):
        """One replan.  Returns f_int as :331-346 builds it: every knot's force held for int(dt / 0.001) samples
        (np.linspace between a knot and itself; the IK's xs has n + 1 entries and the last one carries no force); the
        step sizes L carry over to the next replan."""
        sol = self.optimize_centroidal_batch(np.atleast_2d(x_init), np.atleast_1d(t))
        self.L = sol.L.copy()
        batch = self.last[0]
        n, ne = batch.n_col, 3 * batch.n_eff
        F = sol.F[0]
        dts = np.asarray(self.params.dt_arr, dtype=np.float64)      # create_costs copies params.dt_arr into self.dt_arr, :229
        self.f_int = np.vstack([np.linspace(F[i * ne:(i + 1) * ne], F[i * ne:(i + 1) * ne], int(dts[i] / 0.001))
                                for i in range(n)])
        return self.f_int

    def get_plan_freq(self, t):
        return _table(self.params.plan_freq, t - self.t0)
