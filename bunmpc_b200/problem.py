"""Batched problem container for the centroidal biconvex solve.

`CentroidalBatch` holds, with a leading batch dimension, exactly what the reference's gait generator
hands to one `BiconvexMP` object before `optimize()`:

  set_contact_plan(cnt_plan[i], dt[i])            abstract_cyclic_gen.py:391 -> centroidal.cpp:39-49
  create_bound_constraints(bounds, fx, fy, fz)    abstract_cyclic_gen.py:611-612 -> biconvex.cpp:27-58
  create_cost_X(W_X, W_X_ter, X_ter, X_nom)       abstract_cyclic_gen.py:613 -> biconvex.cpp:60-72
  create_cost_F(W_F)                              abstract_cyclic_gen.py:614 -> biconvex.cpp:74-78
  set_rho(rho), mass m (ctor)                     abstract_cyclic_gen.py:133,146
  optimize(x_init, num_iters)                     kino_dyn.cpp:47 -> biconvex.cpp:80-120

A field whose leading dimension is 1 is shared by every instance (batch stride 0 on the device).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Optional

import numpy as np

L0_F = 506.25    # biconvex.cpp:21  fista_f.set_l0
L0_X = 2.25e6    # biconvex.cpp:20  fista_x.set_l0


@dataclass
class SolverParams:
    """Solver constants; defaults are the reference's member initialisers
    (biconvex.hpp:148-160, fista.hpp:52-60).  None of them is settable from python in the reference."""
    max_outer: int = 100      # kd.optimize(q, v, 100, 1), abstract_cyclic_gen.py:663
    max_inner: int = 150      # biconvex.hpp:156
    tol: float = 1e-5         # biconvex.hpp:158
    exit_tol: float = 1e-3    # biconvex.hpp:160
    beta: float = 1.5         # fista.hpp:54
    mu: float = 1.0           # fista.hpp:60
    slice_outer: int = 0      # scheduling only (time slicing): 0 automatic, <0 off, >0 outer iterations per slice


def _arr(a, tail, name):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == len(tail):
        a = a[None]
    if a.shape[1:] != tuple(tail):
        raise ValueError(f"{name}: expected trailing shape {tuple(tail)}, got {a.shape}")
    return np.ascontiguousarray(a)


@dataclass
class CentroidalBatch:
    n_col: int
    n_eff: int
    m: np.ndarray            # [B|1]
    rho: np.ndarray          # [B|1]
    x_init: np.ndarray       # [B|1, 9]
    cnt_plan: np.ndarray     # [B|1, n_col, n_eff, 4]  rows (c, x, y, z)
    dt: np.ndarray           # [B|1, n_col]
    W_X: np.ndarray          # [B|1, 9 n_col]
    W_X_ter: np.ndarray      # [B|1, 9]
    X_nom: np.ndarray        # [B|1, 9 n_col]
    X_ter: np.ndarray        # [B|1, 9]
    W_F: np.ndarray          # [B|1, 3 n_eff n_col]
    bounds: np.ndarray       # [B|1, n_col, 6]
    L0: np.ndarray = None    # [B|1, 2] FISTA step state (L_f, L_x); fresh object: (506.25, 2.25e6)
    X0: Optional[np.ndarray] = None   # [B, nx] warm starts; None = cold start of kino_dyn.cpp:83-99
    F0: Optional[np.ndarray] = None   # [B, nf]
    P0: Optional[np.ndarray] = None   # [B, nx]
    B: int = field(default=0)

    def __post_init__(self):
        n, e = int(self.n_col), int(self.n_eff)
        nx, nf = 9 * (n + 1), 3 * e * n
        self.m = _arr(self.m, (), "m")
        self.rho = _arr(self.rho, (), "rho")
        self.x_init = _arr(self.x_init, (9,), "x_init")
        self.cnt_plan = _arr(self.cnt_plan, (n, e, 4), "cnt_plan")
        self.dt = _arr(self.dt, (n,), "dt")
        self.W_X = _arr(self.W_X, (9 * n,), "W_X")
        self.W_X_ter = _arr(self.W_X_ter, (9,), "W_X_ter")
        self.X_nom = _arr(self.X_nom, (9 * n,), "X_nom")
        self.X_ter = _arr(self.X_ter, (9,), "X_ter")
        self.W_F = _arr(self.W_F, (nf,), "W_F")
        self.bounds = _arr(self.bounds, (n, 6), "bounds")
        if self.L0 is None:
            self.L0 = np.array([[L0_F, L0_X]])
        self.L0 = _arr(self.L0, (2,), "L0")
        for nm, tail in (("X0", (nx,)), ("F0", (nf,)), ("P0", (nx,))):
            v = getattr(self, nm)
            if v is not None:
                setattr(self, nm, _arr(v, tail, nm))
        sizes = {getattr(self, f).shape[0] for f in self.field_names()
                 if getattr(self, f) is not None}
        sizes.discard(1)
        if len(sizes) > 1:
            raise ValueError(f"inconsistent batch sizes {sorted(sizes)}")
        self.B = sizes.pop() if sizes else 1

    @staticmethod
    def field_names():
        return ("m", "rho", "x_init", "cnt_plan", "dt", "W_X", "W_X_ter", "X_nom", "X_ter", "W_F",
                "bounds", "L0", "X0", "F0", "P0")

    @property
    def nx(self):
        return 9 * (self.n_col + 1)

    @property
    def nf(self):
        return 3 * self.n_eff * self.n_col

    def select(self, idx) -> "CentroidalBatch":
        """Sub-batch (instances idx); shared fields stay shared."""
        kw = {}
        for f in self.field_names():
            v = getattr(self, f)
            kw[f] = v if (v is None or v.shape[0] == 1) else v[idx]
        return CentroidalBatch(self.n_col, self.n_eff, **kw)

    def shard(self, rank: int, world: int) -> "CentroidalBatch":
        """Interleaved shard: instance i -> rank i % world (spreads the iteration-count variance)."""
        return self.select(np.arange(rank, self.B, world))

    def with_state(self, X0=None, F0=None, P0=None, L0=None) -> "CentroidalBatch":
        return replace(self, X0=X0, F0=F0, P0=P0, L0=self.L0 if L0 is None else L0)

    def input_bytes(self) -> int:
        """Bytes that cross host->device for this batch (shared fields counted once)."""
        return int(sum(getattr(self, f).nbytes for f in self.field_names() if getattr(self, f) is not None))


@dataclass
class BatchSolution:
    X: np.ndarray        # [B, nx]   return_opt_x
    F: np.ndarray        # [B, nf]   return_opt_f
    P: np.ndarray        # [B, nx]   return_opt_p
    L: np.ndarray        # [B, 2]    (L_f, L_x) after the solve
    iters: np.ndarray    # [B, 5]    outer, sum inner F, sum inner X, line-search rejections F, X
    viol: np.ndarray     # [B]       ||A_f X - b_f|| at exit
    status: np.ndarray   # [B]       0 converged, 1 max_outer reached, 2 NaN
    m: np.ndarray = None
    cycles: np.ndarray = None      # [B] SM clock cycles spent per instance (in-kernel latency)
    viol_hist: np.ndarray = None   # [B, max_outer] when requested

    def com(self):
        """return_opt_com, biconvex.cpp:122-130: [B, n+1, 3]"""
        return self.X.reshape(self.X.shape[0], -1, 9)[:, :, 0:3].copy()

    def mom(self):
        """return_opt_mom, biconvex.cpp:132-142: [B, n+1, 6] = [m*vcom, amom]"""
        if self.m is None:
            raise ValueError("BatchSolution.mom(): the masses of the batch are not known (pass m to gather_solutions)")
        Xr = self.X.reshape(self.X.shape[0], -1, 9)
        m = np.broadcast_to(np.asarray(self.m, dtype=np.float64).reshape(-1, 1, 1), (Xr.shape[0], 1, 1))
        return np.concatenate([m * Xr[:, :, 3:6], Xr[:, :, 6:9]], axis=2)
