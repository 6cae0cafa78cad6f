"""Batched entry point over the C ABI: `BatchSolver`, `solve_batch`.

`BatchSolver.solve(batch)` is the host path (numpy in, numpy out: copies + kernels + copies, what a caller
of the reference's `BiconvexMP.optimize` sees).  `BatchSolver.upload(batch)` / `solve_resident(dev)` keep
inputs and outputs as torch CUDA tensors and launch on torch's current stream; torch is used only for
device memory and streams, the kernels are libbunmpc.so's.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import _lib
from .problem import BatchSolution, CentroidalBatch, SolverParams


def _c_params(prm: Optional[SolverParams], arith: int) -> _lib.Params:
    prm = prm or SolverParams()
    return _lib.Params(int(prm.max_outer), int(prm.max_inner), float(prm.tol), float(prm.exit_tol),
                       float(prm.beta), float(prm.mu), int(arith), int(prm.slice_outer))


def _in_np(a: Optional[np.ndarray], width: int) -> _lib.In:
    if a is None:
        return _lib.In(None, 0)
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return _lib.In(a.ctypes.data, 0 if a.shape[0] == 1 else width)


@dataclass
class DeviceBatch:
    """A CentroidalBatch resident in HBM (torch CUDA tensors) plus its output tensors."""
    B: int
    fields: Dict[str, "object"]      # name -> torch tensor or None
    out: Dict[str, "object"]
    widths: Dict[str, int]


class BatchSolver:
    """One solver handle = BiConvexMP::BiConvexMP(m, n_col, n_eff) for up to max_batch instances on one GPU
    (biconvex.cpp:6-25).  All device memory is allocated here; solves allocate nothing."""

    def __init__(self, n_col: int, n_eff: int = 4, max_batch: int = 1024, device: int = 0):
        self.n_col, self.n_eff, self.max_batch, self.device = int(n_col), int(n_eff), int(max_batch), int(device)
        self.nx, self.nf = 9 * (self.n_col + 1), 3 * self.n_eff * self.n_col
        self._h = C.c_void_p()
        _lib.check(_lib.lib().bunmpc_create(C.byref(self._h), self.device, self.n_col, self.n_eff, self.max_batch),
                   "bunmpc_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().bunmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection ----
    def launch_count(self) -> int:
        return int(_lib.lib().bunmpc_launch_count(self._h))

    def kernel_info(self) -> dict:
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib().bunmpc_kernel_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(ctas_per_sm=a.value, threads=b.value, smem_bytes=c.value, num_sms=d.value)

    def measure_fp64_peak(self) -> float:
        """TFLOP/s of a register-only DFMA micro-benchmark on this GPU (bench.py's FP64 roofline denominator)."""
        v = C.c_double()
        _lib.check(_lib.lib().bunmpc_measure_fp64_peak(self._h, C.byref(v)), "bunmpc_measure_fp64_peak")
        return float(v.value)

    def selftest_division(self, n_pairs: int = 1 << 24, seed: int = 1) -> int:
        """Mismatches between the kernel's hoisted-reciprocal division and IEEE `/` on n_pairs operand pairs."""
        v = C.c_longlong()
        _lib.check(_lib.lib().bunmpc_selftest_division(self._h, int(n_pairs), int(seed), C.byref(v)), "selftest")
        return int(v.value)

    def _widths(self):
        n, e, nx, nf = self.n_col, self.n_eff, self.nx, self.nf
        return dict(m=1, rho=1, x_init=9, cnt_plan=4 * e * n, dt=n, W_X=9 * n, W_X_ter=9, X_nom=9 * n, X_ter=9,
                    W_F=nf, bounds=6 * n, L0=2, X0=nx, F0=nf, P0=nx, Qx=nx, qx=nx, Qf=nf, qf=nf, lbx=nx, ubx=nx)

    def _check_batch(self, batch: CentroidalBatch):
        if batch.n_col != self.n_col or batch.n_eff != self.n_eff:
            raise ValueError("batch horizon / end-effector count does not match the solver")
        if batch.B > self.max_batch:
            raise ValueError(f"batch {batch.B} > max_batch {self.max_batch}")

    # ---- host path ----
    def solve(self, batch: CentroidalBatch, params: Optional[SolverParams] = None, arith: int = _lib.ARITH_STRICT,
              viol_hist: bool = False, out: Optional[dict] = None) -> BatchSolution:
        """create_bound_constraints + create_cost_X + create_cost_F + optimize for every instance
        (abstract_cyclic_gen.py:611-614,663), host buffers in and out."""
        self._check_batch(batch)
        B, nx, nf = batch.B, self.nx, self.nf
        w = self._widths()
        prob = _lib.CompactProblem()
        prob.batch = B
        keep = []
        for f in _lib.COMPACT_FIELDS:
            a = getattr(batch, f)
            if a is not None:
                a = np.ascontiguousarray(a.reshape(a.shape[0], -1), dtype=np.float64)
                keep.append(a)
            setattr(prob, f, _in_np(a, w[f]))
        prm = _c_params(params, arith)
        o = out if out is not None else {}
        X = o.get("X", np.empty((B, nx))); F = o.get("F", np.empty((B, nf))); P = o.get("P", np.empty((B, nx)))
        L = o.get("L", np.empty((B, 2))); viol = o.get("viol", np.empty(B))
        iters = o.get("iters", np.empty((B, 5), dtype=np.int32)); status = o.get("status", np.empty(B, dtype=np.int32))
        hist = np.empty((B, prm.max_outer)) if viol_hist else None
        cycles = o.get("cycles", np.empty(B, dtype=np.int64))
        sol = _lib.Solution(X.ctypes.data, F.ctypes.data, P.ctypes.data, L.ctypes.data, iters.ctypes.data,
                            viol.ctypes.data, status.ctypes.data, hist.ctypes.data if viol_hist else None,
                            cycles.ctypes.data)
        _lib.check(_lib.lib().bunmpc_solve_compact_host(self._h, C.byref(prob), C.byref(prm), C.byref(sol)),
                   "bunmpc_solve_compact_host")
        res = BatchSolution(X=X, F=F, P=P, L=L, iters=iters, viol=viol, status=status, m=batch.m)
        res.cycles = cycles
        if viol_hist:
            res.viol_hist = hist
        return res

    def solve_expanded(self, m, rho, x_init, cnt_plan, dt, Qx, qx, Qf, qf, lbx, ubx, X0, F0, P0, L0,
                       params: Optional[SolverParams] = None, arith: int = _lib.ARITH_STRICT,
                       viol_hist: bool = False) -> BatchSolution:
        """optimize() on already expanded costs/bounds (set_cost_x/f + set_bounds_x path), host buffers."""
        w = self._widths()
        vals = dict(m=m, rho=rho, x_init=x_init, cnt_plan=cnt_plan, dt=dt, Qx=Qx, qx=qx, Qf=Qf, qf=qf, lbx=lbx,
                    ubx=ubx, L0=L0, X0=X0, F0=F0, P0=P0)
        arrs, B = {}, 1
        for k, v in vals.items():
            a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
            a = a.reshape(1, -1) if a.size == w[k] else a.reshape(-1, w[k])
            arrs[k] = a
            B = max(B, a.shape[0])
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch}")
        prob = _lib.ExpandedProblem()
        prob.batch = B
        for k in _lib.EXPANDED_FIELDS:
            setattr(prob, k, _in_np(arrs[k], w[k]))
        prm = _c_params(params, arith)
        nx, nf = self.nx, self.nf
        X, F, P, L, viol = np.empty((B, nx)), np.empty((B, nf)), np.empty((B, nx)), np.empty((B, 2)), np.empty(B)
        iters, status = np.empty((B, 5), dtype=np.int32), np.empty(B, dtype=np.int32)
        hist = np.empty((B, prm.max_outer)) if viol_hist else None
        sol = _lib.Solution(X.ctypes.data, F.ctypes.data, P.ctypes.data, L.ctypes.data, iters.ctypes.data,
                            viol.ctypes.data, status.ctypes.data, hist.ctypes.data if viol_hist else None, None)
        _lib.check(_lib.lib().bunmpc_solve_expanded_host(self._h, C.byref(prob), C.byref(prm), C.byref(sol)),
                   "bunmpc_solve_expanded_host")
        res = BatchSolution(X=X, F=F, P=P, L=L, iters=iters, viol=viol, status=status, m=arrs["m"])
        if viol_hist:
            res.viol_hist = hist
        return res

    def centroidal_mats(self, m, cnt_plan, dt, X=None, F=None, x_init=None):
        """return_A_x/b_x (needs X) and return_A_f/b_f (needs F, x_init), biconvex.hpp:30-51, one instance."""
        nx, nf = self.nx, self.nf
        cnt = np.ascontiguousarray(cnt_plan, dtype=np.float64)
        dta = np.ascontiguousarray(dt, dtype=np.float64)
        out = {}
        Xa = Fa = xa = None
        ptr = lambda a: a.ctypes.data if a is not None else None
        if X is not None:
            Xa = np.ascontiguousarray(X, dtype=np.float64)
            out["A_x"], out["b_x"] = np.empty((nx, nf)), np.empty(nx)
        if F is not None:
            Fa = np.ascontiguousarray(F, dtype=np.float64)
            xa = np.ascontiguousarray(x_init, dtype=np.float64)
            out["A_f"], out["b_f"] = np.empty((nx, nx)), np.empty(nx)
        _lib.check(_lib.lib().bunmpc_centroidal_mats_host(
            self._h, float(m), ptr(cnt), ptr(dta), ptr(Xa), ptr(Fa), ptr(xa),
            ptr(out.get("A_x")), ptr(out.get("b_x")), ptr(out.get("A_f")), ptr(out.get("b_f"))), "centroidal_mats")
        return out

    # ---- device-resident path (torch tensors) ----
    def upload(self, batch: CentroidalBatch, pin: bool = False) -> DeviceBatch:
        import torch
        self._check_batch(batch)
        dev = torch.device("cuda", self.device)
        w = self._widths()
        fields = {}
        for f in _lib.COMPACT_FIELDS:
            a = getattr(batch, f)
            fields[f] = None if a is None else torch.from_numpy(
                np.ascontiguousarray(a.reshape(a.shape[0], -1))).to(dev)
        B = batch.B
        out = dict(X=torch.empty((B, self.nx), dtype=torch.float64, device=dev),
                   F=torch.empty((B, self.nf), dtype=torch.float64, device=dev),
                   P=torch.empty((B, self.nx), dtype=torch.float64, device=dev),
                   L=torch.empty((B, 2), dtype=torch.float64, device=dev),
                   iters=torch.empty((B, 5), dtype=torch.int32, device=dev),
                   viol=torch.empty((B,), dtype=torch.float64, device=dev),
                   status=torch.empty((B,), dtype=torch.int32, device=dev),
                   cycles=torch.empty((B,), dtype=torch.int64, device=dev))
        return DeviceBatch(B=B, fields=fields, out=out, widths=w)

    def build_device(self, robot, params, com, vcom, amom, foot_pos, t, v_des, w_des, yaw=0.0, amom_des=None,
                     scales=None, L0=None, hip_xy=None, swing_rule=0) -> DeviceBatch:
        """Batched problem builder ON THE DEVICE: contact plan (create_cnt_plan,
        abstract_cyclic_gen.py:159-414) and nominal/terminal references (create_costs, :564-614) from centroidal
        states; only the states cross PCIe.  Same results, bit for bit, as plan_builder.build_batch (numpy).
        swing_rule: plan_builder.SWING_RULE_* (which of the reference's two cyclic generators)."""
        import torch
        from .problem import L0_F, L0_X
        dev = torch.device("cuda", self.device)
        com = np.atleast_2d(np.asarray(com, dtype=np.float64))
        B, n, e = com.shape[0], self.n_col, self.n_eff
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch}")
        if params.horizon() != n:
            raise ValueError("gait horizon does not match the solver")

        def up(a, tail):
            a = np.array(np.broadcast_to(np.asarray(a, dtype=np.float64), (B,) + tuple(tail)), order="C", copy=True)
            return torch.from_numpy(a.reshape(B, -1)).to(dev)

        yaw = np.broadcast_to(np.asarray(yaw, dtype=np.float64), (B,))
        st_t = dict(com=up(com, (3,)), vcom=up(vcom, (3,)), amom=up(amom, (3,)), foot_pos=up(foot_pos, (e, 3)),
                    t=up(t, ()), v_des=up(v_des, (3,)), w_des=up(w_des, ()),
                    cs_yaw=up(np.stack([np.cos(yaw), np.sin(yaw)], axis=1), (2,)),
                    hip_xy=None if hip_xy is None else up(hip_xy, (e, 2)),
                    amom_des=None if amom_des is None else up(amom_des, (3,)),
                    scales=None if scales is None else up(scales, (3,)))
        g = _lib.Gait()
        g.gait_period, g.gait_dt, g.gait_horizon = params.gait_period, params.gait_dt, params.gait_horizon
        for j in range(4):
            g.stance_percent[j] = params.stance_percent[j]
            g.phase_offset[j] = params.phase_offset[j]
            g.hip_offsets[j][0], g.hip_offsets[j][1] = robot.hip_offsets[j][0], robot.hip_offsets[j][1]
        g.foot_size, g.nom_ht, g.I_zz, g.rho = robot.foot_size, params.nom_ht, robot.I_zz, params.rho
        g.swing_rule = int(swing_rule)
        for k in range(3):
            g.ori_correction[k] = params.ori_correction[k]
        for k in range(9):
            g.W_X[k], g.W_X_ter[k] = params.W_X[k], params.W_X_ter[k]
        for k in range(12):
            g.W_F[k] = params.W_F[k]
        st = _lib.States()
        st.batch = B
        widths = dict(com=3, vcom=3, amom=3, foot_pos=3 * e, t=1, v_des=3, w_des=1, cs_yaw=2, hip_xy=2 * e, amom_des=3, scales=3)
        for f in _lib.STATE_FIELDS:
            tt = st_t[f]
            setattr(st, f, _lib.In(None, 0) if tt is None else _lib.In(tt.data_ptr(), widths[f]))
        f64 = dict(dtype=torch.float64, device=dev)
        out = dict(x_init=torch.empty((B, 9), **f64), cnt_plan=torch.empty((B, n * e * 4), **f64),
                   dt=torch.empty((B, n), **f64), X_nom=torch.empty((B, 9 * n), **f64), X_ter=torch.empty((B, 9), **f64))
        if scales is not None:
            out.update(W_X=torch.empty((B, 9 * n), **f64), W_X_ter=torch.empty((B, 9), **f64),
                       W_F=torch.empty((B, self.nf), **f64), rho=torch.empty((B, 1), **f64))
        ptr = lambda k: C.c_void_p(out[k].data_ptr()) if k in out else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().bunmpc_build_problem_device(
            self._h, C.byref(g), C.byref(st), ptr("x_init"), ptr("cnt_plan"), ptr("dt"), ptr("X_nom"), ptr("X_ter"),
            ptr("W_X"), ptr("W_X_ter"), ptr("W_F"), ptr("rho"), C.c_void_p(stream)), "bunmpc_build_problem_device")
        one = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(1, -1))).to(dev)
        fields = dict(out)
        fields["m"] = one([robot.mass])
        if scales is None:
            fields.update(rho=one([params.rho]), W_X=one(np.tile(params.W_X, n)), W_X_ter=one(params.W_X_ter),
                          W_F=one(np.tile(params.W_F, n)))
        fields["bounds"] = one(np.tile([-robot.bx, -robot.by, 0, robot.bx, robot.by, robot.bz], n))
        fields["L0"] = one([L0_F, L0_X]) if L0 is None else up(L0, (2,))
        fields.update(X0=None, F0=None, P0=None)
        res = dict(X=torch.empty((B, self.nx), **f64), F=torch.empty((B, self.nf), **f64),
                   P=torch.empty((B, self.nx), **f64), L=torch.empty((B, 2), **f64),
                   iters=torch.empty((B, 5), dtype=torch.int32, device=dev), viol=torch.empty((B,), **f64),
                   status=torch.empty((B,), dtype=torch.int32, device=dev),
                   cycles=torch.empty((B,), dtype=torch.int64, device=dev))
        self._keep = st_t      # keep the state tensors alive until the (asynchronous) kernel has consumed them
        return DeviceBatch(B=B, fields=fields, out=res, widths=self._widths())

    def build_acyclic_device(self, motion, x_init, t, t0=0.0, L0=None) -> DeviceBatch:
        """The acyclic generator's problem builder ON THE DEVICE (SoloAcyclicGen.create_contact_plan / create_costs,
        abstract_acyclic_gen.py:74-190) for B replans: only x_init [B,9] and the replanning instants t [B] cross PCIe,
        the motion's time tables are uploaded once per (solver, motion).  Same results, bit for bit, as
        acyclic.build_batch (numpy)."""
        import torch
        from .problem import L0_F, L0_X
        dev = torch.device("cuda", self.device)
        x_init = np.atleast_2d(np.asarray(x_init, dtype=np.float64))
        B, n, e = x_init.shape[0], self.n_col, self.n_eff
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch}")
        if int(motion.n_col) != n:
            raise ValueError("the motion's number of knots does not match the solver")
        f64 = dict(dtype=torch.float64, device=dev)
        one = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(1, -1))).to(dev)
        cache = self.__dict__.setdefault("_acyclic_tables", {})
        arrs = [np.ascontiguousarray(np.asarray(getattr(motion, k), dtype=np.float64)) for k in
                ("dt_arr", "cnt_plan", "X_nom", "bounds", "X_ter", "W_X", "W_X_ter", "W_F")]
        key = hash(tuple(a.tobytes() for a in arrs) + (float(motion.rho), float(motion.mass)))   # the record may be edited
        tab = cache.get(key)
        if tab is None:
            tab = dict(dt_arr=one(motion.dt_arr), cnt=one(motion.cnt_plan), nom=one(motion.X_nom), box=one(motion.bounds),
                       X_ter=one(motion.X_ter), rho=one([motion.rho]), m=one([motion.mass]), W_X=one(np.tile(motion.W_X, n)),
                       W_X_ter=one(motion.W_X_ter), W_F=one(np.tile(motion.W_F, n)), motion=motion)
            cache[key] = tab
        xs = torch.from_numpy(np.ascontiguousarray(x_init)).to(dev)
        ts = torch.from_numpy(np.array(np.broadcast_to(np.asarray(t, dtype=np.float64), (B,)), order="C", copy=True)).to(dev)
        m = _lib.AcyclicMotion()
        m.n_cnt, m.n_nom, m.n_box = len(motion.cnt_plan), len(motion.X_nom), len(motion.bounds)
        m.dt_arr, m.cnt_plan, m.X_nom = tab["dt_arr"].data_ptr(), tab["cnt"].data_ptr(), tab["nom"].data_ptr()
        m.bounds, m.X_ter, m.t0 = tab["box"].data_ptr(), tab["X_ter"].data_ptr(), float(t0)
        out = dict(cnt_plan=torch.empty((B, n * e * 4), **f64), dt=torch.empty((B, n), **f64),
                   X_nom=torch.empty((B, 9 * n), **f64), X_ter=torch.empty((B, 9), **f64), bounds=torch.empty((B, 6 * n), **f64))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        xin, tin = _lib.In(xs.data_ptr(), 9), _lib.In(ts.data_ptr(), 1)
        _lib.check(_lib.lib().bunmpc_build_acyclic_device(
            self._h, C.byref(m), B, C.byref(xin), C.byref(tin), *(C.c_void_p(out[k].data_ptr()) for k in
                                                               ("cnt_plan", "dt", "X_nom", "X_ter", "bounds")),
            C.c_void_p(stream)), "bunmpc_build_acyclic_device")
        fields = dict(out)
        fields.update(x_init=xs, m=tab["m"], rho=tab["rho"], W_X=tab["W_X"], W_X_ter=tab["W_X_ter"], W_F=tab["W_F"],
                      L0=one([L0_F, L0_X]) if L0 is None else torch.from_numpy(np.array(
                          np.broadcast_to(np.asarray(L0, dtype=np.float64), (B, 2)), order="C", copy=True)).to(dev),
                      X0=None, F0=None, P0=None)
        res = dict(X=torch.empty((B, self.nx), **f64), F=torch.empty((B, self.nf), **f64),
                   P=torch.empty((B, self.nx), **f64), L=torch.empty((B, 2), **f64),
                   iters=torch.empty((B, 5), dtype=torch.int32, device=dev), viol=torch.empty((B,), **f64),
                   status=torch.empty((B,), dtype=torch.int32, device=dev),
                   cycles=torch.empty((B,), dtype=torch.int64, device=dev))
        self._keep = (xs, ts)
        return DeviceBatch(B=B, fields=fields, out=res, widths=self._widths())

    def solve_resident(self, dev: DeviceBatch, params: Optional[SolverParams] = None,
                       arith: int = _lib.ARITH_STRICT):
        """Expand + solve on device-resident inputs, asynchronous on torch's current stream."""
        import torch
        prob = _lib.CompactProblem()
        prob.batch = dev.B
        for f in _lib.COMPACT_FIELDS:
            t = dev.fields[f]
            setattr(prob, f, _lib.In(None, 0) if t is None else
                    _lib.In(t.data_ptr(), 0 if t.shape[0] == 1 else dev.widths[f]))
        o = dev.out
        sol = _lib.Solution(o["X"].data_ptr(), o["F"].data_ptr(), o["P"].data_ptr(), o["L"].data_ptr(),
                            o["iters"].data_ptr(), o["viol"].data_ptr(), o["status"].data_ptr(), None,
                            o["cycles"].data_ptr())
        prm = _c_params(params, arith)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().bunmpc_solve_compact_device(self._h, C.byref(prob), C.byref(prm), C.byref(sol),
                                                          C.c_void_p(stream)), "bunmpc_solve_compact_device")
        return o


_SOLVERS: Dict[tuple, BatchSolver] = {}


def get_solver(n_col: int, n_eff: int = 4, min_batch: int = 1, device: int = 0) -> BatchSolver:
    """Cached solver handles keyed by (device, n_col, n_eff); regrown when a larger batch arrives."""
    key = (device, n_col, n_eff)
    s = _SOLVERS.get(key)
    if s is None or s.max_batch < min_batch:
        if s is not None:
            s.close()
        cap = 1
        while cap < min_batch:
            cap *= 2
        s = BatchSolver(n_col, n_eff, max_batch=cap, device=device)
        _SOLVERS[key] = s
    return s


def solve_batch(batch: CentroidalBatch, params: Optional[SolverParams] = None, arith: int = _lib.ARITH_STRICT,
                device: int = 0, viol_hist: bool = False) -> BatchSolution:
    """The batched replacement of `for each instance: BiconvexMP(...).optimize(x_init, n)`."""
    return get_solver(batch.n_col, batch.n_eff, batch.B, device).solve(batch, params, arith, viol_hist)
