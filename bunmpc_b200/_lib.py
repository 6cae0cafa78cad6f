"""ctypes binding of libbunmpc.so (include/bunmpc.h).

The library is built in-tree (bunmpc_b200/csrc/Makefile, `python __graft_entry__.py build`).  There is no
CPU fallback: if the shared library is missing this module raises, and every solve runs the sm_100a
kernels.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BUNMPC_LIB selects another build of the same library (profiling / experiment builds); never a fallback
LIB_PATH = os.environ.get("BUNMPC_LIB") or os.path.join(_HERE, "csrc", "libbunmpc.so")

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
CONVERGED, MAX_ITERS, NAN = 0, 1, 2
ARITH_STRICT, ARITH_FMA, ARITH_MIXED = 0, 1, 2

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class Params(C.Structure):
    _fields_ = [("max_outer", C.c_int), ("max_inner", C.c_int), ("tol", C.c_double), ("exit_tol", C.c_double),
                ("beta", C.c_double), ("mu", C.c_double), ("arith", C.c_int), ("slice_outer", C.c_int)]


class In(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("batch_stride", C.c_longlong)]


COMPACT_FIELDS = ("m", "rho", "x_init", "cnt_plan", "dt", "W_X", "W_X_ter", "X_nom", "X_ter", "W_F", "bounds",
                  "L0", "X0", "F0", "P0")
EXPANDED_FIELDS = ("m", "rho", "x_init", "cnt_plan", "dt", "Qx", "qx", "Qf", "qf", "lbx", "ubx", "L0", "X0", "F0", "P0")


class CompactProblem(C.Structure):
    _fields_ = [("batch", C.c_int)] + [(f, In) for f in COMPACT_FIELDS]


class ExpandedProblem(C.Structure):
    _fields_ = [("batch", C.c_int)] + [(f, In) for f in EXPANDED_FIELDS]


class Gait(C.Structure):
    _fields_ = [("gait_period", C.c_double), ("gait_dt", C.c_double), ("gait_horizon", C.c_double),
                ("stance_percent", C.c_double * 4), ("phase_offset", C.c_double * 4),
                ("hip_offsets", (C.c_double * 2) * 4), ("foot_size", C.c_double), ("nom_ht", C.c_double),
                ("ori_correction", C.c_double * 3), ("I_zz", C.c_double), ("W_X", C.c_double * 9),
                ("W_X_ter", C.c_double * 9), ("W_F", C.c_double * 12), ("rho", C.c_double),
                ("swing_rule", C.c_int), ("reserved_", C.c_int)]


STATE_FIELDS = ("com", "vcom", "amom", "foot_pos", "t", "v_des", "w_des", "cs_yaw", "hip_xy", "amom_des", "scales")


class AcyclicMotion(C.Structure):
    _fields_ = [("n_cnt", C.c_int), ("n_nom", C.c_int), ("n_box", C.c_int), ("dt_arr", C.c_void_p), ("cnt_plan", C.c_void_p),
                ("X_nom", C.c_void_p), ("bounds", C.c_void_p), ("X_ter", C.c_void_p), ("t0", C.c_double)]


class States(C.Structure):
    _fields_ = [("batch", C.c_int)] + [(f, In) for f in STATE_FIELDS]


class Solution(C.Structure):
    _fields_ = [("X", C.c_void_p), ("F", C.c_void_p), ("P", C.c_void_p), ("L", C.c_void_p), ("iters", C.c_void_p),
                ("viol", C.c_void_p), ("status", C.c_void_p), ("viol_hist", C.c_void_p), ("cycles", C.c_void_p)]


# every symbol include/bunmpc.h declares
EXPORTS = ("bunmpc_version", "bunmpc_last_error", "bunmpc_default_params", "bunmpc_create", "bunmpc_destroy",
           "bunmpc_launch_count", "bunmpc_kernel_info", "bunmpc_expand_device", "bunmpc_solve_expanded_device",
           "bunmpc_solve_compact_device", "bunmpc_build_problem_device", "bunmpc_build_acyclic_device", "bunmpc_solve_compact_host", "bunmpc_solve_expanded_host",
           "bunmpc_goal_stats_device", "bunmpc_job_counter_create", "bunmpc_job_counter_open", "bunmpc_job_counter_release",
           "bunmpc_set_job_counter", "bunmpc_peer_buffer_create", "bunmpc_peer_buffer_open", "bunmpc_peer_buffer_release",
           "bunmpc_set_peer_results", "bunmpc_centroidal_mats_host", "bunmpc_measure_fp64_peak", "bunmpc_selftest_division", "bunmpc_host_alloc", "bunmpc_host_free")

_lib = None


class BunmpcError(RuntimeError):
    pass


def lib():
    """Load libbunmpc.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BunmpcError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                          "(make -C bunmpc_b200/csrc); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.bunmpc_version.restype = C.c_int
    L.bunmpc_last_error.restype = C.c_char_p
    L.bunmpc_default_params.argtypes = [C.POINTER(Params)]
    L.bunmpc_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]
    L.bunmpc_destroy.argtypes = [C.c_void_p]
    L.bunmpc_destroy.restype = None
    L.bunmpc_launch_count.argtypes = [C.c_void_p]
    L.bunmpc_launch_count.restype = C.c_longlong
    L.bunmpc_kernel_info.argtypes = [C.c_void_p, ip, ip, ip, ip]
    L.bunmpc_expand_device.argtypes = [C.c_void_p, C.POINTER(CompactProblem)] + [C.c_void_p] * 7
    L.bunmpc_solve_expanded_device.argtypes = [C.c_void_p, C.POINTER(ExpandedProblem), C.POINTER(Params),
                                               C.POINTER(Solution), C.c_void_p]
    L.bunmpc_solve_compact_device.argtypes = [C.c_void_p, C.POINTER(CompactProblem), C.POINTER(Params),
                                              C.POINTER(Solution), C.c_void_p]
    L.bunmpc_build_problem_device.argtypes = [C.c_void_p, C.POINTER(Gait), C.POINTER(States)] + [C.c_void_p] * 10
    L.bunmpc_build_acyclic_device.argtypes = [C.c_void_p, C.POINTER(AcyclicMotion), C.c_int, C.POINTER(In), C.POINTER(In)] + [C.c_void_p] * 6
    L.bunmpc_solve_compact_host.argtypes = [C.c_void_p, C.POINTER(CompactProblem), C.POINTER(Params),
                                            C.POINTER(Solution)]
    L.bunmpc_solve_expanded_host.argtypes = [C.c_void_p, C.POINTER(ExpandedProblem), C.POINTER(Params),
                                             C.POINTER(Solution)]
    L.bunmpc_goal_stats_device.argtypes = [C.c_void_p, C.c_int, C.POINTER(In), C.POINTER(In), C.c_void_p, C.c_void_p]
    L.bunmpc_job_counter_create.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_char_p]
    L.bunmpc_job_counter_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    L.bunmpc_job_counter_release.argtypes = [C.c_void_p, C.c_int]
    L.bunmpc_set_job_counter.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.bunmpc_peer_buffer_create.argtypes = [C.c_int, C.c_ulonglong, C.POINTER(C.c_void_p), C.c_char_p]
    L.bunmpc_peer_buffer_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    L.bunmpc_peer_buffer_release.argtypes = [C.c_void_p, C.c_int]
    L.bunmpc_set_peer_results.argtypes = [C.c_void_p, C.c_int, C.POINTER(Solution)]
    L.bunmpc_centroidal_mats_host.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 9
    L.bunmpc_measure_fp64_peak.argtypes = [C.c_void_p, dp]
    L.bunmpc_selftest_division.argtypes = [C.c_void_p, C.c_longlong, C.c_ulonglong, C.POINTER(C.c_longlong)]
    L.bunmpc_host_alloc.argtypes = [C.c_ulonglong]
    L.bunmpc_host_alloc.restype = C.c_void_p
    L.bunmpc_host_free.argtypes = [C.c_void_p]
    L.bunmpc_host_free.restype = None
    _lib = L
    return L


def check(rc: int, what: str = "bunmpc"):
    if rc != OK:
        msg = lib().bunmpc_last_error()
        raise BunmpcError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
