"""`GaitPlanner` of the pybind module gait_planner_cpp (srcpy/gait_planner/py_gait_planner.cpp) over
oracle/_ref/libgait_ref.so = the reference's own src/gait_planner/gait_planner.cpp compiled in place against the Eigen
stand-in (oracle/refshim/Makefile).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_ref", "libgait_ref.so")
_lib = C.CDLL(_SO)
_lib.gait_create.restype = C.c_void_p
_lib.gait_create.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_double]
_lib.gait_destroy.argtypes = [C.c_void_p]
_lib.gait_get_phase.restype = C.c_int
_lib.gait_get_phase.argtypes = [C.c_void_p, C.c_double, C.c_int]
_lib.gait_get_percent_in_phase.restype = C.c_double
_lib.gait_get_percent_in_phase.argtypes = [C.c_void_p, C.c_double, C.c_int]
_lib.gait_get_phi.restype = C.c_double
_lib.gait_get_phi.argtypes = [C.c_void_p, C.c_double, C.c_int]


class GaitPlanner:
    def __init__(self, gait_period, stance_percent, phase_offset, step_height):
        sp = np.ascontiguousarray(stance_percent, dtype=np.float64)
        po = np.ascontiguousarray(phase_offset, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        self._h = C.c_void_p(_lib.gait_create(float(gait_period), sp.ctypes.data_as(dp), po.ctypes.data_as(dp),
                                              int(sp.size), float(step_height)))

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.gait_destroy(self._h)
            self._h = None

    def get_phase(self, t, foot_ID):
        return int(_lib.gait_get_phase(self._h, float(t), int(foot_ID)))

    def get_percent_in_phase(self, t, foot_ID):
        return float(_lib.gait_get_percent_in_phase(self._h, float(t), int(foot_ID)))

    def get_phi(self, t, foot_ID):
        return float(_lib.gait_get_phi(self._h, float(t), int(foot_ID)))
