"""pinocchio.rpy stand-in: R = Rz(y) Ry(p) Rx(r)."""
import numpy as np


def rpyToMatrix(*a):
    r, p, y = (a[0] if len(a) == 1 else a)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def matrixToRpy(R):
    R = np.asarray(R, dtype=np.float64)
    m = np.sqrt(R[2, 1] ** 2 + R[2, 2] ** 2)
    p = np.arctan2(-R[2, 0], m)
    if abs(abs(p) - np.pi / 2) < 1e-3:
        return np.array([0.0, p, -np.arctan2(R[0, 1], R[1, 1])])
    return np.array([np.arctan2(R[2, 1], R[2, 2]), p, np.arctan2(R[1, 0], R[0, 0])])
