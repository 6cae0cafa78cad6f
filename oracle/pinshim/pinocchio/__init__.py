"""Stand-in for the parts of pinocchio that examples/mpc/abstract_cyclic_gen.py touches.  TEST INFRASTRUCTURE ONLY
(oracle/pinshim/README.md).  Kinematic quantities are INJECTED through FakeRobot; the rotation helpers are real maths."""
import numpy as np

from . import rpy  # noqa: F401
from . import utils  # noqa: F401


class Quaternion:
    """Eigen::Quaterniond as pinocchio binds it: constructed from a 4-vector (x, y, z, w) or a 3x3 rotation."""

    def __init__(self, a):
        a = np.asarray(a, dtype=np.float64)
        if a.shape == (3, 3):
            # Eigen's QuaternionBase::operator=(MatrixBase) (Shepperd)
            t = a[0, 0] + a[1, 1] + a[2, 2]
            if t > 0.0:
                t = np.sqrt(t + 1.0)
                w = 0.5 * t
                t = 0.5 / t
                x, y, z = (a[2, 1] - a[1, 2]) * t, (a[0, 2] - a[2, 0]) * t, (a[1, 0] - a[0, 1]) * t
            else:
                i = 0
                if a[1, 1] > a[0, 0]:
                    i = 1
                if a[2, 2] > a[i, i]:
                    i = 2
                j, k = (i + 1) % 3, (i + 2) % 3
                t = np.sqrt(a[i, i] - a[j, j] - a[k, k] + 1.0)
                q = [0.0, 0.0, 0.0]
                q[i] = 0.5 * t
                t = 0.5 / t
                w = (a[k, j] - a[j, k]) * t
                q[j] = (a[j, i] + a[i, j]) * t
                q[k] = (a[k, i] + a[i, k]) * t
                x, y, z = q
            self.x, self.y, self.z, self.w = float(x), float(y), float(z), float(w)
        else:
            self.x, self.y, self.z, self.w = (float(v) for v in a.reshape(4))

    def coeffs(self):
        return np.array([self.x, self.y, self.z, self.w])

    def toRotationMatrix(self):
        x, y, z, w = self.x, self.y, self.z, self.w
        tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
        twx, twy, twz = tx * w, ty * w, tz * w
        txx, txy, txz = tx * x, ty * x, tz * x
        tyy, tyz, tzz = ty * y, tz * y, tz * z
        return np.array([[1.0 - (tyy + tzz), txy - twz, txz + twy],
                         [txy + twz, 1.0 - (txx + tzz), tyz - twx],
                         [txz - twy, tyz + twx, 1.0 - (txx + tyy)]])

    def inverse(self):
        n2 = self.x * self.x + self.y * self.y + self.z * self.z + self.w * self.w
        return Quaternion([-self.x / n2, -self.y / n2, -self.z / n2, self.w / n2])

    def __mul__(self, o):
        a, b = self, o
        return Quaternion([a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                           a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
                           a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
                           a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z])


def log3(R):
    """Rotation vector of a rotation matrix."""
    R = np.asarray(R, dtype=np.float64)
    c = 0.5 * (np.trace(R) - 1.0)
    c = min(1.0, max(-1.0, c))
    theta = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if theta < 1e-9:
        return 0.5 * v
    return (theta / (2.0 * np.sin(theta))) * v


# ---- kinematics: the builder only reads a handful of numbers, which the test harness injects ----
class _SE3:
    def __init__(self, translation):
        self.translation = np.asarray(translation, dtype=np.float64)


class _Inertia:
    def __init__(self, I):
        self.inertia = np.asarray(I, dtype=np.float64)


class FakeModel:
    def __init__(self, frame_names, nv):
        self._ids = {name: i for i, name in enumerate(frame_names)}
        self.nv = nv
        self.nq = nv + 1

    def getFrameId(self, name):
        return self._ids[name]


class FakeData:
    def __init__(self, n_frames):
        self.oMf = [_SE3(np.zeros(3)) for _ in range(n_frames)]
        self.hg = np.zeros(6)               # np.array(rdata.hg): [linear, angular] centroidal momentum
        self.Ycrb = [None, _Inertia(np.zeros((3, 3)))]
        self.com = np.zeros(3)


class FakeRobot:
    """robot.model / robot.data with injected frame placements, CoM, centroidal momentum and composite inertia."""
    FRAMES = ("FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT", "FL_HFE", "FR_HFE", "HL_HFE", "HR_HFE")

    def __init__(self, mass, nv=18):
        self.model = FakeModel(self.FRAMES, nv)
        self.model.mass = float(mass)
        self.data = FakeData(len(self.FRAMES))

    def inject(self, com=None, foot_pos=None, hip_pos=None, hg=None, I_composite=None):
        d = self.data
        if com is not None:
            d.com = np.asarray(com, dtype=np.float64).copy()
        if foot_pos is not None:
            for j in range(4):
                d.oMf[j] = _SE3(np.asarray(foot_pos[j], dtype=np.float64).copy())
        if hip_pos is not None:
            for j in range(4):
                d.oMf[4 + j] = _SE3(np.asarray(hip_pos[j], dtype=np.float64).copy())
        if hg is not None:
            d.hg = np.asarray(hg, dtype=np.float64).copy()
        if I_composite is not None:
            d.Ycrb[1] = _Inertia(I_composite)


# ---- AbstractGaitGen (examples/mpc/abstract_cyclic_gen1.py:28-29) builds its own model from the urdf path: the harness
# registers the FakeRobot that "is" that urdf beforehand ----
_URDF_ROBOTS = {}


def register_urdf(path, robot):
    _URDF_ROBOTS[path] = robot


def JointModelFreeFlyer():
    return None


def buildModelFromUrdf(path, root_joint=None):
    robot = _URDF_ROBOTS[path]
    robot.model.createData = lambda: robot.data
    return robot.model


def forwardKinematics(model, data, q, v=None):
    pass


def updateFramePlacements(model, data):
    pass


def framesForwardKinematics(model, data, q):
    pass


def crba(model, data, q):
    pass


def computeCentroidalMomentum(model, data, *a):
    return data.hg


def centerOfMass(model, data, q=None, v=None):
    return data.com.copy()


def computeTotalMass(model):
    return model.mass


def normalize(model, q):
    q = np.array(q, dtype=np.float64)
    q[3:7] /= np.linalg.norm(q[3:7])
    return q
