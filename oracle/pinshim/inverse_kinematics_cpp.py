"""inert stand-in of the IK pybind module (whole-body IK is out of scope)"""


class InverseKinematics:
    def __getattr__(self, name):
        return lambda *a, **k: None
