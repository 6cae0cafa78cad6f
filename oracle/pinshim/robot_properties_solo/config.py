"""Stand-in for robot_properties_solo.config (TEST INFRASTRUCTURE ONLY, oracle/pinshim/README.md): the reference's
motion files (examples/motions/cyclic/solo12_*.py) call Solo12Config.buildRobotWrapper() at import time and read
model.nv from it."""
import pinocchio as pin


class Solo12Config:
    urdf_path = "solo12.urdf"
    mass = 2.5          # sum of <mass> in robots/solo12/urdf/solo12.urdf
    # the acyclic motion files (examples/motions/acyclic/*.py) read it for their IK regularisation targets only
    # (robot_properties_solo/config.py:246-252); nothing on the centroidal path depends on its value
    initial_configuration = [0.2, 0.0, 0.25, 0.0, 0.0, 0.0, 1.0] + 2 * [0.0, 0.8, -1.6] + 2 * [0.0, -0.8, 1.6]

    @classmethod
    def buildRobotWrapper(cls):
        return pin.FakeRobot(cls.mass, nv=18)
