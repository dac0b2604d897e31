"""LEAP hand (/root/reference/mgs/gripper/leap.py:370-454)."""
from typing import List

import numpy as np

from ..util.geo.transforms import SE3Pose
from .base import MjGripper


class GripperLeap(MjGripper):
    ASSET_DIR = "leap"
    COMPUTE_F64 = True
    REPOSE_ON_CLOSE = 1  # close_gripper_at calls set_pose first (leap.py:406-409)

    def __init__(self, pose: SE3Pose):
        super().__init__(pose, "palm")
        # leap.py:373-392
        self.close_pose = np.array([0.576, 0.0, 1.43, 0.453, 0.856, 0.0, 0.68, 0.826, 0.945, 0.0, 1.3, 0.2, 1.81, 0.258, 0.505, 0.351])

    def base_to_contact_transform(self) -> SE3Pose:  # leap.py:394-398
        return SE3Pose(np.array([0.0, 0.0, 0.0]), np.array([1.0, 0.0, 0.0, 0.0]), type="wxyz")

    def get_actuator_joint_names(self) -> List[str]:  # leap.py:432-450
        return [f"{f}_{j}" for f, js in (("if", ("mcp", "rot", "pip", "dip")), ("mf", ("mcp", "rot", "pip", "dip")),
                                         ("rf", ("mcp", "rot", "pip", "dip")), ("th", ("cmc", "axl", "mcp", "ipl"))) for j in js]

    def open_gripper(self, sim):  # leap.py:400-401: a no-op in the reference
        return

    def close_ctrl(self) -> np.ndarray:
        return np.copy(self.close_pose)
