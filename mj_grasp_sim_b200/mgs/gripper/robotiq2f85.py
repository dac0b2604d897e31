"""Robotiq 2F-85 (/root/reference/mgs/gripper/robotiq2f85.py:228-284)."""
from typing import List

import numpy as np

from ..util.geo.transforms import SE3Pose
from .base import MjGripper


class GripperRobotiq2f85(MjGripper):
    ASSET_DIR = "robotiq2f85"

    def __init__(self, pose: SE3Pose):
        super().__init__(pose, "base_mount")

    def base_to_contact_transform(self) -> SE3Pose:  # robotiq2f85.py:232-235
        return SE3Pose(np.array([0.0, 0.0, -0.15]), np.array([1, 0, 0, 0]), type="wxyz")

    def get_actuator_joint_names(self) -> List[str]:
        # robotiq2f85.py:271-281 verbatim, including the two names that do not exist in the model
        # ("right_spring_link", "left_spring_link"): get_joint_idxs maps them to the LAST joint's address
        return ["right_driver_joint", "right_coupler_joint", "right_spring_link", "right_follower_joint",
                "left_driver_joint", "left_coupler_joint", "left_spring_link", "left_follower_joint"]

    def close_ctrl(self) -> np.ndarray:  # robotiq2f85.py:243
        return np.array([255.0])
