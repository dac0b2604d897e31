"""Allegro hand (/root/reference/mgs/gripper/allegro.py:297-402)."""
from typing import List

import numpy as np

from ..util.geo.transforms import SE3Pose
from .base import MjGripper


class GripperAllegro(MjGripper):
    ASSET_DIR = "allegro"
    COMPUTE_F64 = True
    REPOSE_ON_CLOSE = 1  # close_gripper_at calls set_pose first (allegro.py:354-357)

    def __init__(self, pose: SE3Pose):
        super().__init__(pose, "palm")
        # allegro.py:300-339
        self.open_pose = np.array([-0.08, 0.715, 0.710, 0.95, 0, 0.8, 0.71, 0.67, 0.08, 0.715, 0.710, 0.95, 1.4, 0.55, -0.19, 1.45])
        self.close_pose = np.array([-0.08, 0.95, 1, 0.95, 0, 0.95, 1.2, 0.85, 0.08, 0.95, 1.2, 0.9, 1.4, 0.55, 0.29, 1.45])

    def base_to_contact_transform(self) -> SE3Pose:  # allegro.py:341-347
        theta = -np.pi / 2.0
        q = np.array([np.cos(theta / 2.0), 0.0, np.sin(theta / 2.0), 0.0])
        offset = SE3Pose(np.array([0, 0, 0]), q, type="wxyz") @ SE3Pose(np.array([-0.08, 0.0, 0.01]), np.array([1.0, 0, 0, 0]), type="wxyz")
        return SE3Pose(offset.pos, q, type="wxyz")

    def get_actuator_joint_names(self) -> List[str]:  # allegro.py:380-398
        return [f"{f}j{k}" for f in ("ff", "mf", "rf", "th") for k in range(4)]

    def open_gripper(self, sim):  # allegro.py:349-352
        sim.set_qpos(np.copy(self.open_pose), sim.get_joint_idxs(self.get_actuator_joint_names()))
        sim.data.ctrl[:] = np.copy(self.open_pose)

    def close_ctrl(self) -> np.ndarray:
        return np.copy(self.close_pose)
