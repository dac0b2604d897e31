"""Franka Panda hand (/root/reference/mgs/gripper/panda.py:144-266)."""
from typing import List

import numpy as np

from ..util.geo.transforms import SE3Pose
from .base import MjGripper


class GripperPanda(MjGripper):
    ASSET_DIR = "panda"
    MIN_WIDTH_TARGET, MAX_WIDTH, MIN_WIDTH_CLAMP = 0.0, 0.08, 0.003
    Q1_RANGE, Q2_RANGE = [0.0, 0.04], [-0.04, 0.0]

    def __init__(self, pose: SE3Pose):
        super().__init__(pose, "hand")

    def base_to_contact_transform(self) -> SE3Pose:  # panda.py:190-193
        return SE3Pose(np.array([0, 0, -0.102]), np.array([0.707106781, 0.0, 0.0, 0.707106781]), type="wxyz")

    def get_actuator_joint_names(self) -> List[str]:  # panda.py:243-245
        return ["finger_joint1", "finger_joint2"]

    def close_ctrl(self) -> np.ndarray:  # panda.py:236
        return np.array([0.0, -0.04])

    def open_gripper(self, sim):  # panda.py:195-207
        q = sim.get_joint_idxs(self.get_actuator_joint_names())
        sim.data.qpos[q[0]], sim.data.qpos[q[1]] = self.Q1_RANGE[1], self.Q2_RANGE[1]
        sim.data.ctrl[0], sim.data.ctrl[1] = self.Q1_RANGE[1], self.Q2_RANGE[1]
        sim.mj_forward()

    def width_to_joints(self, width):  # panda.py:217-223
        w = np.clip(width, self.MIN_WIDTH_CLAMP, self.MAX_WIDTH)
        return np.clip(w / 2.0, *self.Q1_RANGE), np.clip(-0.04 + w / 2.0, *self.Q2_RANGE)

    def _clamp_width(self, width):  # panda.py:264-266
        return np.clip(width + 0.025, self.MIN_WIDTH_CLAMP, self.MAX_WIDTH)
