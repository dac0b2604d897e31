"""Interbotix ViperX 300 gripper (/root/reference/mgs/gripper/vx300.py:186-339)."""
from typing import List

import numpy as np

from ..util.geo.transforms import SE3Pose
from .base import MjGripper


class GripperVX300(MjGripper):
    ASSET_DIR = "vx300"
    MIN_WIDTH, MAX_WIDTH, MIN_WIDTH_CLAMP = 0.042, 0.114, 0.003
    Q1_RANGE, Q2_RANGE = [0.021, 0.057], [-0.057, -0.021]

    def __init__(self, pose: SE3Pose):
        super().__init__(pose, "gripper_link")

    def base_to_contact_transform(self) -> SE3Pose:  # vx300.py:242-257
        rot_y = SE3Pose(np.array([0, 0, 0]), np.array([0.707106781, 0, -0.707106781, 0]), type="wxyz")
        rot_z = SE3Pose(np.array([0, 0, 0]), np.array([0.707106781, 0, 0.0, 0.707106781]), type="wxyz")
        rot = rot_z @ rot_y
        rot.pos = np.array([0, 0, -0.12])
        return rot

    def get_actuator_joint_names(self) -> List[str]:  # vx300.py:325-328
        return ["left_finger", "right_finger"]

    def close_ctrl(self) -> np.ndarray:  # vx300.py:306-309
        return np.array([self.Q1_RANGE[0], self.Q2_RANGE[1]])

    def open_gripper(self, sim):  # vx300.py:259-272
        q = sim.get_joint_idxs(self.get_actuator_joint_names())
        sim.data.qpos[q[0]], sim.data.qpos[q[1]] = self.Q1_RANGE[1], self.Q2_RANGE[0]
        sim.data.ctrl[0], sim.data.ctrl[1] = self.Q1_RANGE[1], self.Q2_RANGE[0]
        sim.mj_forward()

    def width_to_joints(self, width):  # vx300.py:284-294
        w = np.clip(width, self.MIN_WIDTH, self.MAX_WIDTH)
        return np.clip(0.5 * w, *self.Q1_RANGE), np.clip(-0.5 * w, *self.Q2_RANGE)

    def _clamp_width(self, width):  # vx300.py:337-339
        return np.clip(width + 0.045, self.MIN_WIDTH_CLAMP, self.MAX_WIDTH)
