"""Shadow hand, right (/root/reference/mgs/gripper/shadow.py:343-455)."""
from typing import List

import numpy as np

from ..util.geo.transforms import SE3Pose
from .base import MjGripper

# the 22-joint closing posture written by close_gripper_at (shadow.py:383-408)
CLOSE_QPOS = np.array([-0.3464, 1.253, 0.7836, -0.001106, 0.01103, 1.475, 0.6181, 0.0155, -0.2083, 1.45, 0.75, 0.0, 0.13, -0.4, 1.5,
                       0.95, 0.35, 0.07708, 1.21, 0.2023, 0.6614, 0.0102])


def qpos_to_ctrl(qpos):
    """22 joint targets -> 18 actuator controls; the J2+J1 pairs share one tendon actuator (shadow.py:444-455)."""
    acc = np.zeros(18)
    acc[:5] = qpos[-5:]
    acc[5:7] = qpos[0:2]
    acc[7] = qpos[2] + qpos[3]
    acc[8:10] = qpos[4:6]
    acc[10] = qpos[6] + qpos[7]
    acc[11:13] = qpos[8:10]
    acc[13] = qpos[10] + qpos[11]
    acc[14:17] = qpos[12:15]
    acc[17] = qpos[15] + qpos[16]
    return acc


class GripperShadowRight(MjGripper):
    ASSET_DIR = "shadow"
    COMPUTE_F64 = True

    def __init__(self, pose: SE3Pose, grasp_type=None):
        super().__init__(pose, "rh_wrist")

    def base_to_contact_transform(self) -> SE3Pose:  # shadow.py:368-371
        return SE3Pose(np.array([0, 0, 0.0]), np.array([1.0, 0.0, 0.0, 0.0]), type="wxyz")

    def get_actuator_joint_names(self) -> List[str]:  # shadow.py:416-442
        return ["rh_FFJ4", "rh_FFJ3", "rh_FFJ2", "rh_FFJ1", "rh_MFJ4", "rh_MFJ3", "rh_MFJ2", "rh_MFJ1", "rh_RFJ4", "rh_RFJ3", "rh_RFJ2",
                "rh_RFJ1", "rh_LFJ5", "rh_LFJ4", "rh_LFJ3", "rh_LFJ2", "rh_LFJ1", "rh_THJ5", "rh_THJ4", "rh_THJ3", "rh_THJ2", "rh_THJ1"]

    def open_gripper(self, sim):  # shadow.py:373-377
        sim.set_qpos(np.zeros(22), sim.get_joint_idxs(self.get_actuator_joint_names()))
        sim.data.ctrl[:] = qpos_to_ctrl(np.zeros(22))

    def close_ctrl(self) -> np.ndarray:
        return qpos_to_ctrl(CLOSE_QPOS)
