"""get_gripper(cfg): name -> gripper instance (/root/reference/mgs/gripper/selector.py:33-66).
`cfg` is anything with a `.name` (the reference passes a Hydra node; Hydra is not required here)."""
import numpy as np

from ..util.geo.transforms import SE3Pose
from .panda import GripperPanda
from .robotiq2f85 import GripperRobotiq2f85
from .vx300 import GripperVX300
from .allegro import GripperAllegro
from .leap import GripperLeap
from .shadow import GripperShadowRight

_REGISTRY = {"PandaGripper": GripperPanda, "Robotiq2f85Gripper": GripperRobotiq2f85, "VXGripper": GripperVX300, "AllegroGripper": GripperAllegro, "LeapGripper": GripperLeap, "ShadowHand": GripperShadowRight}


def get_gripper(cfg, default_pose=None):
    pose = SE3Pose(np.array([0, 0, 0]), np.array([1, 0, 0, 0]), type="wxyz") if default_pose is None else default_pose
    name = cfg if isinstance(cfg, str) else cfg.name
    if name not in _REGISTRY:
        raise ValueError(f"Unknown gripper: {name}")
    return _REGISTRY[name](pose)
