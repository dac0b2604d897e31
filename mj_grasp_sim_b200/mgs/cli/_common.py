"""Shared pieces of the CLI mirrors: Hydra-free `key=value` argument parsing and the candidate file layout
<MGS_INPUT_DIR>/<gripper name>/<object id>/ used by the reference's three filter entry points
(/root/reference/mgs/cli/filter_to_stable.py:27-32)."""
import os
import sys

import numpy as np

from ..env.gravityless_object_grasping import GravitylessObjectGrasping
from ..gripper.selector import get_gripper
from ..obj.selector import get_object
from ..util.geo.transforms import SE3Pose


def parse_kv(argv=None):
    return dict(a.split("=", 1) for a in (sys.argv[1:] if argv is None else argv) if "=" in a)


def candidate_dir(gripper_name: str, object_id: str, file_dir: str | None = None) -> str:
    return os.path.abspath(os.path.join(file_dir or os.getenv("MGS_INPUT_DIR") or ".", gripper_name, object_id))


def single_object_env(gripper_name: str, object_id: str):
    return GravitylessObjectGrasping(get_gripper(gripper_name), get_object(object_id))


def load_candidates(path: str):
    grasps = np.load(path)
    return SE3Pose.from_mat(grasps["pose"], type="wxyz"), grasps["joints"]


def save_grasps(path: str, poses: SE3Pose, joints):
    np.savez(path, pose=poses.to_mat(), joints=joints)
