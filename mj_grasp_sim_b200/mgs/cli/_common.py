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


# ---- Hydra-style configs without Hydra ---------------------------------------------------------------------------
# The reference's entry points take a DictConfig (`cfg.gripper.name`, `cfg.id`, ...) and map `cfg.id` to an object through
# asset/mj-objects/fast_eta_objects.txt (1032 dataset ids, not shipped).  Any attribute- or dict-style object works here;
# the id list is the synthetic stand-in: index 0 = the box primitive, index k = convex hull seed k - 1.
SYNTHETIC_OBJECT_IDS = ["cube"] + [f"hull:{i}" for i in range(1031)]


def cfg_get(cfg, path: str, default=None):
    cur = cfg
    for key in path.split("."):
        if cur is None:
            return default
        cur = cur.get(key, None) if isinstance(cur, dict) else getattr(cur, key, None)
    return default if cur is None else cur


def gripper_name_from_cfg(cfg) -> str:
    g = cfg_get(cfg, "gripper")
    return g if isinstance(g, str) else cfg_get(cfg, "gripper.name")


def object_id_from_cfg(cfg) -> str:
    explicit = cfg_get(cfg, "object")
    if isinstance(explicit, str):
        return explicit
    return SYNTHETIC_OBJECT_IDS[int(cfg_get(cfg, "id", 0))]
