"""filter_to_stable: the single-object hot path's caller (/root/reference/mgs/cli/filter_to_stable.py:15-68).

Reads  <MGS_INPUT_DIR>/<gripper name>/<object id>/candidates.npz   {pose f64[N,4,4], joints[N,nj]}
writes candidates_collision_free.npz and stable_grasps.npz next to it (same keys, same dtypes).
Hydra is not required: `python -m mj_grasp_sim_b200.mgs.cli.filter_to_stable gripper=PandaGripper object=hull:0`
"""
import os
import sys

import numpy as np

from ..env.gravityless_object_grasping import GravitylessObjectGrasping
from ..gripper.selector import get_gripper
from ..obj.selector import get_object
from ..util.geo.transforms import SE3Pose


def run(gripper_name: str, object_id: str, file_dir: str | None = None, enough_stable: int = 1000):
    gripper = get_gripper(gripper_name)
    obj = get_object(object_id)
    env = GravitylessObjectGrasping(gripper, obj)
    file_dir = os.path.abspath(os.path.join(file_dir or os.getenv("MGS_INPUT_DIR") or ".", gripper_name, object_id))
    grasps = np.load(os.path.join(file_dir, "candidates.npz"))
    poses = SE3Pose.from_mat(grasps["pose"], type="wxyz")
    joints = grasps["joints"]
    mask = env.grasp_collision_mask(poses, joints)
    poses_cf, joints_cf = poses[mask], joints[mask]
    print(sum(mask))
    mask2 = env.grasp_stability_evaluation_from_joints(poses_cf, joints_cf, enough_stable=enough_stable)
    poses_st, joints_st = poses_cf[mask2], joints_cf[mask2]
    print(sum(mask2))
    np.savez(os.path.join(file_dir, "candidates_collision_free.npz"), pose=poses_cf.to_mat(), joints=joints_cf)
    np.savez(os.path.join(file_dir, "stable_grasps.npz"), pose=poses_st.to_mat(), joints=joints_st)
    return mask, mask2


if __name__ == "__main__":
    kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
    run(kv.get("gripper", "PandaGripper"), kv.get("object", "cube"), kv.get("dir"))
