"""eval_grasps - the clutter hot path's caller (/root/reference/mgs/cli/eval_grasps.py:13-83).

Contract kept from the reference:
  * `eval_grasps(env_cfg, scene_def, (pose, joints)) -> (success_rate, {"num_objects": n})`: the scene is rebuilt from
    its pickled dictionary, grasp poses arrive in the CONTACT frame and are moved to the gripper base with
    pose @ inv(b2c) (:15-18), candidates that collide are dropped, the rest go through the close + lift rollout, and
    success rate = stable candidates / ALL candidates (:40-43); 0.0 when nothing is collision-free (:26-30);
  * `run` = `main` without Hydra: <MGS_INPUT_DIR>/<gripper name>/<sorted scenes>[id]/{scene.npz, inference_grasps.npz}
    -> grasp_evaluation.json {success_rate, num_objects, scene_id} (:48-79).
Both masks are single batched launches of the rollout kernel.

  python -m mj_grasp_sim_b200.mgs.cli.eval_grasps gripper=ShadowHand id=0 [env=clutter_table] [dir=...]
"""
import json
import os
from copy import deepcopy

import numpy as np

from ..env.selector import get_env_from_dict
from ..util.geo.transforms import SE3Pose
from ._common import cfg_get, gripper_name_from_cfg, parse_kv


def _to_base_frame(env, contact_frame_poses: np.ndarray) -> np.ndarray:
    to_base = env.gripper.base_to_contact_transform().inverse().to_mat()
    return np.einsum("nij,jk->nik", contact_frame_poses, to_base)


def eval_grasps(env_cfg, scene_def, grasps):
    env = get_env_from_dict(env_cfg, deepcopy(scene_def))
    aux = {"num_objects": len(env.object_names)}
    pose, joints = _to_base_frame(env, grasps[0]), grasps[1]
    free = env.grasp_collision_mask(SE3Pose.from_mat(pose.copy(), type="wxyz"), np.array(joints, copy=True))
    if not free.any():
        return 0.0, aux
    stable = env.grasp_stable_mask(SE3Pose.from_mat(pose[free].copy(), type="wxyz"), np.array(joints[free], copy=True),
                                   deepcopy(scene_def["env_state"]["state"]))
    return int(np.sum(stable)) / float(len(pose)), aux


def _scene_dir(input_dir: str, gripper_name: str, index: int):
    root = os.path.join(input_dir, gripper_name)
    scene_id = sorted(os.listdir(root))[index]
    return scene_id, os.path.join(root, scene_id)


def run(gripper_name: str, scene_index: int = 0, env_name: str = "clutter_table", input_dir: str | None = None):
    input_dir = input_dir or os.getenv("MGS_INPUT_DIR")
    assert input_dir is not None, "No input_dir defined!"
    scene_id, where = _scene_dir(input_dir, gripper_name, scene_index)
    scene_def = np.load(os.path.join(where, "scene.npz"), allow_pickle=True)["scene_definition"].item()
    with np.load(os.path.join(where, "inference_grasps.npz")) as g:
        rate, aux = eval_grasps(env_name, scene_def, (g["pose"], g["joints"]))
    report = {"success_rate": float(rate), "num_objects": aux["num_objects"], "scene_id": scene_id}
    target = os.path.join(where, "grasp_evaluation.json")
    with open(target, "w") as fh:
        json.dump(report, fh, indent=2)
    print(f"Evaluation complete: {rate:.2%} success rate")
    print(f"Results saved to {target}")
    return report


def main(cfg):
    """Entry point with the reference's `main(cfg)` shape (a Hydra DictConfig there; any attribute/dict config here)."""
    return run(gripper_name_from_cfg(cfg), int(cfg_get(cfg, "id", 0)), cfg_get(cfg, "env.name", "clutter_table"), cfg_get(cfg, "dir"))


if __name__ == "__main__":
    kv = parse_kv()
    run(kv.get("gripper", "ShadowHand"), int(kv.get("id", 0)), kv.get("env", "clutter_table"), kv.get("dir"))
