"""eval_grasps - the clutter hot path's caller (/root/reference/mgs/cli/eval_grasps.py:13-83).

`eval_grasps(env_name, scene_def, (pose, joints))` rebuilds the scene from its pickled dictionary, moves the
grasp poses from the contact frame to the gripper base (pose @ inv(b2c), :15-18), filters collisions and
evaluates the close+lift rollout; success rate = stable / all candidates (:40-43).  `run` mirrors `main`:
<MGS_INPUT_DIR>/<gripper name>/<scene>/{scene.npz, inference_grasps.npz} -> grasp_evaluation.json
{success_rate, num_objects, scene_id}.

  python -m mj_grasp_sim_b200.mgs.cli.eval_grasps gripper=ShadowHand id=0 [env=clutter_table] [dir=...]
"""
import json
import os
from copy import deepcopy

import numpy as np

from ..env.selector import get_env_from_dict
from ..util.geo.transforms import SE3Pose
from ._common import parse_kv


def eval_grasps(env_cfg, scene_def, grasps):
    env = get_env_from_dict(env_cfg, deepcopy(scene_def))
    b2c = env.gripper.base_to_contact_transform().inverse().to_mat()
    pose, joints = grasps
    pose = np.einsum("nij,jk->nik", pose, b2c)
    collision_free_mask = env.grasp_collision_mask(SE3Pose.from_mat(deepcopy(pose), type="wxyz"), deepcopy(joints))
    aux = {"num_objects": len(env.object_names)}
    if sum(collision_free_mask) == 0:
        return 0.0, aux
    stable_grasp_mask = env.grasp_stable_mask(SE3Pose.from_mat(deepcopy(pose[collision_free_mask]), type="wxyz"),
                                              deepcopy(joints[collision_free_mask]), deepcopy(scene_def["env_state"]["state"]))
    return sum(stable_grasp_mask) / float(len(pose)), aux


def run(gripper_name: str, scene_index: int = 0, env_name: str = "clutter_table", input_dir: str | None = None):
    input_dir = input_dir or os.getenv("MGS_INPUT_DIR")
    assert input_dir is not None, "No input_dir defined!"
    all_scene_dir = os.path.join(input_dir, gripper_name)
    all_scenes = sorted(os.listdir(all_scene_dir))
    scene = all_scenes[scene_index]
    scene_dir = os.path.join(all_scene_dir, scene)
    scene_dict = np.load(os.path.join(scene_dir, "scene.npz"), allow_pickle=True)["scene_definition"].item()
    grasps = np.load(os.path.join(scene_dir, "inference_grasps.npz"))
    success_rate, aux = eval_grasps(env_name, scene_dict, (grasps["pose"], grasps["joints"]))
    results = {"success_rate": float(success_rate), "num_objects": aux["num_objects"], "scene_id": scene}
    results_path = os.path.join(scene_dir, "grasp_evaluation.json")
    with open(results_path, "w") as f:
        json.dump(results, f, indent=2)
    print(f"Evaluation complete: {success_rate:.2%} success rate")
    print(f"Results saved to {results_path}")
    return results


if __name__ == "__main__":
    kv = parse_kv()
    run(kv.get("gripper", "ShadowHand"), int(kv.get("id", 0)), kv.get("env", "clutter_table"), kv.get("dir"))
