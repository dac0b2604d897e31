"""gen_scene - clutter scene generation + per-scene grasp filtering (/root/reference/mgs/cli/gen_scene.py:28-208).

  gen_stable_scene  drop/settle the objects (ClutterTableEnv.gen_clutter), refuse unstable scenes (ValueError, :42-43)
  filter_grasps     per-object stable grasps -> world frame (o2w @ grasp, :58-60) -> batched collision mask ->
                    batched close+lift rollout with enough_stable = min(128, num_objects * 32) (:115-121)
  run               writes <MGS_OUTPUT_DIR>/<gripper name>/<hash>/scene.npz {scene_definition: pickled dict} and one
                    <object id>_<object name>.npz {pose, joints} per object (:178-205)

Per-object grasps come from <MGS_INPUT_DIR>/<gripper name>/<object id>/stable_grasps.npz as in the reference
(:15-25), or are passed in memory (`grasps={object_id: (pose[N,4,4], joints[N,nj])}`).

  python -m mj_grasp_sim_b200.mgs.cli.gen_scene gripper=PandaGripper objects=hull:0,hull:1,hull:2 [seed=0]
"""
import os
from copy import deepcopy

import numpy as np

from ..env.selector import get_env, get_env_from_dict
from ..gripper.selector import get_gripper
from ..obj.selector import get_objects
from ..util.file import generate_unique_hash
from ..util.geo.transforms import SE3Pose
from ._common import parse_kv


def get_grasps(gripper_name, obj_id):
    g = np.load(os.path.join(os.getenv("MGS_INPUT_DIR"), gripper_name, obj_id, "stable_grasps.npz"))
    return g["pose"], g["joints"]


def gen_stable_scene(gripper_name, object_ids, env_name="clutter_table", seed=None):
    objs = get_objects(object_ids)
    gripper = get_gripper(gripper_name, default_pose=SE3Pose(np.array([5.0, 5.0, 1.0]), np.array([1.0, 0.0, 0.0, 0.0]), type="wxyz"))
    env = get_env(env_name, deepcopy(gripper), deepcopy(objs))
    env.set_gripper_pose(np.array([5.0, 5.0, 1.0]))
    env.gen_clutter(seed)
    scene_dict = env.to_dict()
    if not env.is_stable():
        raise ValueError("Scene unstable")
    return scene_dict


def filter_grasps(gripper_name, scene_def, env_name="clutter_table", grasps=None, only_collision_free=False, save_collision_grasps=False,
                  enough_collision_free=128, rng=None):
    env = get_env_from_dict(env_name, deepcopy(scene_def))
    num_objects = len(env.object_names)
    all_poses, all_joints, obj_indices, obj_map = [], [], [], []
    for obj_name, obj_id in zip(env.object_names, env.object_ids):
        poses, joints = grasps[obj_id] if grasps is not None else get_grasps(gripper_name, obj_id)
        if len(poses) == 0:
            continue
        world = (env.get_obj_pose(obj_name) @ SE3Pose.from_mat(deepcopy(poses))).to_mat()
        all_poses.append(world); all_joints.append(joints)
        obj_indices.append(np.full(len(world), len(obj_map), dtype=np.int32))
        obj_map.append((obj_name, obj_id))
    if len(all_poses) == 0:
        raise ValueError("No collision free grasps")
    all_poses, all_joints, obj_indices = np.concatenate(all_poses), np.concatenate(all_joints), np.concatenate(obj_indices)
    free = env.grasp_collision_mask(SE3Pose.from_mat(deepcopy(all_poses), type="wxyz"), deepcopy(all_joints))
    if sum(free) < enough_collision_free:
        raise ValueError("Not enough collision free grasps!")
    res_p, res_j, res_i = all_poses[free], all_joints[free], obj_indices[free]
    if not only_collision_free:
        perm = (rng or np.random.default_rng()).permutation(len(res_p))
        res_p, res_j, res_i = res_p[perm], res_j[perm], res_i[perm]
        enough_stable = min(128, num_objects * 32)
        stable = env.grasp_stable_mask(SE3Pose.from_mat(deepcopy(res_p), type="wxyz"), deepcopy(res_j),
                                       deepcopy(scene_def["env_state"]["state"]), enough_stable=enough_stable)
        if sum(stable) < enough_stable:
            raise ValueError("Not enough stable grasps!")
        res_p, res_j, res_i = res_p[stable], res_j[stable], res_i[stable]
    result, neg_result = [], []
    for k in np.unique(res_i):
        m = res_i == k
        name, oid = obj_map[k]
        result.append({"object_id": oid, "object_name": name, "pose": res_p[m], "joints": res_j[m]})
        if save_collision_grasps:
            cm = obj_indices[~free] == k
            if sum(cm) > 0:
                neg_result.append({"object_id": oid, "object_name": name, "pose": all_poses[~free][cm], "joints": all_joints[~free][cm]})
    return result, neg_result


def run(gripper_name, object_ids, env_name="clutter_table", seed=None, grasps=None, output_dir=None, **kw):
    output_dir = output_dir or os.getenv("MGS_OUTPUT_DIR")
    assert output_dir is not None, "No ouput_dir defined!"
    output_dir = os.path.join(output_dir, gripper_name, generate_unique_hash(16))
    scene_dict = gen_stable_scene(gripper_name, object_ids, env_name, seed)
    valid, invalid = filter_grasps(gripper_name, scene_dict, env_name, grasps=grasps, **kw)
    os.makedirs(output_dir, exist_ok=True)
    np.savez(os.path.join(output_dir, "scene"), scene_definition=scene_dict)
    for g in valid:
        np.savez(os.path.join(output_dir, g["object_id"] + "_" + g["object_name"]), pose=g["pose"], joints=g["joints"])
    for g in invalid:
        np.savez(os.path.join(output_dir, g["object_id"] + "_" + g["object_name"] + "_collision"), pose=g["pose"], joints=g["joints"])
    return output_dir


if __name__ == "__main__":
    kv = parse_kv()
    try:
        print(run(kv.get("gripper", "PandaGripper"), kv.get("objects", "hull:0,hull:1,hull:2").split(","), kv.get("env", "clutter_table"),
                  int(kv["seed"]) if "seed" in kv else None))
    except Exception as e:  # the reference's main swallows and prints (:175-208)
        print(e)
