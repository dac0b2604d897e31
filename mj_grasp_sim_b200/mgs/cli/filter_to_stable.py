"""filter_to_stable: the single-object hot path's caller (/root/reference/mgs/cli/filter_to_stable.py:15-68).

Reads  <MGS_INPUT_DIR>/<gripper name>/<object id>/candidates.npz   {pose [N,4,4], joints [N,nj]}, runs the collision
filter and then the close -> lift -> shake rollout on the survivors (`enough_stable=1000`, :44-48), and writes
candidates_collision_free.npz and stable_grasps.npz next to the input (same keys).  Two batched kernel launches replace
the reference's two per-candidate Python loops.  Hydra is not required:

  python -m mj_grasp_sim_b200.mgs.cli.filter_to_stable gripper=PandaGripper object=hull:0 [dir=...]
"""
import os

from ._common import candidate_dir, cfg_get, gripper_name_from_cfg, load_candidates, object_id_from_cfg, parse_kv, save_grasps, single_object_env


def run(gripper_name: str, object_id: str, file_dir: str | None = None, enough_stable: int = 1000):
    env = single_object_env(gripper_name, object_id)
    where = candidate_dir(gripper_name, object_id, file_dir)
    poses, joints = load_candidates(os.path.join(where, "candidates.npz"))
    free = env.grasp_collision_mask(poses, joints)
    print(sum(free))
    survivors = (poses[free], joints[free])
    stable = env.grasp_stability_evaluation_from_joints(*survivors, enough_stable=enough_stable)
    print(sum(stable))
    save_grasps(os.path.join(where, "candidates_collision_free.npz"), *survivors)
    save_grasps(os.path.join(where, "stable_grasps.npz"), survivors[0][stable], survivors[1][stable])
    return free, stable


def main(cfg):
    """Entry point with the reference's `main(cfg)` shape (a Hydra DictConfig there; any attribute/dict config here)."""
    return run(gripper_name_from_cfg(cfg), object_id_from_cfg(cfg), cfg_get(cfg, "dir"))


if __name__ == "__main__":
    kv = parse_kv()
    run(kv.get("gripper", "PandaGripper"), kv.get("object", "cube"), kv.get("dir"))
