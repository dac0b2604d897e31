"""filter_collision_free_candidates (/root/reference/mgs/cli/filter_collision_free_candidates.py:15-49):
candidates.npz -> candidates_collision_free.npz {pose f64[N,4,4], joints[N,nj]} through one batched
`grasp_collision_mask` launch.

  python -m mj_grasp_sim_b200.mgs.cli.filter_collision_free_candidates gripper=PandaGripper object=hull:0 [dir=...]
"""
import os

from ._common import candidate_dir, cfg_get, gripper_name_from_cfg, load_candidates, object_id_from_cfg, parse_kv, save_grasps, single_object_env


def run(gripper_name: str, object_id: str, file_dir: str | None = None):
    env = single_object_env(gripper_name, object_id)
    d = candidate_dir(gripper_name, object_id, file_dir)
    poses, joints = load_candidates(os.path.join(d, "candidates.npz"))
    mask = env.grasp_collision_mask(poses, joints)
    print(sum(mask))
    save_grasps(os.path.join(d, "candidates_collision_free.npz"), poses[mask], joints[mask])
    return mask


def main(cfg):
    """Entry point with the reference's `main(cfg)` shape (a Hydra DictConfig there; any attribute/dict config here)."""
    return run(gripper_name_from_cfg(cfg), object_id_from_cfg(cfg), cfg_get(cfg, "dir"))


if __name__ == "__main__":
    kv = parse_kv()
    run(kv.get("gripper", "PandaGripper"), kv.get("object", "cube"), kv.get("dir"))
