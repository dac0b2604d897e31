"""gen_grasp_candidates - producer of the hot path's input (/root/reference/mgs/cli/gen_grasp_candidates.py:17-83).

Antipodal candidates for the parallel-jaw grippers the reference serves this way (Panda, VX300: `all_gripper`, :23-28):
poses from `AntipodalGraspGenerator`, finger joints from the contact distance through `_clamp_width` + `width_to_joints`
(:62-64), written as <MGS_OUTPUT_DIR>/<gripper name>/<object id>/candidates.npz {pose [N,4,4], joints [N,2]}.  The LEAP and Shadow
hands get theirs from the contact-based differentiable sampler (`mgs.sampler.contact.ContactBasedDiff`, :30-40, :71-78): palm poses
and joint postures, joints [N, 16 / 22].

  python -m mj_grasp_sim_b200.mgs.cli.gen_grasp_candidates gripper=PandaGripper object=hull:0 [num_grasps=10000] [seed=0]
"""
import os

import numpy as np

from ..gripper.selector import get_gripper
from ..obj.selector import get_object
from ..sampler.antipodal import AntipodalGraspGenerator
from ._common import cfg_get, gripper_name_from_cfg, object_id_from_cfg, parse_kv


HANDS = {"LeapGripper": "leap", "LeapHand": "leap", "ShadowHand": "shadow"}  # hands served by the contact-based sampler


def run(gripper_name: str, object_id: str, num_grasps: int = 10000, output_dir: str | None = None, seed: int | None = None):
    if gripper_name not in ("PandaGripper", "VXGripper") and gripper_name not in HANDS:
        raise NotImplementedError(f"{gripper_name}: Panda / VX300 have antipodal candidates, LEAP / Shadow contact-based ones (gen_grasp_candidates.py:23-40)")
    print(f"Generating grasp candidates for gripper: {gripper_name}")
    obj = get_object(object_id)
    if gripper_name in HANDS:
        from ..sampler.contact import ContactBasedDiff
        from ..sampler.kin import HandKinematics
        Hs, aux = ContactBasedDiff(obj, seed=seed if seed is not None else 0).generate_grasps(num_grasps, HandKinematics(HANDS[gripper_name]))
        joints = aux["joints"]
    else:
        gripper = get_gripper(gripper_name)
        Hs, aux = AntipodalGraspGenerator(obj, seed=seed).generate_grasps(num_grasps)
        j1, j2 = gripper.width_to_joints(gripper._clamp_width(aux["width"]))
        joints = np.stack([j1, j2], axis=-1)
    out = os.path.join(output_dir or os.getenv("MGS_OUTPUT_DIR") or ".", gripper_name, object_id)
    os.makedirs(out, exist_ok=True)
    np.savez(os.path.join(out, "candidates.npz"), pose=Hs, joints=joints)
    print("Done!")
    return Hs, joints


def main(cfg):
    """Entry point with the reference's `main(cfg)` shape (a Hydra DictConfig there; any attribute/dict config here)."""
    return run(gripper_name_from_cfg(cfg), object_id_from_cfg(cfg), int(cfg_get(cfg, "num_grasps", 10000)), cfg_get(cfg, "dir"), cfg_get(cfg, "seed"))


if __name__ == "__main__":
    kv = parse_kv()
    run(kv.get("gripper", "PandaGripper"), kv.get("object", "cube"), int(kv.get("num_grasps", 10000)), kv.get("dir"),
        int(kv["seed"]) if "seed" in kv else None)
