"""Gripper protocol (/root/reference/mgs/gripper/base.py:28-147) for the batched path.

The reference's `close_gripper_at(sim, pose)` is imperative (sets mocap + ctrl on one MjData and steps
3000 times).  On the batched path the same information is declarative: `close_ctrl()` (the control
vector the reference writes) and `REPOSE_ON_CLOSE` (whether its close_gripper_at re-teleports the base
through set_pose, as Allegro/LEAP do)."""
from __future__ import annotations

import os
from typing import Any, Dict, List, Tuple

import numpy as np

from ..util.const import ASSET_PATH
from ..util.geo.transforms import SE3Pose


class MjGripper:
    ASSET_DIR = ""
    FREEJOINT = "freejoint"
    REPOSE_ON_CLOSE = 0
    NSTEP_CLOSE = 3000  # mujoco.mj_step(sim.model, sim.data, nstep=3000) in every close_gripper_at

    def __init__(self, pose: SE3Pose, base_body: str):
        vec = pose.to_vec(layout="pq", type="wxyz")
        self.pos, self.quat, self.base = vec[:3], vec[3:], base_body

    def set_load_pose(self, pose: SE3Pose):
        vec = pose.to_vec(layout="pq", type="wxyz")
        self.pos, self.quat = vec[:3], vec[3:]

    def to_xml(self) -> Tuple[str, Dict[str, Any]]:
        base = os.path.join(ASSET_PATH, self.ASSET_DIR)
        if not os.path.isdir(base):
            raise FileNotFoundError(f"asset directory not found at {base}")
        pos = "{} {} {}".format(*self.pos)
        quat = "{} {} {} {}".format(*self.quat)
        xml = open(os.path.join(base, "template.xml")).read().format(position=pos, quaternion=quat)
        assets = {}
        for fn in os.listdir(base):
            p = os.path.join(base, fn)
            if os.path.isfile(p) and fn != "template.xml":
                with open(p, "rb") as f:
                    assets[fn] = f.read()
        return xml, assets

    def get_freejoint_idxs(self, sim) -> List[int]:
        start = sim.get_joint_idxs([self.FREEJOINT])[0]
        return list(range(start, start + 7))

    # --- per-gripper data -------------------------------------------------------------------
    def base_to_contact_transform(self) -> SE3Pose:
        raise NotImplementedError

    def get_actuator_joint_names(self) -> List[str]:
        raise NotImplementedError

    def close_ctrl(self) -> np.ndarray:
        raise NotImplementedError


MjShakableOpenCloseGripper = MjGripper
