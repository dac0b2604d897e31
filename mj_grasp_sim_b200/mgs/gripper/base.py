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
    # Arithmetic of the rollout kernel for scenes with this gripper: fp32 for the parallel-jaw grippers; the 16+-dof hands run the
    # fp64 build of the same kernel (MuJoCo is fp64; on marginal candidate sets the fp32 hands agree with the oracle on 97.4-98.8 %
    # of the labels, the fp64 build on 98.8-100 %: profiles/label_agreement_r2*.json).  MGS_PRECISION=f32|f64 overrides.
    COMPUTE_F64 = False

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

    # --- imperative protocol on ONE environment (the sim handle's state record; see mgs/core/simualtion.py) -----------
    # The batched entry points of the environments do not go through these: they read close_ctrl() / REPOSE_ON_CLOSE
    # and run the same sequence for every candidate inside the kernel (csrc/mgs_rollout.cuh).  The methods exist so
    # that reference callers that drive a gripper by hand keep working, one environment at a time.
    def set_pose(self, sim, pose: SE3Pose):
        """base.py:48-59: teleport the base (free-joint qpos and the mocap target it is welded to), then mj_forward."""
        idxs = self.get_freejoint_idxs(sim)
        vec = np.asarray(pose.to_vec(layout="pq", type="wxyz"), dtype=np.float64).reshape(-1)
        sim.data.qpos[idxs[0]:idxs[0] + 7] = vec
        sim.data.mocap_pos[0, :] = vec[:3]
        sim.data.mocap_quat[0, :] = vec[3:]
        sim.mj_forward()

    def open_gripper(self, sim):
        """default (robotiq2f85.py:237-238): zero control"""
        sim.data.ctrl[:] = 0.0

    def close_gripper(self, sim):
        sim.data.ctrl[:] = self.close_ctrl()

    def close_gripper_at(self, sim, pose: SE3Pose):
        """panda.py:225-241 and the five siblings: mocap target <- pose (Allegro / LEAP: set_pose again), ctrl <- the close
        signal, then NSTEP_CLOSE steps."""
        if self.REPOSE_ON_CLOSE:
            self.set_pose(sim, pose)
        else:
            vec = np.asarray(pose.to_vec(layout="pq", type="wxyz"), dtype=np.float64).reshape(-1)
            sim.data.mocap_pos[0, :] = vec[:3]
            sim.data.mocap_quat[0, :] = vec[3:]
        sim.data.ctrl[:] = self.close_ctrl()
        sim.mj_step(self.NSTEP_CLOSE)

    def lift_up(self, sim, viewer=None):
        """base.py:61-66"""
        for _ in range(10000):
            sim.data.mocap_pos[0, 2] += 0.00003
            sim.mj_step(1)

    # Shakable (base.py:111-143): the mocap target moves in small increments along the pose's axes
    def _move(self, sim, pose: SE3Pose, axis, step, n):
        d = np.asarray(pose.to_mat(), dtype=np.float64).reshape(4, 4)[:3, :3] @ np.asarray(axis, dtype=np.float64)
        for _ in range(n):
            sim.data.mocap_pos[0, :] += step * d
            sim.mj_step(1)

    def move_back(self, sim, pose: SE3Pose, viewer=None):
        self._move(sim, pose, [0, 0, -1.0], 0.0002, 1000)

    def move_right(self, sim, pose: SE3Pose, viewer=None):
        self._move(sim, pose, [0, 1.0, 0], 0.0005, 500)

    def move_left(self, sim, pose: SE3Pose, viewer=None):
        self._move(sim, pose, [0, -1.0, 0], 0.0005, 500)

    def shake_grasp_at(self, sim, pose: SE3Pose):
        self.move_back(sim, pose)
        self.move_right(sim, pose)
        self.move_left(sim, pose)

    # --- per-gripper data -------------------------------------------------------------------
    def base_to_contact_transform(self) -> SE3Pose:
        raise NotImplementedError

    def get_actuator_joint_names(self) -> List[str]:
        raise NotImplementedError

    def close_ctrl(self) -> np.ndarray:
        raise NotImplementedError


MjShakableOpenCloseGripper = MjGripper
