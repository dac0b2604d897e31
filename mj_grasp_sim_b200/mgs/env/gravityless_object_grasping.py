"""GravitylessObjectGrasping - drop-in for /root/reference/mgs/env/gravityless_object_grasping.py.

Same constructor and the same two public methods with the same argument meaning:
  grasp_collision_mask(poses, joints) -> bool[N]                       (reference :90-125)
  grasp_stability_evaluation_from_joints(poses, joints, nstep_lift, lift_dist, shake_steps,
                                         shake_dist, enough_stable) -> bool[N]      (reference :127-295)
but every candidate is evaluated in ONE batched launch of the sm_100a rollout kernel through the C ABI
(libmgs_b200.so) instead of a Python loop over mujoco.mj_step.  With torch.distributed initialised the
candidates are sharded contiguously over the ranks and only the labels are gathered.
"""
from __future__ import annotations

import numpy as np

from ...compiler.mjcf import compile_mjcf
from ...lib import BatchSim, MgsRolloutCfg
from ...shard import apply_enough_stable, gather_labels, shard_range
from ..core.simualtion import MjSimulation
from ..util.geo.transforms import SE3Pose

# same options, body layout and element order as the reference template (:34-54): the gripper fragment,
# then body:ground / geom:ground, then the object fragment (the contact labels depend on that geom order)
XML = r"""
<mujoco>
    <compiler angle="radian" autolimits="true" discardvisual="false"/>
    <option integrator="implicitfast" timestep="0.001" noslip_iterations="1"/>
    <option><flag multiccd="enable"/></option>
    <option cone="elliptic" impratio="3" noslip_iterations="2" noslip_tolerance="1e-8" tolerance="1e-8" gravity="0 0 0"/>
    {gripper}
    <worldbody>
        <body name="body:ground" pos="0.0 0 -1.0">
           <geom name="geom:ground" pos="0 0 0" rgba="1.0 1.0 1.0 0.0" size="1.0 1.0 0.02" type="box" density="500"/>
        </body>
    </worldbody>
    {object}
</mujoco>
"""


class GravitylessObjectGrasping(MjSimulation):
    def __init__(self, gripper, obj, device: int | None = None, ncon_max: int = 0, nefc_max: int = 0):
        self.gripper, self.obj = gripper, obj
        self.gripper_xml, self.gripper_assets = gripper.to_xml()
        self.object_xml, self.object_assets = obj.to_xml()
        self.model_xml = XML.format(gripper=self.gripper_xml, object=self.object_xml)
        self.model = compile_mjcf(self.model_xml, {**self.gripper_assets, **self.object_assets})
        self._device, self._caps, self._sim = device, (ncon_max, nefc_max), None

    # -- the batched simulator is created on first use (needs a CUDA device; there is no CPU fallback)
    @property
    def sim(self) -> BatchSim:
        if self._sim is None:
            dev = self._device
            if dev is None:
                import torch
                dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
            self._sim = BatchSim(self.model, device=dev, ncon_max=self._caps[0], nefc_max=self._caps[1])
        return self._sim

    def _process(self, poses: SE3Pose, joints: np.ndarray):
        if len(poses) != len(joints):
            raise ValueError(f"Number of poses ({len(poses)}) must match number of joint configurations ({len(joints)}).")
        names = self.gripper.get_actuator_joint_names()
        joints = np.asarray(joints)
        if len(joints) and joints.shape[1] != len(names):
            raise ValueError(f"Joints array has incorrect dimension ({joints.shape[1]}), expected {len(names)}.")
        b2c = self.gripper.base_to_contact_transform()
        processed = poses @ b2c  # SE3Pose semantics: float32 4x4 product, scipy back-conversion (:115-116)
        pose7 = processed.to_vec(layout="pq", type="wxyz").astype(np.float32).reshape(-1, 7)
        return pose7, joints.astype(np.float32).reshape(len(pose7), len(names)), np.array(self.get_joint_idxs(names), dtype=np.int32)

    def _sharded(self, n, fn):
        """Run fn(lo, hi) on this rank's contiguous block and gather the labels from all ranks."""
        try:
            import torch.distributed as dist
            on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        except Exception:
            on = False
        if not on:
            return fn(0, n)
        import torch
        lo, hi = shard_range(n, dist.get_rank(), dist.get_world_size())
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else None
        return gather_labels(fn(lo, hi), n, device=dev)

    def grasp_collision_mask(self, poses: SE3Pose, joints: np.ndarray) -> np.ndarray:
        pose7, joints32, jadr = self._process(poses, joints)
        base = self.gripper.get_freejoint_idxs(self)[0]
        return self._sharded(len(pose7), lambda lo, hi: self.sim.collision_mask(pose7[lo:hi], joints32[lo:hi], jadr, base))

    def grasp_stability_evaluation_from_joints(self, poses: SE3Pose, joints: np.ndarray, nstep_lift: int = 3000, lift_dist: float = 0.1,
                                               shake_steps: int = 500, shake_dist: float = 0.02, enough_stable=None) -> np.ndarray:
        pose7, joints32, jadr = self._process(poses, joints)
        base = self.gripper.get_freejoint_idxs(self)[0]
        cfg = MgsRolloutCfg(self.gripper.NSTEP_CLOSE, nstep_lift, shake_steps, self.gripper.REPOSE_ON_CLOSE, lift_dist, shake_dist)
        ctrl = self.gripper.close_ctrl()
        labels = self._sharded(len(pose7), lambda lo, hi: self.sim.stability(pose7[lo:hi], joints32[lo:hi], jadr, base, ctrl, cfg)[0])
        # `enough_stable`: the reference stops evaluating after that many successes and labels the rest
        # False (:151-156); evaluating everything and masking the tail gives the same array
        return apply_enough_stable(labels, enough_stable)
