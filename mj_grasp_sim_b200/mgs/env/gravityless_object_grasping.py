"""GravitylessObjectGrasping - drop-in for /root/reference/mgs/env/gravityless_object_grasping.py.

Same constructor and the same public methods with the same argument meaning:
  grasp_collision_mask(poses, joints) -> bool[N]                       (reference :90-125)
  grasp_stability_evaluation_from_joints(poses, joints, nstep_lift, lift_dist, shake_steps,
                                         shake_dist, enough_stable) -> bool[N]      (reference :127-295)
  get_object_transform / check_contact / check_contact_with_object       (reference :297-321)
but every candidate is evaluated in ONE batched launch of the sm_100a rollout kernel through the C ABI
(libmgs_b200.so) instead of a Python loop over mujoco.mj_step.  With torch.distributed initialised the
candidates are sharded over the ranks and only the labels are gathered.

Capacity: the kernel keeps a bounded number of contacts / constraint rows per environment in shared memory.  A
candidate that needed more is flagged by the library; this class re-evaluates exactly those candidates on a second
model instance with the largest capacities and warns if any still does not fit - labels computed on truncated
contact sets are never returned silently (`last_overflow` holds the counts of the most recent call).
"""
from __future__ import annotations

import warnings

import numpy as np

from ...compiler.mjcf import compile_mjcf
from ...lib import BatchSim, MgsRolloutCfg
from ...shard import evaluate_sharded
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

# escalation capacities, tried in this order until the model fits one CTA's shared memory: contacts (0 rows = rows to match).
# 256 holds every contact the pair lists of the two-finger grippers can produce (63 pairs x 4 points)
ESCALATION_CAPS = (256, 128, 96, 64)
MAX_CAPS = (ESCALATION_CAPS[0], 0)
# first rung of the ladder: the library's default capacity of a single-object scene.  A model whose first pass was cut below it (the
# fp64 hands: 16 / 12 contacts, so that four warp-environments fit an SM) re-runs its few environments over capacity there before
# the largest capacity is tried for what is still over: a launch of a handful of candidates lasts one rollout whatever its size, and
# a rollout on 32 contacts is much shorter than one on 128
MID_CAPS = (32, 0)


class EscalatingSim:
    """Up to three instances of one model: the first-pass capacities (most environments resident per SM) and, created on first
    need, the default capacity of a single-object scene (when the first pass was cut below it) and the largest one that fits.
    `run(call, n)` evaluates n candidates with `call(sim, index_array_or_None) -> arrays`, then re-runs the candidates whose
    environment overflowed up that ladder until none is left (or the ladder ends: warning)."""

    def __init__(self, make_sim):
        self._make, self._small, self._mid, self._big = make_sim, None, None, None
        self.last_overflow = dict(first_pass=0, after_escalation=0)

    @property
    def small(self) -> BatchSim:
        if self._small is None:
            self._small = self._make(None)
        return self._small

    @property
    def big(self) -> BatchSim:
        if self._big is None:
            err = None
            for nc in ESCALATION_CAPS:
                if nc <= self.small.info.ncon_max:
                    break
                try:
                    self._big = self._make((nc, 0))
                    break
                except Exception as ex:  # does not fit the shared memory of one CTA: next smaller capacity
                    err = ex
            if self._big is None:
                raise err if err is not None else RuntimeError("no larger capacity than the first pass's")
        return self._big

    @property
    def mid(self):
        """the MID_CAPS instance, or None when the first pass already has that capacity (or the model does not fit it)"""
        if self._mid is None and self.small.info.ncon_max < MID_CAPS[0]:
            try:
                self._mid = self._make(MID_CAPS)
            except Exception:
                self._mid = False
        return self._mid or None

    def close(self):
        for s in (self._small, self._mid, self._big):
            if s:
                s.close()
        self._small = self._mid = self._big = None

    def run(self, call, n):
        out = call(self.small, None)
        out = [np.array(o) for o in (out if isinstance(out, tuple) else (out,))]
        aux = self.small.last_aux(n)
        over = np.nonzero(aux["overflow"])[0]
        self.last_overflow = dict(first_pass=int(len(over)), after_escalation=0)
        last = None
        for rung in ("mid", "big"):
            if not len(over):
                break
            sim = None
            if rung == "mid":
                sim = self.mid
            elif self.small.info.ncon_max < MAX_CAPS[0]:
                try:
                    sim = self.big
                except Exception:
                    sim = None
            if sim is None:
                continue
            again = call(sim, over)
            again = again if isinstance(again, tuple) else (again,)
            for o, a in zip(out, again):
                o[over] = a
            aux2 = sim.last_aux(len(over))
            for k in ("bad", "pos_drift", "rot_drift_deg"):
                aux[k][over] = aux2[k]
            aux["overflow"][over] = aux2["overflow"]
            over = over[np.asarray(aux2["overflow"], dtype=bool)]
            self.last_overflow["after_escalation"] = int(len(over))
            last = sim
        if len(over):
            cap = (last or self.small).info.ncon_max
            warnings.warn(f"{len(over)} of {n} environments needed more than {cap} contacts: their labels were computed on a "
                          "truncated contact set", RuntimeWarning, stacklevel=3)
        return (out[0] if len(out) == 1 else tuple(out)), aux


class GravitylessObjectGrasping(MjSimulation):
    def __init__(self, gripper, obj, device: int | None = None, ncon_max: int = 0, nefc_max: int = 0):
        self.gripper, self.obj = gripper, obj
        self.gripper_xml, self.gripper_assets = gripper.to_xml()
        self.object_xml, self.object_assets = obj.to_xml()
        self.model_xml = XML.format(gripper=self.gripper_xml, object=self.object_xml)
        self.model = compile_mjcf(self.model_xml, {**self.gripper_assets, **self.object_assets})
        self._device, self._caps, self._sim = device, (ncon_max, nefc_max), None
        self._esc = EscalatingSim(self._make_sim)
        self._init_state()

    def _make_sim(self, caps):
        if caps is None:
            return self.sim
        dev = self.sim.device
        return BatchSim(self.model, device=dev, ncon_max=caps[0], nefc_max=caps[1], ground_name=self.GROUND_GEOM, f64=self.compute_f64)

    @property
    def last_overflow(self):
        return self._esc.last_overflow

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

    def grasp_collision_mask(self, poses: SE3Pose, joints: np.ndarray) -> np.ndarray:
        pose7, joints32, jadr = self._process(poses, joints)
        base = self.gripper.get_freejoint_idxs(self)[0]

        def run_range(lo, hi):
            sel = lambda idx: (pose7[lo:hi], joints32[lo:hi]) if idx is None else (pose7[lo:hi][idx], joints32[lo:hi][idx])
            return self._esc.run(lambda sim, idx: sim.collision_mask(*sel(idx), jadr, base), hi - lo)[0]
        return evaluate_sharded(len(pose7), run_range)

    def grasp_stability_evaluation_from_joints(self, poses: SE3Pose, joints: np.ndarray, nstep_lift: int = 3000, lift_dist: float = 0.1,
                                               shake_steps: int = 500, shake_dist: float = 0.02, enough_stable=None, return_drift: bool = False):
        """bool[N] like the reference (:295).  `return_drift=True` additionally returns the two arrays the reference computes and
        leaves commented out of its return statement (:175-200, :283-295): positional drift [m] and rotational drift [deg] of the
        object over the close phase, NaN for candidates that were skipped or lost contact while closing."""
        pose7, joints32, jadr = self._process(poses, joints)
        n = len(pose7)
        base = self.gripper.get_freejoint_idxs(self)[0]
        cfg = MgsRolloutCfg(self.gripper.NSTEP_CLOSE, nstep_lift, shake_steps, self.gripper.REPOSE_ON_CLOSE, lift_dist, shake_dist)
        ctrl = self.gripper.close_ctrl()
        pos_drift, rot_drift = np.full(n, np.nan), np.full(n, np.nan)

        def run_range(lo, hi):
            sel = lambda idx: (pose7[lo:hi], joints32[lo:hi]) if idx is None else (pose7[lo:hi][idx], joints32[lo:hi][idx])
            (lab, _), aux = self._esc.run(lambda sim, idx: sim.stability(*sel(idx), jadr, base, ctrl, cfg), hi - lo)
            pos_drift[lo:hi], rot_drift[lo:hi] = aux["pos_drift"], aux["rot_drift_deg"]
            return lab
        # `enough_stable`: the reference stops simulating after that many successes and labels the rest False (:151-156).
        # Kept as an early stop in rounds of one GPU-filling chunk per rank (shard.evaluate_sharded).
        info = self.sim.info
        labels = evaluate_sharded(n, run_range, enough_stable, chunk=info.warps_per_block * info.blocks_per_sm * info.num_sms)
        if return_drift:  # (drifts are per rank: each rank holds those of the candidates it evaluated itself)
            return labels, pos_drift, rot_drift
        return labels

    # ---- single-environment queries of the reference class (operate on the handle's state record) -----------------
    def get_object_transform(self, object_name: str) -> SE3Pose:  # :297-304
        a = int(self.model.jnt_qposadr[self.model.names["joint"][f"{object_name}:joint"]])
        q = self.data.qpos
        return SE3Pose(np.copy(q[a:a + 3]).astype(np.float32), np.copy(q[a + 3:a + 7]).astype(np.float32), "wxyz")

    def check_contact(self) -> bool:  # :306-307
        return self.data.ncon != 0

    def check_contact_with_object(self) -> bool:  # :309-321: a contact whose two geom ids straddle the ground geom's id
        g = int(self.model.names["geom"][self.GROUND_GEOM])
        geoms = self.data.contact.geom
        return bool((((geoms[:, 0] < g) & (geoms[:, 1] > g)) | ((geoms[:, 0] > g) & (geoms[:, 1] < g))).any())

    def idle_grasp(self, pose: SE3Pose, joints: np.ndarray):
        raise NotImplementedError("idle_grasp opens the interactive MuJoCo viewer (:73-88); there is no viewer on the batched path")
