"""Hand kinematics for the contact-based sampler, derived from the COMPILED gripper model.

The reference keeps one hand-written kinematic table per hand (/root/reference/mgs/sampler/kin/leap.py, kin/shadow.py: static link
transforms, joint axes, joint ranges, fingertip joints, pad normals, contact points on the pads, the pre-grasp posture and the
approach alignment) and a differentiable forward kinematics over it (kin/base.py:82-113).  Its static transforms and axes ARE the
body poses and joint axes of the gripper's MJCF (same numbers), so here they are read from the model the MJCF compiler produced
(`compile_mjcf` of the gripper fragment) instead of being typed in again; what remains per hand is a few lines of specification
(`HAND_SPECS`): which joints end a finger, the pad normal in the distal link's frame, the approach alignment.  Contact points on the
pads are taken from the distal link's own collision geometry (its vertices on the pad face), not from a table.

`HandKinematics.fk(theta)` is a batched torch expression (differentiable in theta): link frames relative to the gripper's base body.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class HandSpec:
    gripper: str                 # key of scenes.GRIPPERS
    fingertip_joints: List[str]  # the last joint of every finger (its body carries the pad)
    neg_pad_normal: tuple        # the reference's `fingertip_normals`: MINUS the pad's outward normal, distal-link frame
    align_rot: tuple             # `align_to_approach`: rotation applied to the approach frame ...
    align_pos: tuple             # ... and offset of the base in that frame


# reference: kin/leap.py:28-33,137-147 ; kin/shadow.py:39-44,155-168
HAND_SPECS = {
    "leap": HandSpec("leap", ["if_dip", "mf_dip", "rf_dip", "th_ipl"], (1.0, 0.0, 0.0), ((1.0, 0, 0), (0, 1.0, 0), (0, 0, 1.0)), (0.0, 0.0, 0.0)),
    "shadow": HandSpec("shadow", ["rh_FFJ1", "rh_MFJ1", "rh_RFJ1", "rh_LFJ1", "rh_THJ1"], (0.0, 1.0, 0.0),
                       ((0.0, 0, 1.0), (1.0, 0.0, 0), (0.0, 1.0, 0.0)), (-0.1, 0.0, 0.0)),
}


def _quat_to_mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


class HandKinematics:
    """Kinematic tree of a gripper below its base body, the actuated joints in the gripper class's order."""

    def __init__(self, name: str, n_pad_points: int = 9):
        from ... import scenes
        from ...compiler.mjcf import compile_mjcf
        self.spec = HAND_SPECS[name]
        g = scenes.GRIPPERS[name]
        gx, ga = scenes.gripper_fragment(name)
        m = compile_mjcf("<mujoco><compiler angle='radian' autolimits='true'/>" + gx + "</mujoco>", ga)
        a = m.arr
        self.joint_names = list(g["joints"])
        self.num_dofs = len(self.joint_names)
        jid = [m.names["joint"][j] for j in self.joint_names]
        self.joint_ranges = np.asarray(m.jnt_range)[jid].astype(np.float32)
        self.init_pregrasp_joint = np.asarray(g["open_pose"], dtype=np.float32)
        base = int(m.jnt_bodyid[m.names["joint"][g["freejoint"]]])
        parent = np.asarray(a["body_parentid"]).astype(int)
        # bodies of the gripper tree below the base, parents first (body ids are in document order: a parent precedes its children)
        self.bodies, local = [], {base: -1}
        for b in range(base + 1, m.nbody):
            if parent[b] in local:
                local[b] = len(self.bodies)
                self.bodies.append(b)
        nb = len(self.bodies)
        self.parent = np.array([local[parent[b]] for b in self.bodies])  # -1 = the base
        self.R0 = np.stack([_quat_to_mat(np.asarray(a["body_quat"]).reshape(-1, 4)[b]) for b in self.bodies]).astype(np.float32)
        self.p0 = np.asarray(a["body_pos"]).reshape(-1, 3)[self.bodies].astype(np.float32)
        self.axis = np.zeros((nb, 3), dtype=np.float32)
        self.dof = -np.ones(nb, dtype=int)  # index into theta of the body's hinge, -1: rigidly attached
        jb = np.asarray(m.jnt_bodyid).astype(int)
        for k, j in enumerate(jid):
            lb = local[jb[j]]
            self.axis[lb] = np.asarray(a["jnt_axis"]).reshape(-1, 3)[j]
            self.dof[lb] = k
            assert np.abs(np.asarray(a["jnt_pos"]).reshape(-1, 3)[j]).max() < 1e-12, "hinge anchors are at the body origin in both hands"
        self.fingertip_body = np.array([local[jb[m.names["joint"][j]]] for j in self.spec.fingertip_joints])
        self.neg_normal = np.tile(np.asarray(self.spec.neg_pad_normal, dtype=np.float32), (len(self.fingertip_body), 1))
        self.align_rot = np.asarray(self.spec.align_rot, dtype=np.float32)
        self.align_pos = np.asarray(self.spec.align_pos, dtype=np.float32)
        self.local_fingertip_contact_positions = np.stack([self._pad_points(m, self.bodies[lb], n_pad_points) for lb in self.fingertip_body])

    def _pad_points(self, m, body, k):
        """k points on the pad of a distal link (its frame): vertices of the link's collision geometry that lie on the face
        opposite to `neg_pad_normal`, in the distal 40 % of the link, thinned to k by farthest-point selection."""
        a = m.arr
        n = np.asarray(self.spec.neg_pad_normal, dtype=np.float64)
        pts = []
        for c in range(int(a["ncgeom"])):
            if int(a["cgeom_bodyid"][c]) != body or int(a["cgeom_hullid"][c]) < 0:
                continue
            h = int(a["cgeom_hullid"][c])
            v0, nv = int(a["hull_vertadr"][h]), int(a["hull_vertnum"][h])
            V = np.asarray(a["hull_vert"]).reshape(-1, 3)[v0:v0 + nv]
            R = _quat_to_mat(np.asarray(a["cgeom_quat"]).reshape(-1, 4)[c])
            pts.append(V @ R.T + np.asarray(a["cgeom_pos"]).reshape(-1, 3)[c])
        P = np.concatenate(pts)
        d = P @ n
        face = P[d <= d.min() + 1.5e-3]
        r = np.linalg.norm(face, axis=1)
        face = face[r >= 0.6 * r.max()] if len(face) > k else face
        sel = [int(np.argmax(np.linalg.norm(face, axis=1)))]
        dist = np.full(len(face), np.inf)
        while len(sel) < min(k, len(face)):
            dist = np.minimum(dist, np.linalg.norm(face - face[sel[-1]], axis=1))
            sel.append(int(np.argmax(dist)))
        out = face[sel]
        if len(out) < k:
            out = np.concatenate([out, np.repeat(out[:1], k - len(out), axis=0)])
        return out.astype(np.float32)

    # ---- differentiable forward kinematics -------------------------------------------------------------------------------
    def fk(self, theta):
        """theta [B, num_dofs] (torch) -> (R [B, nbody, 3, 3], p [B, nbody, 3]): link frames relative to the base body.
        Same composition as the reference's forward_kinematic_point_transform (kin/base.py:82-113): parent frame, static link
        transform, then the joint's rotation about its axis."""
        import torch
        B = theta.shape[0]
        dev, dt = theta.device, theta.dtype
        R0, p0, ax = (torch.as_tensor(x, device=dev, dtype=dt) for x in (self.R0, self.p0, self.axis))
        eye = torch.eye(3, device=dev, dtype=dt)
        Rs, ps = [], []
        for b in range(len(self.bodies)):
            if self.parent[b] < 0:
                Rp, pp = eye.expand(B, 3, 3), torch.zeros(B, 3, device=dev, dtype=dt)
            else:
                Rp, pp = Rs[self.parent[b]], ps[self.parent[b]]
            R = Rp @ R0[b]
            p = pp + Rp @ p0[b]
            if self.dof[b] >= 0:
                th = theta[:, self.dof[b]]
                a = ax[b] / ax[b].norm()
                K = torch.tensor([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]], device=dev, dtype=dt)
                Rj = eye + torch.sin(th)[:, None, None] * K + (1.0 - torch.cos(th))[:, None, None] * (K @ K)  # Rodrigues
                R = R @ Rj
            Rs.append(R)
            ps.append(p)
        return torch.stack(Rs, dim=1), torch.stack(ps, dim=1)

    def fingertip_points(self, theta, local_points):
        """local_points [nf, 3] (one per fingertip, distal-link frame) -> [B, nf, 3] in the base frame."""
        import torch
        R, p = self.fk(theta)
        idx = torch.as_tensor(self.fingertip_body, device=theta.device)
        lp = torch.as_tensor(local_points, device=theta.device, dtype=theta.dtype)
        return torch.einsum("bfij,fj->bfi", R[:, idx], lp) + p[:, idx]
