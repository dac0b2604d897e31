"""ContactBasedDiff - the contact-based differentiable grasp sampler of the dexterous hands in PyTorch (SURVEY 8(f) row 4;
reference /root/reference/mgs/sampler/contact.py:161-297 in JAX / Flax / optax).

Same procedure: sample the object surface (area weighted, max(30000, 3 num) points), pick `num` seed points by farthest-point
sampling, give every seed as many contact targets as the hand has fingertips (random seeds within a 10 cm neighbourhood, offset 2 cm
along their normals), start the hand 5 cm above the seed with its approach axis along the seed normal (`align_to_approach` of the
hand), assign fingertips to targets by the best permutation, and run 150 AdamW steps (lr 0.005) on the 6-D rotation, the position
and the joints of every grasp at once, minimising the squared fingertip-to-target distance plus 0.001 x the pad-normal misalignment;
joints are clipped to their ranges after every step.  Output like the reference: `(Hs [num, 4, 4] base pose, {"joints": [num, ndof]})`.

What is kept on purpose: the 6-D rotation convention (rows b1, b2, b3: kin/jax_util.py:150-163), `normalize_vector`'s epsilon,
the cosine term using the targets' normals in their ORIGINAL order while the positions are re-assigned every step (contact.py:123-138),
one pad point per fingertip drawn once per call, optax.adamw's defaults (weight decay 1e-4).  Not reproducible: the reference's random
streams (trimesh, jax.random.PRNGKey(0)); seeds here are explicit.  The hand kinematics come from the compiled gripper model (`kin.py`).
"""
from __future__ import annotations

from itertools import permutations
from typing import Any, Dict, Tuple

import numpy as np

from .base import GraspGenerator
from .kin import HandKinematics

NUM_SURFACE_SAMPLES = 30000
LOCAL_REGION_RADIUS = 0.10
TARGET_OFFSET_DISTANCE = 0.02
POSE_OFFSET_DISTANCE = 0.05
N_STEPS = 150
LEARNING_RATE = 0.005


def normalize_vector(v, eps: float = 1e-6):
    return v / (np.linalg.norm(v, axis=-1, keepdims=True) + eps)


def farthest_point_sampling(x: np.ndarray, n: int) -> np.ndarray:
    """indices of n points, starting from point 0, each the farthest from those chosen so far (kin/jax_util.py:182-199)"""
    idx = np.zeros(n, dtype=np.int64)
    dist = np.full(len(x), np.inf)
    for i in range(1, n):
        dist = np.minimum(dist, ((x - x[idx[i - 1]]) ** 2).sum(axis=1))
        idx[i] = int(np.argmax(dist))
    return idx


def rotation_6d_to_matrix(d6):
    import torch
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = a1 / a1.norm(dim=-1, keepdim=True)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = b2 / b2.norm(dim=-1, keepdim=True)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def matrix_to_rotation_6d(m):
    return m[..., :2, :].reshape(*m.shape[:-2], 6)


def best_assignment(fingertips, targets, perms):
    """targets [B, nf, 3] reordered so that sum_k |fingertip_k - target_perm(k)| is minimal (jax_util.py:202-222)"""
    import torch
    dist = torch.cdist(fingertips, targets)                     # [B, nf, nf]
    k = torch.arange(perms.shape[1], device=perms.device)
    loss = dist[:, k[None, :], perms].sum(-1)                   # [B, nperm]
    best = perms[loss.argmin(dim=1)]                            # [B, nf]
    return torch.gather(targets, 1, best[..., None].expand(-1, -1, 3))


class ContactBasedDiff(GraspGenerator):
    def __init__(self, object, device: str | None = None, seed: int | None = 0):
        super().__init__(object)
        self.verts, self.tris = (np.asarray(x, dtype=np.float64) if i == 0 else np.asarray(x) for i, x in enumerate(object.mesh()))
        self.device = device
        self.rng = np.random.default_rng(seed)
        self.last_losses = None

    def update_object(self, object):
        self.verts, self.tris = np.asarray(object.mesh()[0], dtype=np.float64), np.asarray(object.mesh()[1])
        return self

    def _sample_surface(self, n):
        T = self.verts[self.tris]
        cr = np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0])
        area = 0.5 * np.linalg.norm(cr, axis=1)
        f = self.rng.choice(len(self.tris), size=n, p=area / area.sum())
        r1, r2 = self.rng.uniform(size=n), self.rng.uniform(size=n)
        flip = r1 + r2 > 1.0
        r1[flip], r2[flip] = 1.0 - r1[flip], 1.0 - r2[flip]
        p = T[f, 0] + r1[:, None] * (T[f, 1] - T[f, 0]) + r2[:, None] * (T[f, 2] - T[f, 0])
        return p, normalize_vector(cr[f] / np.maximum(np.linalg.norm(cr[f], axis=1, keepdims=True), 1e-30))

    def generate_grasps(self, num: int, gripper: HandKinematics) -> Tuple[np.ndarray, Dict[str, Any]]:
        import torch
        dev = torch.device(self.device or ("cuda" if torch.cuda.is_available() else "cpu"))
        kin, rng = gripper, self.rng
        points, normals = self._sample_surface(max(NUM_SURFACE_SAMPLES, num * 3))
        fps = farthest_point_sampling(points, num)
        seeds, seed_n = points[fps], normals[fps]
        dists = np.linalg.norm(seeds[:, None, :] - seeds[None, :, :], axis=-1)
        nf = len(kin.fingertip_body)
        # nf random seeds within the local region of every seed become its contact targets (:193-206)
        rv = np.where(dists < LOCAL_REGION_RADIUS, rng.uniform(size=dists.shape), -np.inf)
        chosen = np.argsort(rv, axis=1)[:, -nf:]
        tgt_p = seeds[chosen] + TARGET_OFFSET_DISTANCE * seed_n[chosen]
        tgt_n = seed_n[chosen]
        # approach frame: z = seed normal, x towards the nearest other seed, y = z cross x (:208-222)
        z = seed_n
        x = normalize_vector(seeds[np.argsort(dists, axis=1)[:, 1]] - seeds) if num > 1 else np.tile([1.0, 0.0, 0.0], (num, 1))
        y = np.cross(z, x)
        R_init = np.stack([x, y, z], axis=-1)
        a_pos = np.einsum("nij,j->ni", R_init, kin.align_pos.astype(np.float64))
        R_init = np.einsum("nij,jk->nik", R_init, kin.align_rot.astype(np.float64))
        p_init = seeds + POSE_OFFSET_DISTANCE * seed_n + a_pos

        f32 = dict(device=dev, dtype=torch.float32)
        rot = matrix_to_rotation_6d(torch.as_tensor(R_init, **f32)).clone().requires_grad_(True)
        pos = torch.as_tensor(p_init, **f32).clone().requires_grad_(True)
        joints = torch.as_tensor(kin.init_pregrasp_joint, **f32)[None].repeat(num, 1).requires_grad_(True)
        lo, hi = (torch.as_tensor(kin.joint_ranges[:, k], **f32) for k in (0, 1))
        tp, tn = torch.as_tensor(tgt_p, **f32), torch.as_tensor(tgt_n, **f32)
        perms = torch.as_tensor(list(permutations(range(nf))), device=dev)
        pick = rng.integers(kin.local_fingertip_contact_positions.shape[1], size=nf)  # one pad point per fingertip, once per call (:243-249)
        pad = kin.local_fingertip_contact_positions[np.arange(nf), pick]
        neg_n = kin.neg_normal

        def world(theta, R6, t, local):
            return torch.einsum("bij,bfj->bfi", rotation_6d_to_matrix(R6), kin.fingertip_points(theta, local)) + t[:, None, :]

        opt = torch.optim.AdamW([rot, pos, joints], lr=LEARNING_RATE, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
        losses = None
        for _ in range(N_STEPS):
            opt.zero_grad(set_to_none=True)
            tips = world(joints, rot, pos, pad)
            origin = world(joints, rot, pos, np.zeros_like(pad))
            along = world(joints, rot, pos, neg_n)
            finger_n = along - origin
            loss_cos = (0.5 * (1.0 - (tn * finger_n).sum(-1))).mean(dim=1)
            assigned = best_assignment(tips.detach(), tp, perms)
            losses = ((assigned - tips) ** 2).mean(dim=(1, 2)) + 0.001 * loss_cos  # one independent problem per grasp
            losses.sum().backward()
            opt.step()
            with torch.no_grad():
                joints.copy_(torch.minimum(torch.maximum(joints, lo), hi))
        self.last_losses = losses.detach().cpu().numpy()
        with torch.no_grad():
            Hs = torch.zeros(num, 4, 4, **f32)
            Hs[:, :3, :3] = rotation_6d_to_matrix(rot)
            Hs[:, :3, 3] = pos
            Hs[:, 3, 3] = 1.0
        return Hs.cpu().numpy().astype(np.float64), {"joints": joints.detach().cpu().numpy().astype(np.float64)}
