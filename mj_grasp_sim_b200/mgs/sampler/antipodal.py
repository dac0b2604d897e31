"""AntipodalGraspGenerator - batched restatement of /root/reference/mgs/sampler/antipodal.py:28-298.

The reference walks over the sampled surface points one by one and casts two rays per point through trimesh.  Here all
points are processed at once: area-weighted surface sampling, von Mises-Fisher directions around the inward normal
(closed-form 3-D sampler), Moeller-Trumbore ray/triangle tests for every (ray, face) pair as one torch tensor expression
(on the GPU when there is one), a random valid hit per point, the 10 cm-cube fallback, and the frame construction of
`define_gripper_pose` (:181-298).  Same outputs: `(Hs float64 [num,4,4], {"width": float64 [num]})` in the object's frame.

Reference behaviour that is kept:
  * the mesh is normalised first (unit AABB diagonal, area-weighted centroid at the origin, :60-93) and poses / widths are
    mapped back (:40-58); the offset stored for the way back is the centroid of the UNSCALED mesh, so positions come out
    as (p + c) * scale rather than p * scale + c (:76-79 vs :44) - identical when the centroid is at the origin;
  * eps rejects hits closer than 1e-5 (normalised units) to the ray origin; points without a valid hit get a second
    contact drawn uniformly from a +-0.05 cube around the first (normalised units, :139-144);
  * only the first `num` of the 5 * num sampled points are ever used (every point yields a pair, :113-151).
Not reproducible: the reference's random stream (trimesh + scipy + numpy global state); seeds here are explicit.
"""
from __future__ import annotations

from typing import Any, Dict, Tuple

import numpy as np

from .base import GraspGenerator


def _vmf3(mu: np.ndarray, kappa: float, rng: np.random.Generator) -> np.ndarray:
    """One von Mises-Fisher draw per row of mu (unit vectors, 3-D closed form: w = 1 + log(u + (1-u) e^{-2k}) / k)."""
    n = len(mu)
    u = rng.uniform(size=n)
    w = 1.0 + np.log(u + (1.0 - u) * np.exp(-2.0 * kappa)) / kappa
    phi = rng.uniform(0.0, 2.0 * np.pi, size=n)
    s = np.sqrt(np.clip(1.0 - w * w, 0.0, None))
    # orthonormal basis (a, b, mu) per row
    helper = np.where(np.abs(mu[:, :1]) < 0.9, np.array([[1.0, 0.0, 0.0]]), np.array([[0.0, 1.0, 0.0]]))
    a = np.cross(mu, helper)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = np.cross(mu, a)
    d = w[:, None] * mu + s[:, None] * (np.cos(phi)[:, None] * a + np.sin(phi)[:, None] * b)
    return d / np.linalg.norm(d, axis=1, keepdims=True)


class AntipodalGraspGenerator(GraspGenerator):
    def __init__(self, object, device: str | None = None, seed: int | None = None):
        super().__init__(object)
        self.verts, self.tris = (np.asarray(x) for x in object.mesh())
        self.scale, self.offset = 1.0, np.zeros(3)
        self.device = device
        self.rng = np.random.default_rng(seed)

    # ---- normalisation (:40-93) -------------------------------------------------------------------------------
    @staticmethod
    def _area_centroid(v, t):
        T = v[t]
        area = 0.5 * np.linalg.norm(np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0]), axis=1)
        return (T.mean(axis=1) * area[:, None]).sum(axis=0) / area.sum(), area

    def normalize_load(self):
        v = self.verts.astype(np.float64)
        self.scale = float(np.linalg.norm(v.max(axis=0) - v.min(axis=0)))  # trimesh.Trimesh.scale: AABB diagonal
        c0, _ = self._area_centroid(v, self.tris)
        self.offset = -c0  # centroid of the UNSCALED mesh (reference quirk, see the module docstring)
        vs = v / self.scale
        c1, _ = self._area_centroid(vs, self.tris)
        self.nverts = vs - c1

    def denormalize_points(self, points):
        return (points - self.offset) * self.scale

    def _use_kernel(self) -> bool:
        """CUDA kernel when a device is there and the caller did not ask for the host expression (device="cpu")."""
        if self.device == "cpu":
            return False
        import torch
        return torch.cuda.is_available()

    # ---- ray casting: every ray against every face (host / torch expression; the reference check of the kernel) ---------
    def _ray_hits(self, origins: np.ndarray, dirs: np.ndarray) -> np.ndarray:
        """distance along each ray to each face, inf where the ray misses: float64 [n_rays, n_faces]"""
        import torch
        dev = self.device or ("cuda" if torch.cuda.is_available() else "cpu")
        T = torch.as_tensor(self.nverts[self.tris], dtype=torch.float64, device=dev)
        o = torch.as_tensor(origins, dtype=torch.float64, device=dev)[:, None, :]
        d = torch.as_tensor(dirs, dtype=torch.float64, device=dev)[:, None, :]
        e1, e2 = (T[:, 1] - T[:, 0])[None], (T[:, 2] - T[:, 0])[None]
        p = torch.cross(d.expand(-1, e2.shape[1], -1), e2.expand(d.shape[0], -1, -1), dim=-1)
        det = (e1 * p).sum(-1)
        ok = det.abs() > 1e-14
        inv = torch.where(ok, 1.0 / torch.where(ok, det, torch.ones_like(det)), torch.zeros_like(det))
        tv = o - T[None, :, 0]
        u = (tv * p).sum(-1) * inv
        q = torch.cross(tv, e1.expand(tv.shape[0], -1, -1), dim=-1)
        v = (d * q).sum(-1) * inv
        t = (e2 * q).sum(-1) * inv
        hit = ok & (u >= -1e-12) & (v >= -1e-12) & (u + v <= 1.0 + 1e-12) & (t > 0)
        return torch.where(hit, t, torch.full_like(t, float("inf"))).cpu().numpy()

    # ---- the generator (:96-179) ---------------------------------------------------------------------------------
    def generate_grasps(self, num: int, kappa: float = 10.0, eps: float = 1e-5) -> Tuple[np.ndarray, Dict[str, Any]]:
        self.normalize_load()
        rng = self.rng
        T = self.nverts[self.tris]
        fn = np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0])
        area = 0.5 * np.linalg.norm(fn, axis=1)
        fn /= np.linalg.norm(fn, axis=1, keepdims=True)
        # only the first `num` of the reference's 5 * num samples are ever consumed
        f = rng.choice(len(self.tris), size=num, p=area / area.sum())
        r1, r2 = rng.uniform(size=num), rng.uniform(size=num)
        flip = r1 + r2 > 1.0
        r1[flip], r2[flip] = 1.0 - r1[flip], 1.0 - r2[flip]
        p1 = T[f, 0] + r1[:, None] * (T[f, 1] - T[f, 0]) + r2[:, None] * (T[f, 2] - T[f, 0])
        dirs = _vmf3(-fn[f], kappa, rng)
        # two rays per point (+dir, -dir); valid hits are at least eps away; one valid hit chosen uniformly at random
        pick_u = rng.uniform(size=num)
        if self._use_kernel():
            # the hand-written kernel (csrc/mgs_sampler.cu through mgs_antipodal_hits): one warp per point, triangles in shared memory
            from ...lib import antipodal_hits
            import torch
            dev = torch.cuda.current_device() if self.device in (None, "cuda") else int(str(self.device).split(":")[1])
            st, nvalid = antipodal_hits(p1, dirs, T, eps, pick_u, device=dev)
            p2 = p1 + np.where(nvalid > 0, st, 0.0)[:, None] * dirs
        else:
            t = np.concatenate([self._ray_hits(p1, dirs), self._ray_hits(p1, -dirs)], axis=1)  # [num, 2 * faces]
            sign = np.concatenate([np.ones(len(self.tris)), -np.ones(len(self.tris))])
            valid = np.isfinite(t) & (t >= eps)
            nvalid = valid.sum(axis=1)
            pick = (pick_u * np.maximum(nvalid, 1)).astype(int)
            order = np.argsort(~valid, axis=1, kind="stable")  # valid hits first, in face order
            col = order[np.arange(num), np.minimum(pick, np.maximum(nvalid - 1, 0))]
            p2 = p1 + (sign[col] * t[np.arange(num), col])[:, None] * dirs
        nohit = nvalid == 0
        p2[nohit] = p1[nohit] + rng.uniform(-0.05, 0.05, size=(int(nohit.sum()), 3))
        Hs = self.denorm_grasp_pose(self.define_gripper_pose(p1, p2, rng))
        widths = np.maximum(np.linalg.norm(p2 - p1, axis=1), 0)
        return Hs, {"width": widths * self.scale, "fallback": nohit}

    def denorm_grasp_pose(self, Hs):
        Hs[..., :3, 3] = self.denormalize_points(Hs[..., :3, 3])
        return Hs

    @classmethod
    def define_gripper_pose(cls, contact_one: np.ndarray, contact_two: np.ndarray, rng: np.random.Generator | None = None) -> np.ndarray:
        """x = contact_two - contact_one (normalised; [1,0,0] for coincident contacts), z = x cross a random vector
        (re-drawn while parallel, :236-252), y = z cross x, origin = midpoint (:181-298)."""
        rng = rng or np.random.default_rng()
        c1, c2 = np.atleast_2d(contact_one).astype(np.float64), np.atleast_2d(contact_two).astype(np.float64)
        assert len(c1) == len(c2)
        n = len(c1)
        x = c2 - c1
        norm = np.linalg.norm(x, axis=1, keepdims=True)
        bad = np.isclose(norm, 0.0).flatten()
        x = np.where(bad[:, None], np.array([[1.0, 0.0, 0.0]]), x / np.where(bad[:, None], 1.0, norm))
        z = np.cross(x, rng.normal(size=(n, 3)))
        zn = np.linalg.norm(z, axis=1)
        for _ in range(10):
            again = np.isclose(zn, 0.0)
            if not again.any():
                break
            z[again] = np.cross(x[again], rng.uniform(size=(int(again.sum()), 3)))
            zn = np.linalg.norm(z, axis=1)
        z /= zn[:, None]
        y = np.cross(z, x)
        Hs = np.zeros((n, 4, 4))
        Hs[:, :3, 0], Hs[:, :3, 1], Hs[:, :3, 2], Hs[:, :3, 3], Hs[:, 3, 3] = x, y, z, 0.5 * (c1 + c2), 1.0
        return Hs
