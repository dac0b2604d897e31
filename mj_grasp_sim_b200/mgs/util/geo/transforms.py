"""SE3Pose with the reference's semantics (/root/reference/mgs/util/geo/transforms.py:28-128):
float32 storage, scipy conversions, 4x4 products carried out in float32, unit-norm assertion with
rtol 1e-4, and `inverse()` that also mutates self (reference quirk, :105-106)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.spatial.transform import Rotation as R


def quat_wxyz_to_xyzw(q):
    return np.concatenate([q[..., 1:], q[..., :1]], axis=-1)


def quat_xyzw_to_wxyz(q):
    return np.concatenate([q[..., 3:], q[..., :3]], axis=-1)


@dataclass
class SE3Pose:
    pos: np.ndarray
    quat: np.ndarray
    type: str  # "wxyz" or "xyzw"

    def __post_init__(self):
        self.pos = np.asarray(self.pos)
        self.quat = np.asarray(self.quat)
        assert self.pos.shape[-1] == 3
        assert self.quat.shape[-1] == 4
        assert self.type in ("wxyz", "xyzw")
        self.pos = self.pos.astype(np.float32)
        self.quat = self.quat.astype(np.float32)
        norms = np.sum(self.quat ** 2, axis=-1, keepdims=True)
        assert np.all(np.isclose(norms, np.ones_like(norms), rtol=1e-4))

    def to_vec(self, layout="pq", type=None) -> np.ndarray:
        quat = np.copy(self.quat)
        if type is not None and type != self.type:
            if type == "wxyz":
                quat = quat_xyzw_to_wxyz(quat)
            elif type == "xyzw":
                quat = quat_wxyz_to_xyzw(quat)
            else:
                raise ValueError
        if layout == "pq":
            return np.concatenate([self.pos, quat], axis=-1)
        if layout == "qp":
            return np.concatenate([quat, self.pos], axis=-1)
        return np.array([])

    @classmethod
    def from_vec(cls, vec, type="wxyz", layout="pq"):
        assert vec.shape[-1] == 7
        if layout == "pq":
            return cls(vec[..., -7:-4], vec[..., -4:], type)
        if layout == "qp":
            return cls(vec[..., 4:7], vec[..., 0:4], type)
        raise ValueError

    @classmethod
    def from_mat(cls, mat, type="wxyz"):
        assert mat.shape[-2:] == (4, 4)
        quat = R.from_matrix(mat[..., :3, :3]).as_quat(canonical=False)
        if type != "wxyz":
            raise ValueError
        return cls(mat[..., :3, 3], quat_xyzw_to_wxyz(quat), type)

    def __getitem__(self, idx):
        return self.__class__(self.pos[idx], self.quat[idx], self.type)

    def __matmul__(self, other):
        res = np.einsum("...ij,...jk->...ik", self.to_mat(), other.to_mat())
        return self.__class__.from_mat(res, type=self.type)

    def __len__(self):
        return len(self.pos)

    def inverse(self):
        q = self.quat if self.type == "xyzw" else quat_wxyz_to_xyzw(self.quat)
        rinv = R.from_quat(q).inv()
        inv_q = rinv.as_quat()
        inv_quat = inv_q if self.type == "xyzw" else quat_xyzw_to_wxyz(inv_q)
        inv_pos = -rinv.apply(self.pos)
        self.pos = inv_pos.astype(np.float32)
        self.quat = inv_quat.astype(np.float32)
        return SE3Pose(inv_pos, inv_quat, self.type)

    def to_mat(self) -> np.ndarray:
        q = quat_wxyz_to_xyzw(self.quat) if self.type == "wxyz" else self.quat
        mat = np.zeros((*self.quat.shape[:-1], 4, 4), dtype=np.float32)
        mat[..., :3, :3] = R.from_quat(np.copy(q)).as_matrix()
        mat[..., :3, 3] = self.pos
        mat[..., 3, 3] = 1.0
        return mat

    @classmethod
    def randn_se3(cls, num):
        q = R.random(num).as_quat(canonical=False)
        return cls(np.random.randn(num, 3), quat_xyzw_to_wxyz(q), "wxyz")
