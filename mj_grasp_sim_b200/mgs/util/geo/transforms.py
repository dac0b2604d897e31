"""Rigid poses with the reference's `SE3Pose` contract (/root/reference/mgs/util/geo/transforms.py:28-128).

The contract matters for parity because it fixes the rounding of every initial state that enters the simulator
(SURVEY 8(a) row a10): positions and quaternions are held in float32, pose products go through float32 4x4 matrices and
back through scipy, and a quaternion must be unit length to rtol 1e-4 or construction fails with AssertionError.
`inverse()` returns the inverse AND overwrites the pose it was called on (reference :105-106) - kept.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation

_ORDERS = ("wxyz", "xyzw")


def quat_wxyz_to_xyzw(q):
    return np.roll(np.asarray(q), -1, axis=-1)


def quat_xyzw_to_wxyz(q):
    return np.roll(np.asarray(q), 1, axis=-1)


def _reorder(q, src: str, dst: str):
    """quaternion component order src -> dst (each "wxyz" or "xyzw")"""
    if dst not in _ORDERS:
        raise ValueError
    if src == dst:
        return q
    return quat_wxyz_to_xyzw(q) if src == "wxyz" else quat_xyzw_to_wxyz(q)


class SE3Pose:
    """One pose or a batch of poses: `pos [..., 3]`, `quat [..., 4]` in the component order named by `type`."""

    __slots__ = ("pos", "quat", "type")

    def __init__(self, pos, quat, type):
        pos, quat = np.asarray(pos), np.asarray(quat)
        assert pos.shape[-1] == 3 and quat.shape[-1] == 4
        assert type in _ORDERS
        self.pos, self.quat, self.type = pos.astype(np.float32), quat.astype(np.float32), type
        sq = (self.quat ** 2).sum(axis=-1, keepdims=True)
        assert np.all(np.isclose(sq, np.ones_like(sq), rtol=1e-4))

    def __repr__(self):
        return f"SE3Pose(pos={self.pos!r}, quat={self.quat!r}, type={self.type!r})"

    def __eq__(self, other):
        return isinstance(other, SE3Pose) and self.type == other.type and np.array_equal(self.pos, other.pos) and np.array_equal(self.quat, other.quat)

    def __len__(self):
        return len(self.pos)

    def __getitem__(self, idx):
        return type(self)(self.pos[idx], self.quat[idx], self.type)

    # ---- rotation helpers ------------------------------------------------------------------------------------
    def _scipy(self) -> Rotation:
        return Rotation.from_quat(np.array(_reorder(self.quat, self.type, "xyzw")))

    # ---- conversions -------------------------------------------------------------------------------------------
    def to_mat(self) -> np.ndarray:
        """float32 homogeneous matrices [..., 4, 4]"""
        out = np.zeros(self.quat.shape[:-1] + (4, 4), dtype=np.float32)
        out[..., 3, 3] = 1.0
        out[..., :3, 3] = self.pos
        out[..., :3, :3] = self._scipy().as_matrix()
        return out

    @classmethod
    def from_mat(cls, mat, type="wxyz"):
        assert mat.shape[-2:] == (4, 4)
        if type != "wxyz":
            raise ValueError
        q = Rotation.from_matrix(mat[..., :3, :3]).as_quat(canonical=False)
        return cls(mat[..., :3, 3], _reorder(q, "xyzw", "wxyz"), type)

    def to_vec(self, layout="pq", type=None) -> np.ndarray:
        q = np.copy(self.quat) if type is None else np.array(_reorder(self.quat, self.type, type))
        parts = {"pq": (self.pos, q), "qp": (q, self.pos)}.get(layout)
        return np.array([]) if parts is None else np.concatenate(parts, axis=-1)

    @classmethod
    def from_vec(cls, vec, type="wxyz", layout="pq"):
        assert vec.shape[-1] == 7
        if layout == "pq":
            return cls(vec[..., -7:-4], vec[..., -4:], type)
        if layout == "qp":
            return cls(vec[..., 4:7], vec[..., 0:4], type)
        raise ValueError

    # ---- group operations --------------------------------------------------------------------------------------
    def __matmul__(self, other):
        # float32 product of the two 4x4 matrices, then back through scipy (this is where poses get their rounding)
        return self.from_mat(np.einsum("...ij,...jk->...ik", self.to_mat(), other.to_mat()), type=self.type)

    def inverse(self):
        r_inv = self._scipy().inv()
        pos = (-r_inv.apply(self.pos)).astype(np.float32)
        quat = np.asarray(_reorder(r_inv.as_quat(), "xyzw", self.type)).astype(np.float32)
        self.pos, self.quat = pos, quat  # the reference's inverse() is in-place as well as returning
        return SE3Pose(pos, quat, self.type)

    @classmethod
    def randn_se3(cls, num):
        return cls(np.random.randn(num, 3), _reorder(Rotation.random(num).as_quat(canonical=False), "xyzw", "wxyz"), "wxyz")
