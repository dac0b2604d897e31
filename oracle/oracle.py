"""ctypes wrapper around liboracle.so (the CPU oracle - test infrastructure only).

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never from the
product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mj_grasp_sim_b200.model_desc import MgsModelDesc, make_desc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class RolloutCfg(C.Structure):
    _fields_ = [("nstep_close", C.c_int), ("nstep_lift", C.c_int), ("shake_steps", C.c_int), ("repose_on_close", C.c_int),
                ("lift_dist", C.c_double), ("shake_dist", C.c_double)]


def build(force: bool = False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "mgs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(MgsModelDesc)]
        for name in ("qpos", "qvel", "ctrl", "mocap_pos", "mocap_quat", "qacc_warmstart", "xpos", "xquat", "xmat", "xipos", "gxpos",
                     "gxmat", "M", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "qacc",
                     "qfrc_constraint", "J", "efc_pos", "efc_D", "efc_R", "efc_aref", "efc_force", "efc_jar", "subtree_com", "cdof"):
            f = getattr(L, "orc_" + name)
            f.restype = C.POINTER(C.c_double)
            f.argtypes = [C.c_void_p]
        L.orc_efc_type.restype = C.POINTER(C.c_int)
        L.orc_efc_type.argtypes = [C.c_void_p]
        for name in ("destroy", "reset", "forward", "kinematics_only", "collision_only"):
            getattr(L, "orc_" + name).argtypes = [C.c_void_p]
            getattr(L, "orc_" + name).restype = None
        for name in ("ncon", "nefc", "niter", "bad", "contact_with_object", "ncon_peak", "nefc_peak"):
            getattr(L, "orc_" + name).argtypes = [C.c_void_p]
            getattr(L, "orc_" + name).restype = C.c_int
        L.orc_set_qvel_clip.argtypes = [C.c_void_p, C.c_double]
        L.orc_set_qvel_clip.restype = None
        L.orc_set_analytic.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_analytic.restype = None
        L.orc_step.argtypes = [C.c_void_p, C.c_int]
        L.orc_step.restype = C.c_int
        L.orc_contact.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_place.argtypes = [C.c_void_p, dp, C.c_int, dp, ip, C.c_int]
        L.orc_grasp_collision.argtypes = [C.c_void_p, dp, C.c_int, dp, ip, C.c_int]
        L.orc_grasp_collision.restype = C.c_int
        L.orc_grasp_stability.argtypes = [C.c_void_p, dp, C.c_int, dp, ip, C.c_int, dp, C.POINTER(RolloutCfg), C.POINTER(C.c_longlong)]
        L.orc_grasp_stability.restype = C.c_int
        L.orc_batch.argtypes = [C.POINTER(MgsModelDesc), C.c_int, C.c_int, dp, C.c_int, dp, ip, C.c_int, dp, C.POINTER(RolloutCfg),
                                C.c_int, C.POINTER(C.c_ubyte), C.POINTER(C.c_longlong)]
        L.orc_batch.restype = C.c_int
        L.orc_batch_scene.argtypes = L.orc_batch.argtypes + [dp]
        L.orc_batch_scene.restype = C.c_int
        L.orc_set_record.argtypes = [C.c_void_p, dp]
        L.orc_get_record.argtypes = [C.c_void_p, dp]
        for name in ("gripper_collision", "gripper_contact"):
            getattr(L, "orc_" + name).argtypes = [C.c_void_p]
            getattr(L, "orc_" + name).restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleSim:
    """One fp64 environment.  Attribute names follow MuJoCo's mjData."""

    def __init__(self, model, ground_name="geom:ground"):
        self.model = model
        self.desc, self._keep = make_desc(model, ground_name)
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create(C.byref(self.desc)))
        m = model
        nb, ncg = m.nbody, int(m.arr["ncgeom"])
        self._shapes = dict(qpos=(m.nq,), qvel=(m.nv,), ctrl=(m.nu,), mocap_pos=(int(m.arr["nmocap"]), 3),
                            mocap_quat=(int(m.arr["nmocap"]), 4), qacc_warmstart=(m.nv,), xpos=(nb, 3), xquat=(nb, 4),
                            xmat=(nb, 9), xipos=(nb, 3), gxpos=(ncg, 3), gxmat=(ncg, 9), M=(m.nv, m.nv), qfrc_bias=(m.nv,),
                            qfrc_passive=(m.nv,), qfrc_actuator=(m.nv,), qfrc_smooth=(m.nv,), qacc_smooth=(m.nv,),
                            qacc=(m.nv,), qfrc_constraint=(m.nv,), subtree_com=(nb, 3), cdof=(m.nv, 6))

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def __getattr__(self, k):
        shapes = object.__getattribute__(self, "_shapes")
        if k in shapes:
            p = getattr(self.L, "orc_" + k)(self.h)
            n = int(np.prod(shapes[k]))
            if n == 0:
                return np.zeros(shapes[k])
            return np.ctypeslib.as_array(p, shape=(n,)).reshape(shapes[k])
        raise AttributeError(k)

    def efc(self, name):
        n = self.nefc
        if name == "J":
            p = self.L.orc_J(self.h)
            return np.ctypeslib.as_array(p, shape=(n * self.model.nv,)).reshape(n, self.model.nv).copy() if n else np.zeros((0, self.model.nv))
        if name == "type":
            return np.ctypeslib.as_array(self.L.orc_efc_type(self.h), shape=(n,)).copy() if n else np.zeros(0, dtype=np.int32)
        p = getattr(self.L, "orc_efc_" + name)(self.h)
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0)

    @property
    def ncon(self): return self.L.orc_ncon(self.h)
    @property
    def nefc(self): return self.L.orc_nefc(self.h)
    @property
    def niter(self): return self.L.orc_niter(self.h)
    @property
    def bad(self): return self.L.orc_bad(self.h)

    def contacts(self):
        out = np.zeros((self.ncon, 18))
        for c in range(self.ncon):
            self.L.orc_contact(self.h, c, _dp(out[c]))
        return out

    def reset(self): self.L.orc_reset(self.h)
    def forward(self): self.L.orc_forward(self.h)
    def step(self, n=1): return self.L.orc_step(self.h, n)
    def set_qvel_clip(self, clip): self.L.orc_set_qvel_clip(self.h, float(clip))
    def kinematics(self): self.L.orc_kinematics_only(self.h)
    def collision_only(self): self.L.orc_collision_only(self.h)
    def set_analytic(self, on): self.L.orc_set_analytic(self.h, int(bool(on)))
    def contact_with_object(self): return bool(self.L.orc_contact_with_object(self.h))

    def record_size(self):
        m = self.model
        return m.nq + 2 * m.nv + m.nu + 7 * int(m.arr["nmocap"])

    def get_record(self):
        r = np.zeros(self.record_size())
        self.L.orc_get_record(self.h, _dp(r))
        return r

    def set_record(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        self.L.orc_set_record(self.h, _dp(r))

    def gripper_collision(self): return bool(self.L.orc_gripper_collision(self.h))
    def gripper_contact(self): return bool(self.L.orc_gripper_contact(self.h))

    def place(self, pose7, base_qadr, joints, jadr):
        pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
        joints = np.ascontiguousarray(joints, dtype=np.float64)
        jadr = np.ascontiguousarray(jadr, dtype=np.int32)
        self.L.orc_place(self.h, _dp(pose7), int(base_qadr), _dp(joints), _ip(jadr), len(jadr))


def batch(model, mode, poses7, base_qadr, joints, jadr, close_ctrl, cfg: RolloutCfg, nthreads=1, scene=None, ground_name="geom:ground"):
    """mode 0: collision-free mask, mode 1: stability labels, 2 / 3: clutter-table versions (scene = state
    record qpos|qvel|qacc_warmstart|ctrl|mocap).  Returns (labels bool[N], steps int64[N])."""
    desc, keep = make_desc(model, ground_name)
    poses7 = np.ascontiguousarray(poses7, dtype=np.float64)
    joints = np.ascontiguousarray(joints, dtype=np.float64)
    jadr = np.ascontiguousarray(jadr, dtype=np.int32)
    close_ctrl = np.ascontiguousarray(close_ctrl, dtype=np.float64)
    n = len(poses7)
    labels = np.zeros(n, dtype=np.uint8)
    steps = np.zeros(n, dtype=np.int64)
    sc = np.ascontiguousarray(scene, dtype=np.float64) if scene is not None else None
    lib().orc_batch_scene(C.byref(desc), mode, n, _dp(poses7), int(base_qadr), _dp(joints), _ip(jadr), joints.shape[1], _dp(close_ctrl),
                          C.byref(cfg), nthreads, labels.ctypes.data_as(C.POINTER(C.c_ubyte)), steps.ctypes.data_as(C.POINTER(C.c_longlong)),
                          _dp(sc) if sc is not None else None)
    return labels.astype(bool), steps
