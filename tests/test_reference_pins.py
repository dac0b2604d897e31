"""Pins of the MJCF compiler (and through it of every model the oracle and the kernel consume) to data the reference holds
from REAL MuJoCo runs, extracted by tools/extract_golden.py into tests/golden/reference_model_pins.json:

  * the `segmentation:` lists of mgs/cli/config/gripper/{allegro,leap,panda,shadow,vx300}.yaml - MuJoCo geom ids of the geoms
    that move with each joint: they fix the compiler's geom numbering (visual geoms included, `discardvisual=false`), the
    body tree and the joint -> body map;
  * /root/reference/segments.txt - MuJoCo's own (geom id -> name) table of the LEAP scan scene and the ids of the rendered
    unnamed geoms of the Shadow scene;
  * the second recorded Robotiq closed posture (robotiq_2f_85.yaml:7).
Plus two checks that do NOT share code with the compiler: `body_invweight0` / `dof_invweight0` / `meaninertia` recomputed by
brute force (finite-difference Jacobians of the oracle's kinematics, dense inverse of its mass matrix), and the mesh mass
properties against closed forms.  Together they break the loop "oracle and kernel consume the same compiled model"."""
import json
import os

import numpy as np
import pytest

from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.compiler import mesh as meshlib
from mj_grasp_sim_b200.compiler.mjcf import compile_mjcf
from oracle.oracle import OracleSim

HERE = os.path.dirname(os.path.abspath(__file__))
PINS = json.load(open(os.path.join(HERE, "golden", "reference_model_pins.json")))
# segmentation key -> joint name of the compiled model
JOINT_OF = {
    "allegro": lambda k: k,
    "leap": lambda k: k,
    "panda": lambda k: {"left_finger": "finger_joint1", "right_finger": "finger_joint2"}[k],
    "vx300": lambda k: k,
    "shadow": lambda k: "rh_" + k.split("_")[0].upper() + k.split("_")[1].upper(),
}


def lone_gripper(key):
    gx, ga = scenes.gripper_fragment(key)
    return compile_mjcf("<mujoco><compiler angle='radian' autolimits='true' discardvisual='false'/>" + gx + "</mujoco>", ga)


def subtree_geoms(m, body):
    gb, par = np.asarray(m.arr["geom_bodyid"]), np.asarray(m.arr["body_parentid"])
    out = []
    for g in range(int(m.arr["ngeom"])):
        b = int(gb[g])
        while b != 0 and b != body:
            b = int(par[b])
        if b == body:
            out.append(g)
    return out


@pytest.mark.parametrize("key", ["allegro", "leap", "panda", "shadow", "vx300"])
def test_segmentation_geom_ids(key):
    """For every joint the reference lists the MuJoCo ids of the rendered geoms that move with it: the visual geoms
    (contype = conaffinity = 0) of the joint's body subtree; for LEAP, whose collision geoms are rendered too, all geoms."""
    m = lone_gripper(key)
    visual = (np.asarray(m.arr["geom_contype"]) == 0) & (np.asarray(m.arr["geom_conaffinity"]) == 0)
    checked = 0
    for k, ids in PINS["segmentation"][key].items():
        if k in ("body", "palm"):
            assert max(ids) < int(m.arr["ngeom"])
            continue
        j = m.names["joint"][JOINT_OF[key](k)]
        sub = subtree_geoms(m, int(m.jnt_bodyid[j]))
        want = sub if key == "leap" else [g for g in sub if visual[g]]
        assert sorted(ids) == want, (key, k, ids, want)
        checked += 1
    assert checked == len(PINS["segmentation"][key]) - 1


def test_leap_geom_names_follow_mujoco_numbering():
    m = lone_gripper("leap")
    name_of = {v: k for k, v in m.names["geom"].items()}
    assert int(m.arr["ngeom"]) == 88
    for gid, name in PINS["leap_geom_names"]:
        assert name_of.get(gid) == name, (gid, name, name_of.get(gid))
    assert len(PINS["leap_geom_names"]) >= 60


def test_shadow_rendered_geoms_are_the_unnamed_visual_ones():
    m = lone_gripper("shadow")
    visual = (np.asarray(m.arr["geom_contype"]) == 0) & (np.asarray(m.arr["geom_conaffinity"]) == 0)
    named = set(m.names["geom"].values())
    ids = PINS["shadow_rendered_unnamed_geoms"]
    assert ids == PINS["segmentation"]["shadow"]["palm"]
    for g in ids:
        assert visual[g] and g not in named, g
    # and they are ALL the visual geoms from the wrist plate on (the forearm's are outside the camera frustum)
    assert [g for g in range(int(m.arr["ngeom"])) if visual[g]] == ids


def test_second_recorded_robotiq_closed_posture():
    """robotiq_2f_85.yaml:7 (commented `qpos: [...] # close`): another MuJoCo-recorded closed posture of the 8 joints the gripper
    class lists, taken under conditions the reference does not state (an earlier asset or scene: its follower joints sit 1.6e-2
    rad from the `state_close` record of the same file).  It cannot be reproduced exactly, but it bounds the answer: the oracle's
    settled posture must be no farther from it, joint by joint, than the reference's own other record is (+1e-3), with the same
    signs and the same left/right symmetry."""
    from test_golden_robotiq import SCAN_XML
    gx, ga = scenes.gripper_fragment("robotiq2f85")
    m = compile_mjcf(SCAN_XML.format(gripper=gx), ga)
    s = OracleSim(m)
    s.reset()
    s.qpos[0:3] = [0, 0, -0.15]
    s.mocap_pos[0] = [0, 0, -0.15]
    s.ctrl[:] = 255.0
    s.step(2500)
    names = ["right_driver_joint", "right_coupler_joint", "right_spring_link_joint", "right_follower_joint",
             "left_driver_joint", "left_coupler_joint", "left_spring_link_joint", "left_follower_joint"]
    adr = [int(m.jnt_qposadr[m.names["joint"][n]]) for n in names]
    q = np.array([s.qpos[a] for a in adr])
    rec2 = np.array(PINS["robotiq_close_qpos8"])
    rec1 = np.array(json.load(open(os.path.join(HERE, "golden", "robotiq_2f85_state_close.json")))["state"])[1:1 + m.nq][adr]
    assert (np.abs(q - rec2) <= np.abs(rec1 - rec2) + 1e-3).all(), (q, rec1, rec2)
    assert np.abs(rec1 - rec2).max() < 2e-2
    assert np.abs(q[:4] - q[4:]).max() < 1e-4 and np.abs(rec2[:4] - rec2[4:]).max() < 1e-4
    assert (np.sign(q) == np.sign(rec2)).all()


# ---------------------------------------------------------------------------------------------------------------------------
def _perturbed_qpos(m, q0, dof, eps):
    """qpos0 moved by eps along dof `dof` (free-joint rotations: body-frame axis, quaternion exponential)."""
    q = q0.copy()
    j = int(np.asarray(m.arr["dof_jntid"])[dof])
    k = dof - int(m.jnt_dofadr[j])
    a = int(m.jnt_qposadr[j])
    if int(m.jnt_type[j]) == 0:  # free
        if k < 3:
            q[a + k] += eps
        else:
            w = np.zeros(3)
            w[k - 3] = eps
            dq = np.array([np.cos(0.5 * eps), *(np.sin(0.5 * eps) * w / eps)])
            qa = q[a + 3:a + 7]
            q[a + 3:a + 7] = [qa[0] * dq[0] - qa[1:] @ dq[1:], *(qa[0] * dq[1:] + dq[0] * qa[1:] + np.cross(qa[1:], dq[1:]))]
    else:
        q[a] += eps
    return q


@pytest.mark.parametrize("key", ["panda", "robotiq2f85", "shadow"])
def test_invweight0_and_meaninertia_by_brute_force(key):
    """MuJoCo's set0: body_invweight0[b] = mean diagonal of J M^-1 J' (translation | rotation) with J the Jacobian of the body's
    CoM frame at qpos0; dof_invweight0 = diag(M^-1) averaged per free-joint triple; meaninertia = trace(M) / nv.  Recomputed here
    from central finite differences of the ORACLE's forward kinematics and a dense inverse - no code shared with the compiler."""
    ox, oa, _ = scenes.hull_object_fragment(3, 16)
    model, _ = scenes.build_scene(key, ox, oa)
    s = OracleSim(model)
    nv, nb = model.nv, model.nbody
    q0 = np.array(model.qpos0, dtype=np.float64)

    def frames(q):
        s.reset()
        s.qpos[:] = q
        s.kinematics()
        return s.xipos.copy(), s.xmat.reshape(nb, 3, 3).copy()
    s.reset()
    s.kinematics()
    M = s.M.copy()
    Minv = np.linalg.inv(M)
    eps = 1e-6
    Jt, Jr = np.zeros((nb, 3, nv)), np.zeros((nb, 3, nv))
    for d in range(nv):
        (pp, Rp), (pm, Rm) = frames(_perturbed_qpos(model, q0, d, eps)), frames(_perturbed_qpos(model, q0, d, -eps))
        Jt[:, :, d] = (pp - pm) / (2 * eps)
        dR = np.einsum("bij,bkj->bik", Rp, Rm)  # R+ R-^T = exp([w] 2 eps)
        Jr[:, 0, d], Jr[:, 1, d], Jr[:, 2, d] = (dR[:, 2, 1] - dR[:, 1, 2]) / (4 * eps), (dR[:, 0, 2] - dR[:, 2, 0]) / (4 * eps), (dR[:, 1, 0] - dR[:, 0, 1]) / (4 * eps)
    biw = np.asarray(model.arr["body_invweight0"]).reshape(nb, 2)
    for b in range(1, nb):
        At, Ar = Jt[b] @ Minv @ Jt[b].T, Jr[b] @ Minv @ Jr[b].T
        want = np.array([np.trace(At) / 3, np.trace(Ar) / 3])
        if not (np.abs(Jt[b]).max() > 0 or np.abs(Jr[b]).max() > 0):
            assert np.all(biw[b] == 0)  # static / mocap bodies
            continue
        assert np.allclose(biw[b], want, rtol=2e-5, atol=1e-9), (b, biw[b], want)
    dinv = np.diag(Minv).copy()
    for j in range(int(model.arr["njnt"])):
        if int(model.jnt_type[j]) == 0:
            a = int(model.jnt_dofadr[j])
            dinv[a:a + 3], dinv[a + 3:a + 6] = dinv[a:a + 3].mean(), dinv[a + 3:a + 6].mean()
    assert np.allclose(np.asarray(model.arr["dof_invweight0"]), dinv, rtol=1e-9)
    assert abs(float(model.arr["meaninertia"]) - np.trace(M) / nv) < 1e-12 * max(1.0, np.trace(M))


def test_mesh_mass_properties_closed_forms():
    # box a x b x c: V = abc, C = V/12 diag(a^2, b^2, c^2) about the centre, wherever the box sits
    h = meshlib.box_hull([0.03, 0.02, 0.05])
    shift = np.array([0.4, -0.2, 0.1])
    V, c, C = meshlib.mass_properties(h.verts + shift, h.tri)
    a, b, cc = 0.06, 0.04, 0.10
    assert abs(V - a * b * cc) < 1e-15 and np.allclose(c, shift, atol=1e-12)
    assert np.allclose(C, V / 12 * np.diag([a * a, b * b, cc * cc]), atol=1e-18)
    I = meshlib.cov_to_inertia(C)
    assert np.allclose(np.diag(I), V / 12 * np.array([b * b + cc * cc, a * a + cc * cc, a * a + b * b]))
    # tetrahedron with vertices 0, e1, e2, e3 scaled by s: V = s^3/6, CoM = s/4 (1,1,1), C_ii = 3 s^5/480 ... check via rotation invariance instead
    s = 0.07
    tv = np.array([[0, 0, 0], [s, 0, 0], [0, s, 0], [0, 0, s]], dtype=float)
    tt = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]])
    V, c, C = meshlib.mass_properties(tv, tt)
    assert abs(V - s ** 3 / 6) < 1e-18 and np.allclose(c, [s / 4] * 3)
    # exact: int x^2 over the unit corner tetrahedron = 1/60, int xy = 1/120; about the CoM: 1/60 - 1/96 = 3/480, 1/120 - 1/96 = -1/480
    assert np.allclose(C, s ** 5 * (np.eye(3) * (3 / 480 + 1 / 480) - np.ones((3, 3)) / 480), atol=1e-20)
    from scipy.spatial.transform import Rotation
    Rm = Rotation.from_euler("xyz", [0.3, -1.1, 2.0]).as_matrix()
    V2, c2, C2 = meshlib.mass_properties(tv @ Rm.T + shift, tt)
    assert abs(V2 - V) < 1e-18 and np.allclose(c2, Rm @ c + shift) and np.allclose(C2, Rm @ C @ Rm.T, atol=1e-20)
    # finely tessellated ellipsoid: V -> 4/3 pi abc, C -> V/5 diag(a^2, b^2, c^2)
    rng = np.random.default_rng(0)
    p = rng.normal(size=(6000, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    ax = np.array([0.03, 0.05, 0.02])
    hh = meshlib.build_hull(p * ax)
    V, c, C = meshlib.mass_properties(hh.verts, hh.tri)
    Vx = 4 / 3 * np.pi * ax.prod()
    assert abs(V / Vx - 1) < 5e-3 and np.abs(c).max() < 1e-4
    assert np.allclose(np.diag(C) / (Vx / 5 * ax ** 2), 1.0, atol=1e-2)
