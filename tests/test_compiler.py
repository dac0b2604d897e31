"""Model compiler: dimensions, id ordering, contact mixing."""
import numpy as np
import pytest

from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.compiler import mesh as meshlib
from mj_grasp_sim_b200.compiler.mjcf import compile_mjcf, mass_matrix


def test_panda_cube_dimensions(panda_cube):
    m = panda_cube[0]
    # nq 16, nv 14, nu 2, 7 bodies, 13 gripper colliders + ground + cube
    assert (m.nq, m.nv, m.nu, m.nbody) == (16, 14, 2, 7)
    assert int(m.arr["ncgeom"]) == 15 and int(m.arr["neq"]) == 1 and int(m.arr["ntendon"]) == 1
    assert int((m.dof_frictionloss > 0).sum()) == 2
    # geom-id order the labels rely on: gripper < ground < object
    g = m.names["geom"]
    assert g["panda_col_12"] < g["geom:ground"] < g["geom:cube"]
    # reference segmentation ids (mgs/cli/config/gripper/panda.yaml): finger geoms start at 6 and 14
    assert g["panda_col_1"] == 8 and g["panda_col_7"] == 16 and g["geom:ground"] == 22


def test_panda_inertials_and_actuators(panda_cube):
    m = panda_cube[0]
    b = m.names["body"]
    assert np.isclose(m.body_mass[b["hand"]], 0.73) and np.isclose(m.body_mass[b["left_finger"]], 0.015)
    assert np.isclose(m.body_mass[b["cube"]], 1.0)
    assert np.allclose(m.body_inertia[b["cube"]], 1.0 / 3 * 2 * 0.02 ** 2)  # m/3 (b^2 + c^2)
    assert np.allclose(m.actuator_gainprm[:, 0], 1000) and np.allclose(m.actuator_biasprm[:, 1], -1000)
    assert np.allclose(m.actuator_forcerange, [[-15, 15], [-15, 15]]) and m.actuator_ctrllimited.all()
    assert m.jnt_limited[m.names["joint"]["finger_joint1"]] == 1
    # weld: mocap and hand start at the same pose -> identity relpose
    assert np.allclose(m.eq_data[0], [0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1])


def test_contact_mixing_examples(panda_hull):
    m = panda_hull[0]
    cg = {int(g): i for i, g in enumerate(m.cgeom_geomid)}
    pad = cg[m.names["geom"]["panda_col_2"]]
    obj = int(m.arr["ncgeom"]) - 1
    p = [i for i in range(int(m.arr["npair"])) if m.pair_geom1[i] == pad and m.pair_geom2[i] == obj][0]
    # Appendix C: Panda pad x recipe hull -> condim 4, mu (2.4, .3, .1), solref (0.0105, 1), solimp (.949,.974,.001,.5,2)
    assert m.pair_condim[p] == 4
    assert np.allclose(m.pair_friction[p], [2.4, 2.4, 0.3, 0.1, 0.1])
    assert np.allclose(m.pair_solref[p], [0.0105, 1.0])
    assert np.allclose(m.pair_solimp[p], [0.949, 0.974, 0.001, 0.5, 2.0])


def test_pair_filtering(panda_cube):
    m = panda_cube[0]
    body = m.cgeom_bodyid
    b = m.names["body"]
    pairs = {(int(body[a]), int(body[c])) for a, c in zip(m.pair_geom1, m.pair_geom2)}
    assert (b["hand"], b["left_finger"]) not in pairs and (b["hand"], b["right_finger"]) not in pairs  # <exclude>
    assert (b["left_finger"], b["right_finger"]) in pairs  # siblings do collide
    assert int(m.arr["npair"]) == 36 + 13 + 13 + 1


def test_mass_matrix_spd_and_invweights(panda_cube):
    m = panda_cube[0]
    M, _ = mass_matrix(m, m.qpos0)
    assert np.allclose(M, M.T) and np.linalg.eigvalsh(M).min() > 0
    b = m.names["body"]
    assert np.isclose(m.body_invweight0[b["cube"], 0], 1.0)  # 1/m for a free body
    assert np.isclose(m.body_invweight0[b["cube"], 1], 1 / m.body_inertia[b["cube"], 0])
    assert np.all(m.body_invweight0[b["mocap"]] == 0)


def test_hull_builder_box_and_polygons():
    h = meshlib.box_hull([0.01, 0.02, 0.03])
    assert len(h.verts) == 8 and len(h.face_num) == 6 and (h.face_num == 4).all()
    for f in range(6):
        idx = h.face_vert[h.face_adr[f]:h.face_adr[f] + 4]
        p = h.verts[idx]
        n = np.cross(p[1] - p[0], p[2] - p[1])
        assert np.dot(n, h.face_normal[f]) > 0  # CCW seen from outside
    assert (h.nbr_num >= 3).all()  # cube edges (+ triangulation diagonals; a superset is fine for hill climbing)
    V, com, C = meshlib.mass_properties(h.verts, h.tri)
    assert np.isclose(V, 8 * 0.01 * 0.02 * 0.03) and np.allclose(com, 0, atol=1e-12)
    I = meshlib.cov_to_inertia(C)
    assert np.allclose(np.diag(I), V / 3 * np.array([0.02 ** 2 + 0.03 ** 2, 0.01 ** 2 + 0.03 ** 2, 0.01 ** 2 + 0.02 ** 2]))


def test_include_and_defaults_leak():
    xml = """<mujoco><compiler angle="radian"/><default><geom friction="0.2"/></default>
    <include file="obj.xml"/></mujoco>"""
    inc = b"""<mujoco><worldbody><body name="o"><freejoint name="o:j"/><geom name="g" type="sphere" size="0.01"/></body></worldbody></mujoco>"""
    m = compile_mjcf(xml, {"obj.xml": inc})
    assert np.isclose(m.geom_friction[0, 0], 0.2)  # top-level default applies to included geoms (quirk 11)


def test_pose_processing_matches_se3pose_semantics():
    H = np.eye(4)[None].repeat(2, 0)
    H[1, :3, 3] = [0.01, 0.02, 0.03]
    p = scenes.process_poses(H, "panda")
    assert p.dtype == np.float32
    assert np.allclose(p[0], [0, 0, -0.102, 0.70710677, 0, 0, 0.70710677], atol=1e-7)
    assert np.allclose(p[1, :3], [0.01, 0.02, 0.03 - 0.102], atol=1e-7)


PRIM = """<mujoco><compiler angle="radian"/><worldbody><body name="b" pos="0.1 0.2 0.3"><freejoint/>
<geom type="{t}" size="{s}" density="850" pos="0.01 -0.02 0.03" quat="{q}"/></body></worldbody></mujoco>"""


@pytest.mark.parametrize("gtype,size", [("sphere", (0.03,)), ("box", (0.02, 0.03, 0.05)), ("cylinder", (0.02, 0.04)), ("capsule", (0.015, 0.035))])
def test_primitive_geom_mass_properties_match_closed_forms(gtype, size):
    """Mass, centre of mass and the full inertia tensor of a body made of one rotated primitive geom against the textbook closed forms
    (capsule = cylinder + two hemispheres, each (83 / 320) m r^2 about its own centre of mass, 3 r / 8 from its flat face)."""
    from mj_grasp_sim_b200.compiler.mjcf import quat_to_mat
    q = np.array([0.8, 0.3, -0.4, 0.33]); q /= np.linalg.norm(q)
    m = compile_mjcf(PRIM.format(t=gtype, s=" ".join("%g" % x for x in size), q=" ".join("%.17g" % x for x in q)))
    rho = 850.0
    if gtype == "sphere":
        r, = size
        mass = rho * 4 / 3 * np.pi * r ** 3
        I = np.full(3, 0.4 * mass * r * r)
    elif gtype == "box":
        a, b, c = size
        mass = rho * 8 * a * b * c
        I = mass / 3 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    elif gtype == "cylinder":
        r, h = size
        mass = rho * np.pi * r * r * 2 * h
        I = np.array([mass * (3 * r * r + 4 * h * h) / 12] * 2 + [0.5 * mass * r * r])
    else:
        r, h = size
        mc, ms = rho * np.pi * r * r * 2 * h, rho * 4 / 3 * np.pi * r ** 3
        mass = mc + ms
        side = mc * (3 * r * r + 4 * h * h) / 12 + 2 * (83 / 320 * (ms / 2) * r * r + (ms / 2) * (h + 3 * r / 8) ** 2)
        I = np.array([side, side, 0.5 * mc * r * r + 0.4 * ms * r * r])
    assert np.isclose(m.body_mass[1], mass, rtol=1e-12) and np.allclose(m.body_ipos[1], [0.01, -0.02, 0.03], atol=1e-15)
    Rg, Ri = quat_to_mat(q), quat_to_mat(m.body_iquat[1])
    assert np.allclose(Ri @ np.diag(m.body_inertia[1]) @ Ri.T, Rg @ np.diag(I) @ Rg.T, rtol=1e-10, atol=1e-16)


def test_unsupported_orientation_spellings_are_refused():
    for attr in ('euler="0.1 0.2 0.3"', 'axisangle="0 0 1 0.3"', 'zaxis="0 1 0"', 'xyaxes="1 0 0 0 1 0"'):
        with pytest.raises(NotImplementedError):
            compile_mjcf(PRIM.format(t="box", s="0.02 0.03 0.05", q="1 0 0 0").replace('quat="1 0 0 0"', attr))
