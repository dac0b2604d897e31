"""Closed-form primitive colliders of the oracle (sphere / capsule / box / cylinder pairs, the ones MuJoCo's collision table sends to
engine_collision_primitive.c) checked against two things that share no code with them: the oracle's own MPR on the same shallow
penetrations, and hand-computed known answers.  The kernel's copy of the same routines is checked against the oracle in
tests/test_lane1_vs_oracle.py (1-lane host build) and tests/test_gpu_parity.py (-m gpu)."""
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation as R

from mj_grasp_sim_b200.compiler.mjcf import compile_mjcf
from oracle.oracle import OracleSim

GEOMS = {
    "sphere": ('type="sphere" size="0.012"', 0.012),
    "sphere2": ('type="sphere" size="0.02"', 0.02),
    "capsule": ('type="capsule" size="0.009 0.02"', 0.029),
    "capsule2": ('type="capsule" size="0.012 0.01"', 0.022),
    "cylinder": ('type="cylinder" size="0.0135 0.015"', 0.0202),
    "box": ('type="box" size="0.02 0.03 0.015"', 0.039),
}

PAIR = """<mujoco><compiler angle="radian" autolimits="true"/>
<option integrator="implicitfast" timestep="0.001" cone="elliptic" gravity="0 0 0"/>
<worldbody><body name="a" pos="0 0 0"><freejoint name="ja"/><geom name="ga" {a} mass="0.1"/></body>
<body name="b" pos="0.2 0 0"><freejoint name="jb"/><geom name="gb" {b} mass="0.1"/></body></worldbody></mujoco>"""


def _sim(a, b):
    m = compile_mjcf(PAIR.format(a=GEOMS[a][0], b=GEOMS[b][0]))
    return OracleSim(m)


def _contacts(s, qa, qb, analytic):
    s.reset()
    s.qpos[0:7] = qa
    s.qpos[7:14] = qb
    s.set_analytic(analytic)
    s.collision_only()
    return s.contacts()


def _rand_quat(rng):
    q = R.random(random_state=rng).as_quat()  # xyzw
    return np.array([q[3], q[0], q[1], q[2]])


def _support_height(kind, q, n):
    """h(n) = max over the geom of x . n (world), written from the shape definitions, not from the oracle's support code."""
    size = {"sphere": (0.012,), "sphere2": (0.02,), "capsule": (0.009, 0.02), "capsule2": (0.012, 0.01), "cylinder": (0.0135, 0.015),
            "box": (0.02, 0.03, 0.015)}[kind]
    Rm = R.from_quat(q[[4, 5, 6, 3]]).as_matrix()
    nl = Rm.T @ n
    if kind.startswith("sphere"):
        h = size[0]
    elif kind.startswith("capsule"):
        h = size[0] + size[1] * abs(nl[2])
    elif kind == "cylinder":
        h = size[0] * np.hypot(nl[0], nl[1]) + size[1] * abs(nl[2])
    else:
        h = float(np.abs(nl) @ np.array(size))
    return h + float(q[:3] @ n)


def _overlap(a, qa, b, qb, n):
    """extent of the intersection of the two geoms' projections on n (n from a to b); the penetration depth is its minimum over n"""
    return _support_height(a, qa, n) + _support_height(b, qb, -n)


@pytest.mark.parametrize("a,b", [("sphere", "sphere2"), ("sphere", "capsule"), ("capsule", "sphere"), ("sphere", "cylinder"),
                                 ("sphere", "box"), ("box", "sphere"), ("capsule", "capsule2"), ("capsule", "box"), ("box", "capsule")])
def test_closed_forms_are_the_minimum_penetration(a, b):
    """For two overlapping convex sets the penetration depth is min over unit n of h_a(n) + h_b(-n).  On random ~1 mm penetrations the
    closed form's (normal, dist) must (1) satisfy dist = -(h_a(n) + h_b(-n)), (2) be a minimum: no direction of a dense sample of the
    sphere and no small perturbation of n gives a smaller overlap, and (3) place pos midway between the two support planes."""
    s = _sim(a, b)
    rng = np.random.default_rng(abs(hash((a, b))) % 2**31)
    dirs = rng.normal(size=(500, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    checked = 0
    for trial in range(30):
        qa = np.r_[rng.normal(size=3) * 0.01, _rand_quat(rng)]
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        quat_b = _rand_quat(rng)
        lo, hi = 0.0, GEOMS[a][1] + GEOMS[b][1] + 0.01  # bisection on the centre distance for a target depth
        target = -1e-3
        for _ in range(50):
            mid = 0.5 * (lo + hi)
            c = _contacts(s, qa, np.r_[qa[:3] + d * mid, quat_b], True)
            dist = c[:, 12].min() if len(c) else 1.0
            if dist < target: lo = mid
            else: hi = mid
        qb = np.r_[qa[:3] + d * lo, quat_b]
        ca = _contacts(s, qa, qb, True)
        if not len(ca) or ca[:, 12].min() > 0.5 * target or ca[:, 12].min() < 2 * target:
            continue  # the bisection did not land on a clean shallow contact (deep start): skip
        assert len(ca) <= 2
        k = int(np.argmin(ca[:, 12]))
        n, dist, pos = ca[k, 3:6], ca[k, 12], ca[k, 0:3]
        assert abs(np.linalg.norm(n) - 1) < 1e-12
        ov = _overlap(a, qa, b, qb, n)
        assert abs(ov + dist) < 1e-12, (trial, ov, dist)
        sampled = min(_overlap(a, qa, b, qb, v) for v in dirs)
        assert sampled >= ov - 1e-12, (trial, sampled, ov)
        for _ in range(20):
            v = n + rng.normal(size=3) * 1e-3; v /= np.linalg.norm(v)
            assert _overlap(a, qa, b, qb, v) >= ov - 1e-12
        # midpoint: pos . n lies halfway between a's far plane h_a(n) and b's near plane -h_b(-n)
        assert abs(pos @ n - 0.5 * (_support_height(a, qa, n) - _support_height(b, qb, -n))) < 1e-12
        checked += 1
    assert checked >= 18


def test_mpr_overestimates_at_capsule_ends():
    """Why the closed forms matter: MPR returns the depth along the surface normal where the centre-to-centre ray leaves the Minkowski
    difference.  Off a capsule's end cap that is not the closest direction, and the depth comes out 15-20 % too large."""
    s = _sim("sphere", "capsule")
    ident = np.array([1.0, 0, 0, 0])
    # sphere next to the end cap of the capsule (axis z, half length 0.02, radius 0.009), 1 mm into it along the diagonal
    c0 = np.array([0, 0, 0.02])
    n = np.array([np.sin(1.0), 0, np.cos(1.0)])
    qa = np.r_[c0 + n * (0.009 + 0.012 - 0.001), ident]
    ca, cm = _contacts(s, qa, np.r_[0, 0, 0, ident], True), _contacts(s, qa, np.r_[0, 0, 0, ident], False)
    assert np.allclose(ca[0, 12], -0.001) and np.allclose(ca[0, 3:6], -n)
    assert cm[0, 12] < -0.00105 and cm[0, 3:6] @ (-n) < 0.999


def test_known_answers():
    ident = [1.0, 0, 0, 0]
    # sphere r=0.012 and sphere r=0.02, centres 0.03 apart on x: dist -0.002, normal +x, pos 0.012 - 0.001
    s = _sim("sphere", "sphere2")
    c = _contacts(s, np.r_[0, 0, 0, ident], np.r_[0.03, 0, 0, ident], True)
    assert len(c) == 1 and np.allclose(c[0, 12], -0.002) and np.allclose(c[0, 3:6], [1, 0, 0]) and np.allclose(c[0, 0:3], [0.011, 0, 0])
    # sphere above the middle of a capsule lying along z: closest axis point is the centre
    s = _sim("sphere", "capsule")
    c = _contacts(s, np.r_[0.02, 0, 0.005, ident], np.r_[0, 0, 0, ident], True)
    assert len(c) == 1 and np.allclose(c[0, 12], 0.02 - 0.021) and np.allclose(c[0, 3:6], [-1, 0, 0])
    # ... and beyond its end: the cap sphere at z = +0.02
    c = _contacts(s, np.r_[0, 0, 0.04, ident], np.r_[0, 0, 0, ident], True)
    assert len(c) == 1 and np.allclose(c[0, 12], 0.02 - 0.021) and np.allclose(c[0, 3:6], [0, 0, -1])
    # sphere on the flat cap of a cylinder (half height 0.015), inside the rim radius
    s = _sim("sphere", "cylinder")
    c = _contacts(s, np.r_[0.005, 0, 0.026, ident], np.r_[0, 0, 0, ident], True)
    assert len(c) == 1 and np.allclose(c[0, 12], 0.026 - 0.015 - 0.012) and np.allclose(c[0, 3:6], [0, 0, -1])
    # ... on its side
    c = _contacts(s, np.r_[0.025, 0, 0.0, ident], np.r_[0, 0, 0, ident], True)
    assert len(c) == 1 and np.allclose(c[0, 12], 0.025 - 0.0135 - 0.012) and np.allclose(c[0, 3:6], [-1, 0, 0])
    # ... on its rim: closest point (0.0135, 0, 0.015)
    p = np.array([0.0135 + 0.006, 0, 0.015 + 0.008])
    c = _contacts(s, np.r_[p, ident], np.r_[0, 0, 0, ident], True)
    assert len(c) == 1 and np.allclose(c[0, 12], 0.01 - 0.012) and np.allclose(c[0, 3:6], [-0.6, 0, -0.8])
    # sphere against a box face, an edge, and with its centre inside the box
    s = _sim("sphere", "box")
    c = _contacts(s, np.r_[0.03, 0.01, 0.0, ident], np.r_[0, 0, 0, ident], True)
    assert len(c) == 1 and np.allclose(c[0, 12], 0.01 - 0.012) and np.allclose(c[0, 3:6], [-1, 0, 0])
    c = _contacts(s, np.r_[0.026, 0.038, 0.0, ident], np.r_[0, 0, 0, ident], True)
    assert len(c) == 1 and np.allclose(c[0, 12], 0.01 - 0.012) and np.allclose(c[0, 3:6], [-0.6, -0.8, 0])
    c = _contacts(s, np.r_[0.0, 0.0, 0.013, ident], np.r_[0, 0, 0, ident], True)
    assert len(c) == 1 and np.allclose(c[0, 12], -(0.012 + 0.002)) and np.allclose(c[0, 3:6], [0, 0, -1])
    # capsule (r 0.009, half length 0.02) lying flat on the box's top face (z = 0.015): two points, under the two ends
    s = _sim("capsule", "box")
    qcap = np.r_[0, 0, 0.015 + 0.008, R.from_euler("y", np.pi / 2).as_quat()[[3, 0, 1, 2]]]  # axis along x
    c = _contacts(s, qcap, np.r_[0, 0, 0, ident], True)
    assert len(c) == 2 and np.allclose(c[:, 12], -0.001) and np.allclose(c[:, 3:6], [[0, 0, -1]] * 2)
    assert np.allclose(sorted(c[:, 0]), [-0.02, 0.02]) and np.allclose(c[:, 2], 0.015 - 0.0005)
    # the same capsule standing on the face: one point
    c = _contacts(s, np.r_[0, 0, 0.015 + 0.02 + 0.008, ident], np.r_[0, 0, 0, ident], True)
    assert len(c) == 1 and np.allclose(c[0, 12], -0.001) and np.allclose(c[0, 3:6], [0, 0, -1])
    # capsule lying ACROSS a box edge (axis along y over the edge x = 0.02, z = 0.015 would be parallel; take it along x, overhanging):
    # the end over the face touches, the overhanging end does not -> the contact under the segment's closest point only
    qcap = np.r_[0.03, 0, 0.015 + 0.008, R.from_euler("y", np.pi / 2).as_quat()[[3, 0, 1, 2]]]
    c = _contacts(s, qcap, np.r_[0, 0, 0, ident], True)
    assert 1 <= len(c) <= 2 and np.allclose(c[:, 12].min(), -0.001)
    # parallel capsules side by side: two points; crossed capsules: one
    s = _sim("capsule", "capsule2")
    c = _contacts(s, np.r_[0, 0, 0, ident], np.r_[0.02, 0, 0.0, ident], True)
    assert len(c) == 2 and np.allclose(c[:, 12], 0.02 - 0.021) and np.allclose(c[:, 3:6], [[1, 0, 0]] * 2)
    c = _contacts(s, np.r_[0, 0, 0, ident], np.r_[0.02, 0, 0.0, R.from_euler("x", np.pi / 2).as_quat()[[3, 0, 1, 2]]], True)
    assert len(c) == 1 and np.allclose(c[0, 12], 0.02 - 0.021) and np.allclose(c[0, 0:3], [0.009 - 0.0005, 0, 0])


def test_cylinder_cap_is_a_face_and_its_side_is_not():
    """cylinder-box is a ccd pair in MuJoCo's table (extra points from the multiccd perturbation there): a cylinder standing on a box
    face gets the clipped-face manifold through its cap (the octagon inscribed in the rim), one lying on its side keeps the single
    MPR point."""
    s = _sim("cylinder", "box")
    ident = np.array([1.0, 0, 0, 0])
    c = _contacts(s, np.r_[0.002, 0.001, 0.0295, ident], np.r_[0, 0, 0, ident], True)  # standing, 0.5 mm into the top face
    assert len(c) == 4 and np.allclose(c[:, 12], -0.0005) and np.allclose(c[:, 3:6], [[0, 0, -1]] * 4)
    assert np.all(np.hypot(c[:, 0] - 0.002, c[:, 1] - 0.001) <= 0.0135 + 1e-12) and np.allclose(c[:, 2], 0.015 - 0.00025)
    assert np.linalg.norm(c[:, 0:2].mean(axis=0) - [0.002, 0.001]) < 1e-9  # symmetric about the axis
    assert len(_contacts(s, np.r_[0.002, 0.001, 0.0295, ident], np.r_[0, 0, 0, ident], False)) == 1
    # overhanging the face edge x = 0.02: the cap octagon is clipped by the face
    c = _contacts(s, np.r_[0.018, 0.0, 0.0295, ident], np.r_[0, 0, 0, ident], True)
    assert 3 <= len(c) <= 4 and c[:, 0].max() <= 0.02 + 1e-12 and np.allclose(c[:, 12], -0.0005)
    # tilted by 3 degrees (align 0.9986 < 0.999): an edge-like contact, single point either way
    tilt = R.from_euler("y", np.deg2rad(3)).as_quat()[[3, 0, 1, 2]]
    qa = np.r_[0, 0, 0.0300, tilt]
    ca, cm = _contacts(s, qa, np.r_[0, 0, 0, ident], True), _contacts(s, qa, np.r_[0, 0, 0, ident], False)
    assert len(ca) == 1 and np.array_equal(ca, cm)
    # lying on its side
    side = R.from_euler("y", np.pi / 2).as_quat()[[3, 0, 1, 2]]
    qa = np.r_[0, 0, 0.015 + 0.0135 - 0.0005, side]
    ca, cm = _contacts(s, qa, np.r_[0, 0, 0, ident], True), _contacts(s, qa, np.r_[0, 0, 0, ident], False)
    assert len(ca) == 1 and np.array_equal(ca, cm)


KERNEL_PAIRS = [("sphere", "sphere2"), ("capsule", "sphere"), ("sphere", "cylinder"), ("box", "sphere"), ("capsule", "capsule2"),
                ("capsule", "box"), ("box", "capsule"), ("cylinder", "box"), ("box", "cylinder")]


def _shallow_configs(a, b, s):
    """qpos of random ~1 mm penetrations of the pair (found by bisection on the oracle's closed forms); the first one of capsule-box
    is the two-point case (capsule flat on the top face, tilted by 1 mrad), of cylinder-box the cap manifold (cylinder standing on /
    under a box face).  cylinder-box: only that one - its other contacts are MPR answers, where the kernel's warm-started hill
    climbing and the oracle's cold exhaustive search agree to mpr_tolerance, not to rounding (tests/test_lane1_vs_oracle.py)."""
    rng = np.random.default_rng(abs(hash((b, a))) % 2**31)
    out = []
    for trial in range(1 if {a, b} == {"cylinder", "box"} else 12):
        qa = np.r_[rng.normal(size=3) * 0.01, _rand_quat(rng)]
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        quat_b = _rand_quat(rng)
        if trial == 0 and {a, b} == {"cylinder", "box"}:
            qa = np.r_[0.002, 0.001, 0, 1.0, 0, 0, 0]; d = np.array([0, 0, -1.0 if a == "cylinder" else 1.0]); quat_b = np.array([1.0, 0, 0, 0])
        if trial == 0 and (a, b) == ("capsule", "box"):
            qa = np.r_[0, 0, 0, R.from_euler("y", np.pi / 2 + 1e-3).as_quat()[[3, 0, 1, 2]]]; d = np.array([0, 0, -1.0]); quat_b = np.array([1.0, 0, 0, 0])
        lo, hi = 0.0, GEOMS[a][1] + GEOMS[b][1] + 0.01
        for _ in range(50):
            mid = 0.5 * (lo + hi)
            c = _contacts(s, qa, np.r_[qa[:3] + d * mid, quat_b], True)
            if (c[:, 12].min() if len(c) else 1.0) < -1e-3: lo = mid
            else: hi = mid
        qpos = np.r_[qa, qa[:3] + d * lo, quat_b]
        c = _contacts(s, qpos[:7], qpos[7:], True)
        if not len(c) or c[:, 12].min() < -2e-3:
            continue
        if trial == 0 and (a, b) == ("capsule", "box"):
            assert len(c) == 2
        if trial == 0 and {a, b} == {"cylinder", "box"}:
            assert len(c) == 4
        out.append(qpos)
    assert len(out) >= (1 if {a, b} == {"cylinder", "box"} else 6)
    return out


@pytest.mark.parametrize("a,b", KERNEL_PAIRS)
def test_kernel_source_matches_the_oracle_on_primitive_pairs(a, b):
    """The kernel's own copy of the closed forms (csrc/mgs_collide.cuh prim_pair, here through the fp64 1-lane host build of the
    kernel source) against the oracle: random ~1 mm penetrations per pair, 5 steps each - contact counts equal at every step and
    the states (which feel every contact's position, normal and depth through the solver) equal to 1e-9."""
    from hostsim import lane1
    m = compile_mjcf(PAIR.format(a=GEOMS[a][0], b=GEOMS[b][0]))
    s, k = OracleSim(m), lane1.sim(m, f64=True)
    for qpos in _shallow_configs(a, b, s):
        s.reset(); s.set_analytic(True)
        s.qpos[:] = qpos
        st = k.pack_state(qpos[None], np.zeros((1, 12)))
        for step in range(5):
            s.step(1)
            st, dg = k.step(st, 1, want_diag=True)
            assert int(dg["ncon"][0]) == s.ncon, step
            u = k.unpack_state(st); q2, v2 = u["qpos"], u["qvel"]
            assert np.abs(q2[0] - s.qpos).max() < 1e-9 and np.abs(v2[0] - s.qvel).max() < 1e-7, step


@pytest.mark.gpu
@pytest.mark.parametrize("a,b", KERNEL_PAIRS)
def test_cuda_builds_match_the_oracle_on_primitive_pairs(a, b):
    """The same configurations through the CUDA library (one batch per pair type, every configuration its own environment): the
    fp64 build to 1e-9 like the host build of the source, the fp32 build to 1e-6 in qpos / 1e-3 in qvel after five steps of contact
    response (the 1-lane fp32 host build of the same source measures 2e-7 / 7.5e-5)."""
    from mj_grasp_sim_b200 import lib as mlib
    m = compile_mjcf(PAIR.format(a=GEOMS[a][0], b=GEOMS[b][0]))
    s = OracleSim(m)
    qs = np.array(_shallow_configs(a, b, s))
    ref_q, ref_v, ref_n = [], [], []
    for qpos in qs:
        s.reset(); s.set_analytic(True)
        s.qpos[:] = qpos
        ns = []
        for step in range(5):
            s.step(1); ns.append(s.ncon)
        ref_q.append(np.array(s.qpos)); ref_v.append(np.array(s.qvel)); ref_n.append(ns)
    ref_q, ref_v, ref_n = np.array(ref_q), np.array(ref_v), np.array(ref_n)
    for f64, tol in ((True, 1e-9), (False, 1e-6)):
        if f64 and not os.path.exists(mlib.SO_PATH_F64):
            pytest.fail("the fp64 build is missing")
        G = mlib.BatchSim(m, f64=f64)
        st = G.pack_state(qs, np.zeros((len(qs), 12)))
        for step in range(5):
            st, dg = G.step(st, 1, want_diag=True)
            if f64:
                assert np.array_equal(dg["ncon"], ref_n[:, step]), (step, dg["ncon"], ref_n[:, step])
        u = G.unpack_state(st)
        assert np.abs(u["qpos"] - ref_q).max() < tol and np.abs(u["qvel"] - ref_v).max() < 1000 * tol, (f64, np.abs(u["qpos"] - ref_q).max(), np.abs(u["qvel"] - ref_v).max())
        G.close()


def test_allegro_fingertip_capsules_on_a_cube_kernel_source_vs_oracle():
    """The closed forms on a real hand: the Allegro hand's four fingertip capsules closing on the box primitive (capsule-box pairs
    against the grasped object, the ground box and the hand's own link boxes).  Kernel source (fp64 1-lane build) vs the oracle: same
    contact count at every one of the first 300 steps of the close phase and qpos within 1e-7, capsule-box contacts with the object
    among them; then the labels of a short rollout."""
    from hostsim import lane1
    from mj_grasp_sim_b200 import scenes
    from mj_grasp_sim_b200.lib import MgsRolloutCfg
    from oracle.oracle import RolloutCfg, batch
    m, info, pose7, joints = scenes.workload("allegro", "cube", 0, 6)
    s, L = OracleSim(m), lane1.sim(m, f64=True)
    ty, p1, p2, cg = m.arr["cgeom_type"], m.arr["pair_geom1"], m.arr["pair_geom2"], m.arr["cgeom_geomid"]
    ground = m.names["geom"]["geom:ground"]
    pairmap = {(int(cg[p1[i]]), int(cg[p2[i]])): i for i in range(len(p1))}
    cap_on_object = 0
    for i in (0, 1):
        s.reset()
        s.place(pose7[i].astype(np.float64), info["base_qposadr"], joints[i].astype(np.float64), info["joint_qposadr"])
        s.ctrl[:] = info["close_ctrl"]
        st = L.pack_state(s.qpos.copy(), s.qvel.copy(), ctrl=s.ctrl.copy(), mocap_pos=s.mocap_pos[0].copy(), mocap_quat=s.mocap_quat[0].copy())
        for step in range(300):
            s.step(1)
            st, d = L.step(st, 1, want_diag=True)
            assert int(d["ncon"][0]) == s.ncon, (i, step)
            assert np.abs(L.unpack_state(st)["qpos"][0] - s.qpos).max() < 1e-7, (i, step)
            if step % 10 == 0:
                for c in s.contacts():
                    p = pairmap[(int(c[13]), int(c[14]))]
                    if {int(ty[p1[p]]), int(ty[p2[p]])} == {3, 6} and min(c[13], c[14]) < ground < max(c[13], c[14]):
                        cap_on_object += 1
    assert cap_on_object > 0
    sched = (600, 200, 30, 1, 0.03, 0.02)
    lab, steps = L.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    olab, osteps = batch(m, 1, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"],
                         info["close_ctrl"], RolloutCfg(*sched), 4)
    assert np.array_equal(lab.astype(bool), olab.astype(bool)) and np.array_equal(steps, osteps)
