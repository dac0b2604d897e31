"""The kernel source, built as a 1-lane host program (tests/hostsim), against the fp64 oracle.

This is the CPU-tier differential test: identical algorithmic spec, two independent
implementations (plain C fp64 oracle vs the warp-cooperative CUDA source).  Tolerances: fp64 build
1e-8 abs; fp32 build 1e-4 relative on qpos/qvel over the first 50 steps (north_star)."""
import numpy as np
import pytest

from hostsim import lane1
from mj_grasp_sim_b200.lib import MgsRolloutCfg
from oracle.oracle import OracleSim, RolloutCfg, batch

P7 = np.array([0, 0, -0.102, 0.70710677, 0, 0, 0.70710677])


def _start(s, info, joints):
    s.reset()
    s.place(P7, info["base_qposadr"], np.array(joints), info["joint_qposadr"])
    s.ctrl[:] = info["close_ctrl"]


@pytest.mark.parametrize("f64,tol", [(True, 1e-8), (False, 2e-4)])
def test_forward_quantities(panda_cube, f64, tol):
    m, info = panda_cube[0], panda_cube[1]
    s = OracleSim(m)
    L = lane1.sim(m, f64=f64)
    _start(s, info, [0.0199, -0.0201])  # pads penetrate the cube: contacts present
    st = L.pack_state(s.qpos.copy(), s.qvel.copy(), ctrl=s.ctrl.copy(), mocap_pos=s.mocap_pos[0].copy(), mocap_quat=s.mocap_quat[0].copy())
    s.forward()
    _, d = L.step(st, 0, want_diag=True)
    assert d["ncon"][0] == s.ncon and d["nefc"][0] == s.nefc and s.ncon > 0
    assert np.abs(d["M"][0] - s.M).max() < tol
    assert np.abs(d["xpos"][0] - s.xpos).max() < tol
    scale = max(1.0, np.abs(s.qacc).max())
    assert np.abs(d["qacc_smooth"][0] - s.qacc_smooth).max() < tol * scale * 10
    assert np.abs(d["qacc"][0] - s.qacc).max() < (1e-6 if f64 else 2e-2) * scale
    con = s.contacts()
    assert np.abs(d["contact"][0, : s.ncon, 0] - con[:, 12]).max() < max(tol * 1e-2, 1e-7)  # penetration depths


@pytest.mark.parametrize("f64,tol", [(True, 1e-8), (False, 1e-4)])
def test_first_50_steps(panda_cube, f64, tol):
    m, info = panda_cube[0], panda_cube[1]
    s = OracleSim(m)
    L = lane1.sim(m, f64=f64)
    _start(s, info, [0.0215, -0.0185])  # fingers reach the cube within the window
    st = L.pack_state(s.qpos.copy(), s.qvel.copy(), ctrl=s.ctrl.copy(), mocap_pos=s.mocap_pos[0].copy(), mocap_quat=s.mocap_quat[0].copy())
    for k in range(5):
        s.step(10)
        st = L.step(st, 10)
        u = L.unpack_state(st)
        assert np.abs(u["qpos"][0] - s.qpos).max() <= tol * max(1.0, np.abs(s.qpos).max())
        assert np.abs(u["qvel"][0] - s.qvel).max() <= tol * max(1.0, np.abs(s.qvel).max()) * 10
    assert s.ncon > 0


def test_rollout_labels_short_schedule(panda_hull):
    m, info, pose7, joints = panda_hull
    n = 12
    sched = (400, 200, 40, 0, 0.03, 0.02)
    L = lane1.sim(m)
    free = L.collision_mask(pose7[:n], joints[:n], info["joint_qposadr"], info["base_qposadr"])
    lab, steps = L.stability(pose7[:n], joints[:n], info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    ofree, _ = batch(m, 0, pose7[:n].astype(np.float64), info["base_qposadr"], joints[:n].astype(np.float64), info["joint_qposadr"],
                     info["close_ctrl"], RolloutCfg(*sched), 4)
    olab, osteps = batch(m, 1, pose7[:n].astype(np.float64), info["base_qposadr"], joints[:n].astype(np.float64), info["joint_qposadr"],
                         info["close_ctrl"], RolloutCfg(*sched), 4)
    assert (free == ofree).all()
    assert (lab == olab).mean() >= 11 / 12
    assert np.array_equal(steps[lab == olab], osteps[lab == olab])


@pytest.mark.parametrize("fixture,qtol", [("robotiq_hull", 2e-4), ("vx300_hull", 2e-4)])
def test_other_grippers_first_steps_and_labels(request, fixture, qtol):
    """Robotiq 2F-85 (4-bar linkage: connect + joint equalities, tendon actuator, 800-vertex hulls) and
    ViperX 300 (condim-4 pads, impratio 10) through the same kernel source."""
    m, info, pose7, joints = request.getfixturevalue(fixture)
    s = OracleSim(m)
    L = lane1.sim(m)
    i = 0
    s.reset()
    s.place(pose7[i].astype(np.float64), info["base_qposadr"], joints[i].astype(np.float64), info["joint_qposadr"])
    s.ctrl[:] = info["close_ctrl"]
    st = L.pack_state(s.qpos.copy(), s.qvel.copy(), ctrl=s.ctrl.copy(), mocap_pos=s.mocap_pos[0].copy(), mocap_quat=s.mocap_quat[0].copy())
    for k in range(5):
        s.step(10)
        st, d = L.step(st, 10, want_diag=True)
        u = L.unpack_state(st)
        assert d["overflow"][0] == 0 and d["bad"][0] == 0
        assert np.abs(u["qpos"][0] - s.qpos).max() <= qtol * max(1.0, np.abs(s.qpos).max())
    n = 6
    sched = (500, 150, 30, 0, 0.02, 0.02)
    lab, steps = L.stability(pose7[:n], joints[:n], info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    olab, osteps = batch(m, 1, pose7[:n].astype(np.float64), info["base_qposadr"], joints[:n].astype(np.float64), info["joint_qposadr"],
                         info["close_ctrl"], RolloutCfg(*sched), 4)
    assert (lab == olab).mean() >= 5 / 6


def test_shadow_hand_short_rollouts(shadow_hull):
    """Shadow hand: nv = 34 > 32 lanes (rows are strided over lanes), 18 actuators of which 4 drive fixed
    tendons, cylinders / capsules / spheres / meshes, impratio 10."""
    m, info, pose7, joints = shadow_hull
    assert (m.nq, m.nv, m.nu, int(m.arr["ntendon"])) == (36, 34, 18, 4)
    n = 6
    sched = (400, 120, 20, 0, 0.02, 0.02)
    L = lane1.sim(m, f64=True)
    lab, steps = L.stability(pose7[:n], joints[:n], info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    olab, osteps = batch(m, 1, pose7[:n].astype(np.float64), info["base_qposadr"], joints[:n].astype(np.float64), info["joint_qposadr"],
                         info["close_ctrl"], RolloutCfg(*sched), 4)
    assert (lab == olab).mean() >= 5 / 6


@pytest.mark.parametrize("fixture", ["allegro_hull", "leap_hull"])
def test_dexterous_hands_short_rollouts(request, fixture):
    """Allegro (capsules + boxes, 16 hinge actuators, geom-derived inertias) and LEAP (67 boxes + 4 tip
    meshes, 22 dry-friction dofs, top-level defaults leaking into the scene): nv = 28, tree depth 5."""
    m, info, pose7, joints = request.getfixturevalue(fixture)
    assert (m.nq, m.nv, m.nu) == (30, 28, 16)
    n = 8
    sched = (400, 150, 30, 1, 0.02, 0.02)
    L = lane1.sim(m, f64=True)
    lab, steps = L.stability(pose7[:n], joints[:n], info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    olab, osteps = batch(m, 1, pose7[:n].astype(np.float64), info["base_qposadr"], joints[:n].astype(np.float64), info["joint_qposadr"],
                         info["close_ctrl"], RolloutCfg(*sched), 4)
    assert (lab == olab).mean() >= 7 / 8


@pytest.mark.parametrize("gripper", ["panda", "robotiq2f85", "vx300", "allegro", "leap", "shadow"])
def test_first_50_steps_fp64_every_gripper(gripper, monkeypatch):
    """Kernel source (fp64 1-lane build) vs the oracle, first 50 steps after the close command, 3 collision-free candidates per
    gripper: two independent implementations of the same specification agree to 1e-7 in qpos on every model feature the six
    grippers use (connect / joint equalities, tendon actuators, dry-friction dofs, capsules, cylinders, 1000-vertex hulls)."""
    import ctypes as C
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import first50 as f50
    from mj_grasp_sim_b200 import lib as mlib
    L = mlib.bind(C.CDLL(lane1.build(True)), prefix="l1_")
    monkeypatch.setattr(f50, "BatchSim", lambda model, f64=False: mlib.BatchSim(model, lib=L, prefix="l1_"))
    r = f50.first50(gripper, 3, True, chunk=1)  # one step per launch: cold collision cache, like the oracle
    # a candidate whose contact set differs from the oracle's at a checkpoint (a manifold decision at its threshold: the kernel's
    # warm-started MPR answers within mpr_tolerance of the oracle's cold-started one) is a contact-onset flip, not drift
    assert r["n"] == 3 and r["n_same_contacts"] >= 2 and r["qpos_rel_same_contacts"] <= 1e-7, r
