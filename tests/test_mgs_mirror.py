"""The `mgs` drop-in mirror: SE3Pose semantics, selectors, model equivalence, argument checks (CPU)."""
import numpy as np
import pytest

from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.mgs.env.gravityless_object_grasping import GravitylessObjectGrasping
from mj_grasp_sim_b200.mgs.gripper.selector import get_gripper
from mj_grasp_sim_b200.mgs.obj.selector import get_object
from mj_grasp_sim_b200.mgs.util.geo.transforms import SE3Pose


def test_se3pose_semantics():
    p = SE3Pose(np.array([0.1, 0.2, 0.3]), np.array([1.0, 0, 0, 0]), "wxyz")
    assert p.pos.dtype == np.float32 and p.quat.dtype == np.float32
    assert p.to_mat().dtype == np.float32
    with pytest.raises(AssertionError):
        SE3Pose(np.zeros(3), np.array([2.0, 0, 0, 0]), "wxyz")  # not unit norm
    H = np.eye(4)[None].repeat(3, 0)
    H[:, :3, 3] = np.arange(9).reshape(3, 3) * 0.01
    q = SE3Pose.from_mat(H)
    assert len(q) == 3 and np.allclose(q[1].pos, H[1, :3, 3])
    b2c = get_gripper("PandaGripper").base_to_contact_transform()
    r = q @ b2c
    assert np.allclose(r.pos, H[:, :3, 3] + [0, 0, -0.102], atol=1e-7)
    assert np.allclose(np.abs(r.quat), [[0.70710677, 0, 0, 0.70710677]] * 3, atol=1e-6)
    a = SE3Pose(np.array([0.1, 0.0, 0.0]), np.array([0.70710678, 0, 0, 0.70710678]), "wxyz")
    inv = a.inverse()
    assert np.allclose(a.pos, inv.pos) and np.allclose(a.quat, inv.quat)  # reference quirk: inverse() mutates self
    assert np.allclose(inv.pos, [0, 0.1, 0], atol=1e-6)


def test_env_matches_scene_builder_and_checks_arguments():
    env = GravitylessObjectGrasping(get_gripper("PandaGripper"), get_object("hull:0"))
    m2, info, pose7, joints = scenes.workload("panda", "hull", 0, 8)
    for k in ("body_mass", "body_inertia", "hull_vert", "pair_friction", "qpos0", "eq_data"):
        assert np.allclose(env.model.arr[k], m2.arr[k]), k
    assert env.get_joint_idxs(env.gripper.get_actuator_joint_names()) == list(info["joint_qposadr"])
    assert env.gripper.get_freejoint_idxs(env) == list(range(7))
    assert env.get_joint_idxs(["no_such_joint"]) == [int(env.model.jnt_qposadr[-1])]  # reference quirk (simualtion.py:37-43)
    v, t = env.obj.mesh()
    H, w = scenes.antipodal_candidates(v, t, 8, 0)
    p7, j32, jadr = env._process(SE3Pose.from_mat(H), scenes.panda_width_to_joints(w))
    assert np.array_equal(p7, pose7) and np.allclose(j32, joints)
    with pytest.raises(ValueError):
        env._process(SE3Pose.from_mat(H), joints[:3])
    with pytest.raises(ValueError):
        env._process(SE3Pose.from_mat(H), np.zeros((8, 5)))


def test_selectors_reject_unknown():
    with pytest.raises(ValueError):
        get_gripper("NoSuchGripper")
    with pytest.raises(ValueError):
        get_object("003_cracker_box")  # YCB assets are not shipped offline
    cube = get_object("cube")
    xml, assets = cube.to_xml()
    assert 'name="geom:cube"' in xml and assets == {}


def test_robotiq_and_vx300_mirror_classes():
    from mj_grasp_sim_b200 import scenes
    for name, key in (("Robotiq2f85Gripper", "robotiq2f85"), ("VXGripper", "vx300"), ("AllegroGripper", "allegro"), ("LeapGripper", "leap"), ("ShadowHand", "shadow")):
        g = get_gripper(name)
        env = GravitylessObjectGrasping(g, get_object("hull:0"))
        m2, info, _, _ = scenes.workload(key, "hull", 0, 2)
        assert np.allclose(env.model.body_mass, m2.body_mass)
        assert env.get_joint_idxs(g.get_actuator_joint_names()) == list(info["joint_qposadr"])  # incl. the misnamed Robotiq joints
        b2c = g.base_to_contact_transform()
        assert np.allclose(b2c.pos, scenes.GRIPPERS[key]["b2c_pos"]) and np.allclose(b2c.quat, scenes.GRIPPERS[key]["b2c_quat"], atol=1e-6)
        assert np.allclose(g.close_ctrl(), scenes.GRIPPERS[key]["close_ctrl"])


def test_cli_pipeline_end_to_end_on_the_host_build(tmp_path, monkeypatch):
    """gen_grasp_candidates -> candidates.npz -> filter_to_stable -> candidates_collision_free.npz / stable_grasps.npz, the
    reference's single-object flow (README 'gen_grasps' stages), with the env's simulator routed to the 1-lane HOST build of
    the kernel source (test harness: the product path needs a CUDA device) and the labels checked against the oracle."""
    import ctypes as C
    import os
    from hostsim import lane1
    from mj_grasp_sim_b200 import lib as mlib
    from mj_grasp_sim_b200.mgs.cli import filter_to_stable, gen_grasp_candidates
    from mj_grasp_sim_b200.mgs.env import gravityless_object_grasping as gog
    from oracle.oracle import RolloutCfg, batch
    L = mlib.bind(C.CDLL(lane1.build(False)), prefix="l1_")
    from mj_grasp_sim_b200.mgs.core import simualtion as simmod
    host = lambda model, device=0, ncon_max=0, nefc_max=0, ground_name="geom:ground", f64=False: mlib.BatchSim(model, lib=L, prefix="l1_", ground_name=ground_name)
    monkeypatch.setattr(simmod, "BatchSim", host)
    monkeypatch.setattr(gog, "BatchSim", host)
    H, joints = gen_grasp_candidates.run("PandaGripper", "hull:0", 6, str(tmp_path), seed=4)
    free, stable = filter_to_stable.run("PandaGripper", "hull:0", str(tmp_path))
    d = tmp_path / "PandaGripper" / "hull:0"
    cf, st = np.load(d / "candidates_collision_free.npz"), np.load(d / "stable_grasps.npz")
    assert cf["pose"].shape == (int(free.sum()), 4, 4) and st["pose"].shape == (int(stable.sum()), 4, 4) and st["joints"].shape[1] == 2
    # oracle on the same candidates
    env = gog.GravitylessObjectGrasping(get_gripper("PandaGripper"), get_object("hull:0"))
    pose7, j32, jadr = env._process(SE3Pose.from_mat(H, type="wxyz"), joints)
    base = env.gripper.get_freejoint_idxs(env)[0]
    sched = RolloutCfg(3000, 3000, 500, 0, 0.1, 0.02)
    ofree, _ = batch(env.model, 0, pose7.astype(np.float64), base, j32.astype(np.float64), jadr, env.gripper.close_ctrl(), sched, os.cpu_count() or 1)
    assert np.array_equal(free, ofree)
    olab, _ = batch(env.model, 1, pose7[free].astype(np.float64), base, j32[free].astype(np.float64), jadr, env.gripper.close_ctrl(), sched,
                    os.cpu_count() or 1)
    assert (stable == olab).mean() >= 0.8


def test_cli_main_accepts_hydra_style_configs(monkeypatch):
    """`main(cfg)` of every CLI mirror maps the reference's config fields (cfg.gripper.name, cfg.id, ...) to `run`."""
    from types import SimpleNamespace as NS
    from mj_grasp_sim_b200.mgs.cli import (_common, eval_grasps, filter_collision_free_candidates, filter_stable_grasps, filter_to_stable,
                                           gen_grasp_candidates)
    cfg = NS(gripper=NS(name="VXGripper"), id=3, env=NS(name="clutter_table"), num_grasps=17)
    assert _common.object_id_from_cfg(cfg) == "hull:2" and _common.object_id_from_cfg(NS(id=0)) == "cube"
    assert _common.object_id_from_cfg({"object": "hull:9", "id": 1}) == "hull:9"
    seen = {}
    for mod in (filter_to_stable, filter_stable_grasps, filter_collision_free_candidates, gen_grasp_candidates, eval_grasps):
        monkeypatch.setattr(mod, "run", lambda *a, _m=mod.__name__: seen.setdefault(_m.rsplit(".", 1)[1], a))
        mod.main(cfg)
    assert seen["filter_to_stable"][:2] == ("VXGripper", "hull:2") and seen["gen_grasp_candidates"][:3] == ("VXGripper", "hull:2", 17)
    assert seen["eval_grasps"][:3] == ("VXGripper", 3, "clutter_table")
