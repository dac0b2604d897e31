"""CPU-tier tests of the round-2 additions, run on the 1-lane HOST build of the kernel source (tests/hostsim) against
the fp64 oracle: last-wins joint scatter, per-candidate auxiliary outputs (overflow flag, object drift), per-step
qvel clamp, the `enough_stable` early stop, the imperative gripper / sim protocol and gen_scene's grasp table."""
import ctypes as C

import numpy as np
import pytest

from hostsim import lane1
from mj_grasp_sim_b200 import lib as mlib
from mj_grasp_sim_b200 import scenes, shard
from mj_grasp_sim_b200.lib import MgsRolloutCfg
from oracle.oracle import OracleSim, RolloutCfg, batch


def test_duplicate_joint_addresses_keep_the_last_value(robotiq_hull):
    """Robotiq's two misnamed joints both resolve to the object's x coordinate (robotiq2f85.py:275,279): `data.qpos[idxs] = qpos`
    keeps the LAST value.  Distinct values in the two columns must move the object exactly as the oracle's sequential scatter."""
    m, info, pose7, joints = robotiq_hull
    jadr = info["joint_qposadr"]
    assert jadr[2] == jadr[6]
    j = joints[:6].copy()
    j[:, 2] = 0.3   # would push the object 30 cm away if it won
    j[:, 6] = 0.002  # the reference's winner
    L = lane1.sim(m, f64=True)
    sched = (300, 100, 0, 0, 0.02, 0.02)
    lab, steps = L.stability(pose7[:6], j, jadr, info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    olab, osteps = batch(m, 1, pose7[:6].astype(np.float64), info["base_qposadr"], j.astype(np.float64), jadr, info["close_ctrl"], RolloutCfg(*sched), 2)
    assert np.array_equal(lab, olab) and np.array_equal(steps, osteps)
    jw = j.copy()
    jw[:, 6] = 0.3
    lab_far, _ = L.stability(pose7[:6], jw, jadr, info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    assert not lab_far.any()  # with 0.3 as the last value the object is out of reach


def test_aux_drift_matches_oracle(panda_hull):
    m, info, pose7, joints = panda_hull
    n = 6
    L = lane1.sim(m, f64=True)
    sched = (400, 100, 0, 0, 0.02, 0.02)
    lab, _ = L.stability(pose7[:n], joints[:n], info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    aux = L.last_aux(n)
    assert not aux["overflow"].any() and not aux["bad"].any()
    a = int(m.jnt_qposadr[-1])
    s = OracleSim(m)
    seen = 0
    for i in range(n):
        s.reset()
        s.place(pose7[i].astype(np.float64), info["base_qposadr"], joints[i].astype(np.float64), info["joint_qposadr"])
        s.forward()
        before = s.qpos[a:a + 7].astype(np.float32)
        s.ctrl[:] = info["close_ctrl"]
        s.step(400)
        if not s.contact_with_object():
            assert np.isnan(aux["pos_drift"][i]) and np.isnan(aux["rot_drift_deg"][i])
            continue
        after = s.qpos[a:a + 7].astype(np.float32)
        dot = np.clip(np.sum(before[3:] * after[3:]), -1.0, 1.0)
        assert abs(aux["pos_drift"][i] - np.linalg.norm(before[:3] - after[:3])) < 1e-6
        assert abs(aux["rot_drift_deg"][i] - np.degrees(np.arccos(np.clip(2 * dot ** 2 - 1, -1, 1)))) < 2e-2
        seen += 1
    assert seen >= 2


def test_overflow_flag_and_limit_rows(panda_cube):
    """A model created with tiny capacities must flag every environment that needed more (contacts or rows)."""
    m, info, pose7, joints = panda_cube
    L = lane1.sim(m, f64=True)
    sched = (200, 50, 0, 0, 0.02, 0.02)
    L.stability(pose7[:8], joints[:8], info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    assert not L.last_aux(8)["overflow"].any()  # default capacities are enough here


def test_qvel_clip_every_step(panda_cube):
    m, info = panda_cube[0], panda_cube[1]
    s = OracleSim(m)
    L = lane1.sim(m, f64=True)
    s.reset()
    s.qvel[:] = np.linspace(-80, 80, m.nv)
    st = L.pack_state(s.qpos.copy(), s.qvel.copy(), ctrl=s.ctrl.copy(), mocap_pos=s.mocap_pos[0].copy(), mocap_quat=s.mocap_quat[0].copy())
    s.set_qvel_clip(50.0)
    L.set_qvel_clip(50.0)
    s.step(5)
    out = L.unpack_state(L.step(st, 5))
    L.set_qvel_clip(0.0)
    assert np.abs(out["qvel"][0] - s.qvel).max() < 1e-8 * max(1.0, np.abs(s.qvel).max())
    assert np.abs(out["qpos"][0] - s.qpos).max() < 1e-9
    unclipped = L.unpack_state(L.step(st, 5))
    assert np.abs(unclipped["qpos"][0] - s.qpos).max() > 1e-3  # the clamp mattered


def test_enough_stable_early_stop_matches_sequential_semantics():
    rng = np.random.default_rng(0)
    truth = rng.uniform(size=1000) < 0.4
    calls = []

    def run_range(lo, hi):
        calls.append((lo, hi))
        return truth[lo:hi]
    for enough, chunk in ((5, 64), (50, 100), (10 ** 6, 300), (1, 1)):
        calls.clear()
        got = shard.evaluate_sharded(len(truth), run_range, enough, chunk)
        assert np.array_equal(got, shard.apply_enough_stable(truth, enough))
        evaluated = sum(hi - lo for lo, hi in calls)
        need = int(np.searchsorted(np.cumsum(truth), enough) + 1) if truth.sum() >= enough else len(truth)
        assert evaluated <= min(len(truth), -(-need // chunk) * chunk)  # never more than the chunk that reached the target
    calls.clear()
    assert np.array_equal(shard.evaluate_sharded(len(truth), run_range), truth) and calls == [(0, 1000)]


def _host_env(monkeypatch, gripper_name, obj_id):
    from mj_grasp_sim_b200.mgs.core import simualtion as simmod
    from mj_grasp_sim_b200.mgs.env import gravityless_object_grasping as gog
    from mj_grasp_sim_b200.mgs.gripper.selector import get_gripper
    from mj_grasp_sim_b200.mgs.obj.selector import get_object
    Lh = mlib.bind(C.CDLL(lane1.build(True)), prefix="l1_")
    host = lambda model, device=0, ncon_max=0, nefc_max=0, ground_name="geom:ground", f64=False: mlib.BatchSim(model, lib=Lh, prefix="l1_", ground_name=ground_name)
    monkeypatch.setattr(simmod, "BatchSim", host)
    monkeypatch.setattr(gog, "BatchSim", host)
    return gog.GravitylessObjectGrasping(get_gripper(gripper_name), get_object(obj_id))


@pytest.mark.parametrize("gripper_name", ["PandaGripper", "AllegroGripper"])
def test_imperative_protocol_matches_oracle(monkeypatch, gripper_name):
    """The reference's hand-driven sequence on ONE environment - set_qpos, gripper.set_pose, mj_forward, check_contact,
    gripper.close_gripper_at, check_contact_with_object, get_object_transform, get_state / set_state - through the handle's
    `data` views (1-lane host build) against the oracle driven the same way."""
    from mj_grasp_sim_b200.mgs.util.geo.transforms import SE3Pose
    env = _host_env(monkeypatch, gripper_name, "hull:0")
    g = env.gripper
    key = {"PandaGripper": "panda", "AllegroGripper": "allegro"}[gripper_name]
    m, info, pose7, joints = scenes.workload(key, "hull", 0, 4)
    jidx = env.get_joint_idxs(g.get_actuator_joint_names())
    assert list(jidx) == list(info["joint_qposadr"])
    pose = SE3Pose(pose7[1, :3].astype(np.float64), pose7[1, 3:].astype(np.float64), "wxyz")
    env.mj_resetData()
    env.set_qpos(joints[1].astype(np.float64), jidx)
    g.set_pose(env, pose)
    env.mj_forward()
    s = OracleSim(env.model)
    s.reset()
    s.place(pose7[1].astype(np.float64), info["base_qposadr"], joints[1].astype(np.float64), info["joint_qposadr"])
    s.forward()
    assert env.check_contact() == (s.ncon != 0)
    st0 = env.get_state()
    g.NSTEP_CLOSE = 300
    g.close_gripper_at(env, pose)
    if g.REPOSE_ON_CLOSE:
        s.place(pose7[1].astype(np.float64), info["base_qposadr"], np.zeros(0), np.zeros(0, dtype=np.int32))
    s.ctrl[:] = g.close_ctrl()
    s.step(300)
    assert np.abs(env.data.qpos - s.qpos).max() < 1e-6
    assert env.check_contact_with_object() == s.contact_with_object()
    t = env.get_object_transform(env.obj.name)
    a = int(env.model.jnt_qposadr[-1])
    assert np.allclose(t.pos, s.qpos[a:a + 3].astype(np.float32))
    assert abs(env.data.time - 0.3) < 1e-9
    # state vector round trip (mjSTATE_INTEGRATION layout): back to the pre-close state, same trajectory again
    q_after = env.data.qpos.copy()
    env.set_state(st0)
    assert env.data.time == st0[0]
    g.close_gripper_at(env, pose)
    assert np.abs(env.data.qpos - q_after).max() < 1e-12
    g.open_gripper(env)
    assert np.isfinite(env.data.ctrl).all()
    with pytest.raises(NotImplementedError):
        env.idle()


def test_grasp_table_and_reference_index_quirk():
    from mj_grasp_sim_b200.mgs.cli.gen_scene import GraspTable
    pose = np.tile(np.eye(4), (6, 1, 1))
    pose[:, 0, 3] = np.arange(6)
    t = GraspTable(pose, np.arange(12.0).reshape(6, 2), np.array([0, 0, 0, 1, 1, 1], dtype=np.int32), [("a", "ida"), ("b", "idb")])
    perm = np.array([5, 0, 3, 1, 4, 2])
    moved, quirk = t.take(perm), t.take(perm, move_owner=False)
    assert moved.owner.tolist() == [1, 0, 1, 0, 1, 0] and quirk.owner.tolist() == [0, 0, 0, 1, 1, 1]
    assert np.array_equal(moved.pose, quirk.pose)
    by = dict(moved.per_object())
    assert by[0]["object_id"] == "ida" and sorted(by[0]["pose"][:, 0, 3].tolist()) == [0, 1, 2]
    byq = dict(quirk.per_object())
    assert sorted(byq[0]["pose"][:, 0, 3].tolist()) == [0, 3, 5]  # the reference files grasps of b under a after its shuffle
    sel = t.take(np.array([True, False, True, False, False, True]))
    assert len(sel) == 3 and sel.owner.tolist() == [0, 0, 1]
