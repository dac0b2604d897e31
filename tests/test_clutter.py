"""Clutter table (config 5 path): oracle vs the kernel source (lane-1 build) on the
restore-scene / place / close / lift program, and the mgs.env.ClutterTableEnv host logic."""
import ctypes as C

import numpy as np
import pytest

from hostsim import lane1
from mj_grasp_sim_b200 import lib as mlib
from mj_grasp_sim_b200 import scenes
from oracle.oracle import OracleSim, RolloutCfg, batch


@pytest.fixture(scope="module")
def clutter():
    m, info = scenes.build_clutter_scene("panda", [0, 1, 2])
    s = OracleSim(m, ground_name="geom:table")

    def step_fn(rec, n):
        s.set_record(rec)
        s.step(n)
        assert s.bad == 0
        return s.get_record()

    rec = scenes.gen_clutter(m, info, step_fn, 0)
    H, w = scenes.clutter_candidates(m, info, rec, 12, 0)
    return m, info, rec, scenes.process_poses(H, "panda"), scenes.panda_width_to_joints(w).astype(np.float32)


def test_scene_layout_and_settling(clutter):
    m, info, rec, _, _ = clutter
    g = m.names["geom"]
    # geom order the contact tests rely on: gripper < table < camera < origin < walls < objects
    assert g["panda_col_12"] < g["geom:table"] < g["geom:camera"] < g["geom:base_origin"] < g["geom:wall_top"] < g["geom:wall_left"]
    assert (m.nq, m.nv) == (9 + 7 + 21, 8 + 6 + 18)
    assert m.body_gravcomp[m.names["body"]["body:camera"]] == 1.0
    cam = m.jnt_qposadr[m.names["joint"]["camera:joint"]]
    assert np.allclose(rec[cam:cam + 3], [0, 0, -1.0], atol=1e-6)           # gravity-compensated camera hovers
    for a in info["object_qposadr"]:
        assert 0.005 < rec[a + 2] < 0.06 and np.abs(rec[a:a + 2]).max() < 0.3  # objects rest on the table near the drop point
    assert np.abs(rec[m.nq:m.nq + m.nv]).max() < 1e-3                        # settled
    assert abs(rec[2] - 1.5) < 2e-3                                           # gripper hangs on its weld (sags < 2 mm)


@pytest.mark.parametrize("f64", [True, False])
def test_clutter_programs_lane1_vs_oracle(clutter, f64):
    m, info, rec, pose7, joints = clutter
    sched = (500, 300, 0, 0, 0.05, 0.0)
    ofree, _ = batch(m, 2, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"],
                     RolloutCfg(*sched), 4, scene=rec, ground_name="geom:table")
    olab, osteps = batch(m, 3, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"],
                         RolloutCfg(*sched), 4, scene=rec, ground_name="geom:table")
    L = mlib.BatchSim(m, lib=mlib.bind(C.CDLL(lane1.build(f64)), prefix="l1_"), prefix="l1_", ground_name="geom:table")
    free = L.clutter_collision_mask(rec, pose7, joints, info["joint_qposadr"], info["base_qposadr"])
    lab, steps = L.clutter_stable_mask(rec, pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], mlib.MgsRolloutCfg(*sched))
    assert (free == ofree).all()
    assert (lab == olab).mean() >= (1.0 if f64 else 10 / 12)
    assert set(np.unique(osteps)) <= {500 + 100 * k for k in range(1, 4)}  # early break only at (t+1) % 100 == 0


def test_state_vector_roundtrip_and_bounds():
    from mj_grasp_sim_b200.mgs.env.clutter_table import ClutterTableEnv
    from mj_grasp_sim_b200.mgs.gripper.selector import get_gripper
    from mj_grasp_sim_b200.mgs.obj.hull import ObjectConvexHull
    from mj_grasp_sim_b200.mgs.util.geo.transforms import SE3Pose
    objs = []
    for i in range(2):
        pts, mass = scenes.random_hull_points(i, 16)
        objs.append(ObjectConvexHull(SE3Pose(np.array([-8.0, -8.0 + 0.5 * i, 0.06]), np.array([1.0, 0, 0, 0]), "wxyz"), f"o{i}", [pts], mass))
    env = ClutterTableEnv(get_gripper("PandaGripper"), objs)
    m = env.model
    st = env.get_state()
    # mjSTATE_INTEGRATION: time, qpos, qvel, act, qacc_warmstart, ctrl, qfrc_applied, xfrc_applied, eq_active, mocap_pos, mocap_quat
    assert len(st) == 1 + m.nq + m.nv + m.nv + m.nu + m.nv + 6 * m.nbody + int(m.arr["neq"]) + 7
    st2 = st.copy()
    st2[1:4] = [0.1, 0.2, 0.9]
    env.set_state(st2)
    assert np.allclose(env.get_state(), st2)
    with pytest.raises(ValueError):
        env.set_state(st[:-1])
    d = env.to_dict()
    assert set(d) == {"gripper", "objects", "env_state"} and set(d["env_state"]) == {"geom_conaffinity", "geom_contype", "geom_rgba", "body_gravcomp", "state"}
    env2 = ClutterTableEnv.from_dict(d)
    assert np.allclose(env2.get_state(), st2)


def test_batched_scene_generation_equals_single(clutter):
    """scenes.gen_clutter_batch on the 1-lane host build: scene k of a batch is bit-identical to
    gen_clutter(seed k) - an environment's trajectory does not depend on what else is in the batch."""
    m, info, _, _, _ = clutter
    L = lane1.sim(m, f64=False)
    single_step = lambda rec, n: L.step(rec[None].astype(np.float32), n)[0].astype(np.float64)
    batch_step = lambda recs, n: L.step(recs.astype(np.float32), n).astype(np.float64)
    seeds = [3, 4]
    recs = scenes.gen_clutter_batch(m, info, batch_step, seeds)
    one = scenes.gen_clutter(m, info, single_step, seeds[1])
    assert recs.shape == (2, len(one)) and np.array_equal(recs[1], one)
    ok, after = scenes.scenes_stable(m, info, batch_step, recs.copy())
    assert ok.shape == (2,) and after.shape == recs.shape


def test_remove_obj_and_mask_roundtrip():
    """ClutterTableEnv.remove_obj (reference :146-155) on the compiled model + from_dict restoring edited masks (:394-397):
    the removed object stops colliding and floats (gravcomp 1) while the others keep falling; the scene dictionary carries it."""
    from mj_grasp_sim_b200.mgs.env.clutter_table import ClutterTableEnv
    from mj_grasp_sim_b200.mgs.gripper.selector import get_gripper
    from mj_grasp_sim_b200.mgs.obj.selector import get_objects
    env = ClutterTableEnv(get_gripper("PandaGripper"), get_objects(["hull:50:16", "hull:51:16"]))
    m = env.model
    npair0 = int(m.arr["npair"])
    gone, kept = env.objects[0], env.objects[1]
    env.remove_obj(env.get_object(gone.name))
    bid = m.names["body"][gone.name]
    assert int(m.arr["npair"]) < npair0 and m.body_gravcomp[bid] == 1.0
    cg_gone = np.nonzero(m.cgeom_bodyid == bid)[0]
    assert not np.isin(m.pair_geom1, cg_gone).any() and not np.isin(m.pair_geom2, cg_gone).any()
    assert (m.geom_contype[m.geom_bodyid == bid] == 0).all()
    s = OracleSim(m, ground_name="geom:table")
    s.reset()
    a_gone = int(m.jnt_qposadr[m.names["joint"][f"{gone.name}:joint"]])
    a_kept = int(m.jnt_qposadr[m.names["joint"][f"{kept.name}:joint"]])
    z0 = s.qpos[[a_gone + 2, a_kept + 2]].copy()
    s.step(150)
    assert abs(s.qpos[a_gone + 2] - z0[0]) < 1e-9          # compensated, nothing touches it
    assert s.qpos[a_kept + 2] < z0[1] - 1e-3               # the other one falls onto the table
    # scene dictionary round trip: masks, gravcomp and the reduced pair list come back
    env2 = ClutterTableEnv.from_dict(env.to_dict())
    assert int(env2.model.arr["npair"]) == int(m.arr["npair"]) and np.array_equal(env2.model.geom_contype, m.geom_contype)
    assert np.array_equal(env2.model.body_gravcomp, m.body_gravcomp) and np.array_equal(env2.model.pair_geom1, m.pair_geom1)


def test_shadow_hand_over_clutter_first_steps_lane1_vs_oracle():
    """The Shadow hand (spheres, capsules, cylinders, boxes, hulls) closing over a settled two-object scene: kernel source (fp64 1-lane
    build) vs the oracle over the first 150 steps of the close phase - same contact count at every step, qpos within 1e-6 - with the
    closed-form pairs (sphere / capsule against box) and MPR pairs both among the contacts."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import clutter_shadow_bench as csb
    m, info = scenes.build_clutter_scene("shadow", [0, 1])
    S = OracleSim(m, ground_name="geom:table")

    def ostep(rec, k):
        S.set_record(rec); S.step(k)
        return S.get_record()
    rec = scenes.gen_clutter(m, info, ostep, 7)
    pose7, joints = csb.make_inputs(scenes, m, info, rec, 8)
    L = mlib.BatchSim(m, lib=mlib.bind(C.CDLL(lane1.build(True)), prefix="l1_"), prefix="l1_", ground_name="geom:table")
    nq, nv, nu = m.nq, m.nv, m.nu
    ty, p1, p2, cg = m.arr["cgeom_type"], m.arr["pair_geom1"], m.arr["pair_geom2"], m.arr["cgeom_geomid"]
    pairmap = {(int(cg[p1[i]]), int(cg[p2[i]])): i for i in range(len(p1))}
    kinds = set()
    for cand in (1, 5):
        st = rec.copy()
        b = info["base_qposadr"]
        st[b:b + 7] = pose7[cand]
        for k, a in enumerate(info["joint_qposadr"]):
            st[a] = joints[cand, k]
        st[nq + 2 * nv:nq + 2 * nv + nu] = info["close_ctrl"]
        st[nq + 2 * nv + nu:nq + 2 * nv + nu + 7] = pose7[cand]
        sa = st[None].copy()
        S.set_record(st)
        for step in range(150):  # (candidate 5 meets its first manifold decision at a threshold at step 173: DESIGN.md 5)
            sa, d = L.step(sa, 1, want_diag=True)
            S.step(1)
            assert int(d["ncon"][0]) == S.ncon, (cand, step)
            assert np.abs(sa[0, :nq] - S.qpos).max() < 1e-6, (cand, step)
            if step % 20 == 0:
                for c in S.contacts():
                    p = pairmap[(int(c[13]), int(c[14]))]
                    kinds.add(tuple(sorted((int(ty[p1[p]]), int(ty[p2[p]])))))
    GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX, GEOM_MESH = 2, 3, 6, 7
    assert kinds & {(GEOM_SPHERE, GEOM_BOX), (GEOM_CAPSULE, GEOM_BOX), (GEOM_CAPSULE, GEOM_CAPSULE), (GEOM_SPHERE, GEOM_CAPSULE)}, kinds
    assert kinds & {(GEOM_BOX, GEOM_MESH), (GEOM_MESH, GEOM_MESH), (GEOM_CAPSULE, GEOM_MESH), (GEOM_BOX, GEOM_BOX)}, kinds
