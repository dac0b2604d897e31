"""Environment-per-CTA ("wide") kernel variant vs the warp-per-environment variants on the same models, and its throughput on
the config-5 scene.  GPU box:  python tools/wide_check.py [parity|race|clutter] ...
  parity : Panda-on-cube, Robotiq, Shadow-on-hull - 50 steps of qpos/qvel and short-schedule labels, wide vs warp vs oracle
  race   : a tiny wide launch (for `compute-sanitizer --tool racecheck`)
  clutter: Shadow + 10 objects (nv = 94): gen_clutter + close/lift throughput, wide vs warp, a few candidates vs the oracle"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim, MgsRolloutCfg


def sim(m, variant, **kw):
    if variant:
        os.environ["MGS_KERNEL_VARIANT"] = variant
    else:
        os.environ.pop("MGS_KERNEL_VARIANT", None)
    try:
        return BatchSim(m, **kw)
    finally:
        os.environ.pop("MGS_KERNEL_VARIANT", None)


def start_states(G, m, info, pose7, joints, n):
    qpos = np.tile(m.qpos0, (n, 1))
    b = info["base_qposadr"]
    qpos[:, b:b + 7] = pose7[:n]
    for k, a in enumerate(info["joint_qposadr"]):
        qpos[:, a] = joints[:n, k]
    return G.pack_state(qpos, np.zeros((n, m.nv)), ctrl=np.tile(info["close_ctrl"], (n, 1)), mocap_pos=pose7[:n, :3], mocap_quat=pose7[:n, 3:7])


def parity(f64=False):
    from oracle import oracle as orc
    for gripper, kind, n, sched in (("panda", "cube", 24, (300, 100, 20, 0, 0.02, 0.02)), ("robotiq2f85", "hull", 16, (300, 100, 20, 0, 0.02, 0.02)),
                                    ("shadow", "hull", 12, (300, 100, 20, 0, 0.02, 0.02))):
        m, info, pose7, joints = scenes.workload(gripper, kind, 0, n)
        A, B = sim(m, None, f64=f64), sim(m, "wide", f64=f64)
        assert B.info.lanes_per_env == 256 and A.info.lanes_per_env == 32, (A.info.lanes_per_env, B.info.lanes_per_env)
        sa, sb = start_states(A, m, info, pose7, joints, n), start_states(B, m, info, pose7, joints, n)
        for k in range(5):
            sa, da = A.step(sa, 10, want_diag=True)
            sb, db = B.step(sb, 10, want_diag=True)
            ua, ub = A.unpack_state(sa), B.unpack_state(sb)
            print(gripper, "step", 10 * (k + 1), "ncon equal", bool((da["ncon"] == db["ncon"]).all()), "qpos diff %.2e" % np.abs(ua["qpos"] - ub["qpos"]).max(),
                  "qvel diff %.2e" % np.abs(ua["qvel"] - ub["qvel"]).max(), "bad", int(db["bad"].sum()), "ovf", int(db["overflow"].sum()), flush=True)
        args = (pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
        la, sta = A.stability(*args)
        t = time.time(); lb, stb = B.stability(*args); tb = time.time() - t
        olab, osteps = orc.batch(m, 1, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"],
                                 orc.RolloutCfg(*sched), os.cpu_count() or 1)
        fa, fb = A.collision_mask(*args[:4]), B.collision_mask(*args[:4])
        print(gripper, "labels wide==warp", float((la == lb).mean()), "steps equal", float((sta == stb).mean()), "| wide vs oracle", float((lb == olab).mean()),
              "warp vs oracle", float((la == olab).mean()), "| free equal", bool((fa == fb).all()), "| wide %.2fs" % tb, flush=True)
        A.close(); B.close()


def race():
    m, info, pose7, joints = scenes.workload("panda", "cube", 0, 4)
    B = sim(m, "wide")
    lab, st = B.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(30, 10, 2, 0, 0.02, 0.02))
    print("race run done", lab, st)


def clutter(n=296, n_or=8, caps=(48, 240)):
    from oracle import oracle as orc
    m, info = scenes.build_clutter_scene("shadow", list(range(10)))
    rec = None
    for variant in (("wide", "w12") if caps == (48, 240) else ("wide",)):
        G = sim(m, variant, ground_name="geom:table", ncon_max=caps[0], nefc_max=caps[1])
        print(variant, "model nv", m.nv, "pairs", len(m.pair_geom1), "| caps", G.info.ncon_max, G.info.nefc_max, "smem/env", G.info.smem_bytes_per_env,
              "lanes/env", G.info.lanes_per_env, "envs/SM", G.info.warps_per_block * G.info.blocks_per_sm, flush=True)
        if rec is None:
            step_fn = lambda r, k: G.step(r[None].astype(np.float32), k)[0].astype(np.float64)
            t = time.time(); rec = scenes.gen_clutter(m, info, step_fn, 7); print("gen_clutter (1 env, 11700 steps): %.1fs" % (time.time() - t), flush=True)
            H, w = scenes.clutter_candidates(m, info, rec, n, 2)
            g = scenes.GRIPPERS["shadow"]
            Rt = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
            T = np.eye(4); T[:3, :3] = Rt; T[:3, 3] = -Rt @ np.array([0.01, -0.06, 0.12])
            pose7 = scenes.process_poses(H @ T, "shadow")
            jid = [m.names["joint"][j] for j in g["joints"]]
            joints = np.clip(np.asarray(g["open_pose"])[None] + np.random.default_rng(3).normal(scale=0.05, size=(n, 22)), m.jnt_range[jid, 0], m.jnt_range[jid, 1]).astype(np.float32)
        sched = (300, 200, 0, 0, 0.02, 0.0)
        free = G.clutter_collision_mask(rec, pose7, joints, info["joint_qposadr"], info["base_qposadr"])
        t = time.time(); lab, steps = G.clutter_stable_mask(rec, pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched)); dt = time.time() - t
        print(f"{variant}: stable mask {int(steps.sum())} env-steps in {dt:.2f}s = {steps.sum() / dt:.4g} env-steps/s, free {free.mean():.2f} stable {lab.mean():.2f}, overflowed {G.overflow_count()} of {n}", flush=True)
        if variant == "wide":
            k = min(n_or, n)
            a = (pose7[:k].astype(np.float64), info["base_qposadr"], joints[:k].astype(np.float64), info["joint_qposadr"], info["close_ctrl"], orc.RolloutCfg(*sched), os.cpu_count())
            t = time.time()
            ofree, _ = orc.batch(m, 2, *a, scene=rec, ground_name="geom:table")
            olab, osteps = orc.batch(m, 3, *a, scene=rec, ground_name="geom:table")
            to = time.time() - t
            print(f"oracle on the first {k}: {to:.1f}s ({osteps.sum() / to:.4g} env-steps/s on {os.cpu_count()} threads) | free agree {(free[:k] == ofree).mean():.3f} "
                  f"stable agree {(lab[:k] == olab).mean():.3f}", flush=True)
            lab_w = lab
        else:
            print("labels wide==warp", float((lab == lab_w).mean()), flush=True)
        G.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if what == "parity":
        parity("f64" in sys.argv)
    elif what == "race":
        race()
    else:
        a = [int(x) for x in sys.argv[2:]]
        clutter(*a[:2], caps=tuple(a[2:4]) if len(a) >= 4 else (48, 240))
