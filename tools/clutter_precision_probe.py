"""Is the clutter label disagreement arithmetic precision?  Shadow hand over a k-object clutter scene (k small enough for the fp64 build
to fit one SM), the same candidates through the fp32 build, the fp64 build and the oracle.  GPU box:
  python tools/clutter_precision_probe.py [k_objects] [n_candidates] [close] [lift]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim, MgsRolloutCfg
from oracle import oracle as orc
import clutter_shadow_bench as csb

k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
close, lift = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (300, 200)
gripper = sys.argv[5] if len(sys.argv) > 5 else "shadow"
m, info = scenes.build_clutter_scene(gripper, list(range(k)))
S = orc.OracleSim(m, ground_name="geom:table")
def ostep(rec, kk):
    S.set_record(rec); S.step(kk); return S.get_record()
t = time.time(); rec = scenes.gen_clutter(m, info, ostep, 7); print("oracle scene %.1fs, nv %d" % (time.time() - t, m.nv), flush=True)
if gripper == "shadow":
    pose7, joints = csb.make_inputs(scenes, m, info, rec, n)
else:
    H, w = scenes.clutter_candidates(m, info, rec, n, 2)
    pose7 = scenes.process_poses(H, gripper).astype(np.float32); joints = scenes.panda_width_to_joints(w).astype(np.float32)
sched = (close, lift, 0, 0, 0.3 * lift / 3000.0, 0.0)
a = (pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"], orc.RolloutCfg(*sched), os.cpu_count() or 1)
olab, osteps = orc.batch(m, 3, *a, scene=rec, ground_name="geom:table")
print("oracle stable fraction %.2f" % olab.mean(), flush=True)
for f64 in (False, True):
    G = None
    for nc in (64, 48, 40, 32):
        try:
            G = BatchSim(m, ground_name="geom:table", f64=f64, ncon_max=nc)
            break
        except Exception as ex:
            err = ex
    if G is None:
        print("f64" if f64 else "f32", "does not fit:", err); continue
    lab, steps = G.clutter_stable_mask(rec, pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    over = G.last_aux(n)["overflow"]
    print("f64" if f64 else "f32", "lanes/env", G.info.lanes_per_env, "caps", G.info.ncon_max, G.info.nefc_max, "| agree %.3f" % (lab == olab).mean(),
          "false_pos", int((lab & ~olab).sum()), "false_neg", int((~lab & olab).sum()), "overflow", int(over.sum()), "stable fraction %.2f" % lab.mean(),
          "| steps equal where labels equal %.3f" % (steps[lab == olab] == osteps[lab == olab]).mean(), flush=True)
    # scene drift: the settled scene itself stepped 500 steps, max object displacement (should be ~0: the scene is at rest)
    st = G.step(rec[None].astype(G.real), 500)[0].astype(np.float64)
    d = max(np.abs(st[q:q + 3] - rec[q:q + 3]).max() for q in info["object_qposadr"])
    S.set_record(rec); S.step(500); so = S.get_record()
    do = max(np.abs(so[q:q + 3] - rec[q:q + 3]).max() for q in info["object_qposadr"])
    print("   rest drift over 500 steps: kernel %.2e m, oracle %.2e m" % (d, do), flush=True)
    G.close()
