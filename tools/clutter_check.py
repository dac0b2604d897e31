import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim, MgsRolloutCfg
from oracle import oracle as orc
m, info = scenes.build_clutter_scene("panda", [10, 11, 12])
G = BatchSim(m, ground_name="geom:table")
print("info", G.info.ncon_max, G.info.nefc_max, G.info.smem_bytes_per_env, G.info.warps_per_block, G.info.blocks_per_sm)
def step_fn(rec, n):
    return G.step(rec[None].astype(np.float32), n)[0].astype(np.float64)
t = time.time(); rec = scenes.gen_clutter(m, info, step_fn, 3); print("gen_clutter on GPU", round(time.time() - t, 1), "s")
s = orc.OracleSim(m, ground_name="geom:table")
def step_o(rec, n):
    s.set_record(rec); s.step(n); return s.get_record()
t = time.time(); rec_o = scenes.gen_clutter(m, info, step_o, 3); print("gen_clutter on oracle", round(time.time() - t, 1), "s")
for a in info["object_qposadr"]: print(" gpu obj", rec[a:a + 3].round(4), " oracle obj", rec_o[a:a + 3].round(4))
H, w = scenes.clutter_candidates(m, info, rec, 96, 1)
pose7 = scenes.process_poses(H, "panda"); joints = scenes.panda_width_to_joints(w).astype(np.float32)
sched = (3000, 1000, 0, 0, 0.1, 0.0)
ofree, _ = orc.batch(m, 2, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"], orc.RolloutCfg(*sched), os.cpu_count(), scene=rec, ground_name="geom:table")
olab, osteps = orc.batch(m, 3, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"], orc.RolloutCfg(*sched), os.cpu_count(), scene=rec, ground_name="geom:table")
for f64 in (False, True):
    G2 = BatchSim(m, ground_name="geom:table", f64=f64)
    free = G2.clutter_collision_mask(rec, pose7, joints, info["joint_qposadr"], info["base_qposadr"])
    t = time.time(); lab, steps = G2.clutter_stable_mask(rec, pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched)); dt = time.time() - t
    print("f64" if f64 else "f32", "free agree", (free == ofree).mean(), "stable agree", (lab == olab).mean(), "oracle stable", olab.mean(), "gpu stable", lab.mean(),
          "overflow", G2.overflow_count(), "steps", steps.sum(), osteps.sum(), "time", round(dt, 2))
    print("  mismatches", np.nonzero(lab != olab)[0], steps[lab != olab], osteps[lab != olab])
