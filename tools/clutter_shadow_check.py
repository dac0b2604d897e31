"""Config 5 at full size: Shadow hand + 10-object clutter table (nv = 94).  Generates one scene with the step mode, evaluates a
batch of candidates with a shortened close+lift schedule, compares a few of them with the oracle and prints the throughput.
Run on the GPU box: python tools/clutter_shadow_check.py [n_candidates] [n_oracle]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim, MgsRolloutCfg
from oracle import oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n_or = int(sys.argv[2]) if len(sys.argv) > 2 else 16
m, info = scenes.build_clutter_scene("shadow", list(range(10)))
G = BatchSim(m, ground_name="geom:table", ncon_max=48, nefc_max=240)
print("model nq", m.nq, "nv", m.nv, "pairs", len(m.pair_geom1), "| caps", G.info.ncon_max, G.info.nefc_max, "smem/env", G.info.smem_bytes_per_env,
      "envs/SM", G.info.warps_per_block * G.info.blocks_per_sm, flush=True)
step_fn = lambda rec, k: G.step(rec[None].astype(np.float32), k)[0].astype(np.float64)
t = time.time(); rec = scenes.gen_clutter(m, info, step_fn, 7); print("gen_clutter (1 env, 11700 steps):", round(time.time() - t, 1), "s", flush=True)
H, w = scenes.clutter_candidates(m, info, rec, n, 2)
g = scenes.GRIPPERS["shadow"]
Rt = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
T = np.eye(4); T[:3, :3] = Rt; T[:3, 3] = -Rt @ np.array([0.01, -0.06, 0.12])
pose7 = scenes.process_poses(H @ T, "shadow")
jid = [m.names["joint"][j] for j in g["joints"]]
joints = np.clip(np.asarray(g["open_pose"])[None] + np.random.default_rng(3).normal(scale=0.05, size=(n, 22)), m.jnt_range[jid, 0], m.jnt_range[jid, 1]).astype(np.float32)
sched = (300, 200, 0, 0, 0.02, 0.0)
t = time.time(); free = G.clutter_collision_mask(rec, pose7, joints, info["joint_qposadr"], info["base_qposadr"]); t_free = time.time() - t
t = time.time(); lab, steps = G.clutter_stable_mask(rec, pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched)); dt = time.time() - t
print(f"collision mask: {n} candidates in {t_free:.2f}s, free {free.mean():.2f} | stable mask: {int(steps.sum())} env-steps in {dt:.2f}s = {steps.sum() / dt:.4g} env-steps/s, "
      f"stable {lab.mean():.2f}, overflowed {G.overflow_count()}", flush=True)
k = min(n_or, n)
t = time.time()
ofree, _ = orc.batch(m, 2, pose7[:k].astype(np.float64), info["base_qposadr"], joints[:k].astype(np.float64), info["joint_qposadr"], info["close_ctrl"], orc.RolloutCfg(*sched), os.cpu_count(), scene=rec, ground_name="geom:table")
olab, osteps = orc.batch(m, 3, pose7[:k].astype(np.float64), info["base_qposadr"], joints[:k].astype(np.float64), info["joint_qposadr"], info["close_ctrl"], orc.RolloutCfg(*sched), os.cpu_count(), scene=rec, ground_name="geom:table")
to = time.time() - t
print(f"oracle on the first {k}: {to:.1f}s ({osteps.sum() / to:.4g} env-steps/s on {os.cpu_count()} threads) | free agree {(free[:k] == ofree).mean():.3f} stable agree {(lab[:k] == olab).mean():.3f}")
