"""First GPU bring-up: step parity vs the oracle, label agreement, and a rough timing."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim, MgsRolloutCfg
from oracle.oracle import OracleSim, RolloutCfg, batch

m, info, pose7, joints = scenes.workload("panda", "cube", 0, 64)
G = BatchSim(m)
print("info", {k: getattr(G.info, k) for k, _ in G.info._fields_}, flush=True)
s = OracleSim(m)
p7 = np.array([0, 0, -0.102, 0.70710677, 0, 0, 0.70710677])
s.reset(); s.place(p7, 0, np.array([0.0325, -0.0075]), info["joint_qposadr"]); s.ctrl[:] = [0, -0.04]
st = G.pack_state(s.qpos.copy(), s.qvel.copy(), ctrl=s.ctrl.copy(), mocap_pos=s.mocap_pos[0].copy(), mocap_quat=s.mocap_quat[0].copy())
st = np.repeat(st, 8, axis=0)
for k in range(6):
    s.step(50)
    st, d = G.step(st, 50, want_diag=True)
    u = G.unpack_state(st)
    print("t", 50 * (k + 1), "qpos err", np.abs(u["qpos"] - s.qpos).max(), "qvel err", np.abs(u["qvel"] - s.qvel).max(), "ncon", s.ncon, d["ncon"],
          "niter", s.niter, d["niter"], "bad", d["bad"].max(), flush=True)
cfgo = RolloutCfg(3000, 3000, 500, 0, 0.1, 0.02)
cfg = MgsRolloutCfg(3000, 3000, 500, 0, 0.1, 0.02)
for kind, seed in (("cube", 0), ("hull", 0)):
    m, info, pose7, joints = scenes.workload("panda", kind, seed, 64)
    G = BatchSim(m)
    t = time.time()
    fo, _ = batch(m, 0, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"], cfgo, os.cpu_count())
    lo, so = batch(m, 1, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"], cfgo, os.cpu_count())
    to = time.time() - t
    t = time.time()
    fg = G.collision_mask(pose7, joints, info["joint_qposadr"], info["base_qposadr"])
    lg, sg = G.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], cfg)
    tg = time.time() - t
    print(kind, seed, "free agree", (fo == fg).mean(), "stable agree", (lo == lg).mean(), "oracle stable", lo.mean(), "gpu stable", lg.mean(),
          "steps", so.sum(), sg.sum(), "t_oracle", round(to, 2), "t_gpu", round(tg, 2), flush=True)
# throughput at 4096 candidates
m, info, pose7, joints = scenes.workload("panda", "hull", 0, 4096)
G = BatchSim(m)
for rep in range(2):
    t = time.time()
    lg, sg = G.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], cfg)
    tg = time.time() - t
    print("4096 hull candidates: stable", lg.mean(), "steps", sg.sum(), "time", round(tg, 3), "env-steps/s", sg.sum() / tg, "grasps/s", 4096 / tg, flush=True)
