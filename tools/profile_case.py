"""Short, representative launch of the rollout kernel for ncu (Panda on a hull object)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim, MgsRolloutCfg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
close = int(sys.argv[2]) if len(sys.argv) > 2 else 150
m, info, pose7, joints = scenes.workload("panda", "hull", 0, n)
G = BatchSim(m)
cfg = MgsRolloutCfg(close, 40, 10, 0, 0.01, 0.01)
for rep in range(2):
    t = time.time()
    lab, steps = G.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], cfg)
    dt = time.time() - t
    print("n", n, "stable", lab.mean(), "steps", steps.sum(), "time", round(dt, 3), "env-steps/s", steps.sum() / dt, flush=True)
