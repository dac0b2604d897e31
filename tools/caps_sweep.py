"""Capacity sweep: per gripper, full 8000-step stability rollout of N candidates at several (ncon_max, nefc_max); prints environments
per SM, overflowed environments, time and label differences vs the default capacities.
  [MGS_SWEEP_F64=1] [MGS_SWEEP_CAPS=0:0,16:0,12:0] python tools/caps_sweep.py [n] [grippers]   (nefc 0 = the rows that match the contact capacity)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim, MgsRolloutCfg, MgsError

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
SWEEP = {"vx300": [(0, 0), (24, 100), (20, 84)], "allegro": [(0, 0), (32, 150), (24, 120), (16, 90)], "leap": [(0, 0), (32, 170), (24, 140), (16, 110)],
         "shadow": [(0, 0), (32, 180), (24, 150), (16, 120)]}
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
F64 = os.environ.get("MGS_SWEEP_F64", "0") == "1"
if os.environ.get("MGS_SWEEP_CAPS"):
    SWEEP = {g: [tuple(int(x) for x in c.split(":")) for c in os.environ["MGS_SWEEP_CAPS"].split(",")] for g in SWEEP}
for g, caps in SWEEP.items():
    if only and g not in only:
        continue
    m, info, pose7, joints = scenes.workload(g, "hull", 0, n)
    sched = MgsRolloutCfg(3000, 3000, 500, scenes.GRIPPERS[g]["repose"], 0.1, 0.02)
    ref = None
    for c in caps:
        try:
            G = BatchSim(m, f64=F64, ncon_max=c[0], nefc_max=c[1])
        except MgsError as ex:
            print(g, c, "rejected:", ex); continue
        G.stability(pose7[:64], joints[:64], info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(50, 10, 5, 0, 0.01, 0.01))  # warm-up
        t = time.time()
        lab, steps = G.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], sched)
        dt = time.time() - t
        if ref is None:
            ref = lab
        print(f"{g:8s} {'f64' if F64 else 'f32'} lanes/env {G.info.lanes_per_env:3d} caps {G.info.ncon_max:3d}/{G.info.nefc_max:3d} smem/env {G.info.smem_bytes_per_env:6d} envs/SM {G.info.warps_per_block * G.info.blocks_per_sm:2d} "
              f"overflowed {G.overflow_count():4d} time {dt:6.2f}s  {steps.sum() / dt:.4g} env-steps/s  stable {lab.mean():.3f}  labels differ from default caps: {(lab != ref).sum()}", flush=True)
        G.close()
