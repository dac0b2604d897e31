"""Steady-state launch of the rollout kernel for ncu: N environments are first stepped through the transient
of the close phase (launch 1, not profiled: use `ncu --launch-skip 1 --launch-count 1`), then `nstep` more
steps of the HOLD phase run as launch 2 - the regime that makes up >90 % of the 8000-step schedule.

usage: [MGS_STEADY_F64=1] python tools/profile_steady.py [gripper] [n_env] [settle_steps] [nstep] [ncon_max nefc_max]
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim

gripper = sys.argv[1] if len(sys.argv) > 1 else "panda"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
settle = int(sys.argv[3]) if len(sys.argv) > 3 else 1200
nstep = int(sys.argv[4]) if len(sys.argv) > 4 else 200
m, info, pose7, joints = scenes.workload(gripper, "hull", 0, n)
# capacities: argv[5], argv[6] (default: bench.py's for the two bench grippers)
caps = {"panda": dict(ncon_max=24, nefc_max=100), "robotiq2f85": dict(ncon_max=24, nefc_max=110)}.get(gripper, {})
if len(sys.argv) > 6:
    caps = dict(ncon_max=int(sys.argv[5]), nefc_max=int(sys.argv[6]))
G = BatchSim(m, f64=os.environ.get("MGS_STEADY_F64", "0") == "1", **caps)  # MGS_STEADY_F64=1: the fp64 build (product path of the hands)
qpos = np.tile(m.qpos0, (n, 1))
b = info["base_qposadr"]
qpos[:, b:b + 7] = pose7
for k, a in enumerate(info["joint_qposadr"]):
    qpos[:, a] = joints[:, k]
st = G.pack_state(qpos, np.zeros((n, m.nv)), ctrl=np.tile(info["close_ctrl"], (n, 1)), mocap_pos=pose7[:, :3], mocap_quat=pose7[:, 3:7])
t = time.time()
st = G.step(st, settle)
print("settle", settle, "steps:", round(time.time() - t, 3), "s", flush=True)
reps = int(os.environ.get("MGS_STEADY_REPS", "5"))  # under ncu use MGS_STEADY_REPS=1
times = []
for _ in range(reps):
    t = time.time()
    st2, d = G.step(st, nstep, want_diag=True)
    times.append(time.time() - t)
dt = min(times)
print(f"steady {gripper} n={n} nstep={nstep}: best {dt:.3f}s of {[round(x, 3) for x in times]}  env-steps/s {n * nstep / dt:.4g}  ncon mean {d['ncon'].mean():.2f} "
      f"nefc mean {d['nefc'].mean():.1f} niter mean {d['niter'].mean():.2f} max {d['niter'].max()} hist {np.bincount(d['niter'])[:8].tolist()} "
      f"bad {int(d['bad'].sum())} overflowed {G.overflow_count()} caps {G.info.ncon_max}/{G.info.nefc_max}  envs/SM {G.info.warps_per_block * G.info.blocks_per_sm}", flush=True)
