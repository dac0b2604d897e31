"""Steady-state launch of the environment-per-CTA kernel on the config-5 scene (Shadow hand + 10 objects) for ncu.
N candidates are placed over a settled scene and stepped through `settle` steps of the close phase (not profiled), then
`nstep` more steps run as the LAST launch - profile it with
  ncu -k regex:mgs_rollout_kernel_wide --launch-skip <printed count> --launch-count 1 ...
usage: python tools/profile_clutter.py [n_env] [settle] [nstep] [ncon_max]   (MGS_SCENE_STEPS shortens the scene generation)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim, load

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 300
nstep = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ncon = int(sys.argv[4]) if len(sys.argv) > 4 else 48
m, info = scenes.build_clutter_scene("shadow", list(range(10)))
G = BatchSim(m, ground_name="geom:table", ncon_max=ncon)
print("caps", G.info.ncon_max, G.info.nefc_max, "smem/env", G.info.smem_bytes_per_env, "lanes/env", G.info.lanes_per_env, flush=True)
step_fn = lambda r, k: G.step(r[None].astype(np.float32), k)[0].astype(np.float64)
rec = scenes.gen_clutter(m, info, step_fn, 7)
H, w = scenes.clutter_candidates(m, info, rec, n, 2)
g = scenes.GRIPPERS["shadow"]
Rt = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
T = np.eye(4); T[:3, :3] = Rt; T[:3, 3] = -Rt @ np.array([0.01, -0.06, 0.12])
pose7 = scenes.process_poses(H @ T, "shadow")
jid = [m.names["joint"][j] for j in g["joints"]]
joints = np.clip(np.asarray(g["open_pose"])[None] + np.random.default_rng(3).normal(scale=0.05, size=(n, 22)), m.jnt_range[jid, 0], m.jnt_range[jid, 1])
st = np.tile(rec, (n, 1))
b = info["base_qposadr"]
st[:, b:b + 7] = pose7
for k, a in enumerate(info["joint_qposadr"]):
    st[:, a] = joints[:, k]
nq, nv, nu = m.nq, m.nv, m.nu
st[:, nq + 2 * nv:nq + 2 * nv + nu] = info["close_ctrl"]
st[:, nq + 2 * nv + nu:nq + 2 * nv + nu + 7] = pose7
st = st.astype(np.float32)
st = G.step(st, settle)
print("launches before the profiled one:", load().mgs_launch_count(), flush=True)
t = time.time()
st2, d = G.step(st, nstep, want_diag=True)
dt = time.time() - t
print(f"steady clutter n={n} nstep={nstep}: {dt:.3f}s  env-steps/s {n * nstep / dt:.4g}  ncon mean {d['ncon'].mean():.1f} max {d['ncon'].max()} nefc mean {d['nefc'].mean():.1f} "
      f"niter mean {d['niter'].mean():.2f} bad {int(d['bad'].sum())} overflowed {G.overflow_count()}", flush=True)
