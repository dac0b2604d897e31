"""A/B of two builds of the library on IDENTICAL inputs (config-5 scene and candidates made once): max |difference| of the state
after 1, 2, 5, 10, 20 steps.  GPU box: python tools/ab_wide.py build_ab/lib_a.so build_ab/lib_b.so"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from mj_grasp_sim_b200 import scenes, lib as mlib
import clutter_shadow_bench as csb
m, info = scenes.build_clutter_scene("shadow", list(range(10)))
sims = [mlib.BatchSim(m, ground_name="geom:table", ncon_max=80, lib=mlib.bind(C.CDLL(os.path.abspath(p)))) for p in sys.argv[1:3]]
A, B = sims
step_fn = lambda r, k: A.step(r[None].astype(np.float32), k)[0].astype(np.float64)
rec = scenes.gen_clutter(m, info, step_fn, 7)
n = 16
pose7, joints = csb.make_inputs(scenes, m, info, rec, 8 * n)
free = A.clutter_collision_mask(rec, pose7, joints, info["joint_qposadr"], info["base_qposadr"])
pose7, joints = pose7[free][:n], joints[free][:n]
n = len(pose7)
nq, nv, nu = m.nq, m.nv, m.nu
st = np.tile(rec, (n, 1))
b = info["base_qposadr"]
st[:, b:b + 7] = pose7
for k, a in enumerate(info["joint_qposadr"]):
    st[:, a] = joints[:, k]
st[:, nq + 2 * nv:nq + 2 * nv + nu] = info["close_ctrl"]
st[:, nq + 2 * nv + nu:nq + 2 * nv + nu + 7] = pose7
sa = sb = st.astype(np.float32)
done = 0
for upto in (1, 2, 5, 10, 20):
    sa, da = A.step(sa, upto - done, want_diag=True); sb, db = B.step(sb, upto - done, want_diag=True); done = upto
    print("step", upto, "max |dqpos| %.3e" % np.abs(sa[:, :nq] - sb[:, :nq]).max(), "max |dqvel| %.3e" % np.abs(sa[:, nq:nq + nv] - sb[:, nq:nq + nv]).max(),
          "qacc diff %.3e" % np.abs(da["qacc"] - db["qacc"]).max(), "ncon equal", bool((da["ncon"] == db["ncon"]).all()), "niter", da["niter"].mean(), db["niter"].mean(), flush=True)
