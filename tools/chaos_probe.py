"""How chaotic are the config-5 labels?  Oracle vs oracle: the same candidates of the same settled Shadow + 10-object scene, evaluated
twice by the fp64 oracle - once as given and once with the base poses perturbed at fp32 rounding level (1e-7 relative).  The fraction
of labels that survive that perturbation bounds the agreement ANY float32 implementation can reach with the oracle on this workload.
CPU only:  python tools/chaos_probe.py [n_candidates] [close] [lift]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from mj_grasp_sim_b200 import scenes
from oracle import oracle as orc
import clutter_shadow_bench as csb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
close, lift = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (300, 200)
m, info = scenes.build_clutter_scene("shadow", list(range(10)))
S = orc.OracleSim(m, ground_name="geom:table")

def step_fn(rec, k):
    S.set_record(rec)
    S.step(k)
    return S.get_record()

t = time.time(); rec = scenes.gen_clutter(m, info, step_fn, 7); print("oracle gen_clutter %.1fs" % (time.time() - t), flush=True)
pose7, joints = csb.make_inputs(scenes, m, info, rec, n)
sched = (close, lift, 0, 0, 0.3 * lift / 3000.0, 0.0)
thr = os.cpu_count() or 1
run = lambda p: orc.batch(m, 3, p.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"], orc.RolloutCfg(*sched),
                          thr, scene=rec, ground_name="geom:table")[0]
base = run(pose7)
rng = np.random.default_rng(0)
out = []
for eps in (1e-7, 1e-6):
    p = pose7.astype(np.float64) * (1.0 + eps * rng.standard_normal(pose7.shape))
    p[:, 3:] /= np.linalg.norm(p[:, 3:], axis=1, keepdims=True)
    lab = run(p)
    out.append((eps, float((lab == base).mean())))
    print(f"relative pose perturbation {eps:g}: {int((lab == base).sum())}/{n} labels unchanged (oracle stable fraction {base.mean():.2f})", flush=True)
