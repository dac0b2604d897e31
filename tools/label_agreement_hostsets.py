"""CPU proxy of tools/label_agreement.py on the MARGINAL sets: the kernel SOURCE (1-lane fp32 / fp64 host builds, tests/hostsim) against
the cached oracle labels of tests/golden/labels_r2, full 8000-step rollouts, split over the host cores.  Same source and arithmetic as
the CUDA builds up to the order of reductions (and without capacities); says nothing about the CUDA build itself.
  python tools/label_agreement_hostsets.py [grippers] [n_per_object]  -> JSON on stdout"""
import json, os, sys, time
from concurrent.futures import ProcessPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))


def one(a):
    g, seed, f64, lo, hi = a
    from hostsim import lane1
    from mj_grasp_sim_b200 import scenes
    from mj_grasp_sim_b200.lib import MgsRolloutCfg
    import label_agreement as la
    m, info, pose7, joints = scenes.workload(g, "hull", seed, 512, marginal=True)
    _, olab, _ = la.load_oracle(g, "hull", seed, 512, pose7, joints)
    L = lane1.sim(m, f64=f64)
    lab, _ = L.stability(pose7[lo:hi], joints[lo:hi], info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*la.SCHED(g)))
    lab, ol = lab.astype(bool), olab[lo:hi]
    return g, seed, f64, hi - lo, int((lab == ol).sum()), int((lab & ~ol).sum()), int((~lab & ol).sum()), int(ol.sum())


if __name__ == "__main__":
    grippers = sys.argv[1].split(",") if len(sys.argv) > 1 else ["allegro", "leap", "shadow"]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    from hostsim import lane1
    lane1.build(False); lane1.build(True)
    t = time.time()
    jobs = [(g, seed, f64, lo, min(n, lo + 32)) for g in grippers for seed in (0, 1) for f64 in (False, True) for lo in range(0, n, 32)]
    with ProcessPoolExecutor(os.cpu_count() or 1) as ex:
        rows = list(ex.map(one, jobs))
    out = {}
    for g, seed, f64, k, eq, fp, fn, st in rows:
        r = out.setdefault(g, {}).setdefault("f64" if f64 else "f32", dict(n=0, equal=0, false_pos=0, false_neg=0, oracle_stable=0))
        r["n"] += k; r["equal"] += eq; r["false_pos"] += fp; r["false_neg"] += fn; r["oracle_stable"] += st
    for g in out:
        for r in out[g].values():
            r["stable_agree"] = r["equal"] / r["n"]; r["oracle_stable"] = r["oracle_stable"] / r["n"]
    print(json.dumps(dict(what="kernel source (1-lane host builds) vs cached oracle labels, marginal sets, full schedule", per_gripper=out,
                          seconds=round(time.time() - t, 1)), indent=1))
