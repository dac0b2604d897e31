"""Label agreement of the kernel SOURCE (1-lane fp32 / fp64 host builds, tests/hostsim) with the oracle over full 8000-step rollouts.
CPU proxy of tools/label_agreement.py for when no GPU is available: same source and arithmetic as the CUDA build up to the
order of reductions; says nothing about the CUDA build itself.  python tools/label_agreement_host.py [n] -> profiles-style JSON on stdout"""
import json, os, sys, time
from concurrent.futures import ProcessPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(args):
    gripper, kind, seed, n, f64 = args
    from hostsim import lane1
    from mj_grasp_sim_b200 import scenes
    from mj_grasp_sim_b200.lib import MgsRolloutCfg
    from oracle.oracle import RolloutCfg, batch
    m, info, pose7, joints = scenes.workload(gripper, kind, seed, n)
    sched = (3000, 3000, 500, scenes.GRIPPERS[gripper]["repose"], 0.1, 0.02)
    t = time.time()
    L = lane1.sim(m, f64=f64)
    free = L.collision_mask(pose7, joints, info["joint_qposadr"], info["base_qposadr"])
    lab, _ = L.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
    a = (pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"], RolloutCfg(*sched), 1)
    ofree, _ = batch(m, 0, *a)
    olab, _ = batch(m, 1, *a)
    return dict(gripper=gripper, object=f"{kind}:{seed}", n=n, build="f64" if f64 else "f32", oracle_stable=float(olab.mean()),
                free_agree=float((free == ofree).mean()), stable_agree=float((lab == olab).mean()), seconds=round(time.time() - t, 1))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    from hostsim import lane1
    lane1.build(False); lane1.build(True)
    jobs = [(g, k, s, n, f) for g, k, s in (("panda", "cube", 0), ("panda", "hull", 0), ("vx300", "hull", 0), ("robotiq2f85", "hull", 0), ("allegro", "hull", 0),
                                            ("leap", "hull", 0), ("shadow", "hull", 0)) for f in (False, True)]
    jobs.sort(key=lambda j: -{"shadow": 5, "leap": 4, "allegro": 3, "robotiq2f85": 2}.get(j[0], 1))
    with ProcessPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        rows = list(ex.map(one, jobs))
    print(json.dumps(rows, indent=1))
