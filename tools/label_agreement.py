"""Label agreement of the CUDA path (fp32 product build and the fp64 ablation) with the fp64 oracle, full
8000-step schedule, per gripper.  Run on the GPU box: python tools/label_agreement.py [n_per_gripper]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim, MgsRolloutCfg, SO_PATH_F64
from oracle.oracle import RolloutCfg, batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None  # e.g. "shadow,leap"
rows = []
for gripper, kind, seeds in (("panda", "cube", [0]), ("panda", "hull", [0, 1]), ("vx300", "hull", [0, 1]), ("robotiq2f85", "hull", [0, 1]),
                             ("allegro", "hull", [0]), ("leap", "hull", [0]), ("shadow", "hull", [0])):
    if only is not None and gripper not in only:
        continue
    for seed in seeds:
        m, info, pose7, joints = scenes.workload(gripper, kind, seed, n)
        rep = scenes.GRIPPERS[gripper]["repose"]
        sched = (3000, 3000, 500, rep, 0.1, 0.02)
        t = time.time()
        ofree, _ = batch(m, 0, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"], RolloutCfg(*sched), os.cpu_count())
        olab, osteps = batch(m, 1, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"], RolloutCfg(*sched), os.cpu_count())
        t_or = time.time() - t
        row = dict(gripper=gripper, object=f"{kind}:{seed}", n=n, oracle_stable=float(olab.mean()), oracle_free=float(ofree.mean()), oracle_s=round(t_or, 1))
        for tag, f64 in (("f32", False), ("f64", True)):
            if f64 and not os.path.exists(SO_PATH_F64):
                continue
            G = BatchSim(m, f64=f64)
            t = time.time()
            free = G.collision_mask(pose7, joints, info["joint_qposadr"], info["base_qposadr"])
            lab, steps = G.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], MgsRolloutCfg(*sched))
            row[f"{tag}_free_agree"] = float((free == ofree).mean())
            row[f"{tag}_stable_agree"] = float((lab == olab).mean())
            row[f"{tag}_overflow"] = G.overflow_count()
            row[f"{tag}_s"] = round(time.time() - t, 1)
            row[f"{tag}_env_steps"] = int(steps.sum())
            G.close()
        rows.append(row)
        print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", os.environ.get("MGS_LABELS_OUT", "label_agreement.json")), "w"), indent=1)
