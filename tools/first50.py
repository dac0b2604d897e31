"""North-star trajectory tolerance: qpos / qvel of the CUDA path (fp32 product build and fp64 ablation) vs the fp64 oracle over the
first 50 `mj_step` after placing the gripper and commanding the close signal, per gripper, `n` candidates each.
Relative error = max |x - x_oracle| / max(1, max |x_oracle|), maximum over candidates and over steps 10, 20, .., 50.
Run on the GPU box: python tools/first50.py [n]  ->  gpurun_out/first50.json"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes
from mj_grasp_sim_b200.lib import BatchSim, SO_PATH_F64
from oracle.oracle import OracleSim, RolloutCfg, batch


def first50(gripper, n, f64, kind="hull", free_only=True, chunk=10):
    """free_only: keep the candidates the ORACLE labels collision-free - the ones the reference pipeline hands to the stability
    rollout (filter_to_stable.py:39-44); otherwise the first n candidates, penetrating starts included.
    chunk: steps per launch.  The per-pair collision cache (MPR warm start) lives for one launch, so chunk=1 runs every step with a
    cold cache - the oracle's situation - and isolates arithmetic differences from the warm start's within-mpr_tolerance ones."""
    m, info, pose7, joints = scenes.workload(gripper, kind, 0, 6 * n if free_only else n)
    if free_only:
        free, _ = batch(m, 0, pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"],
                        RolloutCfg(1, 1, 1, 0, 0.0, 0.0), os.cpu_count() or 1)
        keep = np.nonzero(free)[0][:n]
        pose7, joints, n = pose7[keep], joints[keep], len(keep)
    G = BatchSim(m, f64=f64)
    sims = []
    for i in range(n):
        s = OracleSim(m)
        s.reset()
        s.place(pose7[i].astype(np.float64), info["base_qposadr"], joints[i].astype(np.float64), info["joint_qposadr"])
        s.ctrl[:] = info["close_ctrl"]
        sims.append(s)
    st = np.concatenate([G.pack_state(s.qpos.copy(), s.qvel.copy(), ctrl=s.ctrl.copy(), mocap_pos=s.mocap_pos[0].copy(), mocap_quat=s.mocap_quat[0].copy()) for s in sims])
    eq_i, ev_i, same_i = np.zeros(n), np.zeros(n), np.ones(n, dtype=bool)
    for k in range(50 // chunk):
        st, d = G.step(st, chunk, want_diag=True)
        u = G.unpack_state(st)
        for i, s in enumerate(sims):
            s.step(chunk)
            eq_i[i] = max(eq_i[i], float(np.abs(u["qpos"][i] - s.qpos).max() / max(1.0, np.abs(s.qpos).max())))
            ev_i[i] = max(ev_i[i], float(np.abs(u["qvel"][i] - s.qvel).max() / max(1.0, np.abs(s.qvel).max())))
            same_i[i] &= int(d["ncon"][i]) == s.ncon
    G.close()
    eq, ev, ncon_equal = float(eq_i.max()) if n else 0.0, float(ev_i.max()) if n else 0.0, bool(same_i.all())
    # candidates whose contact COUNT equals the oracle's at every checkpoint: the drift figure without contact-onset flips
    qpos_same = float(eq_i[same_i].max()) if same_i.any() else 0.0
    return dict(steps_per_launch=chunk, n_same_contacts=int(same_i.sum()), qpos_rel_same_contacts=qpos_same, gripper=gripper, object=kind, n=n, candidates="collision-free" if free_only else "all", build="f64" if f64 else "f32", qpos_rel=eq, qvel_rel=ev, ncon_equal_every_10_steps=bool(ncon_equal),
                in_contact=int(sum(s.ncon > 0 for s in sims)))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rows = []
    for g in ("panda", "robotiq2f85", "vx300", "allegro", "leap", "shadow"):
        for f64 in (False, True):
            if f64 and not os.path.exists(SO_PATH_F64):
                continue
            for free_only in (True, False):
                rows.append(first50(g, n, f64, free_only=free_only))
                print(json.dumps(rows[-1]), flush=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "first50.json"), "w"), indent=1)
