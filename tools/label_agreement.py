"""Label agreement of the CUDA path (fp32 product build and the fp64 ablation) with the fp64 oracle over the full 8000-step
schedule, per gripper, on MARGINAL candidate sets (scenes.workload(..., marginal=True): oracle stable fraction 0.2-0.6, so that a
constant predictor cannot score well).

  python tools/label_agreement.py --make-oracle [n]   (CPU, anywhere)  oracle labels -> tests/golden/labels_r2/*.npz
  python tools/label_agreement.py [n] [grippers]      (GPU box)        CUDA labels vs the cached oracle labels
                                                                       -> gpurun_out/label_agreement_r2.json (copy to profiles/)
The cache keys on a checksum of the candidate arrays: if a workload generator changes, the cache is refused, not silently reused.
"""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mj_grasp_sim_b200 import scenes

GOLD = os.path.join(ROOT, "tests", "golden", "labels_r2")
# (gripper, object kind, object seeds): n candidates per object
SETS = (("panda", "hull", (0, 1)), ("vx300", "hull", (0, 1)), ("robotiq2f85", "hull", (0, 1)), ("allegro", "hull", (0, 1)),
        ("leap", "hull", (0, 1)), ("shadow", "hull", (0, 1)))
SCHED = lambda g: (3000, 3000, 500, scenes.GRIPPERS[g]["repose"], 0.1, 0.02)


def checksum(pose7, joints):
    return hashlib.sha256(np.ascontiguousarray(pose7).tobytes() + np.ascontiguousarray(joints).tobytes()).hexdigest()[:16]


def cache_path(gripper, kind, seed, n):
    return os.path.join(GOLD, f"{gripper}_{kind}{seed}_{n}.npz")


def load_oracle(gripper, kind, seed, n, pose7, joints):
    d = np.load(cache_path(gripper, kind, seed, n))
    if str(d["checksum"]) != checksum(pose7, joints):
        raise RuntimeError(f"oracle label cache {cache_path(gripper, kind, seed, n)} was made for different candidates: rerun --make-oracle")
    return d["free"].astype(bool), d["stable"].astype(bool), d["steps"]


def make_oracle(n, only):
    from oracle.oracle import RolloutCfg, batch
    os.makedirs(GOLD, exist_ok=True)
    for gripper, kind, seeds in SETS:
        if only and gripper not in only:
            continue
        for seed in seeds:
            m, info, pose7, joints = scenes.workload(gripper, kind, seed, n, marginal=True)
            a = (pose7.astype(np.float64), info["base_qposadr"], joints.astype(np.float64), info["joint_qposadr"], info["close_ctrl"],
                 RolloutCfg(*SCHED(gripper)), os.cpu_count() or 1)
            t = time.time()
            free, _ = batch(m, 0, *a)
            lab, steps = batch(m, 1, *a)
            np.savez_compressed(cache_path(gripper, kind, seed, n), free=free, stable=lab, steps=steps.astype(np.int32),
                                checksum=checksum(pose7, joints))
            print(json.dumps(dict(gripper=gripper, object=f"{kind}:{seed}", n=n, oracle_free=float(free.mean()), oracle_stable=float(lab.mean()),
                                  seconds=round(time.time() - t, 1))), flush=True)


def run_escalated(G_small, make_big, call):
    """Labels with NO truncated contact set: candidates whose environment overflowed the first-pass capacities are re-run up the
    ladder mgs.env's EscalatingSim climbs - the default capacity of a single-object scene (32 contacts) when the first pass was
    cut below it, then the largest capacity that fits one CTA's shared memory.  Returns (arrays..., n_overflow_first, n_overflow_left)."""
    out = [np.array(o) for o in call(G_small, None)]
    n = len(out[0])
    over = np.nonzero(G_small.last_aux(n)["overflow"])[0]
    first = int(len(over))
    for rung in ((32,), (256, 128, 96, 64)):
        if not len(over):
            break
        big = None
        for nc in rung:
            if nc <= G_small.info.ncon_max:
                continue
            try:
                big = make_big(nc)
                break
            except Exception:
                continue
        if big is None:
            continue
        again = call(big, over)
        for o, a in zip(out, again):
            o[over] = a
        over = over[np.asarray(big.last_aux(len(over))["overflow"], dtype=bool)]
        big.close()
    return out, first, int(len(over))


def measure_one(gripper, kind, seed, n, f64):
    """CUDA labels (one build) of one cached marginal set vs the oracle's.  Returns a dict of figures."""
    from mj_grasp_sim_b200.lib import BatchSim, MgsRolloutCfg
    m, info, pose7, joints = scenes.workload(gripper, kind, seed, n, marginal=True)
    ofree, olab, osteps = load_oracle(gripper, kind, seed, n, pose7, joints)
    G = BatchSim(m, f64=f64)
    t = time.time()
    free = G.collision_mask(pose7, joints, info["joint_qposadr"], info["base_qposadr"])
    sel = lambda idx: (pose7, joints) if idx is None else (pose7[idx], joints[idx])
    (lab, steps), over0, over1 = run_escalated(
        G, lambda nc: BatchSim(m, f64=f64, ncon_max=nc), lambda sim, idx: sim.stability(*sel(idx), info["joint_qposadr"], info["base_qposadr"],
                                                                                  info["close_ctrl"], MgsRolloutCfg(*SCHED(gripper))))
    lab = lab.astype(bool)
    G.close()
    return dict(free_agree=float((free == ofree).mean()), stable_agree=float((lab == olab).mean()), stable_fraction=float(lab.mean()),
                false_pos=int((lab & ~olab).sum()), false_neg=int((~lab & olab).sum()), overflow_first_pass=over0, overflow=over1,
                s=round(time.time() - t, 1), env_steps=int(steps.sum()), oracle_stable=float(olab.mean()), oracle_free=float(ofree.mean()))


def measure(n, only):
    from mj_grasp_sim_b200.lib import SO_PATH_F64
    rows = []
    for gripper, kind, seeds in SETS:
        if only and gripper not in only:
            continue
        for seed in seeds:
            row = dict(gripper=gripper, object=f"{kind}:{seed}", n=n, marginal=scenes.MARGINAL[gripper],
                       product="f64" if gripper in scenes.F64_GRIPPERS else "f32")
            for tag, f64 in (("f32", False), ("f64", True)):
                if f64 and (not os.path.exists(SO_PATH_F64) or os.environ.get("MGS_LABELS_SKIP_F64")):
                    continue
                r = measure_one(gripper, kind, seed, n, f64)
                row["oracle_stable"], row["oracle_free"] = r.pop("oracle_stable"), r.pop("oracle_free")
                row.update({f"{tag}_{k}": v for k, v in r.items()})
            rows.append(row)
            print(json.dumps(row), flush=True)
    # per-gripper totals; "product" = the build the precision policy selects for that gripper (mgs/gripper/base.py COMPUTE_F64)
    tot = {}
    for r in rows:
        t = tot.setdefault(r["gripper"], dict(n=0, f32=0.0, f64=0.0, stable=0.0, product=r["product"], overflow=0))
        t["n"] += r["n"]; t["f32"] += r["f32_stable_agree"] * r["n"]; t["f64"] += r.get("f64_stable_agree", 0.0) * r["n"]; t["stable"] += r["oracle_stable"] * r["n"]
        t["overflow"] += r.get(r["product"] + "_overflow", 0)
    summary = {g: dict(n=t["n"], oracle_stable=t["stable"] / t["n"], f32_stable_agree=t["f32"] / t["n"], f64_stable_agree=t["f64"] / t["n"],
                       product=t["product"], product_stable_agree=t[t["product"]] / t["n"], product_overflow=t["overflow"]) for g, t in tot.items()}
    print(json.dumps(summary), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(rows=rows, per_gripper=summary), open(os.path.join(ROOT, "gpurun_out", os.environ.get("MGS_LABELS_OUT", "label_agreement_r2.json")), "w"), indent=1)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if args else 512
    only = set(args[1].split(",")) if len(args) > 1 else None
    (make_oracle if "--make-oracle" in sys.argv else measure)(n, only)
