#!/usr/bin/env python
"""bench.py - grasp-evaluation rollout throughput on B200 (and the CPU arm beside it).

Workload (config.workload), default = BASELINE.json configs[1]: Robotiq 2F-85 on one synthetic convex-hull
object per rank, 4096 antipodal-style grasp candidates per object, the reference's stability rollout (close
3000 + lift 3000 + shake 2000 `mj_step` at dt = 1 ms; failed candidates stop early) - the loop of
/root/reference/mgs/env/gravityless_object_grasping.py:127-295.  The same line carries, under "also", the
Panda-on-convex-objects measurement (the gripper the north-star's target sentence names); `--workload panda`
makes it the primary.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload robotiq|panda]

A "step" is one full pass of the hot path over one batch of candidates.
  value  = env-steps/s with inputs resident in HBM, timed with CUDA events on the launch stream
  e2e    = the same metric through the host-pointer C-ABI call (pinned H2D + kernel + D2H inside)
  --impl reference = the CPU arm: the fp64 oracle port (MuJoCo is not installable here) on all
           host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CAND = 4096
ROLLOUT = dict(nstep_close=3000, nstep_lift=3000, shake_steps=500, repose_on_close=0, lift_dist=0.1, shake_dist=0.02)
# --workload: gripper, explicit per-environment capacities (contacts, constraint rows; 0 = the model's default), description
WORKLOADS = {
    "robotiq": ("robotiq2f85", (24, 110), "configs[1]: robotiq 2f-85 gripper, 1 synthetic 32-vertex convex-hull object per GPU (ycb recipe), 4096 antipodal candidates per object, close3000+lift3000+shake2000"),
    "panda": ("panda", (24, 100), "panda gripper, 1 synthetic 32-vertex convex-hull object per GPU (ycb recipe), 4096 antipodal candidates, close3000+lift3000+shake2000"),
}


def b_step(model):
    """Algorithmic HBM bytes per env-step: fp32 state in and out once per step."""
    return 4 * (2 * model.nq + 4 * model.nv + model.nu + 7) + 1


class ClockSampler:
    def __init__(self, index=0):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm)}


def cpu_arm(model, info, pose7, joints, target_seconds, threads):
    """Time the oracle port on a bounded sample; returns (env_steps_per_s, grasps_per_s, n_sample, seconds)."""
    from oracle.oracle import RolloutCfg, batch
    cfg = RolloutCfg(ROLLOUT["nstep_close"], ROLLOUT["nstep_lift"], ROLLOUT["shake_steps"], ROLLOUT["repose_on_close"],
                     ROLLOUT["lift_dist"], ROLLOUT["shake_dist"])
    chunk = max(threads * 2, 8)
    done, steps, t0 = 0, 0, time.perf_counter()
    while done < len(pose7):
        sl = slice(done, min(len(pose7), done + chunk))
        _, st = batch(model, 1, pose7[sl].astype(np.float64), info["base_qposadr"], joints[sl].astype(np.float64), info["joint_qposadr"],
                      info["close_ctrl"], cfg, threads)
        steps += int(st.sum())
        done = sl.stop
        if time.perf_counter() - t0 >= target_seconds:
            break
    dt = time.perf_counter() - t0
    return steps / dt, done / dt, done, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    from mj_grasp_sim_b200 import scenes
    from oracle import oracle as orc
    orc.build()
    gripper, _, WORKLOAD = WORKLOADS[args.workload]
    model, info, pose7, joints = scenes.workload(gripper, "hull", 0, N_CAND)
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_arm(model, info, pose7[:threads], joints[:threads], 0.0, threads)
    per_step = max(5.0, min(30.0, 90.0 / max(1, args.steps)))
    vals, gps, ns, t_total = [], [], 0, 0.0
    for k in range(args.steps):
        off = (k * 997) % (N_CAND // 2)
        v, g, n, dt = cpu_arm(model, info, pose7[off:], joints[off:], per_step, threads)
        vals.append(v); gps.append(g); ns += n; t_total += dt
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{ns} candidates over {args.steps} bounded steps"},
            "grasps_per_s": float(np.mean(gps)),
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                             "sample": f"oracle port (fp64 C restatement, MuJoCo not installable), {ns} candidates, ~{per_step:.0f}s per step"},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    from mj_grasp_sim_b200 import scenes
    from mj_grasp_sim_b200.lib import BatchSim, MgsRolloutCfg, load
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the rollout path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = load()
    dev = torch.device("cuda", local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()
    cfg = MgsRolloutCfg(**ROLLOUT)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # dram__bytes_read.sum + dram__bytes_write.sum of one rollout launch of each workload, captured once with ncu
    # (profiles/traffic_r1.json, written by tools/ncu_traffic.py from the ncu CSV)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))
    except Exception:
        pass

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(key, with_cpu):
        """One workload: W warm-up + K timed passes on device-resident inputs (CUDA events), then K passes through
        the host-pointer C ABI.  Returns the JSON fields of that workload (rank 0) or None."""
        gripper, (ncon_max, nefc_max), workload = WORKLOADS[key]
        if args.caps and key == args.workload:
            ncon_max, nefc_max = (int(x) for x in args.caps.split(","))
        model, info, pose7, joints = scenes.workload(gripper, "hull", rank, N_CAND)  # one object per rank (weak scaling)
        # per-environment capacities (contacts / constraint rows) bound shared memory per environment; the library
        # counts every environment that would have needed more (config.capacity.envs_overflowed)
        sim = BatchSim(model, device=local_rank, ncon_max=ncon_max, nefc_max=nefc_max)
        d_pose = torch.from_numpy(pose7).to(dev)
        d_joint = torch.from_numpy(joints).to(dev)
        d_lab = torch.zeros(N_CAND, dtype=torch.uint8, device=dev)
        d_steps = torch.zeros(N_CAND, dtype=torch.int32, device=dev)

        def one_step():
            sim.rollout_device(2, N_CAND, d_pose.data_ptr(), d_joint.data_ptr(), joints.shape[1], info["joint_qposadr"], info["base_qposadr"],
                               info["close_ctrl"], cfg, d_lab.data_ptr(), d_steps.data_ptr(), stream.cuda_stream)

        for _ in range(args.warmup):
            flush.fill_(1)
            one_step()
        barrier()
        launches0 = lib.mgs_launch_count()
        kern_ms, total_steps = [], 0
        with ClockSampler(local_rank) as clk:
            barrier()
            t_begin = time.perf_counter()
            for _ in range(args.steps):
                flush.fill_(1)  # L2 flush between timed iterations
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                one_step()
                e1.record(stream)
                e1.synchronize()
                kern_ms.append(e0.elapsed_time(e1))
                total_steps += int(d_steps.sum().item())
            barrier()
            t_wall = time.perf_counter() - t_begin
        launches = lib.mgs_launch_count() - launches0
        overflowed = sim.overflow_count()
        dev_s = sum(kern_ms) / 1e3
        labels_dev = d_lab.clone()
        # end-to-end through the host-pointer C ABI (pinned staging, H2D + kernel + D2H inside the call)
        barrier()
        e2e_steps, t0 = 0, time.perf_counter()
        for _ in range(args.steps):
            lab, st = sim.stability(pose7, joints, info["joint_qposadr"], info["base_qposadr"], info["close_ctrl"], cfg)
            e2e_steps += int(st.sum())
        barrier()
        e2e_s = time.perf_counter() - t0
        stats = torch.tensor([dev_s, e2e_s, float(total_steps), float(e2e_steps)], dtype=torch.float64, device=dev)
        if dist is not None:
            mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            # the only data-path exchange: gather the success labels (uint8[N] per rank) on every rank
            gathered = torch.empty(world * N_CAND, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(gathered, labels_dev)
            dev_s, e2e_s, total_steps, e2e_steps = mx[0].item(), mx[1].item(), sm[2].item(), sm[3].item()
            stable_frac = gathered.float().mean().item()
        else:
            stable_frac = labels_dev.float().mean().item()
        out = None
        if rank == 0:
            value = total_steps / dev_s
            bs = b_step(model)
            per_rank_steps = total_steps / world
            achieved = bs * (per_rank_steps / args.steps) / (np.mean(kern_ms) / 1e3) / 1e9
            out = {"metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
                   "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak",
                   "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                   "config": {"workload": workload, "candidates_per_gpu": N_CAND, "l2": "flushed between timed iterations (256 MiB fill)",
                              "stable_fraction": stable_frac,
                              "capacity": {"ncon_max": sim.info.ncon_max, "nefc_max": sim.info.nefc_max, "envs_overflowed": overflowed},
                              "envs_per_sm": sim.info.warps_per_block * sim.info.blocks_per_sm, "smem_bytes_per_env": sim.info.smem_bytes_per_env},
                   "grasps_per_s": world * N_CAND * args.steps / dev_s,
                   "e2e": {"value": e2e_steps / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": int(pose7.nbytes + joints.nbytes),
                           "d2h_bytes_per_step": int(N_CAND * 5), "grasps_per_s": world * N_CAND * args.steps / e2e_s},
                   "gpu_launches": int(launches),
                   "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                "traffic": traffic.get(key, {}).get("gbs"), "traffic_bytes_per_launch": traffic.get(key, {}).get("bytes_per_launch"),
                                "algorithmic_bytes_per_launch": bs * per_rank_steps / args.steps,
                                "traffic_source": traffic.get(key, {}).get("source"),
                                "kernel": "mgs_rollout_kernel", "bytes_per_env_step": bs,
                                "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                                "note": "state stays in shared memory for the whole rollout; the kernel is issue/latency bound, not HBM bound"},
                   "clocks": clk.summary(), "wall_s": t_wall}
            # secondary roofline: warp-instruction issue slots (the bound that actually applies).  Instructions per env-step are
            # the ncu count of the steady hold phase (profiles/ncu_r1_m_*_steady.txt: smsp__inst_executed.sum / env-steps);
            # peak = 148 SMs x 4 schedulers x 1 warp-instruction per cycle at the SM clock sampled during the run
            ipe = {"panda": 26.1e3, "robotiq": 40.4e3}.get(key)
            mhz = out["clocks"].get("sm_mhz") or 1965.0
            if ipe:
                peak_issue = 148 * 4 * mhz * 1e6
                out["issue_slots"] = {"warp_instr_per_env_step": ipe, "source": "ncu r1_m steady capture", "achieved_ginst_s": ipe * value / world / 1e9,
                                      "peak_ginst_s": peak_issue / 1e9, "frac": ipe * value / world / peak_issue}
            if with_cpu:
                from oracle import oracle as orc
                orc.build()
                threads = os.cpu_count() or 1
                v, g, n, dt = cpu_arm(model, info, pose7, joints, 12.0, threads)
                out["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port", "grasps_per_s": g,
                                       "sample": f"oracle port (fp64 C restatement; MuJoCo not installable offline), first {n} of {N_CAND} candidates, {dt:.1f}s"}
        sim.close()
        return out

    with_cpu = world == 1 and not args.no_cpu
    line = measure(args.workload, with_cpu)
    # secondary workload in the same line: Panda on convex objects, the gripper the north-star's target sentence names
    also = measure("panda", with_cpu) if (args.workload != "panda" and not args.no_also) else None
    if rank == 0:
        if also is not None:
            line["also"] = {"panda_on_convex": {k: also[k] for k in ("value", "unit", "ms_per_step", "grasps_per_s", "e2e", "config", "roofline", "issue_slots", "gpu_launches")
                                                 + (("cpu_baseline",) if "cpu_baseline" in also else ())}}
            line["gpu_launches"] += also["gpu_launches"]
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="robotiq", choices=sorted(WORKLOADS), help="robotiq = BASELINE.json configs[1] (default)")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary Panda-on-convex measurement")
    ap.add_argument("--caps", default="", help="override the primary workload's per-environment capacities: ncon_max,nefc_max")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
