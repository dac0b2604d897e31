"""Summarise an .ncu-rep of mgs_rollout_kernel: headline metrics + per-function instruction/stall shares.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [libmgs_b200.so]  (needs ncu, cuobjdump, nvdisasm; no GPU)
       MGS_VARIANT=w16|w12|wide selects the kernel variant whose SASS is joined (default w16)
       MGS_ISSUE_JSON=profiles/issue_slots_r2.json MGS_ISSUE_KEY=robotiq MGS_ENV_STEPS=<env-steps of the captured launch>
         additionally records warp-instructions per env-step (smsp__inst_executed.sum / env-steps) under that key: bench.py's
         issue_slots block reads the file, so the figure is derived by this script from the committed capture, not typed in.
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mj_grasp_sim_b200", "libmgs_b200.so")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__stack_size"]
print("== headline")
for h, u, v in zip(hdr, units, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.05):
        print(f"{h} [{u}] {v}")
if os.environ.get("MGS_ISSUE_JSON"):
    import json
    path, key, steps = os.environ["MGS_ISSUE_JSON"], os.environ["MGS_ISSUE_KEY"], float(os.environ["MGS_ENV_STEPS"])
    d = {}
    try:
        d = json.load(open(path))
    except Exception:
        pass
    m = dict(zip(hdr, vals))
    inst = float(m["smsp__inst_executed.sum"].replace(",", ""))
    d[key] = {"warp_instr_per_env_step": inst / steps, "env_steps_of_capture": steps, "smsp__inst_executed.sum": inst,
              "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
              "thread_inst_per_inst": float(m["smsp__thread_inst_executed_per_inst_executed.ratio"]),
              "source": f"{os.path.basename(rep)}: steady hold-phase launch (tools/profile_steady.py), summarised by tools/ncu_summary.py"}
    json.dump(d, open(path, "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
shdr = srows[1]
ci, si, ti = shdr.index("Instructions Executed"), shdr.index("# Samples"), shdr.index("Thread Instructions Executed")
stall_cols = {k: shdr.index(k) for k in ("stall_long_sb", "stall_wait", "stall_barrier", "stall_short_sb", "stall_branch_resolving", "stall_no_inst", "stall_math", "stall_lg", "stall_mio")}
asp_i, op_i = shdr.index("Address Space"), shdr.index("Access Operation")
data = srows[2:]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
    # one cubin per kernel variant (translation unit): MGS_VARIANT=w16 (default) | w12 selects the profiled one
    tag = "mgs_rollout_kernel_" + os.environ.get("MGS_VARIANT", "w16")
    dis = ""
    for cub in sorted(f for f in os.listdir(td) if f.endswith(".cubin")):
        d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
        if tag in d or "mgs_rollout_kernelv" in d:
            dis = d
            break
func, line, seq = None, None, []
for l in dis.split("\n"):
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"^(\$?[_A-Za-z][\w$]*):\s*$", l)
    if m and not m.group(1).startswith(".L"):
        func = m.group(1)
    if re.match(r"^\s+/\*[0-9a-f]{4,5}\*/", l):
        seq.append((func, line))
if len(seq) != len(data):
    print(f"WARNING: SASS length mismatch ({len(seq)} vs {len(data)}): the .so does not match the profiled build")
byf = collections.defaultdict(lambda: [0, 0, 0])
byl = collections.defaultdict(lambda: [0, 0, 0])
stf = collections.defaultdict(lambda: collections.Counter())
stl = collections.defaultdict(lambda: collections.Counter())
memf = collections.defaultdict(lambda: collections.Counter())
for (f, ln), r in zip(seq, data):
    n, s, t = int(r[ci]), int(r[si]), int(r[ti])
    for d, k in ((byf, f), (byl, ln)):
        d[k][0] += n; d[k][1] += s; d[k][2] += t
    for k, i in stall_cols.items():
        v = int(r[i] or 0)
        stf[f][k] += v; stl[ln][k] += v
    if r[asp_i]:
        memf[f][r[asp_i] + ":" + r[op_i]] += n
tot, tots = sum(v[0] for v in byf.values()), sum(v[1] for v in byf.values())
print(f"== by function (total warp-instructions {tot}, samples {tots}): %inst %samples lane-efficiency")
for k, v in sorted(byf.items(), key=lambda kv: -kv[1][1])[:28]:
    name = re.sub(r"^\$?_Z\d+mgs_rollout_kernel\w*?v\$|_ZN\d+_INTERNAL_[0-9a-f]{8}_\d+_\w+?_cu_[0-9a-f]{8}\d+", "", str(k))
    print(f"{100*v[0]/tot:6.2f} {100*v[1]/tots:6.2f} {v[2]/max(1,v[0])/32:5.2f}  {name}")
print("== top source lines: %inst %samples lane-efficiency")
for k, v in sorted(byl.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{100*v[0]/tot:6.2f} {100*v[1]/tots:6.2f} {v[2]/max(1,v[0])/32:5.2f}  {k}")

print("== stall mix by function (% of all samples): " + " ".join(k.replace("stall_", "") for k in stall_cols))
for k, v in sorted(byf.items(), key=lambda kv: -kv[1][1])[:28]:
    name = re.sub(r"^\$?_Z\d+mgs_rollout_kernel\w*?v\$|_ZN\d+_INTERNAL_[0-9a-f]{8}_\d+_\w+?_cu_[0-9a-f]{8}\d+", "", str(k))
    print(" ".join(f"{100*stf[k][c]/tots:6.2f}" for c in stall_cols) + "  " + name)
print("== stall mix, top source lines")
for k, v in sorted(byl.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{100*v[1]/tots:6.2f} | " + " ".join(f"{100*stl[k][c]/tots:6.2f}" for c in stall_cols) + f"  {k}")
print("== memory instructions by function (% of all warp-instructions): space:op")
for k, v in sorted(byf.items(), key=lambda kv: -kv[1][1])[:28]:
    name = re.sub(r"^\$?_Z\d+mgs_rollout_kernel\w*?v\$|_ZN\d+_INTERNAL_[0-9a-f]{8}_\d+_\w+?_cu_[0-9a-f]{8}\d+", "", str(k))
    print("  " + name + "  " + "  ".join(f"{a} {100*c/tot:.2f}" for a, c in sorted(memf[k].items(), key=lambda ac: -ac[1])))
