"""DRAM traffic of one full rollout launch per bench workload -> profiles/traffic_r1.json (read by bench.py's roofline.traffic).

Run on the GPU box, one workload at a time (ncu replays the kernel once per metric pass):
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mgs_rollout \\
      --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/traffic_<workload>.csv \\
      python bench.py --workload <workload> --no-also --no-cpu --steps 1 --warmup 3
  python tools/ncu_traffic.py gpurun_out/traffic_robotiq.csv:robotiq gpurun_out/traffic_panda.csv:panda
"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for arg in sys.argv[1:]:
    path, key = arg.split(":")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    mi, vi, ui = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    vals = {}
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "second": 1}.get(u, 1)
        vals[r[mi]] = v * scale
    b = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
    t = vals["gpu__time_duration.sum"]
    out[key] = {"bytes_per_launch": b, "seconds_under_ncu": t, "gbs": b / t / 1e9, "read_bytes": vals["dram__bytes_read.sum"], "write_bytes": vals["dram__bytes_write.sum"],
                "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of one full 4096-candidate rollout launch ({os.path.basename(path)})"}
json.dump(out, open(os.path.join(ROOT, "profiles", os.environ.get("MGS_TRAFFIC_OUT", "traffic_r2.json")), "w"), indent=1)
print(json.dumps(out, indent=1))
