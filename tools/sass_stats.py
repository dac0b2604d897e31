"""Static SASS statistics per device function of libmgs_b200.so: instructions, local loads/stores (LDL/STL).
usage: python tools/sass_stats.py [path/to/lib.so]   (needs cuobjdump + nvdisasm; no GPU)"""
import collections, os, re, subprocess, sys, tempfile
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mj_grasp_sim_b200", "libmgs_b200.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
    # one cubin per kernel variant (translation unit): MGS_VARIANT=w16 (default) | w12
    tag = "mgs_rollout_kernel_" + os.environ.get("MGS_VARIANT", "w16")
    dis = ""
    for cub in sorted(f for f in os.listdir(td) if f.endswith(".cubin")):
        d = subprocess.run(["nvdisasm", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
        if tag in d or "mgs_rollout_kernelv" in d:
            dis = d
            break
func, st = None, collections.defaultdict(lambda: [0, 0, 0])
for l in dis.split("\n"):
    m = re.match(r"^(\$?[_A-Za-z][\w$]*):\s*$", l)
    if m and not m.group(1).startswith(".L"):
        func = m.group(1)
    if re.match(r"^\s+/\*[0-9a-f]{4,5}\*/", l):
        st[func][0] += 1
        st[func][1] += " LDL" in l
        st[func][2] += " STL" in l
tot = [sum(v[i] for v in st.values()) for i in range(3)]
print(f"total: {tot[0]} instructions ({tot[0] * 16 / 1024:.0f} KB), {tot[1]} LDL, {tot[2]} STL")
for k, v in sorted(st.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]:6d} {v[1]:4d} {v[2]:4d}  " + re.sub(r"^\$?_Z\d+mgs_rollout_kernel\w*?v\$|_ZN\d+_INTERNAL_[0-9a-f]{8}_\d+_\w+?_cu_[0-9a-f]{8}\d+", "", str(k)))
