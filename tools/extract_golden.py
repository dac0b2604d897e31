"""Extract the reference's only MuJoCo-produced numeric artefact on the hot path into tests/golden/.

`/root/reference/mgs/cli/config/gripper/robotiq_2f_85.yaml:11` (`state_close`) is an
mjSTATE_INTEGRATION vector (203 doubles) captured from a GripperScanEnv run of the real MuJoCo 3.2.2:
a lone Robotiq 2F-85 in zero gravity, ctrl = 255, mocap at (0, 0, -0.15), t = 223 s (steady state).
Run: python tools/extract_golden.py [/root/reference]
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
txt = open(os.path.join(ref, "mgs/cli/config/gripper/robotiq_2f_85.yaml")).read()
m = re.search(r"^state_close:\s*\[(.*?)\]", txt, re.S | re.M)
vals = [float(x) for x in m.group(1).replace("\n", " ").split(",")]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "robotiq_2f85_state_close.json")
json.dump({"source": "mgs/cli/config/gripper/robotiq_2f_85.yaml:11 (state_close), MuJoCo 3.2.2 mjSTATE_INTEGRATION",
           "layout": "time, qpos[22], qvel[20], act[0], qacc_warmstart[20], ctrl[1], qfrc_applied[20], xfrc_applied[6*18], eq_active[4], mocap_pos[3], mocap_quat[4]",
           "state": vals}, open(out, "w"), indent=0)
print(len(vals), "values ->", out)

# ---- round 2: every other reference-held artefact that constrains the compiled models -------------------------------------
# (a) the per-gripper `segmentation:` geom-id lists (mgs/cli/config/gripper/{allegro,leap,panda,shadow,vx300}.yaml): MuJoCo geom
#     ids of the geoms that move with each joint, written by the reference's author from MuJoCo segmentation renders;
# (b) segments.txt: the (geom id -> geom name) table MuJoCo printed for the LEAP scan scene (lines 1-88) and the ids of the
#     rendered, unnamed geoms of the Shadow scan scene (the block that follows);
# (c) the second recorded Robotiq closed posture (robotiq_2f_85.yaml:7, commented out): the 8 actuator-joint angles.
import yaml

seg = {}
for key in ("allegro", "leap", "panda", "shadow", "vx300"):
    y = yaml.safe_load(open(os.path.join(ref, f"mgs/cli/config/gripper/{key}.yaml")))
    seg[key] = {k: [int(i) for i in v] for k, v in y["segmentation"].items()}
lines = open(os.path.join(ref, "segments.txt")).read().splitlines()
named, unnamed, block = [], [], 0
for ln in lines:
    mm = re.match(r"Segment ID (\d+) \(geom\) -> Model name: (.*)$", ln)
    if not mm:
        continue
    gid, name = int(mm.group(1)), mm.group(2).strip()
    if name:
        named.append([gid, name])
    else:
        unnamed.append(gid)
m2 = re.search(r"^# qpos:\s*\[(.*?)\]\s*# close", txt, re.M)
close8 = [float(x) for x in m2.group(1).split(",")]
out2 = os.path.join(os.path.dirname(out), "reference_model_pins.json")
json.dump({"source": "mgs/cli/config/gripper/*.yaml `segmentation:`; segments.txt; robotiq_2f_85.yaml:7",
           "segmentation": seg, "leap_geom_names": named, "shadow_rendered_unnamed_geoms": unnamed,
           "robotiq_close_qpos8": close8}, open(out2, "w"), indent=0)
print(len(named), "named geoms,", len(unnamed), "unnamed,", sum(len(v) for v in seg.values()), "segmentation keys ->", out2)
