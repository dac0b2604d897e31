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
