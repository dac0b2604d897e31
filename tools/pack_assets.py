"""Build the in-repo asset pack from the read-only reference checkout.

The GPU box has no /root/reference, so everything the six in-scope grippers need at run time is
regenerated here into `mj_grasp_sim_b200/assets/<gripper>/`:

* `template.xml`  - the MJCF template string of `/root/reference/mgs/gripper/<gripper>.py`
  (MuJoCo-Menagerie-derived model data, Apache-2.0/BSD, see the reference's
  3rd-party-licenses.txt), extracted textually because the module cannot be imported
  (it imports `mujoco`).
* one file per mesh in `/root/reference/asset/<gripper>/`, same basename, holding only the
  convex hull (MuJoCo collides hulls, so collision geometry is unchanged; ~100x smaller).
* `massprops.json` - volume / CoM / second moments of the ORIGINAL meshes, used by the model
  compiler where a body has no <inertial> (Allegro, Robotiq base_mount).

Run: python tools/pack_assets.py [/root/reference]
"""
import json
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_grasp_sim_b200.compiler import mesh as M  # noqa: E402

GRIPPERS = {"panda": "panda", "vx300": "vx300", "robotiq2f85": "robotiq2f85",
            "allegro": "allegro", "leap": "leap", "shadow": "shadow"}


def main(ref="/root/reference"):
    out_root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                            "mj_grasp_sim_b200", "assets")
    for mod, adir in GRIPPERS.items():
        src = open(os.path.join(ref, "mgs", "gripper", mod + ".py")).read()
        m = re.search(r'^XML = r?"""(.*?)"""', src, re.S | re.M)
        out = os.path.join(out_root, adir)
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "template.xml"), "w") as f:
            f.write(m.group(1))
        props = {}
        for root, _, files in os.walk(os.path.join(ref, "asset", adir)):
            for fn in sorted(files):
                if not fn.lower().endswith((".stl", ".obj")):
                    continue
                data = open(os.path.join(root, fn), "rb").read()
                if len(data) < 200:  # missing LFS blob
                    print("skip (blob missing)", fn)
                    continue
                v, f = M.load_mesh(fn, data)
                V, com, C = M.mass_properties(v, f)
                props[fn] = {"volume": V, "com": com.tolist(), "cov": C.tolist()}
                h = M.build_hull(v)
                blob = M.write_stl(h.verts, h.tri) if fn.lower().endswith(".stl") else M.write_obj(h.verts, h.tri)
                with open(os.path.join(out, fn), "wb") as g:
                    g.write(blob)
                print(f"{adir}/{fn}: {len(v)} verts -> hull {len(h.verts)} verts, {len(h.face_num)} faces, {len(blob)} B")
        with open(os.path.join(out, "massprops.json"), "w") as g:
            json.dump(props, g, indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
