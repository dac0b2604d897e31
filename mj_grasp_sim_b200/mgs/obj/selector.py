"""get_object(object_id) (/root/reference/mgs/obj/selector.py:33-52).  YCB/GSO ids cannot be served
offline; "cube" and "hull:<seed>[:<n_vertices>]" select the synthetic objects of SURVEY.md 8(d)."""
import numpy as np

from ...scenes import random_hull_points
from ..util.geo.transforms import SE3Pose
from .cube import ObjectCube
from .hull import ObjectConvexHull


def get_object(object_id: str):
    pose = SE3Pose(np.array([0, 0, 0]), np.array([1, 0, 0, 0]), type="wxyz")
    if object_id in ("cube", "Cube"):
        return ObjectCube(pose, name="cube", size=0.02)
    if object_id.startswith("hull:"):
        parts = object_id.split(":")
        seed, n_v = int(parts[1]), int(parts[2]) if len(parts) > 2 else 32
        pts, mass = random_hull_points(seed, n_v)
        o = ObjectConvexHull(pose, name=f"hull{seed}", hull_points=[pts], weight=mass)
        o.object_id = object_id  # the id that selects it again (file layout: <gripper>/<object id>/stable_grasps.npz)
        return o
    raise ValueError(f"object '{object_id}' is not available offline (YCB/GSO assets are not shipped); use 'cube' or 'hull:<seed>'")


def get_objects(object_ids):
    """get_objects (/root/reference/mgs/obj/selector.py:54-246 picks random dataset objects and parks them on a grid);
    here the caller names the synthetic objects, e.g. ["hull:3", "hull:7:24"].  Names are made unique per scene."""
    objs = []
    for i, oid in enumerate(object_ids):
        o = get_object(oid)
        o.name = f"{o.name}_{i}"
        # parked away from the workspace until gen_clutter drops them (obj/selector.py:207-246 uses the same grid)
        o.pos = np.array([-8.0 + 0.5 * (i // 10), -8.0 + 0.5 * (i % 10), 0.06])
        objs.append(o)
    return objs
