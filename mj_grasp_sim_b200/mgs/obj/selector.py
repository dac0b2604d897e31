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
        return ObjectConvexHull(pose, name=f"hull{seed}", hull_points=[pts], weight=mass)
    raise ValueError(f"object '{object_id}' is not available offline (YCB/GSO assets are not shipped); use 'cube' or 'hull:<seed>'")
