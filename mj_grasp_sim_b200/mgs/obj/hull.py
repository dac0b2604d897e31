"""ObjectConvexHull: synthetic stand-in for ObjectYCB / ObjectGSO (the datasets are not available
offline).  One or more convex sub-meshes emitted with the reference's object recipe
(/root/reference/mgs/obj/ycb.py:70-160): per-hull mesh geoms with mass = weight * proportion, condim 4,
friction 1.0/0.3/0.1, solimp .998 .998 .001, solref .001 1, free joint with damping 1e-4, delivered as
an <include> whose file travels in the asset dict (ycb.py:60-68)."""
from typing import Any, Dict, List, Tuple

import numpy as np

from ...compiler import mesh as meshlib
from ..util.geo.transforms import SE3Pose
from .base import CollisionMeshObject


class ObjectConvexHull(CollisionMeshObject):
    def __init__(self, pose: SE3Pose, name: str, hull_points: List[np.ndarray], weight: float):
        vec = pose.to_vec(layout="pq", type="wxyz")
        self.pos, self.quat, self.name, self.object_id = vec[:3], vec[3:], name, name
        self.hulls = [meshlib.build_hull(np.asarray(p, dtype=np.float64)) for p in hull_points]
        vols = np.array([meshlib.mass_properties(h.verts, h.tri)[0] for h in self.hulls])
        self.props = vols / vols.sum()
        self.weight = float(weight)

    def mesh(self):
        """(verts, triangles) of the union of hulls - what a candidate sampler would be given."""
        verts, tris, off = [], [], 0
        for h in self.hulls:
            verts.append(h.verts); tris.append(h.tri + off); off += len(h.verts)
        return np.concatenate(verts), np.concatenate(tris)

    def to_xml(self) -> Tuple[str, Dict[str, Any]]:
        assets: Dict[str, Any] = {}
        meshes, geoms = [], []
        for i, (h, prop) in enumerate(zip(self.hulls, self.props)):
            fn = f"{self.name}_coll_{i}.obj"
            assets[fn] = meshlib.write_obj(h.verts, h.tri)
            meshes.append(f'<mesh name="{self.name}_coll_{i}" file="{fn}"/>')
            geoms.append(f'<geom mesh="{self.name}_coll_{i}" mass="{self.weight * prop}" group="3" type="mesh" conaffinity="1" '
                         f'contype="1" condim="4" rgba="1 1 1 1" friction="1.0 0.3 0.1" solimp="0.998 0.998 0.001" solref="0.001 1"/>')
        doc = (f'<mujoco model="{self.name}"><asset>{"".join(meshes)}</asset><worldbody>'
               f'<body name="{self.name}" pos="{" ".join(map(str, self.pos))}" quat="{" ".join(map(str, self.quat))}">'
               f'{"".join(geoms)}<joint damping="0.0001" name="{self.name}:joint" type="free"/></body></worldbody></mujoco>')
        inc = f"{self.name}_model.xml"
        assets[inc] = doc.encode()
        return f'<include file="{inc}" />', assets
