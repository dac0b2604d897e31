"""ObjectCube: box primitive, mass 1.0, default contact parameters (/root/reference/mgs/obj/cube.py:22-56)."""
from typing import Any, Dict, Tuple

from ..util.geo.transforms import SE3Pose
from .base import CollisionMeshObject

_TEMPLATE = """
    <worldbody>
        <body name="{name}" pos="{position}" quat="{quaternion}">
            <freejoint name="{name}:joint"/>
            <geom name="geom:{name}" rgba="1.0 0.32 0.32 1" size="{size} {size} {size}" type="box" mass="1.0"/>
        </body>
    </worldbody>
"""


class ObjectCube(CollisionMeshObject):
    def __init__(self, pose: SE3Pose, name: str, size: float):
        vec = pose.to_vec(layout="pq", type="wxyz")
        self.pos, self.quat, self.name, self.size, self.object_id = vec[:3], vec[3:], name, size, name

    def mesh(self):
        """(verts, triangles) of the box - what a candidate sampler is given"""
        from ...compiler import mesh as meshlib
        h = meshlib.box_hull([self.size, self.size, self.size])
        return h.verts, h.tri

    def to_xml(self) -> Tuple[str, Dict[str, Any]]:
        return _TEMPLATE.format(position="{} {} {}".format(*self.pos), quaternion="{} {} {} {}".format(*self.quat),
                                name=self.name, size=self.size), {}
