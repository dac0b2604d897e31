"""CollisionMeshObject base (/root/reference/mgs/obj/base.py:22-35)."""


class CollisionMeshObject:
    name: str
    object_id: str
