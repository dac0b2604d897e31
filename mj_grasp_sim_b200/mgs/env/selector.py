"""get_env (/root/reference/mgs/env/selector.py:23-40): only the zero-gravity single-object env is built so far."""
from .gravityless_object_grasping import GravitylessObjectGrasping


def get_env(cfg, gripper, obj):
    name = cfg if isinstance(cfg, str) else cfg.name
    if name in ("GravitylessObjectGrasping", "gravityless_object_grasping", "gravityless"):
        return GravitylessObjectGrasping(gripper, obj)
    raise NotImplementedError(f"environment '{name}' is not built yet (clutter_table is a later SURVEY 8 row)")
