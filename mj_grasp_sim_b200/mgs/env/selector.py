"""get_env / get_env_from_dict (/root/reference/mgs/env/selector.py:23-40)."""
from .clutter_table import ClutterTableEnv
from .gravityless_object_grasping import GravitylessObjectGrasping


def get_env(cfg, gripper, objects):
    name = cfg if isinstance(cfg, str) else cfg.name
    if name in ("GravitylessObjectGrasping", "gravityless_object_grasping", "gravityless"):
        return GravitylessObjectGrasping(gripper, objects if not isinstance(objects, (list, tuple)) else objects[0])
    if name in ("ClutterTable", "clutter_table", "ClutterTableEnv"):
        return ClutterTableEnv(gripper, list(objects))
    raise NotImplementedError(f"environment '{name}' is not built (bin_picking is stale in the reference, SURVEY 2 row 3b)")


def get_env_from_dict(cfg, state_dict):
    name = cfg if isinstance(cfg, str) else cfg.name
    if name in ("ClutterTable", "clutter_table", "ClutterTableEnv"):
        return ClutterTableEnv.from_dict(state_dict)
    raise NotImplementedError(name)
