"""ASSET_PATH: where gripper templates and (pre-hulled) meshes live.  Mirrors
/root/reference/mgs/util/const.py:27-29, pointing at this package's asset pack."""
import os

PACKAGE_PATH = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GIT_PATH = os.path.dirname(os.path.dirname(PACKAGE_PATH))
ASSET_PATH = os.path.join(os.path.dirname(PACKAGE_PATH), "assets")
