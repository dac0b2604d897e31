"""Scene assembly + synthetic workloads (objects and grasp candidates) for tests and bench.py.

Scene template: same options and body layout as the reference's zero-gravity grasping scene
(`/root/reference/mgs/env/gravityless_object_grasping.py:34-54`): gripper fragment, then the
ground body with `geom:ground`, then the object fragment - the geom-id ordering the contact
labels rely on (`:309-321`).  Synthetic inputs follow SURVEY.md 8(d): YCB/GSO meshes are not
available offline, so objects are a box primitive or seeded random convex hulls emitted with the
reference's object recipe (`/root/reference/mgs/obj/ycb.py:138-158`).
"""
from __future__ import annotations

import os

import numpy as np
from scipy.spatial import ConvexHull
from scipy.spatial.transform import Rotation as R

from .compiler import mesh as meshlib
from .compiler.mjcf import compile_mjcf

ASSET_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

GRAVITYLESS_XML = """<mujoco>
  <compiler angle="radian" autolimits="true" discardvisual="false"/>
  <option integrator="implicitfast" timestep="0.001" cone="elliptic" impratio="3" noslip_iterations="2"
          noslip_tolerance="1e-8" tolerance="1e-8" gravity="0 0 0"><flag multiccd="enable"/></option>
  {gripper}
  <worldbody>
    <body name="body:ground" pos="0.0 0 -1.0">
      <geom name="geom:ground" pos="0 0 0" size="1.0 1.0 0.02" type="box" density="500"/>
    </body>
  </worldbody>
  {object}
</mujoco>"""

CUBE_XML = """<worldbody><body name="{name}" pos="0 0 0" quat="1 0 0 0"><freejoint name="{name}:joint"/>
<geom name="geom:{name}" size="{size} {size} {size}" type="box" mass="1.0"/></body></worldbody>"""

# gripper name -> (asset dir, base free joint, actuated joints, close ctrl, b2c pos, b2c quat wxyz, repose_on_close)
GRIPPERS = {
    "panda": dict(dir="panda", freejoint="freejoint", joints=["finger_joint1", "finger_joint2"], close_ctrl=[0.0, -0.04],
                  b2c_pos=[0, 0, -0.102], b2c_quat=[0.707106781, 0.0, 0.0, 0.707106781], repose=0),
    # the reference lists two misnamed joints ("right_spring_link", "left_spring_link": robotiq2f85.py:275,279);
    # mj_name2id -> -1 -> jnt_qposadr[-1] = the LAST joint (the object's free joint): reproduced via None
    "robotiq2f85": dict(dir="robotiq2f85", freejoint="freejoint",
                        joints=["right_driver_joint", "right_coupler_joint", None, "right_follower_joint",
                                "left_driver_joint", "left_coupler_joint", None, "left_follower_joint"],
                        close_ctrl=[255.0], b2c_pos=[0, 0, -0.15], b2c_quat=[1.0, 0.0, 0.0, 0.0], repose=0),
    # b2c = Rz(90 deg) * Ry(-90 deg), offset (0,0,-0.12)  (vx300.py:242-257)
    "vx300": dict(dir="vx300", freejoint="freejoint", joints=["left_finger", "right_finger"], close_ctrl=[0.021, -0.021],
                  b2c_pos=[0, 0, -0.12], b2c_quat=[0.5, 0.5, -0.5, 0.5], repose=0),
    # allegro.py:300-347 (open/close poses, b2c = Ry(-90 deg) with offset Ry(-90 deg) * (-0.08, 0, 0.01))
    "allegro": dict(dir="allegro", freejoint="freejoint", joints=[f"{f}j{k}" for f in ("ff", "mf", "rf", "th") for k in range(4)],
                    close_ctrl=[-0.08, 0.95, 1, 0.95, 0, 0.95, 1.2, 0.85, 0.08, 0.95, 1.2, 0.9, 1.4, 0.55, 0.29, 1.45],
                    open_pose=[-0.08, 0.715, 0.710, 0.95, 0, 0.8, 0.71, 0.67, 0.08, 0.715, 0.710, 0.95, 1.4, 0.55, -0.19, 1.45],
                    b2c_pos=[-0.01, 0.0, -0.08], b2c_quat=[0.70710678, 0.0, -0.70710678, 0.0], repose=1),
    # leap.py:373-398; pre-grasp joints of the reference's contact sampler (mgs/sampler/kin/leap.py:462-484)
    "leap": dict(dir="leap", freejoint="freejoint",
                 joints=[f"{f}_{j}" for f, js in (("if", ("mcp", "rot", "pip", "dip")), ("mf", ("mcp", "rot", "pip", "dip")),
                                                  ("rf", ("mcp", "rot", "pip", "dip")), ("th", ("cmc", "axl", "mcp", "ipl"))) for j in js],
                 close_ctrl=[0.576, 0.0, 1.43, 0.453, 0.856, 0.0, 0.68, 0.826, 0.945, 0.0, 1.3, 0.2, 1.81, 0.258, 0.505, 0.351],
                 open_pose=[0.785, 0, 0, 0] * 4, b2c_pos=[0, 0, 0], b2c_quat=[1.0, 0.0, 0.0, 0.0], repose=1),
    # shadow.py:368-455; close ctrl = _qpos_to_qacc(22-vector) -> 18 actuators; pre-grasp joints of the reference's
    # contact sampler (mgs/sampler/kin/shadow.py:200-227)
    "shadow": dict(dir="shadow", freejoint="freejoint",
                   joints=["rh_FFJ4", "rh_FFJ3", "rh_FFJ2", "rh_FFJ1", "rh_MFJ4", "rh_MFJ3", "rh_MFJ2", "rh_MFJ1", "rh_RFJ4", "rh_RFJ3",
                           "rh_RFJ2", "rh_RFJ1", "rh_LFJ5", "rh_LFJ4", "rh_LFJ3", "rh_LFJ2", "rh_LFJ1", "rh_THJ5", "rh_THJ4", "rh_THJ3",
                           "rh_THJ2", "rh_THJ1"],
                   close_ctrl=[0.07708, 1.21, 0.2023, 0.6614, 0.0102, -0.3464, 1.253, 0.782494, 0.01103, 1.475, 0.6336, -0.2083, 1.45,
                               0.75, 0.13, -0.4, 1.5, 1.3],
                   open_pose=[-0.350, 0.425, 0.015, 0.005, -0.095, 0.415, 0.010, 0.0, -0.075, 0.435, 0.015, 0.005, 0.0, -0.220, 0.255,
                              0.0, 0.0, -0.480, 1.05, -0.19, -0.080, 0.45],
                   b2c_pos=[0, 0, 0], b2c_quat=[1.0, 0.0, 0.0, 0.0], repose=0),
}


# precision policy of the product path per gripper (mgs/gripper/base.py COMPUTE_F64)
F64_GRIPPERS = ("allegro", "leap", "shadow")


def gripper_fragment(name: str):
    g = GRIPPERS[name]
    d = os.path.join(ASSET_PATH, g["dir"])
    xml = open(os.path.join(d, "template.xml")).read().format(position="0 0 0", quaternion="1 0 0 0")
    assets = {f: open(os.path.join(d, f), "rb").read() for f in os.listdir(d) if f != "template.xml"}
    return xml, assets


def cube_fragment(name="cube", size=0.02):
    return CUBE_XML.format(name=name, size=size), {}


def random_hull_points(seed: int, n_v: int = 32):
    """n_v points on an ellipsoid with semi-axes ~U(0.015, 0.05) m."""
    rng = np.random.default_rng(1000 + seed)
    ax = rng.uniform(0.015, 0.05, size=3)
    p = rng.normal(size=(n_v, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    mass = rng.uniform(0.05, 0.5)
    return p * ax, mass


def hull_object_fragment(seed: int, n_v: int = 32, name="obj", mass_scale: float = 1.0):
    """Convex-hull object with the reference's object recipe: condim 4, friction 1/.3/.1,
    solimp .998 .998 .001, solref .001 1, free joint damping 1e-4 (ycb.py:138-157)."""
    pts, mass = random_hull_points(seed, n_v)
    mass *= mass_scale
    h = meshlib.build_hull(pts)
    fn = f"{name}_hull_{seed}.obj"
    xml = f"""<asset><mesh name="{name}_coll_0" file="{fn}"/></asset>
<worldbody><body name="{name}" pos="0 0 0" quat="1 0 0 0">
<geom mesh="{name}_coll_0" mass="{mass}" group="3" type="mesh" conaffinity="1" contype="1" condim="4"
      friction="1.0 0.3 0.1" solimp="0.998 0.998 0.001" solref="0.001 1"/>
<joint damping="0.0001" name="{name}:joint" type="free"/></body></worldbody>"""
    return xml, {fn: meshlib.write_obj(h.verts, h.tri)}, (h.verts, h.tri)


def build_scene(gripper: str, obj_xml: str, obj_assets: dict):
    gx, ga = gripper_fragment(gripper)
    model = compile_mjcf(GRAVITYLESS_XML.format(gripper=gx, object=obj_xml), {**ga, **obj_assets})
    g = GRIPPERS[gripper]
    info = dict(base_qposadr=int(model.jnt_qposadr[model.names["joint"][g["freejoint"]]]),
                joint_qposadr=np.array([model.jnt_qposadr[model.names["joint"][j] if j is not None else -1] for j in g["joints"]], dtype=np.int32),
                close_ctrl=np.array(g["close_ctrl"], dtype=np.float64), repose=g["repose"], gripper=gripper)
    return model, info


# ---- SE3Pose-compatible pose preprocessing (fp32, like the reference's SE3Pose.__matmul__) -----
def pose_to_mat32(pos, quat_wxyz):
    T = np.zeros((4, 4), dtype=np.float32)
    T[:3, :3] = R.from_quat([quat_wxyz[1], quat_wxyz[2], quat_wxyz[3], quat_wxyz[0]]).as_matrix()
    T[:3, 3] = pos
    T[3, 3] = 1
    return T


def process_poses(H: np.ndarray, gripper: str) -> np.ndarray:
    """candidate 4x4 matrices (f64 [N,4,4]) -> base pose7 float32 [N,7] = SE3Pose.from_mat(H) @ b2c.

    Mirrors /root/reference/mgs/util/geo/transforms.py:78-128: from_mat -> float32 pos/quat,
    to_mat -> float32 4x4, einsum in float32, from_mat again (scipy), cast float32."""
    g = GRIPPERS[gripper]
    b2c = pose_to_mat32(np.asarray(g["b2c_pos"], dtype=np.float32), np.asarray(g["b2c_quat"], dtype=np.float32))
    H = np.asarray(H)
    q = R.from_matrix(H[:, :3, :3]).as_quat()  # xyzw
    q32 = q.astype(np.float32)
    pos32 = H[:, :3, 3].astype(np.float32)
    M = np.zeros((len(H), 4, 4), dtype=np.float32)
    M[:, :3, :3] = R.from_quat(q32).as_matrix()
    M[:, :3, 3] = pos32
    M[:, 3, 3] = 1
    P = np.einsum("nij,jk->nik", M, b2c)
    q2 = R.from_matrix(P[:, :3, :3]).as_quat()
    out = np.concatenate([P[:, :3, 3], q2[:, [3, 0, 1, 2]]], axis=1).astype(np.float32)
    return out


def panda_width_to_joints(width):
    """gen_grasp_candidates.py:62-64 with GripperPanda._clamp_width / width_to_joints (panda.py:217-223,264-266)."""
    w = np.clip(np.asarray(width) + 0.025, 0.003, 0.08)
    w = np.clip(w, 0.003, 0.08)
    return np.stack([np.clip(w / 2, 0, 0.04), np.clip(-0.04 + w / 2, -0.04, 0.0)], axis=-1)


def antipodal_candidates(verts, tri, n: int, seed: int):
    """Antipodal-style candidates on a convex mesh (frames as mgs/sampler/antipodal.py:181-298):
    x = p2 - p1 normalised, z random perpendicular to x, y = z cross x, origin = midpoint."""
    rng = np.random.default_rng(2000 + seed)
    T = verts[tri]
    area = 0.5 * np.linalg.norm(np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0]), axis=1)
    fn = np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0])
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    d0 = np.einsum("ij,ij->i", fn, T[:, 0])
    f = rng.choice(len(tri), size=n, p=area / area.sum())
    u, v = rng.uniform(size=n), rng.uniform(size=n)
    flip = u + v > 1
    u[flip], v[flip] = 1 - u[flip], 1 - v[flip]
    p1 = T[f, 0] + u[:, None] * (T[f, 1] - T[f, 0]) + v[:, None] * (T[f, 2] - T[f, 0])
    # inward direction: -normal perturbed (stand-in for the vMF draw, kappa=10)
    dirs = -fn[f] + rng.normal(size=(n, 3)) * 0.3
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    # exit point of the ray p1 + t*dir through the convex body: min positive t over faces
    denom = dirs @ fn.T  # [n, nf]
    num = d0[None, :] - p1 @ fn.T
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(denom > 1e-9, num / denom, np.inf)
    t_exit = t.min(axis=1)
    p2 = p1 + dirs * t_exit[:, None]
    # degenerate rays (start on an edge): the reference's fallback - a random second contact inside a
    # 10 cm cube around the first (antipodal.py:139-144)
    bad = ~np.isfinite(t_exit) | (t_exit < 1e-5)
    p2[bad] = p1[bad] + rng.uniform(-0.05, 0.05, size=(int(bad.sum()), 3))
    x = p2 - p1
    width = np.linalg.norm(x, axis=1)
    x /= width[:, None]
    rv = rng.normal(size=(n, 3))
    z = np.cross(x, rv)
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    y = np.cross(z, x)
    H = np.zeros((n, 4, 4))
    H[:, :3, 0], H[:, :3, 1], H[:, :3, 2], H[:, :3, 3], H[:, 3, 3] = x, y, z, 0.5 * (p1 + p2), 1.0
    return H, width


def box_mesh(size):
    h = meshlib.box_hull([size, size, size])
    return h.verts, h.tri


# "marginal" candidate sets (label-agreement measurements need an informative label distribution: on the plain antipodal
# sets 93-98 % of the parallel-jaw candidates are stable, so a constant predictor scores 95 %).  Per gripper: how far (m) the
# grasp frame is pulled back along its approach axis (uniform in [0.5, 1] x retreat: the fingertips barely reach the
# object), lateral noise sigma = retreat / 4, object 10 x heavier.  Tuned so that the oracle's stable fraction is 0.3-0.6.
MARGINAL = {"panda": 0.03, "vx300": 0.03, "robotiq2f85": 0.025, "allegro": 0.0, "leap": 0.0, "shadow": 0.06}


def workload(gripper: str, kind: str, seed: int, n: int, n_v: int = 32, marginal: bool = False):
    """(model, info, pose7 float32 [n,7], joints float32 [n,nj]) for one synthetic object."""
    if kind == "cube":
        ox, oa = cube_fragment()
        verts, tri = box_mesh(0.02)
    else:
        ox, oa, (verts, tri) = hull_object_fragment(seed, n_v, mass_scale=10.0 if marginal else 1.0)
    model, info = build_scene(gripper, ox, oa)
    H, width = antipodal_candidates(verts, tri, n, seed)
    if marginal and MARGINAL[gripper] > 0:
        rng = np.random.default_rng(6000 + seed)
        r = MARGINAL[gripper]
        H = H.copy()
        H[:, :3, 3] += rng.normal(scale=0.25 * r, size=(n, 3))
        H[:, :3, 3] -= H[:, :3, 2] * (r * rng.uniform(0.5, 1.0, size=(n, 1)))
    if gripper == "leap":
        # The reference's LEAP candidates come from its contact sampler (palm poses, identity b2c); that
        # sampler is out of scope, so the harness turns the antipodal frame into a palm pose: fingers point
        # along -z of the palm and close along its x axis around the point c (palm frame), hence
        # palm = frame * [Rx(180 deg), -Rx(180 deg) c].
        Rx = np.diag([1.0, -1.0, -1.0])
        T = np.eye(4)
        T[:3, :3] = Rx
        T[:3, 3] = -Rx @ np.array([0.0, -0.035, -0.09])
        H = H @ T
    if gripper == "shadow":
        # same idea for the Shadow wrist frame: fingers extend along +z, the palm faces -y and the thumb
        # opposes the fingers along z, closing around c = (0.01, -0.06, 0.12)
        Rt = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
        T = np.eye(4)
        T[:3, :3] = Rt
        T[:3, 3] = -Rt @ np.array([0.01, -0.06, 0.12])
        H = H @ T
    pose7 = process_poses(H, gripper)
    if gripper == "panda":
        joints = panda_width_to_joints(width)
    elif gripper == "vx300":
        w = np.clip(np.clip(width + 0.045, 0.003, 0.114), 0.042, 0.114)  # _clamp_width then width_to_joints (vx300.py:284-294,337-339)
        joints = np.stack([np.clip(0.5 * w, 0.021, 0.057), np.clip(-0.5 * w, -0.057, -0.021)], axis=-1)
    elif gripper in ("allegro", "leap", "shadow"):
        # pre-grasp posture + N(0, 0.05), clipped to the joint ranges
        rng = np.random.default_rng(3000 + seed)
        op = np.asarray(GRIPPERS[gripper]["open_pose"])
        joints = op[None] + rng.normal(scale=0.05, size=(n, len(op)))
        jid = [model.names["joint"][j] for j in GRIPPERS[gripper]["joints"]]
        joints = np.clip(joints, model.jnt_range[jid, 0], model.jnt_range[jid, 1])
    elif gripper == "robotiq2f85":
        joints = np.zeros((n, 8))  # open; the two misnamed columns then write 0 to the object's x - a no-op
    else:
        raise NotImplementedError(gripper)
    return model, info, pose7, joints.astype(np.float32)


# ---------------------------------------------------------------------------------------------------
# Clutter table (reference: /root/reference/mgs/env/clutter_table.py:41-79).  Same options, same body and
# geom order: gripper, table, camera (free joint, gravcomp 1, parked under the table), base_origin, four
# walls, then the objects - the contact tests compare geom ids with geom:table's.
CLUTTER_XML = """<mujoco>
  <compiler angle="radian" autolimits="true" discardvisual="false"/>
  <option integrator="implicitfast" timestep="0.001" cone="elliptic" gravity="0 0 -9.81" impratio="3" noslip_iterations="3"
          noslip_tolerance="1e-10" tolerance="1e-10"><flag multiccd="enable"/></option>
  {gripper}
  <worldbody>
    <body name="body:table" pos="0.0 0 -0.02">
      <geom name="geom:table" pos="0 0 0" size="10 10 0.02" type="box" density="500" friction="1.0 0.1 0.1"/>
    </body>
    <body name="body:camera" pos="0.0 0.0 -1.0" quat="1.0 0.0 0 0" gravcomp="1">
      <freejoint name="camera:joint"/>
      <geom name="geom:camera" size="0.01"/>
    </body>
    <body name="base_origin" pos="0.0 0.0 -0.025" quat="1.0 0.0 0 0"><geom name="geom:base_origin" size="0.01"/></body>
    <body name="body:wall_top" pos="0.0 1.0 0.1"><geom name="geom:wall_top" size="1.0 0.02 0.2" type="box" density="500"/></body>
    <body name="body:wall_right" pos="1.0 0.0 0.1"><geom name="geom:wall_right" size="0.02 1.0 0.2" type="box" density="500"/></body>
    <body name="body:wall_bottom" pos="0.0 -1.0 0.1"><geom name="geom:wall_bottom" size="1.0 0.02 0.2" type="box" density="500"/></body>
    <body name="body:wall_left" pos="-1.0 0.0 0.1"><geom name="geom:wall_left" size="0.02 1.0 0.2" type="box" density="500"/></body>
  </worldbody>
  {objects}
</mujoco>"""


def clutter_object_fragment(seed: int, index: int, n_v: int = 24):
    """Hull object `index` of a clutter scene, parked on the reference's far-away grid (obj/selector.py:207-246)."""
    pts, mass = random_hull_points(seed, n_v)
    h = meshlib.build_hull(pts)
    name = f"obj{index}"
    fn = f"{name}_hull.obj"
    x, y = -8.0 + 0.5 * (index // 10), -8.0 + 0.5 * (index % 10)
    xml = f"""<asset><mesh name="{name}_coll_0" file="{fn}"/></asset>
<worldbody><body name="{name}" pos="{x} {y} 0.06" quat="1 0 0 0">
<geom mesh="{name}_coll_0" mass="{mass}" group="3" type="mesh" conaffinity="1" contype="1" condim="4"
      friction="1.0 0.3 0.1" solimp="0.998 0.998 0.001" solref="0.001 1"/>
<joint damping="0.0001" name="{name}:joint" type="free"/></body></worldbody>"""
    return xml, {fn: meshlib.write_obj(h.verts, h.tri)}, name


def build_clutter_scene(gripper: str, object_seeds):
    gx, ga = gripper_fragment(gripper)
    oxml, oassets, names = "", {}, []
    for i, sd in enumerate(object_seeds):
        x, a, nm = clutter_object_fragment(sd, i)
        oxml += x
        oassets.update(a)
        names.append(nm)
    model = compile_mjcf(CLUTTER_XML.format(gripper=gx, objects=oxml), {**ga, **oassets})
    g = GRIPPERS[gripper]
    info = dict(base_qposadr=int(model.jnt_qposadr[model.names["joint"][g["freejoint"]]]),
                joint_qposadr=np.array([model.jnt_qposadr[model.names["joint"][j] if j is not None else -1] for j in g["joints"]], dtype=np.int32),
                close_ctrl=np.array(g["close_ctrl"], dtype=np.float64), repose=g["repose"], gripper=gripper, ground_name="geom:table",
                object_qposadr=[int(model.jnt_qposadr[model.names["joint"][f"{n}:joint"]]) for n in names],
                object_dofadr=[int(model.jnt_dofadr[model.names["joint"][f"{n}:joint"]]) for n in names])
    return model, info


def record_from_model(model):
    """Initial scene record (mj_resetData): qpos0 | qvel 0 | qacc_warmstart 0 | ctrl 0 | mocap pose."""
    return np.concatenate([model.qpos0, np.zeros(2 * model.nv + model.nu), model.mocap_pos0.reshape(-1), model.mocap_quat0.reshape(-1)])


def gen_clutter(model, info, step_fn, seed: int, park_gripper=(0.0, 0.0, 1.5)):
    """ClutterTableEnv.gen_clutter (clutter_table.py:197-222): drop the objects one by one from (0, 0, 0.8) with
    a random orientation, 900 steps each with qvel clipped to +-50, then 9000 settling steps.  `step_fn(record,
    nstep) -> record` advances a scene record (the oracle in tests, the GPU library in the mgs mirror).  The
    gripper is parked above the workspace, as gen_scene does before generating (gen_scene.py:28-45)."""
    rng = np.random.default_rng(4000 + seed)
    nq, nv = model.nq, model.nv
    rec = record_from_model(model)
    b = info["base_qposadr"]
    rec[b:b + 3] = park_gripper
    mo = nq + 2 * nv + model.nu
    rec[mo:mo + 3] = park_gripper
    q = R.random(random_state=np.random.RandomState(int(rng.integers(1 << 31)))).as_quat()
    drop_quat = np.array([q[3], q[0], q[1], q[2]])
    for a in info["object_qposadr"]:
        rec[a:a + 3] = [0.0, 0.0, 0.8]
        rec[a + 3:a + 7] = drop_quat
        rec[nq:nq + nv] = 0.0
        for _ in range(9):
            rec[nq:nq + nv] = np.clip(rec[nq:nq + nv], -50.0, 50.0)
            rec = step_fn(rec, 100)
    for _ in range(90):
        rec[nq:nq + nv] = np.clip(rec[nq:nq + nv], -50.0, 50.0)
        rec = step_fn(rec, 100)
    return rec


def clutter_candidates(model, info, rec, n: int, seed: int):
    """Top-down-ish grasp candidates over the settled objects: antipodal frames around each object's
    position with the approach axis within ~35 degrees of straight down (harness-generated inputs)."""
    rng = np.random.default_rng(5000 + seed)
    H = np.zeros((n, 4, 4))
    width = rng.uniform(0.02, 0.07, size=n)
    for i in range(n):
        a = info["object_qposadr"][int(rng.integers(len(info["object_qposadr"])))]
        c = rec[a:a + 3] + rng.normal(scale=0.01, size=3)
        z = np.array([0.0, 0.0, -1.0]) + rng.normal(scale=0.3, size=3)
        z /= np.linalg.norm(z)
        x = np.cross(rng.normal(size=3), z)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        H[i, :3, 0], H[i, :3, 1], H[i, :3, 2], H[i, :3, 3], H[i, 3, 3] = x, y, z, c, 1.0
    return H, width


def gen_clutter_batch(model, info, step_fn, seeds, park_gripper=(0.0, 0.0, 1.5)):
    """Many clutter scenes of ONE model at once: the same drop / settle schedule as
    `gen_clutter`, every scene with its own seed, all scenes advanced by the same batched launches.
    `step_fn(records[n, stride], nstep) -> records`.  Scene k equals gen_clutter(..., seeds[k]) exactly: an
    environment's trajectory does not depend on what else is in the batch."""
    seeds = list(seeds)
    n, nq, nv = len(seeds), model.nq, model.nv
    rec = np.tile(record_from_model(model), (n, 1))
    b = info["base_qposadr"]
    mo = nq + 2 * nv + model.nu
    rec[:, b:b + 3] = park_gripper
    rec[:, mo:mo + 3] = park_gripper
    quats = np.zeros((n, 4))
    for k, sd in enumerate(seeds):
        rng = np.random.default_rng(4000 + sd)
        q = R.random(random_state=np.random.RandomState(int(rng.integers(1 << 31)))).as_quat()
        quats[k] = [q[3], q[0], q[1], q[2]]

    def advance(r, times):
        for _ in range(times):
            r[:, nq:nq + nv] = np.clip(r[:, nq:nq + nv], -50.0, 50.0)
            r = step_fn(r, 100)
        return r

    for a in info["object_qposadr"]:
        rec[:, a:a + 3] = [0.0, 0.0, 0.8]
        rec[:, a + 3:a + 7] = quats
        rec[:, nq:nq + nv] = 0.0
        rec = advance(rec, 9)
    return advance(rec, 90)


def scenes_stable(model, info, step_fn, rec):
    """ClutterTableEnv.is_stable (clutter_table.py:160-195) for a batch of scene records: summed |displacement| of every
    object over 10 x 100 steps below 5 mm.  Returns (stable bool[n], records after the 1000 steps)."""
    adr = info["object_qposadr"]
    stats = np.zeros((len(rec), len(adr)))
    for _ in range(10):
        start = np.stack([rec[:, a:a + 3] for a in adr], axis=1)
        rec = step_fn(rec, 100)
        stats += np.abs(np.stack([rec[:, a:a + 3] for a in adr], axis=1) - start).sum(axis=2)
    return (stats.max(axis=1) < 5e-3) if len(adr) else np.ones(len(rec), dtype=bool), rec
