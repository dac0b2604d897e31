"""MJCF-subset model compiler (host side).

Takes the same `(xml_str, assets)` pair the reference passes to `MjModel.from_xml_string`
(`/root/reference/mgs/env/gravityless_object_grasping.py:61-69`) and produces flat arrays
(`Model`) for the oracle and the CUDA kernels.  Covers the feature subset listed in SURVEY.md
section 2.2: default classes / childclass, multiple spliced top-level sections, <include> from the
asset dict, free/slide/hinge joints, mocap bodies, box/sphere/capsule/cylinder/mesh geoms with
convex hulls, explicit or geom-derived inertials, weld/connect/joint equalities, fixed tendons,
position/general actuators, <exclude>, contact-parameter mixing, and the qpos0-time constants
(body_invweight0, dof_invweight0, meaninertia, weld relpose, connect anchors).
"""
from __future__ import annotations

import json
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

from . import mesh as meshlib

# ---- enums shared with the C side (include/mgs_b200.h) --------------------------------------
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
EQ_CONNECT, EQ_WELD, EQ_JOINT = 0, 1, 2
TRN_JOINT, TRN_TENDON = 0, 1
GEOM_TYPES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
              "ellipsoid": GEOM_ELLIPSOID, "cylinder": GEOM_CYLINDER, "box": GEOM_BOX, "mesh": GEOM_MESH}
MINVAL = 1e-15


def _f(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    a = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and len(a) < n and default is not None:
        d = np.array(default, dtype=np.float64)
        d[: len(a)] = a
        return d
    return a


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw])


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def mat_to_quat(R):
    from scipy.spatial.transform import Rotation
    q = Rotation.from_matrix(R).as_quat()  # xyzw
    return np.array([q[3], q[0], q[1], q[2]])


def quat_norm(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    return q / n if n > 0 else np.array([1.0, 0, 0, 0])


@dataclass
class Model:
    """Flat model arrays (float64/int32 numpy).  Names follow MuJoCo's mjModel."""
    opt: dict = field(default_factory=dict)
    names: dict = field(default_factory=dict)  # {"body": {name:id}, "joint":…, "geom":…, …}
    arr: dict = field(default_factory=dict)

    def __getattr__(self, k):
        arr = object.__getattribute__(self, "arr")
        if k in arr:
            return arr[k]
        raise AttributeError(k)

    @property
    def nq(self): return int(self.arr["nq"])
    @property
    def nv(self): return int(self.arr["nv"])
    @property
    def nu(self): return int(self.arr["nu"])
    @property
    def nbody(self): return int(self.arr["nbody"])


# ---------------------------------------------------------------------------------------------
class _Defaults:
    """Default-class tree.  MJCF sections are merged by type before compilation, so top-level
    <default> children of every spliced fragment all land in class "main" (SURVEY quirk 11)."""

    def __init__(self):
        self.cls = {"main": {}}
        self.parent = {"main": None}

    def load(self, node, parent="main", top=True):
        name = "main" if top else node.get("class")
        if name not in self.cls:
            self.cls[name] = {}
            self.parent[name] = parent if name != "main" else None
        for ch in node:
            if ch.tag == "default":
                self.load(ch, name, top=False)
            else:
                self.cls[name].setdefault(ch.tag, {}).update(ch.attrib)

    def resolve(self, cls, tags):
        """Merged attribute dict for `tags` (a tag or a list of equivalent tags) in class `cls`,
        ancestors first."""
        chain = []
        c = cls
        while c is not None:
            chain.append(c)
            c = self.parent.get(c)
        out = {}
        for c in reversed(chain):
            for t in tags:
                if t in self.cls[c]:
                    out.update(_norm_actuator(t, self.cls[c][t]) if t in _ACT_TAGS else self.cls[c][t])
        return out


_ACT_TAGS = ("general", "position", "motor", "velocity")


def _norm_actuator(tag, attrib):
    """Express position/motor shortcuts as <general> attributes."""
    a = dict(attrib)
    if tag == "position":
        if "kp" in a:
            a["_kp"] = a.pop("kp")
        if "kv" in a:
            a["_kv"] = a.pop("kv")
        a["_position"] = "1"
    elif tag == "motor":
        a["_motor"] = "1"
    return a


def _expand_includes(root, assets):
    for parent in list(root.iter()):
        for i, ch in enumerate(list(parent)):
            if ch.tag == "include":
                key = os.path.basename(ch.get("file"))
                data = assets.get(ch.get("file"), assets.get(key))
                if data is None:
                    raise ValueError(f"include file not in assets: {ch.get('file')}")
                sub = ET.fromstring(data if isinstance(data, (bytes, str)) else bytes(data))
                _expand_includes(sub, assets)
                idx = list(parent).index(ch)
                parent.remove(ch)
                for k, sc in enumerate(list(sub)):
                    parent.insert(idx + k, sc)


def compile_mjcf(xml: str, assets: dict | None = None, massprops: dict | None = None) -> Model:
    assets = assets or {}
    massprops = dict(massprops or {})
    if "massprops.json" in assets:
        massprops.update(json.loads(assets["massprops.json"]))
    root = ET.fromstring(xml)
    _expand_includes(root, assets)

    comp = {"angle": "degree", "autolimits": "true", "meshdir": ""}
    opt = {"timestep": 0.002, "gravity": np.array([0, 0, -9.81]), "integrator": "Euler", "cone": "pyramidal",
           "impratio": 1.0, "tolerance": 1e-8, "iterations": 100, "ls_iterations": 50, "ls_tolerance": 0.01,
           "noslip_iterations": 0, "noslip_tolerance": 1e-6, "mpr_iterations": 50, "mpr_tolerance": 1e-6,
           "multiccd": False, "solver": "Newton"}
    dfl = _Defaults()
    for sec in root:
        if sec.tag == "compiler":
            comp.update(sec.attrib)
        elif sec.tag == "option":
            for k, v in sec.attrib.items():
                if k == "gravity":
                    opt[k] = _f(v)
                elif k in ("integrator", "cone", "solver", "jacobian"):
                    opt[k] = v
                elif k in ("iterations", "ls_iterations", "noslip_iterations", "mpr_iterations"):
                    opt[k] = int(v)
                else:
                    opt[k] = float(v)
            for fl in sec.findall("flag"):
                if fl.get("multiccd") == "enable":
                    opt["multiccd"] = True
        elif sec.tag == "default":
            dfl.load(sec)
    if comp["angle"] != "radian":
        raise NotImplementedError("only angle=radian models are supported")
    autolimits = comp.get("autolimits", "true") == "true"

    # ---- assets: meshes -------------------------------------------------------------------
    meshes = {}
    for sec in root.findall("asset"):
        for m in sec.findall("mesh"):
            a = dfl.resolve(m.get("class", "main"), ["mesh"])
            a.update(m.attrib)
            fn = a["file"]
            name = a.get("name", os.path.splitext(os.path.basename(fn))[0])
            meshes[name] = {"file": fn, "scale": _f(a.get("scale"), 3, [1, 1, 1])}
    mesh_cache = {}

    def get_mesh(name):
        if name in mesh_cache:
            return mesh_cache[name]
        md = meshes[name]
        key = os.path.basename(md["file"])
        data = assets.get(md["file"], assets.get(key))
        if data is None:
            raise ValueError(f"mesh file not in assets: {md['file']}")
        v, f = meshlib.load_mesh(key, data)
        s = md["scale"]
        v = v * s
        if np.prod(s) < 0:
            f = f[:, ::-1]
        if key in massprops:
            mp = massprops[key]
            S = np.diag(s)
            V = mp["volume"] * abs(np.prod(s))
            com = S @ np.array(mp["com"])
            C = abs(np.prod(s)) * S @ np.array(mp["cov"]) @ S
        else:
            V, com, C = meshlib.mass_properties(v, f)
        hull = meshlib.build_hull(v)
        hV, hcom, _ = meshlib.mass_properties(hull.verts, hull.tri)
        hull.verts = hull.verts - hcom  # geom frame origin := hull centroid (interior point for MPR)
        out = {"V": V, "com": com, "C": C, "hull": hull, "hull_center": hcom}
        mesh_cache[name] = out
        return out

    # ---- bodies ------------------------------------------------------------------------------
    B = {k: [] for k in ("name", "parent", "pos", "quat", "mocap", "gravcomp", "inertial", "cls")}
    J = {k: [] for k in ("name", "type", "body", "pos", "axis", "range", "limited", "stiffness", "springref",
                         "ref", "armature", "damping", "frictionloss", "solreflimit", "solimplimit",
                         "solreffriction", "solimpfriction", "margin")}
    G = {k: [] for k in ("name", "type", "body", "pos", "quat", "size", "mesh", "contype", "conaffinity", "condim",
                         "priority", "friction", "solref", "solimp", "solmix", "margin", "gap", "mass", "density",
                         "group")}
    B["name"].append("world"); B["parent"].append(0); B["pos"].append(np.zeros(3)); B["quat"].append(np.array([1., 0, 0, 0]))
    B["mocap"].append(False); B["gravcomp"].append(0.0); B["inertial"].append(None); B["cls"].append("main")

    def quat_of(attrs, what):
        """orientation of a body / geom / inertial: `quat` only - the other MJCF spellings are refused loudly, not read as identity"""
        for k in ("euler", "axisangle", "xyaxes", "zaxis"):
            if k in attrs:
                raise NotImplementedError(f"{what}: orientation given as '{k}' (this compiler reads 'quat' only; the six gripper templates use nothing else)")
        return quat_norm(_f(attrs.get("quat"), 4, [1, 0, 0, 0]))

    def add_geom(g, bid, childclass):
        cls = g.get("class", childclass or "main")
        a = dfl.resolve(cls, ["geom"])
        a.update(g.attrib)
        gtype = GEOM_TYPES[a.get("type", "sphere")]
        if "mesh" in a and "type" not in a:
            gtype = GEOM_MESH
        if "fromto" in a:
            raise NotImplementedError("geom fromto")
        G["name"].append(a.get("name")); G["type"].append(gtype); G["body"].append(bid)
        G["pos"].append(_f(a.get("pos"), 3, [0, 0, 0])); G["quat"].append(quat_of(a, "geom"))
        G["size"].append(_f(a.get("size"), 3, [0, 0, 0])); G["mesh"].append(a.get("mesh"))
        G["contype"].append(int(a.get("contype", 1))); G["conaffinity"].append(int(a.get("conaffinity", 1)))
        G["condim"].append(int(a.get("condim", 3))); G["priority"].append(int(a.get("priority", 0)))
        G["friction"].append(_f(a.get("friction"), 3, [1, 0.005, 0.0001]))
        G["solref"].append(_f(a.get("solref"), 2, [0.02, 1])); G["solimp"].append(_f(a.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2]))
        G["solmix"].append(float(a.get("solmix", 1))); G["margin"].append(float(a.get("margin", 0))); G["gap"].append(float(a.get("gap", 0)))
        G["mass"].append(float(a["mass"]) if "mass" in a else None); G["density"].append(float(a.get("density", 1000)))
        G["group"].append(int(a.get("group", 0)))

    def add_joint(j, bid, childclass, free=False):
        cls = j.get("class", childclass or "main")
        a = {} if free else dfl.resolve(cls, ["joint"])  # <freejoint> takes no defaults
        a.update(j.attrib)
        jt = JNT_FREE if free else {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[a.get("type", "hinge")]
        if jt == JNT_BALL:
            raise NotImplementedError("ball joints")
        rng = _f(a.get("range"), 2, [0, 0])
        lim = a.get("limited", "auto")
        limited = (lim == "true") or (lim == "auto" and autolimits and "range" in a)
        J["name"].append(a.get("name")); J["type"].append(jt); J["body"].append(bid)
        J["pos"].append(_f(a.get("pos"), 3, [0, 0, 0]))
        ax = _f(a.get("axis"), 3, [0, 0, 1]); J["axis"].append(ax / max(np.linalg.norm(ax), MINVAL))
        J["range"].append(rng); J["limited"].append(bool(limited) and jt in (JNT_SLIDE, JNT_HINGE))
        J["stiffness"].append(float(a.get("stiffness", 0))); J["springref"].append(float(a.get("springref", 0)))
        J["ref"].append(float(a.get("ref", 0))); J["armature"].append(float(a.get("armature", 0)))
        J["damping"].append(float(a.get("damping", 0))); J["frictionloss"].append(float(a.get("frictionloss", 0)))
        J["solreflimit"].append(_f(a.get("solreflimit"), 2, [0.02, 1])); J["solimplimit"].append(_f(a.get("solimplimit"), 5, [0.9, 0.95, 0.001, 0.5, 2]))
        J["solreffriction"].append(_f(a.get("solreffriction"), 2, [0.02, 1])); J["solimpfriction"].append(_f(a.get("solimpfriction"), 5, [0.9, 0.95, 0.001, 0.5, 2]))
        J["margin"].append(float(a.get("margin", 0)))

    def add_body(node, parent, childclass):
        bid = len(B["name"])
        cc = node.get("childclass", childclass)
        B["name"].append(node.get("name")); B["parent"].append(parent)
        B["pos"].append(_f(node.get("pos"), 3, [0, 0, 0])); B["quat"].append(quat_of(node.attrib, "body"))
        B["mocap"].append(node.get("mocap") == "true"); B["gravcomp"].append(float(node.get("gravcomp", 0)))
        B["cls"].append(cc)
        inert = node.find("inertial")
        if inert is not None:
            if "fullinertia" in inert.attrib:
                fi = _f(inert.get("fullinertia"))
                Im = np.array([[fi[0], fi[3], fi[4]], [fi[3], fi[1], fi[5]], [fi[4], fi[5], fi[2]]])
                w, Q = np.linalg.eigh(Im)
                if np.linalg.det(Q) < 0:
                    Q[:, 2] = -Q[:, 2]
                iq, di = mat_to_quat(Q), w
            else:
                iq = quat_of(inert.attrib, "inertial")
                di = _f(inert.get("diaginertia"), 3, [0, 0, 0])
            B["inertial"].append({"mass": float(inert.get("mass")), "pos": _f(inert.get("pos"), 3, [0, 0, 0]),
                                  "quat": iq, "diag": di})
        else:
            B["inertial"].append(None)
        for ch in node:
            if ch.tag == "joint":
                add_joint(ch, bid, cc)
            elif ch.tag == "freejoint":
                add_joint(ch, bid, cc, free=True)
            elif ch.tag == "geom":
                add_geom(ch, bid, cc)
        for ch in node:
            if ch.tag == "body":
                add_body(ch, bid, cc)
        return bid

    # MuJoCo appends world-level elements of every <worldbody> section to the one world body, then
    # numbers bodies depth-first; geoms are numbered in body order.
    for wb in root.findall("worldbody"):
        for ch in wb:
            if ch.tag == "geom":
                add_geom(ch, 0, None)
    for wb in root.findall("worldbody"):
        for ch in wb:
            if ch.tag == "body":
                add_body(ch, 0, wb.get("childclass"))

    nbody = len(B["name"])
    # geoms must be grouped by body id in id order
    order = sorted(range(len(G["body"])), key=lambda i: (G["body"][i], i))
    for k in G:
        G[k] = [G[k][i] for i in order]
    ngeom = len(G["body"])
    njnt = len(J["body"])
    # joints are already in body order because add_body visits depth-first, but child bodies are
    # added after all joints of the parent, so the order is body-major as MuJoCo's.
    jorder = sorted(range(njnt), key=lambda i: (J["body"][i], i))
    for k in J:
        J[k] = [J[k][i] for i in jorder]

    # ---- per-geom mass properties -> body inertials ------------------------------------------
    geom_hull = [None] * ngeom
    geom_massinfo = [None] * ngeom
    G["pos"] = [p.copy() for p in G["pos"]]
    geom_rbound = np.zeros(ngeom)
    for g in range(ngeom):
        t, sz = G["type"][g], G["size"][g]
        R = quat_to_mat(G["quat"][g])
        if t == GEOM_MESH:
            md = get_mesh(G["mesh"][g])
            V, com_l, C_l = md["V"], md["com"], md["C"]
            com = G["pos"][g] + R @ com_l
            Ig = R @ meshlib.cov_to_inertia(C_l) @ R.T  # unit density, about com
            geom_hull[g] = md["hull"]
            G["pos"][g] = G["pos"][g] + R @ md["hull_center"]
            geom_rbound[g] = np.linalg.norm(md["hull"].verts, axis=1).max()
        else:
            com = G["pos"][g]
            if t == GEOM_BOX:
                a, b, c = sz
                V = 8 * a * b * c
                Il = V / 3.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
                geom_rbound[g] = np.linalg.norm(sz)
            elif t == GEOM_SPHERE:
                r = sz[0]
                V = 4.0 / 3.0 * np.pi * r ** 3
                Il = np.full(3, 0.4 * V * r * r)
                geom_rbound[g] = r
            elif t == GEOM_CAPSULE:
                r, h = sz[0], sz[1]
                Vc, Vs = 2 * np.pi * r * r * h, 4.0 / 3.0 * np.pi * r ** 3
                V = Vc + Vs
                izz = Vc * r * r / 2 + Vs * 0.4 * r * r
                ixx = Vc * (3 * r * r + 4 * h * h) / 12 + Vs * (0.4 * r * r + h * h + 0.75 * r * h)
                Il = np.array([ixx, ixx, izz])
                geom_rbound[g] = r + h
            elif t == GEOM_CYLINDER:
                r, h = sz[0], sz[1]
                V = 2 * np.pi * r * r * h
                Il = np.array([V * (3 * r * r + 4 * h * h) / 12] * 2 + [V * r * r / 2])
                geom_rbound[g] = np.hypot(r, h)
            else:
                raise NotImplementedError(f"geom type {t}")
            Ig = R @ np.diag(Il) @ R.T
        m = G["mass"][g] if G["mass"][g] is not None else G["density"][g] * V
        geom_massinfo[g] = (m, com, Ig * (m / V if V > 0 else 0.0))

    body_mass = np.zeros(nbody); body_ipos = np.zeros((nbody, 3)); body_iquat = np.tile([1., 0, 0, 0], (nbody, 1))
    body_inertia = np.zeros((nbody, 3))
    for b in range(1, nbody):
        ine = B["inertial"][b]
        if ine is not None:
            body_mass[b], body_ipos[b], body_iquat[b], body_inertia[b] = ine["mass"], ine["pos"], ine["quat"], ine["diag"]
            continue
        gs = [g for g in range(ngeom) if G["body"][g] == b]
        m = sum(geom_massinfo[g][0] for g in gs)
        if m <= 0:
            continue
        c = sum(geom_massinfo[g][0] * geom_massinfo[g][1] for g in gs) / m
        I = np.zeros((3, 3))
        for g in gs:
            mg, pg, Ig = geom_massinfo[g]
            d = pg - c
            I += Ig + mg * ((d @ d) * np.eye(3) - np.outer(d, d))
        w, Q = np.linalg.eigh(I)
        if np.linalg.det(Q) < 0:
            Q[:, 2] = -Q[:, 2]
        body_mass[b], body_ipos[b], body_iquat[b], body_inertia[b] = m, c, mat_to_quat(Q), w

    # ---- kinematic tree bookkeeping ----------------------------------------------------------
    body_parent = np.array(B["parent"], dtype=np.int32)
    jnt_type = np.array(J["type"], dtype=np.int32)
    jnt_body = np.array(J["body"], dtype=np.int32)
    jnt_qposadr = np.zeros(njnt, dtype=np.int32); jnt_dofadr = np.zeros(njnt, dtype=np.int32)
    nq = nv = 0
    for j in range(njnt):
        jnt_qposadr[j], jnt_dofadr[j] = nq, nv
        nq += 7 if jnt_type[j] == JNT_FREE else 1
        nv += 6 if jnt_type[j] == JNT_FREE else 1
    body_jntadr = -np.ones(nbody, dtype=np.int32); body_jntnum = np.zeros(nbody, dtype=np.int32)
    body_dofadr = -np.ones(nbody, dtype=np.int32); body_dofnum = np.zeros(nbody, dtype=np.int32)
    for j in range(njnt):
        b = jnt_body[j]
        if body_jntnum[b] == 0:
            body_jntadr[b], body_dofadr[b] = j, jnt_dofadr[j]
        body_jntnum[b] += 1
        body_dofnum[b] += 6 if jnt_type[j] == JNT_FREE else 1
    dof_body = np.zeros(nv, dtype=np.int32); dof_jnt = np.zeros(nv, dtype=np.int32); dof_parent = -np.ones(nv, dtype=np.int32)
    dof_armature = np.zeros(nv); dof_damping = np.zeros(nv); dof_frictionloss = np.zeros(nv)
    for j in range(njnt):
        n = 6 if jnt_type[j] == JNT_FREE else 1
        for k in range(n):
            d = jnt_dofadr[j] + k
            dof_body[d], dof_jnt[d] = jnt_body[j], j
            dof_armature[d], dof_damping[d], dof_frictionloss[d] = J["armature"][j], J["damping"][j], J["frictionloss"][j]
    last_dof_of_body = -np.ones(nbody, dtype=np.int32)
    for b in range(1, nbody):
        p = body_parent[b]
        inherit = last_dof_of_body[p]
        if body_dofnum[b] > 0:
            for k in range(body_dofnum[b]):
                d = body_dofadr[b] + k
                dof_parent[d] = inherit if k == 0 else d - 1
            last_dof_of_body[b] = body_dofadr[b] + body_dofnum[b] - 1
        else:
            last_dof_of_body[b] = inherit
    body_weldid = np.zeros(nbody, dtype=np.int32); body_rootid = np.zeros(nbody, dtype=np.int32)
    body_mocapid = -np.ones(nbody, dtype=np.int32)
    nmocap = 0
    for b in range(1, nbody):
        p = body_parent[b]
        body_weldid[b] = b if body_dofnum[b] > 0 else body_weldid[p]
        body_rootid[b] = b if p == 0 else body_rootid[p]
        if B["mocap"][b]:
            body_mocapid[b] = nmocap
            nmocap += 1
    body_pos = np.array(B["pos"]); body_quat = np.array(B["quat"])

    qpos0 = np.zeros(nq); qpos_spring = np.zeros(nq)
    # world pose of bodies at qpos0 (all hinge/slide at ref) for free-joint qpos0
    xpos0 = np.zeros((nbody, 3)); xquat0 = np.tile([1., 0, 0, 0], (nbody, 1))
    for b in range(1, nbody):
        p = body_parent[b]
        xpos0[b] = xpos0[p] + quat_to_mat(xquat0[p]) @ body_pos[b]
        xquat0[b] = quat_mul(xquat0[p], body_quat[b])
    for j in range(njnt):
        a = jnt_qposadr[j]
        if jnt_type[j] == JNT_FREE:
            qpos0[a:a + 3], qpos0[a + 3:a + 7] = xpos0[jnt_body[j]], xquat0[jnt_body[j]]
            qpos_spring[a:a + 7] = qpos0[a:a + 7]
        else:
            qpos0[a] = J["ref"][j]
            qpos_spring[a] = J["springref"][j]

    # ---- tendons, actuators, equalities ------------------------------------------------------
    jid = {n: i for i, n in enumerate(J["name"]) if n is not None}
    bidx = {n: i for i, n in enumerate(B["name"]) if n is not None}
    ten_names, ten_adr, ten_num, wrap_dof, wrap_qadr, wrap_coef = [], [], [], [], [], []
    for sec in root.findall("tendon"):
        for t in sec.findall("fixed"):
            ten_names.append(t.get("name")); ten_adr.append(len(wrap_dof)); ten_num.append(0)
            for w in t.findall("joint"):
                j = jid[w.get("joint")]
                wrap_dof.append(jnt_dofadr[j]); wrap_qadr.append(jnt_qposadr[j]); wrap_coef.append(float(w.get("coef", 1)))
                ten_num[-1] += 1
        if sec.findall("spatial"):
            raise NotImplementedError("spatial tendons")
    tid = {n: i for i, n in enumerate(ten_names) if n is not None}

    A = {k: [] for k in ("name", "trntype", "trnid", "gain", "bias", "ctrlrange", "ctrllimited", "forcerange", "forcelimited", "gear")}
    for sec in root.findall("actuator"):
        for a_el in sec:
            if a_el.tag not in _ACT_TAGS:
                raise NotImplementedError(f"actuator <{a_el.tag}>")
            a = dfl.resolve(a_el.get("class", "main"), list(_ACT_TAGS))
            a.update(_norm_actuator(a_el.tag, a_el.attrib))
            if a.get("dyntype", "none") != "none":
                raise NotImplementedError("actuator dynamics")
            gain = _f(a.get("gainprm"), 3, [1, 0, 0]); bias = _f(a.get("biasprm"), 3, [0, 0, 0])
            if a.get("biastype", "none") == "none" and a_el.tag != "position":
                bias = np.zeros(3)
            if a_el.tag == "position":
                kp = float(a.get("_kp", 1)); kv = float(a.get("_kv", 0))
                gain = np.array([kp, 0, 0]); bias = np.array([0, -kp, -kv])
            elif a_el.tag == "motor":
                gain = np.array([1.0, 0, 0]); bias = np.zeros(3)
            if "joint" in a:
                A["trntype"].append(TRN_JOINT); A["trnid"].append(jid[a["joint"]])
            elif "tendon" in a:
                A["trntype"].append(TRN_TENDON); A["trnid"].append(tid[a["tendon"]])
            else:
                raise NotImplementedError("actuator transmission")
            A["name"].append(a.get("name")); A["gain"].append(gain); A["bias"].append(bias)
            cl = a.get("ctrllimited", "auto"); fl = a.get("forcelimited", "auto")
            A["ctrlrange"].append(_f(a.get("ctrlrange"), 2, [0, 0])); A["forcerange"].append(_f(a.get("forcerange"), 2, [0, 0]))
            A["ctrllimited"].append(cl == "true" or (cl == "auto" and autolimits and "ctrlrange" in a))
            A["forcelimited"].append(fl == "true" or (fl == "auto" and autolimits and "forcerange" in a))
            A["gear"].append(float(a.get("gear", "1").split()[0]))
    nu = len(A["name"])

    E = {k: [] for k in ("type", "obj1", "obj2", "data", "solref", "solimp", "active")}
    for sec in root.findall("equality"):
        for e in sec:
            a = dfl.resolve(e.get("class", "main"), ["equality"])
            a.update(e.attrib)
            data = np.zeros(11)
            if e.tag == "weld":
                E["type"].append(EQ_WELD); b1 = bidx[a["body1"]]; b2 = bidx[a["body2"]] if "body2" in a else 0
                data[0:3] = _f(a.get("anchor"), 3, [0, 0, 0]); data[10] = float(a.get("torquescale", 1))
                if "relpose" in a:
                    raise NotImplementedError("weld relpose")
                R1 = quat_to_mat(xquat0[b1]); R2 = quat_to_mat(xquat0[b2])
                data[3:6] = R1.T @ (xpos0[b2] + R2 @ data[0:3] - xpos0[b1])
                data[6:10] = quat_mul(quat_conj(xquat0[b1]), xquat0[b2])
                E["obj1"].append(b1); E["obj2"].append(b2)
            elif e.tag == "connect":
                E["type"].append(EQ_CONNECT); b1 = bidx[a["body1"]]; b2 = bidx[a["body2"]] if "body2" in a else 0
                data[0:3] = _f(a.get("anchor"), 3, [0, 0, 0])
                R1 = quat_to_mat(xquat0[b1]); R2 = quat_to_mat(xquat0[b2])
                data[3:6] = R2.T @ (xpos0[b1] + R1 @ data[0:3] - xpos0[b2])
                E["obj1"].append(b1); E["obj2"].append(b2)
            elif e.tag == "joint":
                E["type"].append(EQ_JOINT)
                E["obj1"].append(jid[a["joint1"]]); E["obj2"].append(jid[a["joint2"]] if "joint2" in a else -1)
                data[0:5] = _f(a.get("polycoef"), 5, [0, 1, 0, 0, 0])
            else:
                raise NotImplementedError(f"equality <{e.tag}>")
            E["data"].append(data)
            E["solref"].append(_f(a.get("solref"), 2, [0.02, 1])); E["solimp"].append(_f(a.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2]))
            E["active"].append(a.get("active", "true") == "true")
    neq = len(E["type"])

    excludes = set()
    for sec in root.findall("contact"):
        for e in sec.findall("exclude"):
            b1, b2 = bidx[e.get("body1")], bidx[e.get("body2")]
            excludes.add((min(b1, b2), max(b1, b2)))
        if sec.findall("pair"):
            raise NotImplementedError("explicit contact pairs")

    # ---- collision geoms, hull table, candidate pairs ---------------------------------------
    cg = [g for g in range(ngeom) if (G["contype"][g] | G["conaffinity"][g]) != 0]
    hulls, hull_key = [], {}
    cg_hull = []
    for g in cg:
        t = G["type"][g]
        if t == GEOM_MESH:
            key = ("mesh", G["mesh"][g])
            h = geom_hull[g]
        elif t == GEOM_BOX:
            key = ("box",) + tuple(np.round(G["size"][g], 12))
            h = None
        else:
            cg_hull.append(-1)
            continue
        if key not in hull_key:
            hull_key[key] = len(hulls)
            hulls.append(h if h is not None else meshlib.box_hull(G["size"][g]))
        cg_hull.append(hull_key[key])
    hv_adr, hv_num, hf_adr, hf_num = [], [], [], []
    hverts, hfn, hfadr, hfnum, hfvert, hnadr, hnnum, hnbr = [], [], [], [], [], [], [], []
    for h in hulls:
        hv_adr.append(len(hverts)); hv_num.append(len(h.verts)); hf_adr.append(len(hfn)); hf_num.append(len(h.face_num))
        base_fv, base_nb = len(hfvert), len(hnbr)
        hverts.extend(h.verts.tolist()); hfn.extend(h.face_normal.tolist())
        hfadr.extend((h.face_adr + base_fv).tolist()); hfnum.extend(h.face_num.tolist()); hfvert.extend(h.face_vert.tolist())
        hnadr.extend((h.nbr_adr + base_nb).tolist()); hnnum.extend(h.nbr_num.tolist()); hnbr.extend(h.nbr.tolist())

    def mix(g1, g2):
        p1, p2 = G["priority"][g1], G["priority"][g2]
        fr = np.zeros(5)
        if p1 != p2:
            gp = g1 if p1 > p2 else g2
            condim, f3, solref, solimp = G["condim"][gp], G["friction"][gp], G["solref"][gp].copy(), G["solimp"][gp].copy()
        else:
            condim = max(G["condim"][g1], G["condim"][g2])
            f3 = np.maximum(G["friction"][g1], G["friction"][g2])
            s1, s2 = G["solmix"][g1], G["solmix"][g2]
            if s1 >= MINVAL and s2 >= MINVAL:
                mx = s1 / (s1 + s2)
            elif s1 < MINVAL and s2 < MINVAL:
                mx = 0.5
            else:
                mx = 0.0 if s1 < MINVAL else 1.0
            r1, r2 = G["solref"][g1], G["solref"][g2]
            if r1[0] > 0 and r2[0] > 0:
                solref = mx * r1 + (1 - mx) * r2
            else:
                solref = np.minimum(r1, r2)
            solimp = mx * G["solimp"][g1] + (1 - mx) * G["solimp"][g2]
        fr[:] = [f3[0], f3[0], f3[1], f3[2], f3[2]]
        return condim, fr, solref, solimp, max(G["margin"][g1], G["margin"][g2]), max(G["gap"][g1], G["gap"][g2])

    pairs = []
    for ia in range(len(cg)):
        for ib in range(ia + 1, len(cg)):
            g1, g2 = cg[ia], cg[ib]
            b1, b2 = G["body"][g1], G["body"][g2]
            if b1 == b2:
                continue
            w1, w2 = body_weldid[b1], body_weldid[b2]
            if w1 == w2:
                continue
            if not ((G["contype"][g1] & G["conaffinity"][g2]) or (G["contype"][g2] & G["conaffinity"][g1])):
                continue
            if (min(b1, b2), max(b1, b2)) in excludes:
                continue
            pw1, pw2 = body_weldid[body_parent[w1]], body_weldid[body_parent[w2]]
            if w1 != 0 and w2 != 0 and (w1 == pw2 or w2 == pw1):
                continue
            pairs.append((ia, ib) + mix(g1, g2))
    # Order of the candidate list = order of narrowphase lanes = order of contacts.  Pairs that involve a
    # free-floating single-body tree (the grasped objects) go first: they are the pairs that actually
    # collide during a grasp, and packing them into the same batch of 32 lanes lets the other batches
    # (gripper self-pairs, ground) finish after one or two support queries.
    def _is_object(b):
        return body_parent[b] == 0 and body_dofnum[b] == 6 and not any(body_parent[c] == b for c in range(nbody))
    pairs.sort(key=lambda p: 0 if (_is_object(G["body"][cg[p[0]]]) or _is_object(G["body"][cg[p[1]]])) else 1)
    npair = len(pairs)

    m = Model()
    m.opt = opt
    m.names = {"body": bidx, "joint": jid, "geom": {n: i for i, n in enumerate(G["name"]) if n is not None},
               "tendon": tid, "actuator": {n: i for i, n in enumerate(A["name"]) if n is not None}}
    ar = m.arr
    ar.update(nq=nq, nv=nv, nu=nu, nbody=nbody, njnt=njnt, ngeom=ngeom, neq=neq, nmocap=nmocap, ntendon=len(ten_names),
              ncgeom=len(cg), npair=npair, nhull=len(hulls))
    ar.update(body_parentid=body_parent, body_rootid=body_rootid, body_weldid=body_weldid, body_mocapid=body_mocapid,
              body_jntadr=body_jntadr, body_jntnum=body_jntnum, body_dofadr=body_dofadr, body_dofnum=body_dofnum,
              body_pos=body_pos, body_quat=body_quat, body_ipos=body_ipos, body_iquat=body_iquat,
              body_mass=body_mass, body_inertia=body_inertia, body_gravcomp=np.array(B["gravcomp"]))
    ar.update(jnt_type=jnt_type, jnt_bodyid=jnt_body, jnt_qposadr=jnt_qposadr, jnt_dofadr=jnt_dofadr,
              jnt_pos=np.array(J["pos"]).reshape(njnt, 3), jnt_axis=np.array(J["axis"]).reshape(njnt, 3),
              jnt_range=np.array(J["range"]).reshape(njnt, 2), jnt_limited=np.array(J["limited"], dtype=np.int32),
              jnt_stiffness=np.array(J["stiffness"]), jnt_solref=np.array(J["solreflimit"]).reshape(njnt, 2),
              jnt_solimp=np.array(J["solimplimit"]).reshape(njnt, 5), jnt_margin=np.array(J["margin"]),
              qpos0=qpos0, qpos_spring=qpos_spring)
    dof_solref = np.array([J["solreffriction"][j] for j in dof_jnt]).reshape(nv, 2)
    dof_solimp = np.array([J["solimpfriction"][j] for j in dof_jnt]).reshape(nv, 5)
    ar.update(dof_bodyid=dof_body, dof_jntid=dof_jnt, dof_parentid=dof_parent, dof_armature=dof_armature,
              dof_damping=dof_damping, dof_frictionloss=dof_frictionloss, dof_solref=dof_solref, dof_solimp=dof_solimp)
    ar.update(geom_type=np.array(G["type"], dtype=np.int32), geom_bodyid=np.array(G["body"], dtype=np.int32),
              geom_contype=np.array(G["contype"], dtype=np.int32), geom_conaffinity=np.array(G["conaffinity"], dtype=np.int32),
              geom_condim=np.array(G["condim"], dtype=np.int32), geom_priority=np.array(G["priority"], dtype=np.int32),
              geom_friction=np.array(G["friction"]).reshape(ngeom, 3), geom_solref=np.array(G["solref"]).reshape(ngeom, 2),
              geom_solimp=np.array(G["solimp"]).reshape(ngeom, 5), geom_pos=np.array(G["pos"]).reshape(ngeom, 3),
              geom_quat=np.array(G["quat"]).reshape(ngeom, 4), geom_size=np.array(G["size"]).reshape(ngeom, 3),
              geom_rbound=geom_rbound)
    cga = np.array(cg, dtype=np.int32)
    ar.update(cgeom_geomid=cga, cgeom_type=ar["geom_type"][cga], cgeom_bodyid=ar["geom_bodyid"][cga],
              cgeom_pos=ar["geom_pos"][cga], cgeom_quat=ar["geom_quat"][cga], cgeom_size=ar["geom_size"][cga],
              cgeom_rbound=geom_rbound[cga], cgeom_hullid=np.array(cg_hull, dtype=np.int32))
    ar.update(hull_vertadr=np.array(hv_adr, dtype=np.int32), hull_vertnum=np.array(hv_num, dtype=np.int32),
              hull_faceadr=np.array(hf_adr, dtype=np.int32), hull_facenum=np.array(hf_num, dtype=np.int32),
              hull_vert=np.array(hverts).reshape(-1, 3), hull_facenormal=np.array(hfn).reshape(-1, 3),
              hull_facevertadr=np.array(hfadr, dtype=np.int32), hull_facevertnum=np.array(hfnum, dtype=np.int32),
              hull_facevert=np.array(hfvert, dtype=np.int32), hull_nbradr=np.array(hnadr, dtype=np.int32),
              hull_nbrnum=np.array(hnnum, dtype=np.int32), hull_nbr=np.array(hnbr, dtype=np.int32))
    ar.update(pair_geom1=np.array([p[0] for p in pairs], dtype=np.int32), pair_geom2=np.array([p[1] for p in pairs], dtype=np.int32),
              pair_condim=np.array([p[2] for p in pairs], dtype=np.int32),
              pair_friction=np.array([p[3] for p in pairs]).reshape(npair, 5),
              pair_solref=np.array([p[4] for p in pairs]).reshape(npair, 2),
              pair_solimp=np.array([p[5] for p in pairs]).reshape(npair, 5),
              pair_margin=np.array([p[6] for p in pairs]), pair_gap=np.array([p[7] for p in pairs]))
    ar.update(tendon_adr=np.array(ten_adr, dtype=np.int32), tendon_num=np.array(ten_num, dtype=np.int32),
              wrap_dofadr=np.array(wrap_dof, dtype=np.int32), wrap_qposadr=np.array(wrap_qadr, dtype=np.int32),
              wrap_coef=np.array(wrap_coef))
    ar.update(actuator_trntype=np.array(A["trntype"], dtype=np.int32), actuator_trnid=np.array(A["trnid"], dtype=np.int32),
              actuator_gainprm=np.array(A["gain"]).reshape(nu, 3), actuator_biasprm=np.array(A["bias"]).reshape(nu, 3),
              actuator_ctrlrange=np.array(A["ctrlrange"]).reshape(nu, 2), actuator_ctrllimited=np.array(A["ctrllimited"], dtype=np.int32),
              actuator_forcerange=np.array(A["forcerange"]).reshape(nu, 2), actuator_forcelimited=np.array(A["forcelimited"], dtype=np.int32),
              actuator_gear=np.array(A["gear"]))
    ar.update(eq_type=np.array(E["type"], dtype=np.int32), eq_obj1id=np.array(E["obj1"], dtype=np.int32),
              eq_obj2id=np.array(E["obj2"], dtype=np.int32), eq_data=np.array(E["data"]).reshape(neq, 11),
              eq_solref=np.array(E["solref"]).reshape(neq, 2), eq_solimp=np.array(E["solimp"]).reshape(neq, 5),
              eq_active=np.array(E["active"], dtype=np.int32))
    mocap_pos0 = np.array([body_pos[b] for b in range(nbody) if body_mocapid[b] >= 0]).reshape(nmocap, 3)
    mocap_quat0 = np.array([body_quat[b] for b in range(nbody) if body_mocapid[b] >= 0]).reshape(nmocap, 4)
    ar.update(mocap_pos0=mocap_pos0, mocap_quat0=mocap_quat0)
    _set_const(m)
    return m


# ---------------------------------------------------------------------------------------------
def forward_kinematics(m: Model, qpos, mocap_pos=None, mocap_quat=None):
    """Reference FK in numpy (compile-time constants and tests only)."""
    nb = m.nbody
    xpos = np.zeros((nb, 3)); xquat = np.tile([1., 0, 0, 0], (nb, 1))
    xanchor = np.zeros((m.arr["njnt"], 3)); xaxis = np.zeros((m.arr["njnt"], 3))
    mocap_pos = m.mocap_pos0 if mocap_pos is None else mocap_pos
    mocap_quat = m.mocap_quat0 if mocap_quat is None else mocap_quat
    for b in range(1, nb):
        p = m.body_parentid[b]
        if m.body_mocapid[b] >= 0:
            xpos[b] = mocap_pos[m.body_mocapid[b]]; xquat[b] = quat_norm(mocap_quat[m.body_mocapid[b]])
        else:
            xpos[b] = xpos[p] + quat_to_mat(xquat[p]) @ m.body_pos[b]
            xquat[b] = quat_mul(xquat[p], m.body_quat[b])
        for j in range(m.body_jntadr[b], m.body_jntadr[b] + m.body_jntnum[b]):
            a = m.jnt_qposadr[j]
            t = m.jnt_type[j]
            if t == JNT_FREE:
                xpos[b] = qpos[a:a + 3]; xquat[b] = quat_norm(qpos[a + 3:a + 7])
                xanchor[j] = xpos[b]; xaxis[j] = [0, 0, 1]
                continue
            R = quat_to_mat(xquat[b])
            xaxis[j] = R @ m.jnt_axis[j]
            xanchor[j] = xpos[b] + R @ m.jnt_pos[j]
            dq = qpos[a] - m.qpos0[a]
            if t == JNT_SLIDE:
                xpos[b] = xpos[b] + xaxis[j] * dq
            else:
                ql = np.concatenate([[np.cos(dq / 2)], np.sin(dq / 2) * m.jnt_axis[j]])
                xquat[b] = quat_mul(xquat[b], ql)
                xpos[b] = xanchor[j] - quat_to_mat(xquat[b]) @ m.jnt_pos[j]
    return xpos, xquat, xanchor, xaxis


def body_jacobian(m: Model, xpos, xquat, xanchor, xaxis, body, point):
    """6 x nv Jacobian [jacp; jacr] of world `point` attached to `body`."""
    nv = m.nv
    Jm = np.zeros((6, nv))
    b = body
    while b > 0:
        for j in range(m.body_jntadr[b], m.body_jntadr[b] + m.body_jntnum[b]):
            d = m.jnt_dofadr[j]
            t = m.jnt_type[j]
            if t == JNT_FREE:
                Jm[0:3, d:d + 3] = np.eye(3)
                R = quat_to_mat(xquat[b])
                for k in range(3):
                    Jm[3:6, d + 3 + k] = R[:, k]
                    Jm[0:3, d + 3 + k] = np.cross(R[:, k], point - xpos[b])
            elif t == JNT_SLIDE:
                Jm[0:3, d] = xaxis[j]
            else:
                Jm[3:6, d] = xaxis[j]
                Jm[0:3, d] = np.cross(xaxis[j], point - xanchor[j])
        b = m.body_parentid[b]
    return Jm


def mass_matrix(m: Model, qpos):
    xpos, xquat, xanchor, xaxis = forward_kinematics(m, qpos)
    nv = m.nv
    Mm = np.diag(m.dof_armature.astype(np.float64)).copy()
    for b in range(1, m.nbody):
        if m.body_mass[b] <= 0 and not np.any(m.body_inertia[b] > 0):
            continue
        R = quat_to_mat(xquat[b])
        xipos = xpos[b] + R @ m.body_ipos[b]
        Ri = R @ quat_to_mat(m.body_iquat[b])
        Iw = Ri @ np.diag(m.body_inertia[b]) @ Ri.T
        Jb = body_jacobian(m, xpos, xquat, xanchor, xaxis, b, xipos)
        Mm += m.body_mass[b] * Jb[:3].T @ Jb[:3] + Jb[3:].T @ Iw @ Jb[3:]
    return Mm, (xpos, xquat, xanchor, xaxis)


def _set_const(m: Model):
    """qpos0-time constants: dof_invweight0, body_invweight0, meaninertia, tendon_invweight0."""
    nv, nb = m.nv, m.nbody
    ar = m.arr
    if nv == 0:
        ar.update(dof_invweight0=np.zeros(0), body_invweight0=np.zeros((nb, 2)), meaninertia=1.0,
                  tendon_invweight0=np.zeros(ar["ntendon"]))
        return
    Mm, (xpos, xquat, xanchor, xaxis) = mass_matrix(m, m.qpos0)
    Minv = np.linalg.inv(Mm)
    dinv = np.diag(Minv).copy()
    for j in range(ar["njnt"]):
        if m.jnt_type[j] == JNT_FREE:
            d = m.jnt_dofadr[j]
            dinv[d:d + 3] = dinv[d:d + 3].mean(); dinv[d + 3:d + 6] = dinv[d + 3:d + 6].mean()
    biw = np.zeros((nb, 2))
    for b in range(1, nb):
        if m.body_weldid[b] == 0:
            continue
        R = quat_to_mat(xquat[b])
        Jb = body_jacobian(m, xpos, xquat, xanchor, xaxis, b, xpos[b] + R @ m.body_ipos[b])
        Ab = Jb @ Minv @ Jb.T
        biw[b] = [np.trace(Ab[:3, :3]) / 3, np.trace(Ab[3:, 3:]) / 3]
    tiw = np.zeros(ar["ntendon"])
    for t in range(ar["ntendon"]):
        Jt = np.zeros(nv)
        for w in range(m.tendon_adr[t], m.tendon_adr[t] + m.tendon_num[t]):
            Jt[m.wrap_dofadr[w]] += m.wrap_coef[w]
        tiw[t] = Jt @ Minv @ Jt
    ar.update(dof_invweight0=dinv, body_invweight0=biw, meaninertia=float(np.trace(Mm) / nv), tendon_invweight0=tiw)
