"""Mesh loading and convex-hull preprocessing for the model compiler.

The reference hands MuJoCo raw STL/OBJ bytes in the asset dict
(`/root/reference/mgs/gripper/panda.py:156-188`); MuJoCo's MJCF compiler then builds a convex
hull (qhull) for every colliding mesh geom and derives mass properties.  This module does the
same job on the host with numpy/scipy: binary STL + OBJ parsing, `scale`, hull with coplanar
facets merged into polygons (needed for face clipping in the narrowphase), the vertex adjacency
graph (hill-climbing support function) and volume/CoM/inertia.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np
from scipy.spatial import ConvexHull

MAX_POLY = 8  # maximum polygon size kept per hull face (device clipping buffers are sized to this)


def load_stl(data: bytes) -> tuple[np.ndarray, np.ndarray]:
    """Binary STL -> (verts[nv,3] float64 with duplicates merged, faces[nf,3] int)."""
    if data[:5].lower() == b"solid" and b"facet" in data[:1000]:
        return _load_stl_ascii(data)
    (ntri,) = struct.unpack_from("<I", data, 80)
    rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
    tris = np.frombuffer(data, dtype=rec, count=ntri, offset=84)["v"].astype(np.float64)
    flat = tris.reshape(-1, 3)
    verts, inv = np.unique(flat, axis=0, return_inverse=True)
    return verts, inv.reshape(-1, 3)


def _load_stl_ascii(data: bytes):
    pts = []
    for line in data.decode("ascii", "ignore").splitlines():
        s = line.split()
        if len(s) == 4 and s[0] == "vertex":
            pts.append([float(s[1]), float(s[2]), float(s[3])])
    flat = np.asarray(pts, dtype=np.float64)
    verts, inv = np.unique(flat, axis=0, return_inverse=True)
    return verts, inv.reshape(-1, 3)


def load_obj(data: bytes) -> tuple[np.ndarray, np.ndarray]:
    """Wavefront OBJ -> (verts, triangle faces); polygons are fan-triangulated."""
    verts, faces = [], []
    for line in data.decode("utf-8", "ignore").splitlines():
        if line.startswith("v "):
            s = line.split()
            verts.append([float(s[1]), float(s[2]), float(s[3])])
        elif line.startswith("f "):
            idx = []
            for tok in line.split()[1:]:
                k = int(tok.split("/")[0])
                idx.append(k - 1 if k > 0 else len(verts) + k)
            for j in range(1, len(idx) - 1):
                faces.append([idx[0], idx[j], idx[j + 1]])
    return np.asarray(verts, dtype=np.float64), np.asarray(faces, dtype=np.int64).reshape(-1, 3)


def load_mesh(name: str, data: bytes):
    low = name.lower()
    if low.endswith(".stl"):
        return load_stl(data)
    if low.endswith(".obj"):
        return load_obj(data)
    raise ValueError(f"unsupported mesh format: {name}")


def mass_properties(verts: np.ndarray, faces: np.ndarray):
    """Volume, centre of mass and unit-density second-moment matrix C = int (x-c)(x-c)^T dV.

    Signed tetrahedra against the surface centroid, absolute volumes summed: MuJoCo 3.2.2's
    default ("legacy") mesh inertia.  For a convex mesh this equals the exact value.
    """
    tri = verts[faces]  # [nf,3,3]
    # area-weighted surface centroid as the tetrahedron apex
    e1, e2 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    area = 0.5 * np.linalg.norm(np.cross(e1, e2), axis=1)
    if area.sum() <= 0:
        raise ValueError("degenerate mesh")
    cen = (tri.mean(axis=1) * area[:, None]).sum(0) / area.sum()
    a, b, c = tri[:, 0] - cen, tri[:, 1] - cen, tri[:, 2] - cen
    vol = np.abs(np.einsum("ij,ij->i", a, np.cross(b, c))) / 6.0
    V = vol.sum()
    tet_c = (a + b + c) / 4.0
    com_rel = (tet_c * vol[:, None]).sum(0) / V
    # second moments of each tetrahedron (apex at origin): integral of x x^T
    # = vol/20 * (sum_i p_i p_i^T + (sum_i p_i)(sum_i p_i)^T) with p_0 = 0
    s = a + b + c
    P = (np.einsum("ni,nj->nij", a, a) + np.einsum("ni,nj->nij", b, b)
         + np.einsum("ni,nj->nij", c, c) + np.einsum("ni,nj->nij", s, s))
    C = (P * (vol / 20.0)[:, None, None]).sum(0)  # covariance about `cen`
    C -= V * np.outer(com_rel, com_rel)  # shift to CoM
    return V, cen + com_rel, C


def cov_to_inertia(C: np.ndarray) -> np.ndarray:
    return np.trace(C) * np.eye(3) - C


@dataclass
class Hull:
    """Convex polytope in its own frame (origin = where the caller put it)."""
    verts: np.ndarray  # [n,3]
    # polygon faces: face_adr[f]..+face_num[f] index into face_vert (CCW seen from outside)
    face_normal: np.ndarray  # [nf,3]
    face_adr: np.ndarray
    face_num: np.ndarray
    face_vert: np.ndarray
    # vertex adjacency (CSR)
    nbr_adr: np.ndarray
    nbr_num: np.ndarray
    nbr: np.ndarray
    tri: np.ndarray = field(default=None)  # triangulated hull surface (for mass properties)


def _order_polygon(pts: np.ndarray, n: np.ndarray) -> np.ndarray:
    c = pts.mean(0)
    u = pts[0] - c
    u -= n * (u @ n)
    if np.linalg.norm(u) < 1e-14:
        u = np.cross(n, [1.0, 0, 0])
        if np.linalg.norm(u) < 1e-8:
            u = np.cross(n, [0, 1.0, 0])
    u /= np.linalg.norm(u)
    w = np.cross(n, u)
    ang = np.arctan2((pts - c) @ w, (pts - c) @ u)
    return np.argsort(ang)


def _decimate_polygon(idx: list[int], verts: np.ndarray, k: int) -> list[int]:
    """Drop the vertices whose removal loses the least area until k remain."""
    idx = list(idx)
    while len(idx) > k:
        n = len(idx)
        best, besta = 0, np.inf
        for i in range(n):
            p0, p1, p2 = verts[idx[i - 1]], verts[idx[i]], verts[idx[(i + 1) % n]]
            a = np.linalg.norm(np.cross(p1 - p0, p2 - p0))
            if a < besta:
                best, besta = i, a
        idx.pop(best)
    return idx


def build_hull(points: np.ndarray, coplanar_tol: float = 1e-7) -> Hull:
    """Convex hull with merged coplanar facets and the vertex adjacency graph."""
    points = np.asarray(points, dtype=np.float64)
    scale = max(1e-12, np.abs(points - points.mean(0)).max())
    ch = ConvexHull(points)
    used = ch.vertices  # indices into points
    remap = -np.ones(len(points), dtype=np.int64)
    remap[used] = np.arange(len(used))
    verts = points[used]
    simp = remap[ch.simplices]
    eq = ch.equations  # [nf,4], outward normal . x + d <= 0 inside
    # orient triangles outward
    tri = simp.copy()
    for f in range(len(tri)):
        a, b, c = verts[tri[f]]
        if np.dot(np.cross(b - a, c - a), eq[f, :3]) < 0:
            tri[f, 1], tri[f, 2] = tri[f, 2], tri[f, 1]
    # group coplanar facets
    nf = len(tri)
    group = -np.ones(nf, dtype=np.int64)
    groups = []
    for f in range(nf):
        if group[f] >= 0:
            continue
        same = np.where((group < 0)
                        & (np.abs(eq[:, :3] @ eq[f, :3] - 1.0) < 1e-9)
                        & (np.abs(eq[:, 3] - eq[f, 3]) < coplanar_tol * max(1.0, scale / 1e-2)))[0]
        group[same] = len(groups)
        groups.append(same)
    face_normal, face_adr, face_num, face_vert = [], [], [], []
    edges = set()
    for same in groups:
        n = eq[same[0], :3] / np.linalg.norm(eq[same[0], :3])
        vid = np.unique(tri[same].reshape(-1))
        # keep only boundary vertices of the merged polygon = hull of the coplanar set (2-D)
        order = _order_polygon(verts[vid], n)
        poly = [int(v) for v in vid[order]]
        # drop vertices interior to edges / inside (collinear within tolerance)
        changed = True
        while changed and len(poly) > 3:
            changed = False
            for i in range(len(poly)):
                p0, p1, p2 = verts[poly[i - 1]], verts[poly[i]], verts[poly[(i + 1) % len(poly)]]
                if np.dot(np.cross(p1 - p0, p2 - p1), n) <= 1e-14 * scale:
                    poly.pop(i)
                    changed = True
                    break
        for i in range(len(poly)):
            a, b = poly[i], poly[(i + 1) % len(poly)]
            edges.add((min(a, b), max(a, b)))
        poly = _decimate_polygon(poly, verts, MAX_POLY)
        face_normal.append(n)
        face_adr.append(len(face_vert))
        face_num.append(len(poly))
        face_vert.extend(poly)
    # also keep the triangulation edges so every vertex has neighbours even if decimated
    for t in tri:
        for i in range(3):
            a, b = int(t[i]), int(t[(i + 1) % 3])
            edges.add((min(a, b), max(a, b)))
    nbrs = [[] for _ in range(len(verts))]
    for a, b in sorted(edges):
        nbrs[a].append(b)
        nbrs[b].append(a)
    nbr_adr, nbr_num, nbr = [], [], []
    for lst in nbrs:
        nbr_adr.append(len(nbr))
        nbr_num.append(len(lst))
        nbr.extend(lst)
    return Hull(
        verts=verts,
        face_normal=np.asarray(face_normal),
        face_adr=np.asarray(face_adr, dtype=np.int32),
        face_num=np.asarray(face_num, dtype=np.int32),
        face_vert=np.asarray(face_vert, dtype=np.int32),
        nbr_adr=np.asarray(nbr_adr, dtype=np.int32),
        nbr_num=np.asarray(nbr_num, dtype=np.int32),
        nbr=np.asarray(nbr, dtype=np.int32),
        tri=tri,
    )


def box_hull(size) -> Hull:
    sx, sy, sz = size
    pts = np.array([[x * sx, y * sy, z * sz] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64)
    return build_hull(pts)


def write_stl(verts: np.ndarray, tri: np.ndarray) -> bytes:
    out = bytearray(b"mgs-b200 convex hull".ljust(80, b" "))
    out += struct.pack("<I", len(tri))
    for t in tri:
        a, b, c = verts[t]
        n = np.cross(b - a, c - a)
        ln = np.linalg.norm(n)
        n = n / ln if ln > 0 else n
        out += struct.pack("<12fH", *n, *a, *b, *c, 0)
    return bytes(out)


def write_obj(verts: np.ndarray, tri: np.ndarray) -> bytes:
    lines = ["# mgs-b200 convex hull"]
    lines += ["v %.9g %.9g %.9g" % tuple(v) for v in verts]
    lines += ["f %d %d %d" % tuple(t + 1) for t in tri]
    return ("\n".join(lines) + "\n").encode()
