"""`MgsModelDesc`: the flat, pointer-based model description that crosses the C ABI.

One field table drives (a) the ctypes Structure used here and (b) the generated C declaration in
`include/mgs_model_desc.h` (run `python -m mj_grasp_sim_b200.model_desc` to regenerate), so the
two cannot drift.  All reals are float64 and all integers int32 on the host side of the boundary;
the CUDA library converts to its compute type when it uploads.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

I, D, PI, PD = "int", "double", "int*", "double*"

FIELDS = [
    # sizes
    ("nq", I), ("nv", I), ("nu", I), ("nbody", I), ("njnt", I), ("ngeom", I), ("neq", I), ("nmocap", I),
    ("ntendon", I), ("nwrap", I), ("ncgeom", I), ("npair", I), ("nhull", I), ("nhullvert", I), ("nhullface", I),
    ("nhullfacevert", I), ("nhullnbr", I),
    # options
    ("cone_elliptic", I), ("iterations", I), ("ls_iterations", I), ("noslip_iterations", I), ("mpr_iterations", I),
    ("timestep", D), ("impratio", D), ("tolerance", D), ("ls_tolerance", D), ("noslip_tolerance", D),
    ("mpr_tolerance", D), ("meaninertia", D), ("gravity", "double3"),
    # bodies
    ("body_parentid", PI), ("body_rootid", PI), ("body_weldid", PI), ("body_mocapid", PI), ("body_jntadr", PI),
    ("body_jntnum", PI), ("body_dofadr", PI), ("body_dofnum", PI), ("body_pos", PD), ("body_quat", PD),
    ("body_ipos", PD), ("body_iquat", PD), ("body_mass", PD), ("body_inertia", PD), ("body_gravcomp", PD),
    ("body_invweight0", PD),
    # joints / dofs
    ("jnt_type", PI), ("jnt_bodyid", PI), ("jnt_qposadr", PI), ("jnt_dofadr", PI), ("jnt_limited", PI),
    ("jnt_pos", PD), ("jnt_axis", PD), ("jnt_range", PD), ("jnt_stiffness", PD), ("jnt_solref", PD),
    ("jnt_solimp", PD), ("jnt_margin", PD), ("qpos0", PD), ("qpos_spring", PD),
    ("dof_bodyid", PI), ("dof_jntid", PI), ("dof_parentid", PI), ("dof_armature", PD), ("dof_damping", PD),
    ("dof_frictionloss", PD), ("dof_solref", PD), ("dof_solimp", PD), ("dof_invweight0", PD),
    # collision geoms (only geoms with contype|conaffinity != 0), hulls, candidate pairs
    ("cgeom_geomid", PI), ("cgeom_type", PI), ("cgeom_bodyid", PI), ("cgeom_hullid", PI), ("cgeom_pos", PD),
    ("cgeom_quat", PD), ("cgeom_size", PD), ("cgeom_rbound", PD),
    ("hull_vertadr", PI), ("hull_vertnum", PI), ("hull_faceadr", PI), ("hull_facenum", PI), ("hull_vert", PD),
    ("hull_facenormal", PD), ("hull_facevertadr", PI), ("hull_facevertnum", PI), ("hull_facevert", PI),
    ("hull_nbradr", PI), ("hull_nbrnum", PI), ("hull_nbr", PI),
    ("pair_geom1", PI), ("pair_geom2", PI), ("pair_condim", PI), ("pair_friction", PD), ("pair_solref", PD),
    ("pair_solimp", PD), ("pair_margin", PD), ("pair_gap", PD),
    # tendons, actuators, equalities
    ("tendon_adr", PI), ("tendon_num", PI), ("wrap_dofadr", PI), ("wrap_qposadr", PI), ("wrap_coef", PD),
    ("actuator_trntype", PI), ("actuator_trnid", PI), ("actuator_ctrllimited", PI), ("actuator_forcelimited", PI),
    ("actuator_gainprm", PD), ("actuator_biasprm", PD), ("actuator_ctrlrange", PD), ("actuator_forcerange", PD),
    ("actuator_gear", PD),
    ("eq_type", PI), ("eq_obj1id", PI), ("eq_obj2id", PI), ("eq_active", PI), ("eq_data", PD), ("eq_solref", PD),
    ("eq_solimp", PD),
    ("mocap_pos0", PD), ("mocap_quat0", PD),
    # label logic: geom id of "geom:ground"/table in the full geom numbering (-1 if absent)
    ("ground_geomid", I),
]

_CT = {I: C.c_int, D: C.c_double, PI: C.POINTER(C.c_int), PD: C.POINTER(C.c_double), "double3": C.c_double * 3}


class MgsModelDesc(C.Structure):
    _fields_ = [(n, _CT[t]) for n, t in FIELDS]


def make_desc(model, ground_name: str = "geom:ground"):
    """ctypes MgsModelDesc for a compiled `Model`; returns (desc, keepalive list)."""
    d = MgsModelDesc()
    keep = []
    ar = model.arr
    derived = {
        "nwrap": len(ar["wrap_coef"]), "nhullvert": len(ar["hull_vert"]), "nhullface": len(ar["hull_facenormal"]),
        "nhullfacevert": len(ar["hull_facevert"]), "nhullnbr": len(ar["hull_nbr"]),
        "cone_elliptic": int(model.opt["cone"] == "elliptic"),
        "ground_geomid": int(model.names["geom"].get(ground_name, -1)),
    }
    if model.opt["integrator"] != "implicitfast":
        raise NotImplementedError("only integrator=implicitfast is implemented (the reference's setting)")
    for n, t in FIELDS:
        if t == I:
            v = derived[n] if n in derived else (model.opt[n] if n in model.opt else ar[n])
            setattr(d, n, int(v))
        elif t == D:
            setattr(d, n, float(model.opt[n] if n in model.opt else ar[n]))
        elif t == "double3":
            setattr(d, n, (C.c_double * 3)(*[float(x) for x in model.opt[n]]))
        else:
            a = np.ascontiguousarray(ar[n], dtype=np.int32 if t == PI else np.float64).reshape(-1)
            if a.size == 0:
                a = np.zeros(1, dtype=a.dtype)
            keep.append(a)
            setattr(d, n, a.ctypes.data_as(_CT[t]))
    return d, keep


def c_declaration() -> str:
    lines = ["/* GENERATED by python -m mj_grasp_sim_b200.model_desc - do not edit. */",
             "#ifndef MGS_MODEL_DESC_H", "#define MGS_MODEL_DESC_H", "",
             "/* Flat model description produced by the host MJCF compiler.  It replaces, for the hot path,",
             " * the mjModel the reference builds with MjModel.from_xml_string",
             " * (/root/reference/mgs/env/gravityless_object_grasping.py:67).  Arrays are row-major, reals are",
             " * double, ids are int; pointers are borrowed for the duration of the call that receives them. */",
             "typedef struct MgsModelDesc {"]
    for n, t in FIELDS:
        if t == "double3":
            lines.append(f"  double {n}[3];")
        elif t in (PI, PD):
            lines.append(f"  const {t[:-1]} *{n};")
        else:
            lines.append(f"  {t} {n};")
    lines += ["} MgsModelDesc;", "", "#endif", ""]
    return "\n".join(lines)


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mgs_model_desc.h")
    with open(out, "w") as f:
        f.write(c_declaration())
    print("wrote", out)
