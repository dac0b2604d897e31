// mgs_solver.cuh - constraint rows, primal Newton solver, noslip, implicitfast (warp per env).
//
// Row order and arithmetic follow SURVEY.md 8(a-MJ): equality -> dof friction -> limits ->
// contacts; soft constraints from solref/solimp (refsafe), R from body/dof invweight0, elliptic
// cones with impratio; Newton on qacc with H = M + J'DJ + cone blocks factored by a
// warp-cooperative Cholesky in shared memory; exact 1-D line search on the convex cost; noslip
// Gauss-Seidel on friction rows with the unregularised A = J M^-1 J'.
#pragma once
#include "mgs_collide.cuh"

// accumulator types of the precision-critical reductions of the Newton solve (ablation switches)
#ifdef MGS_ACC64_JAR
typedef double acc_jar_t;
#else
typedef real acc_jar_t;
#endif
#ifdef MGS_ACC64_GRAD
typedef double acc_grad_t;
#else
typedef real acc_grad_t;
#endif
#ifdef MGS_ACC64_LS
typedef double acc_ls_t;
#else
typedef real acc_ls_t;
#endif
#ifdef MGS_ACC64_HESS
typedef double acc_hess_t;
#else
typedef real acc_hess_t;
#endif

// constraint row tag: type (2 bits) | state (3 bits) | id
#define EFC_TAG(i) (IARR(EF(efc_tsi))[i])
#define EFC_TYPE(i) (EFC_TAG(i) & 3)
#define EFC_STATE(i) ((EFC_TAG(i) >> 2) & 7)
#define EFC_ID(i) (EFC_TAG(i) >> 5)
#define EFC_SET_STATE(i, st) (EFC_TAG(i) = (EFC_TAG(i) & ~(7 << 2)) | ((st) << 2))

MGS_DEVN real impedance_f(const real *solimp, real pos) {
  real dmin = fmin(R_(0.9999), fmax(R_(0.0001), solimp[0])), dmax = fmin(R_(0.9999), fmax(R_(0.0001), solimp[1]));
  real width = fmax(MGS_MINVAL, solimp[2]), mid = fmin(R_(0.9999), fmax(R_(0.0001), solimp[3])), power = fmax(R_(1.0), solimp[4]);
  if (dmin == dmax || width <= MGS_MINVAL) return R_(0.5) * (dmin + dmax);
  real x = fabs(pos) / width, y;
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}

// does dof d move a point fixed to `body`? (host-built ancestry mask, one bit per dof)
MGS_DEV int body_has_dof(int body, int d) { return (int)((LDG(MD.body_dofmask + body * MD.dofmask_words + (d >> 5)) >> (d & 31)) & 1u); }

MGS_DEV void kbi_from_solref(const real *solref, const real *solimp, real pos, int friction_row, real *k, real *b, real *imp) {
  *imp = impedance_f(solimp, pos);
  real dmax = fmin(R_(0.9999), fmax(R_(0.0001), solimp[1]));
  if (solref[0] > 0) {
    real tc = fmax(solref[0], 2 * MD.timestep), dr = solref[1];
    *k = R_(1.0) / fmax(MGS_MINVAL, dmax * dmax * tc * tc * dr * dr);
    *b = R_(2.0) / fmax(MGS_MINVAL, dmax * tc);
  } else {
    *k = -solref[0] / fmax(MGS_MINVAL, dmax * dmax);
    *b = -solref[1] / fmax(MGS_MINVAL, dmax);
  }
  if (friction_row) *k = 0;
}

// Row bookkeeping while the Jacobian rows are written: tag + a scratch value (violation `pos` for
// equality/limit rows, frictionloss for dof-friction rows).  Impedance, regularisation and the
// reference acceleration are then computed one ROW PER LANE in finalize_rows_w.
MGS_DEV void tag_row(Env &e, int r, int type, int id, real aux) {
  EFC_TAG(r) = type | (id << 5);
  EF(efc_aux)[r] = aux;
}

// mj_makeImpedance + mj_referenceConstraint for every row (lane per row)
MGS_DEVN void finalize_rows_w(Env &e) {
  const int nv = MD.nv;
  #pragma unroll 1
  PFOR(r, EH.nefc) {
    const int type = EFC_TYPE(r), id = EFC_ID(r);
    real sr[2], si[5], pos = 0, diagApprox, floss = 0;
    int friction_row = 0;
    if (type == CT_EQUALITY) {
      sr[0] = LDG(MD.eq_solref + 2 * id); sr[1] = LDG(MD.eq_solref + 2 * id + 1);
      for (int k = 0; k < 5; k++) si[k] = LDG(MD.eq_solimp + 5 * id + k);
      pos = EF(efc_aux)[r];
      if (LDG(MD.eq_type + id) == EQ_JOINT) {
        const int j1 = LDG(MD.eq_obj1id + id), j2 = LDG(MD.eq_obj2id + id);
        diagApprox = LDG(MD.dof_invweight0 + LDG(MD.jnt_dofadr + j1)) + (j2 >= 0 ? LDG(MD.dof_invweight0 + LDG(MD.jnt_dofadr + j2)) : R_(0.0));
      } else {
        const int b1 = LDG(MD.eq_obj1id + id), b2 = LDG(MD.eq_obj2id + id), k = r - LDG(MD.eq_rowadr + id);
        diagApprox = LDG(MD.body_invweight0 + 2 * b1 + (k >= 3)) + LDG(MD.body_invweight0 + 2 * b2 + (k >= 3));
      }
    } else if (type == CT_FRICTION_DOF) {
      sr[0] = LDG(MD.dof_solref + 2 * id); sr[1] = LDG(MD.dof_solref + 2 * id + 1);
      for (int k = 0; k < 5; k++) si[k] = LDG(MD.dof_solimp + 5 * id + k);
      floss = EF(efc_aux)[r];
      diagApprox = LDG(MD.dof_invweight0 + id);
      friction_row = 1;
    } else if (type == CT_LIMIT) {
      sr[0] = LDG(MD.jnt_solref + 2 * id); sr[1] = LDG(MD.jnt_solref + 2 * id + 1);
      for (int k = 0; k < 5; k++) si[k] = LDG(MD.jnt_solimp + 5 * id + k);
      pos = EF(efc_aux)[r];
      diagApprox = LDG(MD.dof_invweight0 + LDG(MD.jnt_dofadr + id));
    } else {
      const int p = IARR(EF(con_pair))[id], k = r - IARR(EF(con_efc))[id];
      sr[0] = LDG(MD.pair_solref + 2 * p); sr[1] = LDG(MD.pair_solref + 2 * p + 1);
      for (int q = 0; q < 5; q++) si[q] = LDG(MD.pair_solimp + 5 * p + q);
      pos = EF(con_dist)[id];
      const int b1 = LDG(MD.cgeom_bodyid + LDG(MD.pair_geom1 + p)), b2 = LDG(MD.cgeom_bodyid + LDG(MD.pair_geom2 + p));
      diagApprox = LDG(MD.body_invweight0 + 2 * b1 + (k >= 3)) + LDG(MD.body_invweight0 + 2 * b2 + (k >= 3));
      friction_row = k > 0;
    }
    real kk, b, imp, vel = 0;
    kbi_from_solref(sr, si, pos, friction_row, &kk, &b, &imp);
    MGS_UNROLL_INNER
    for (int d = 0; d < nv; d++) vel += EF(J)[r * nv + d] * EF(qvel)[d];
    EF(efc_aux)[r] = floss;
    const real R = fmax(MGS_MINVAL, (1 - imp) * diagApprox / imp);
    EF(efc_R)[r] = R;
    EF(efc_D)[r] = R_(1.0) / R;
    EF(efc_aref)[r] = -b * vel - kk * imp * (friction_row ? R_(0.0) : pos);
  }
  WSYNC();
  // elliptic cones: friction-row regularisation from the normal row and impratio; regularised mu
  #pragma unroll 1
  PFOR(c, EH.ncon) {
    const int r0 = IARR(EF(con_efc))[c];
    if (r0 < 0) continue;
    const int p = IARR(EF(con_pair))[c], dim = LDG(MD.pair_condim + p);
    real fr[5];
    for (int k = 0; k < 5; k++) fr[k] = LDG(MD.pair_friction + 5 * p + k);
    real mu = fr[0];
    if (MD.cone_elliptic && dim >= 3) {
      const real R0 = EF(efc_R)[r0], R1 = R0 / fmax(MGS_MINVAL, MD.impratio);
      EF(efc_R)[r0 + 1] = R1;
      mu = fr[0] * sqrt(R1 / R0);
      #pragma unroll 1
      for (int j = 2; j < dim; j++) EF(efc_R)[r0 + j] = R1 * fr[0] * fr[0] / fmax(MGS_MINVAL, fr[j - 1] * fr[j - 1]);
      #pragma unroll 1
      for (int j = 1; j < dim; j++) EF(efc_D)[r0 + j] = R_(1.0) / EF(efc_R)[r0 + j];
    }
    EF(con_mu)[c] = mu;
  }
  WSYNC();
}

// mj_makeConstraint: Jacobian rows in MuJoCo's order (equality, dof friction, limits, contacts)
MGS_DEVN void make_constraint_w(Env &e) {
  const int nv = MD.nv;
  // --- row bookkeeping (warp-uniform)
  int ne = MD.ne_rows, nf = MD.nf_rows, nl = 0;
  // limits: count with a scan so that rows come out in joint order
  {
    int cnt_base = 0;
    #pragma unroll 1
    for (int j0 = 0; j0 < MD.njnt; j0 += LANES) {
      int j = j0 + MGS_LANE, cnt = 0;
      real dist0 = 0, dist1 = 0, mg = 0;
      if (j < MD.njnt && LDG(MD.jnt_limited + j) && LDG(MD.jnt_type + j) != JNT_FREE) {
        real q = EF(qpos)[LDG(MD.jnt_qposadr + j)];
        mg = LDG(MD.jnt_margin + j);
        dist0 = q - LDG(MD.jnt_range + 2 * j);
        dist1 = LDG(MD.jnt_range + 2 * j + 1) - q;
        cnt = (dist0 < mg) + (dist1 < mg);
      }
      int total, off = wscan_excl(cnt, &total);
      if (cnt) {
        int r = ne + nf + cnt_base + off, dof = LDG(MD.jnt_dofadr + j);
        if (dist0 < mg && r < LY.nefc_max) {
          #pragma unroll 1
          for (int d = 0; d < nv; d++) EF(J)[r * nv + d] = 0;
          EF(J)[r * nv + dof] = 1;
          tag_row(e, r, CT_LIMIT, j, dist0 - mg);
          r++;
        }
        if (dist1 < mg && r < LY.nefc_max) {
          #pragma unroll 1
          for (int d = 0; d < nv; d++) EF(J)[r * nv + d] = 0;
          EF(J)[r * nv + dof] = -1;
          tag_row(e, r, CT_LIMIT, j, dist1 - mg);
        }
      }
      cnt_base += total;
    }
    nl = cnt_base;
    if (ne + nf + nl > LY.nefc_max) {  // limit rows that did not fit were dropped above: flag it and keep the row count in range
      if (MGS_LANE == 0) EH.overflow += 1;
      nl = LY.nefc_max - ne - nf;
    }
  }
  const int row_con0 = ne + nf + nl;
  // contact rows: prefix sum of condim over contacts (capacity-limited)
  int nefc;
  {
    int base = row_con0;
    #pragma unroll 1
    for (int c0 = 0; c0 < EH.ncon; c0 += LANES) {
      int c = c0 + MGS_LANE, dim = 0;
      if (c < EH.ncon) dim = LDG(MD.pair_condim + IARR(EF(con_pair))[c]);
      int total, off = wscan_excl(dim, &total);
      if (c < EH.ncon) IARR(EF(con_efc))[c] = (base + off + dim <= LY.nefc_max) ? base + off : -1;
      base += total;
    }
    WSYNC();
    if (base > LY.nefc_max) {
      // drop the contacts that do not fit (flagged); rows of the kept ones stay contiguous
      if (MGS_LANE == 0) EH.overflow += 1;
      int keep = row_con0;
      #pragma unroll 1
      for (int c = 0; c < EH.ncon; c++) {
        int r = IARR(EF(con_efc))[c];
        if (r >= 0) keep = r + LDG(MD.pair_condim + IARR(EF(con_pair))[c]);
      }
      base = keep;
    }
    nefc = base;
  }
  // Jacobian rows, LANE PER DOF: constraints are visited one after the other (there are few), every lane owns the
  // column of one dof and writes each entry exactly once (no zero-fill, no accumulation).  Whether dof d moves a
  // point of body b is one bit of a host-built ancestry mask.  (The lane-per-constraint version walked the dof chain
  // of both bodies with 1-4 lanes active: ncu r1_j, 1.7 k of the 3.2 k make_constraint instructions per step.)
  #pragma unroll 1
  PFOR(k, nf * nv) EF(J)[ne * nv + k] = 0;  // dof-friction rows are unit rows: zero, then one store each (below)
  // --- equality rows
  #pragma unroll 1
  for (int q = 0; q < MD.neq; q++) {
    if (!LDG(MD.eq_active + q)) continue;
    const int r0 = LDG(MD.eq_rowadr + q), type = LDG(MD.eq_type + q);
    real data[11];
#pragma unroll
    for (int k = 0; k < 11; k++) data[k] = LDG(MD.eq_data + 11 * q + k);
    if (type == EQ_JOINT) {
      const int j1 = LDG(MD.eq_obj1id + q), j2 = LDG(MD.eq_obj2id + q);
      const int qa1 = LDG(MD.jnt_qposadr + j1), d1 = LDG(MD.jnt_dofadr + j1);
      real q1 = EF(qpos)[qa1] - LDG(MD.qpos0 + qa1), pos, deriv = 0;
      int d2 = -1;
      if (j2 >= 0) {
        const int qa2 = LDG(MD.jnt_qposadr + j2);
        d2 = LDG(MD.jnt_dofadr + j2);
        real dif = EF(qpos)[qa2] - LDG(MD.qpos0 + qa2);
        real poly = data[0] + dif * (data[1] + dif * (data[2] + dif * (data[3] + dif * data[4])));
        deriv = data[1] + dif * (2 * data[2] + dif * (3 * data[3] + dif * 4 * data[4]));
        pos = q1 - poly;
      } else pos = q1 - data[0];
      #pragma unroll 1
      PFOR(d, nv) EF(J)[r0 * nv + d] = (d == d1) ? R_(1.0) : (d == d2 ? -deriv : R_(0.0));
      PFOR(k, 1) tag_row(e, r0, CT_EQUALITY, q, pos);
      continue;
    }
    const int b1 = LDG(MD.eq_obj1id + q), b2 = LDG(MD.eq_obj2id + q), weld = (type == EQ_WELD);
    const real *a1 = weld ? data + 3 : data, *a2 = weld ? data : data + 3;
    real p1[3], p2[3], cpos[6] = {0, 0, 0, 0, 0, 0}, off1[3], off2[3];
    mulmatvec3(p1, EF(xmat) + 9 * b1, a1); add3(p1, p1, EF(xpos) + 3 * b1);
    mulmatvec3(p2, EF(xmat) + 9 * b2, a2); add3(p2, p2, EF(xpos) + 3 * b2);
    sub3(cpos, p1, p2);
    sub3(off1, p1, EF(rootcom) + 3 * LDG(MD.body_rootid + b1));
    sub3(off2, p2, EF(rootcom) + 3 * LDG(MD.body_rootid + b2));
    real ts = 0, quat[4] = {1, 0, 0, 0}, quat1[4] = {1, 0, 0, 0};
    if (weld) {
      real quat2[4];
      ts = data[10];
      mulquat(quat, EF(xquat) + 4 * b1, data + 6);
      quat1[0] = EF(xquat)[4 * b2]; quat1[1] = -EF(xquat)[4 * b2 + 1]; quat1[2] = -EF(xquat)[4 * b2 + 2]; quat1[3] = -EF(xquat)[4 * b2 + 3];
      mulquat(quat2, quat1, quat);
      cpos[3] = ts * quat2[1]; cpos[4] = ts * quat2[2]; cpos[5] = ts * quat2[3];
    }
    #pragma unroll 1
    PFOR(d, nv) {
      const int m1 = body_has_dof(b1, d), m2 = body_has_dof(b2, d);
      real jt[3] = {0, 0, 0}, jr[3] = {0, 0, 0};
      if (m1 | m2) {
        const real *c = EF(cdof) + 6 * d;
        const real c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4], c5 = c[5];
        if (m1) {
          jt[0] += c1 * off1[2] - c2 * off1[1] + c3; jt[1] += c2 * off1[0] - c0 * off1[2] + c4; jt[2] += c0 * off1[1] - c1 * off1[0] + c5;
          jr[0] += c0; jr[1] += c1; jr[2] += c2;
        }
        if (m2) {
          jt[0] -= c1 * off2[2] - c2 * off2[1] + c3; jt[1] -= c2 * off2[0] - c0 * off2[2] + c4; jt[2] -= c0 * off2[1] - c1 * off2[0] + c5;
          jr[0] -= c0; jr[1] -= c1; jr[2] -= c2;
        }
      }
      EF(J)[r0 * nv + d] = jt[0]; EF(J)[(r0 + 1) * nv + d] = jt[1]; EF(J)[(r0 + 2) * nv + d] = jt[2];
      if (weld) {
        if (jr[0] != 0 || jr[1] != 0 || jr[2] != 0) {
          real ax[4] = {0, jr[0], jr[1], jr[2]}, quat2[4], quat3[4];
          mulquat(quat2, quat1, ax);
          mulquat(quat3, quat2, quat);
          jr[0] = R_(0.5) * ts * quat3[1]; jr[1] = R_(0.5) * ts * quat3[2]; jr[2] = R_(0.5) * ts * quat3[3];
        }
        EF(J)[(r0 + 3) * nv + d] = jr[0]; EF(J)[(r0 + 4) * nv + d] = jr[1]; EF(J)[(r0 + 5) * nv + d] = jr[2];
      }
    }
    PFOR(k0, 1) {
#pragma unroll
      for (int k = 0; k < 6; k++)
        if (k < (weld ? 6 : 3)) tag_row(e, r0 + k, CT_EQUALITY, q, cpos[k]);
    }
  }
  // --- dof friction rows (lane per dof; row index = rank among friction dofs)
  WSYNC();
  #pragma unroll 1
  PFOR(d, nv) {
    real fl = LDG(MD.dof_frictionloss + d);
    if (fl <= 0) continue;
    const int r = ne + LDG(MD.dof_frictionrank + d);
    EF(J)[r * nv + d] = 1;
    tag_row(e, r, CT_FRICTION_DOF, d, fl);
  }
  // --- contact rows
  #pragma unroll 1
  for (int c = 0; c < EH.ncon; c++) {
    const int r0 = IARR(EF(con_efc))[c];
    if (r0 < 0) continue;
    const int p = IARR(EF(con_pair))[c], dim = LDG(MD.pair_condim + p);
    const int b1 = LDG(MD.cgeom_bodyid + LDG(MD.pair_geom1 + p)), b2 = LDG(MD.cgeom_bodyid + LDG(MD.pair_geom2 + p));
    real axes[9], off1[3], off2[3];
    copy3(axes, EF(con_normal) + 3 * c);
    make_frame(axes);  // tangents are a pure function of the normal: not stored
    sub3(off1, EF(con_pos) + 3 * c, EF(rootcom) + 3 * LDG(MD.body_rootid + b1));
    sub3(off2, EF(con_pos) + 3 * c, EF(rootcom) + 3 * LDG(MD.body_rootid + b2));
    #pragma unroll 1
    PFOR(d, nv) {
      const int m1 = body_has_dof(b1, d), m2 = body_has_dof(b2, d);
      real lin[3] = {0, 0, 0}, ang[3] = {0, 0, 0};
      if (m1 | m2) {
        const real *cd = EF(cdof) + 6 * d;
        const real c0 = cd[0], c1 = cd[1], c2 = cd[2], c3 = cd[3], c4 = cd[4], c5 = cd[5];
        if (m2) {
          lin[0] += c1 * off2[2] - c2 * off2[1] + c3; lin[1] += c2 * off2[0] - c0 * off2[2] + c4; lin[2] += c0 * off2[1] - c1 * off2[0] + c5;
          ang[0] += c0; ang[1] += c1; ang[2] += c2;
        }
        if (m1) {
          lin[0] -= c1 * off1[2] - c2 * off1[1] + c3; lin[1] -= c2 * off1[0] - c0 * off1[2] + c4; lin[2] -= c0 * off1[1] - c1 * off1[0] + c5;
          ang[0] -= c0; ang[1] -= c1; ang[2] -= c2;
        }
      }
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (k < dim) EF(J)[(r0 + k) * nv + d] = axes[3 * k] * lin[0] + axes[3 * k + 1] * lin[1] + axes[3 * k + 2] * lin[2];
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (3 + k < dim) EF(J)[(r0 + 3 + k) * nv + d] = axes[3 * k] * ang[0] + axes[3 * k + 1] * ang[1] + axes[3 * k + 2] * ang[2];
    }
    #pragma unroll 1
    PFOR(k, dim) tag_row(e, r0 + k, CT_CONTACT, c, 0);
  }
  EH.ne = ne; EH.nf = nf; EH.nl = nl; EH.nefc = nefc;
  WSYNC();
  finalize_rows_w(e);
}

// ---------------------------------------------------------------------------------- Newton
// per-row (or per-contact) cost/force/state at jar; returns this lane's partial cost
MGS_DEVN real constraint_update_w(Env &e) {
  real cost = 0;
  #pragma unroll 1
  PFOR(i, EH.nefc) {
    int type = EFC_TYPE(i);
    real D = EF(efc_D)[i], jar = EF(efc_jar)[i];
    if (type == CT_EQUALITY) {
      EF(efc_force)[i] = -D * jar; cost += R_(0.5) * D * jar * jar; EFC_SET_STATE(i, ST_QUADRATIC);
    } else if (type == CT_FRICTION_DOF) {
      real f = EF(efc_aux)[i], R = EF(efc_R)[i];
      if (jar <= -R * f) { EF(efc_force)[i] = f; cost += R_(-0.5) * R * f * f - f * jar; EFC_SET_STATE(i, ST_LINEARNEG); }
      else if (jar >= R * f) { EF(efc_force)[i] = -f; cost += R_(-0.5) * R * f * f + f * jar; EFC_SET_STATE(i, ST_LINEARPOS); }
      else { EF(efc_force)[i] = -D * jar; cost += R_(0.5) * D * jar * jar; EFC_SET_STATE(i, ST_QUADRATIC); }
    } else if (type == CT_LIMIT) {
      if (jar < 0) { EF(efc_force)[i] = -D * jar; cost += R_(0.5) * D * jar * jar; EFC_SET_STATE(i, ST_QUADRATIC); }
      else { EF(efc_force)[i] = 0; EFC_SET_STATE(i, ST_SATISFIED); }
    } else {
      int c = EFC_ID(i);
      if (IARR(EF(con_efc))[c] != i) continue;  // the contact's first row does the whole cone
      int p = IARR(EF(con_pair))[c], dim = LDG(MD.pair_condim + p);
      if (dim < 3 || !MD.cone_elliptic) {
        if (jar < 0) { EF(efc_force)[i] = -D * jar; cost += R_(0.5) * D * jar * jar; EFC_SET_STATE(i, ST_QUADRATIC); }
        else { EF(efc_force)[i] = 0; EFC_SET_STATE(i, ST_SATISFIED); }
#pragma unroll
        for (int j = 1; j < 4; j++) if (j < dim) { EF(efc_force)[i + j] = 0; EFC_SET_STATE(i + j, ST_SATISFIED); }
        continue;
      }
      // condim is 1, 3 or 4 (checked when the model is built): all loops over the cone's rows are unrolled to 4
      real mu = EF(con_mu)[c], fr[3], U[4] = {0, 0, 0, 0}, T = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) fr[k] = LDG(MD.pair_friction + 5 * p + k);
      U[0] = jar * mu;
#pragma unroll
      for (int j = 1; j < 4; j++) if (j < dim) { U[j] = EF(efc_jar)[i + j] * fr[j - 1]; T += U[j] * U[j]; }
      real N = U[0];
      T = sqrt(T);
      int st;
      if (N >= mu * T || (T <= 0 && N >= 0)) {
#pragma unroll
        for (int j = 0; j < 4; j++) if (j < dim) EF(efc_force)[i + j] = 0;
        st = ST_SATISFIED;
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
#pragma unroll
        for (int j = 0; j < 4; j++) if (j < dim) {
          real x = EF(efc_jar)[i + j], Dj = EF(efc_D)[i + j];
          EF(efc_force)[i + j] = -Dj * x;
          cost += R_(0.5) * Dj * x * x;
        }
        st = ST_QUADRATIC;
      } else {
        real Dm = D / fmax(MGS_MINVAL, mu * mu * (1 + mu * mu)), NmT = N - mu * T;
        cost += R_(0.5) * Dm * NmT * NmT;
        real f0 = -Dm * NmT * mu;
        EF(efc_force)[i] = f0;
        const real f0T = -f0 / T;
#pragma unroll
        for (int j = 1; j < 4; j++) if (j < dim) EF(efc_force)[i + j] = f0T * U[j] * fr[j - 1];
        st = ST_CONE;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) if (j < dim) EFC_SET_STATE(i + j, st);
    }
  }
  return cost;
}

// derivative pair of the cost along the search direction at alpha (this lane's partial sums)
MGS_DEVN void ls_eval_w(const Env &e, real alpha, real *d1, real *d2) {
  real a = 0, h = 0;
  #pragma unroll 1
  PFOR(i, EH.nefc) {
    int type = EFC_TYPE(i);
    real D = EF(efc_D)[i], jv = EF(efc_jv)[i], x = EF(efc_jar)[i] + alpha * jv;
    if (type == CT_EQUALITY) { a += D * x * jv; h += D * jv * jv; }
    else if (type == CT_FRICTION_DOF) {
      real f = EF(efc_aux)[i], R = EF(efc_R)[i];
      if (x <= -R * f) a += -f * jv;
      else if (x >= R * f) a += f * jv;
      else { a += D * x * jv; h += D * jv * jv; }
    } else if (type == CT_LIMIT) {
      if (x < 0) { a += D * x * jv; h += D * jv * jv; }
    } else {
      int c = EFC_ID(i);
      if (IARR(EF(con_efc))[c] != i) continue;
      int p = IARR(EF(con_pair))[c], dim = LDG(MD.pair_condim + p);
      if (dim < 3 || !MD.cone_elliptic) { if (x < 0) { a += D * x * jv; h += D * jv * jv; } continue; }
      real mu = EF(con_mu)[c], T = 0, UV = 0, VV = 0;
      real N = x * mu, N1 = jv * mu;
#pragma unroll
      for (int j = 1; j < 4; j++) if (j < dim) {
        real fj = LDG(MD.pair_friction + 5 * p + j - 1);
        real Uj = (EF(efc_jar)[i + j] + alpha * EF(efc_jv)[i + j]) * fj, Vj = EF(efc_jv)[i + j] * fj;
        T += Uj * Uj; UV += Uj * Vj; VV += Vj * Vj;
      }
      T = sqrt(T);
      if (N >= mu * T || (T <= 0 && N >= 0)) {
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
#pragma unroll
        for (int j = 0; j < 4; j++) if (j < dim) {
          real xj = EF(efc_jar)[i + j] + alpha * EF(efc_jv)[i + j], vj = EF(efc_jv)[i + j], Dj = EF(efc_D)[i + j];
          a += Dj * xj * vj; h += Dj * vj * vj;
        }
      } else {
        real Dm = D / fmax(MGS_MINVAL, mu * mu * (1 + mu * mu)), NmT = N - mu * T;
        real T1 = UV / T, T2 = VV / T - UV * UV / (T * T * T), s = N1 - mu * T1;
        a += Dm * NmT * s;
        h += Dm * (s * s - NmT * mu * T2);
      }
    }
  }
  *d1 = a; *d2 = h;
}

// jar = J*qacc - aref ; Ma = M*qacc
MGS_DEVN void eval_point_w(Env &e, const real *qacc) {
  const int nv = MD.nv;
  #pragma unroll 1
  PFOR(i, EH.nefc) {
    acc_jar_t t = -(acc_jar_t)EF(efc_aref)[i];
    MGS_UNROLL_INNER
    for (int d = 0; d < nv; d++) t += (acc_jar_t)EF(J)[i * nv + d] * (acc_jar_t)qacc[d];
    EF(efc_jar)[i] = (real)t;
  }
  #pragma unroll 1
  PFOR(i, nv) {
    int ro, lo, tn;
    blk_row(i, ro, lo, tn);
    const int hi = lo + tn;
    const real *Mrow = EF(M) + ro;
    acc_jar_t t = 0;
    MGS_UNROLL_INNER
    for (int j = lo; j < hi; j++) t += (acc_jar_t)Mrow[j] * (acc_jar_t)qacc[j];
    EF(Ma)[i] = (real)t;
  }
  WSYNC();
}
MGS_DEV real gauss_cost_w(const Env &e, const real *qacc) {
  real g = 0;
  #pragma unroll 1
  PFOR(d, MD.nv) g += (EF(Ma)[d] - EF(qfrc_smooth)[d]) * (qacc[d] - EF(qacc_smooth)[d]);
  return R_(0.5) * g;
}

// dim x dim Hessian block of a contact in the cone's middle zone (recomputed from jar: not stored)
MGS_DEVN void cone_hessian(const Env &e, int c, int i, int dim, real *h) {
  const int p = IARR(EF(con_pair))[c];
  real mu = EF(con_mu)[c], fr[5], U[4], scl[4], T = 0;
  for (int k = 0; k < 5; k++) fr[k] = LDG(MD.pair_friction + 5 * p + k);
  U[0] = EF(efc_jar)[i] * mu;
  scl[0] = mu;
  for (int j = 1; j < 4; j++) {
    if (j < dim) { U[j] = EF(efc_jar)[i + j] * fr[j - 1]; T += U[j] * U[j]; scl[j] = fr[j - 1]; } else { U[j] = 0; scl[j] = 0; }
  }
  real N = U[0];
  T = sqrt(T);
  real Dm = EF(efc_D)[i] / fmax(MGS_MINVAL, mu * mu * (1 + mu * mu));
  real muNT3 = mu * N / (T * T * T), dg = mu * mu - mu * N / T;
  for (int j = 0; j < 4; j++)
    for (int k = 0; k < 4; k++) {
      real v;
      if (j == 0 && k == 0) v = 1;
      else if (j == 0) v = -mu * U[k] / T;
      else if (k == 0) v = -mu * U[j] / T;
      else v = muNT3 * U[j] * U[k] + (j == k ? dg : R_(0.0));
      h[j * 4 + k] = (j < dim && k < dim) ? v * Dm * scl[j] * scl[k] : R_(0.0);
    }
}

#ifdef MGS_WIDE
// Rows of the constraint Jacobian touch at most two kinematic trees (a contact between two objects: 12 of the 94 columns of the
// config-5 scene), so H = M + J' D J only receives row i in the tile (tree, tree') of its two trees.  Rows are bucketed by their
// tree pair once per step (deterministic counting sort: key = lower tree * ntree + upper tree); an entry (a, b) of H then walks
// only the rows of the bucket(s) that can contribute.  Dense, the product was 19 % of the config-5 step (stage clocks).
#define MGS_SPARSE_H_MAXTREE 16
MGS_DEV int row_key(const Env &e, int r) {
  const int type = EFC_TYPE(r), id = EFC_ID(r);
  int ta = -1, tb = -1;
  if (type == CT_EQUALITY) {
    if (LDG(MD.eq_type + id) == EQ_JOINT) {
      const int j2 = LDG(MD.eq_obj2id + id);
      ta = LDG(MD.dof_treeid + LDG(MD.jnt_dofadr + LDG(MD.eq_obj1id + id)));
      if (j2 >= 0) tb = LDG(MD.dof_treeid + LDG(MD.jnt_dofadr + j2));
    } else { ta = LDG(MD.body_treeid + LDG(MD.eq_obj1id + id)); tb = LDG(MD.body_treeid + LDG(MD.eq_obj2id + id)); }
  } else if (type == CT_FRICTION_DOF) ta = LDG(MD.dof_treeid + id);
  else if (type == CT_LIMIT) ta = LDG(MD.dof_treeid + LDG(MD.jnt_dofadr + id));
  else {
    const int p = IARR(EF(con_pair))[id];
    ta = LDG(MD.body_treeid + LDG(MD.cgeom_bodyid + LDG(MD.pair_geom1 + p)));
    tb = LDG(MD.body_treeid + LDG(MD.cgeom_bodyid + LDG(MD.pair_geom2 + p)));
  }
  if (ta < 0) { ta = tb; tb = -1; }
  if (ta < 0) return 0;  // a row between two static bodies has no dofs: any bucket will do (its Jacobian row is zero)
  if (tb < 0) tb = ta;
  if (ta > tb) { const int t = ta; ta = tb; tb = t; }
  return ta * MD.ntree + tb;
}
MGS_DEVN void bucket_rows_w(Env &e) {
  const int nefc = EH.nefc, nkey = MD.ntree * MD.ntree;
  int *key = IARR(EF(efc_key)), *list = IARR(EF(efc_list)), *start = IARR(EF(nsS));
  #pragma unroll 1
  PFOR(r, nefc) key[r] = row_key(e, r);
  WSYNC();
  // (nkey <= LANES: one key per thread) rows per key, then the exclusive scan
  int cnt = 0;
  if (MGS_LANE < nkey) {
    #pragma unroll 4
    for (int r = 0; r < nefc; r++) cnt += (key[r] == MGS_LANE);
  }
  int total;
  const int off = wscan_excl(cnt, &total);
  if (MGS_LANE < nkey) start[MGS_LANE] = off;
  if (MGS_LANE == 0) start[nkey] = total;
  WSYNC();
  // position of row r = start of its bucket + number of earlier rows with the same key (row order is kept inside a bucket)
  #pragma unroll 1
  PFOR(r, nefc) {
    const int k = key[r];
    int pos = start[k];
    #pragma unroll 4
    for (int q = 0; q < r; q++) pos += (key[q] == k);
    list[pos] = r;
  }
  WSYNC();
}
MGS_DEV real hess_bucket(const Env &e, int k, int a, int b, const real *W, real s) {
  const int *list = IARR(EF(efc_list)), *start = IARR(EF(nsS));
  const int q1 = start[k + 1], nv = MD.nv;
  #pragma unroll 2
  for (int q = start[k]; q < q1; q++) {
    const int i = list[q];
    s += W[i] * EF(J)[i * nv + a] * EF(J)[i * nv + b];
  }
  return s;
}
#endif

MGS_DEVN void newton_hessian_w(Env &e) {
  const int nv = MD.nv, npairs = nv * (nv + 1) / 2, nefc = EH.nefc;
  // row weights (D for rows in their quadratic zone, else 0) make the accumulation loop branch-free; efc_jv is
  // free here (it is rewritten right after the Hessian solve)
  real *W = EF(efc_jv);
  #pragma unroll 1
  PFOR(i, nefc) W[i] = (EFC_STATE(i) == ST_QUADRATIC) ? EF(efc_D)[i] : R_(0.0);
  WSYNC();
#ifdef MGS_WIDE
  if (MD.ntree <= MGS_SPARSE_H_MAXTREE) {
    const int T = MD.ntree;
    #pragma unroll 1
    PFOR(idx, npairs) {
      const int ab = LDG(MD.tri_ab + idx), a = ab >> 8, b = ab & 255;
      const int madr = LDG(MD.tri_madr + idx);
      const int ta = LDG(MD.dof_treeid + a), tb = LDG(MD.dof_treeid + b);  // ta >= tb (a >= b, trees are contiguous dof ranges)
      real s = madr >= 0 ? EF(M)[madr] : R_(0.0);
      if (ta != tb) s = hess_bucket(e, tb * T + ta, a, b, W, s);
      else {
        #pragma unroll 1
        for (int x = 0; x < T; x++) s = hess_bucket(e, x <= ta ? x * T + ta : ta * T + x, a, b, W, s);
      }
      EF(H)[a * nv + b] = s;
    }
  } else
#endif
  #pragma unroll 1
  PFOR(idx, npairs) {
    const int ab = LDG(MD.tri_ab + idx), a = ab >> 8, b = ab & 255;  // lower-triangle index table
    const int madr = LDG(MD.tri_madr + idx);  // M is block diagonal: zero between different kinematic trees
    acc_hess_t s = madr >= 0 ? EF(M)[madr] : R_(0.0);
    const real *Ja = EF(J) + a, *Jb = EF(J) + b;
    MGS_UNROLL_INNER
    for (int i = 0; i < nefc; i++) s += (acc_hess_t)W[i] * (acc_hess_t)Ja[i * nv] * (acc_hess_t)Jb[i * nv];
    EF(H)[a * nv + b] = (real)s;
  }
  // contacts on the cone surface (usually few): every lane rebuilds the small block, lanes split (a,b)
#ifdef MGS_WIDE
  WSYNC();  // the entries change hands below (per-contact enumeration): the products above must have landed
#endif
  #pragma unroll 1
  for (int c = 0; c < EH.ncon; c++) {
    const int i = IARR(EF(con_efc))[c];
    if (i < 0 || EFC_STATE(i) != ST_CONE) continue;
    const int dim = LDG(MD.pair_condim + IARR(EF(con_pair))[c]);
    real h[16];
    cone_hessian(e, c, i, dim, h);
#ifdef MGS_WIDE
    {
      // only the entries (a, b) whose two dofs lie in the contact's (at most two) kinematic trees can change: enumerate the lower
      // triangle of that index set instead of all nv (nv + 1) / 2 entries (config 5: <= 820 of 4465, usually 78 or 21)
      const int p = IARR(EF(con_pair))[c];
      int t1 = LDG(MD.body_treeid + LDG(MD.cgeom_bodyid + LDG(MD.pair_geom1 + p))), t2 = LDG(MD.body_treeid + LDG(MD.cgeom_bodyid + LDG(MD.pair_geom2 + p)));
      if (t1 < 0) { t1 = t2; t2 = -1; }
      if (t1 >= 0) {
        if (t2 == t1) t2 = -1;
        if (t2 >= 0 && t2 < t1) { const int t = t1; t1 = t2; t2 = t; }
        // first dof / size of the two trees: found from the contact's own Jacobian support is not needed - the tree tables give them
        int lo1 = 0, n1 = 0, lo2 = 0, n2 = 0;
        #pragma unroll 1
        for (int d = 0; d < nv; d += LDG(MD.dof_treenum + d)) {
          const int t = LDG(MD.dof_treeid + d);
          if (t == t1) { lo1 = d; n1 = LDG(MD.dof_treenum + d); }
          if (t == t2) { lo2 = d; n2 = LDG(MD.dof_treenum + d); }
        }
        const int m = n1 + n2, ne2 = m * (m + 1) / 2;
        #pragma unroll 1
        PFOR(e2, ne2) {
          int u = (int)((sqrtf(8.0f * (float)e2 + 1.0f) - 1.0f) * 0.5f);
          while (u * (u + 1) / 2 > e2) u--;
          while ((u + 1) * (u + 2) / 2 <= e2) u++;
          const int v = e2 - u * (u + 1) / 2;
          const int a = u < n1 ? lo1 + u : lo2 + (u - n1), b = v < n1 ? lo1 + v : lo2 + (v - n1);
          real s = 0;
          #pragma unroll 1
          for (int j = 0; j < dim; j++) {
            const real ja = EF(J)[(i + j) * nv + a];
            if (ja == 0) continue;
            real t = 0;
            for (int k = 0; k < 4; k++) t += (k < dim) ? h[j * 4 + k] * EF(J)[(i + k) * nv + b] : R_(0.0);
            s += ja * t;
          }
          EF(H)[a * nv + b] += s;
        }
      }
      WSYNC();  // the next contact enumerates the entries differently: another thread may own (a, b) then
    }
#else
    #pragma unroll 1
    PFOR(idx, npairs) {
      const int ab = LDG(MD.tri_ab + idx), a = ab >> 8, b = ab & 255;
      real s = 0;
      #pragma unroll 1
      for (int j = 0; j < dim; j++) {
        const real ja = EF(J)[(i + j) * nv + a];
        if (ja == 0) continue;
        real t = 0;
        for (int k = 0; k < 4; k++) t += (k < dim) ? h[j * 4 + k] * EF(J)[(i + k) * nv + b] : R_(0.0);
        s += ja * t;
      }
      EF(H)[a * nv + b] += s;
    }
#endif
  }
  WSYNC();
  MGS_CLK(4);
  chol_factor_w(EF(H), nv, 0, EF(nsB));  // (nsB: 3 nv words of noslip scratch, free during the Newton solve)
}

MGS_DEVN void solve_newton_w(Env &e) {
  const int nv = MD.nv;
  const real scale = R_(1.0) / (MD.meaninertia * (nv > 1 ? nv : 1));
  EH.niter = 0;
  // warm start: cheaper of qacc_warmstart and qacc_smooth.  The warm start is evaluated LAST: it wins on almost every
  // step, and then the row states / forces / jar / Ma left behind are already those of the starting point.
  eval_point_w(e, EF(qacc_smooth));
  const real cs = wsum(constraint_update_w(e) + gauss_cost_w(e, EF(qacc_smooth)));
  WSYNC();
  eval_point_w(e, EF(qacc_ws));
  real cost = wsum(constraint_update_w(e) + gauss_cost_w(e, EF(qacc_ws)));
  WSYNC();
  const int use_ws = cost < cs;
  #pragma unroll 1
  PFOR(d, nv) EF(qacc)[d] = use_ws ? EF(qacc_ws)[d] : EF(qacc_smooth)[d];
  WSYNC();
  if (!use_ws) {
    eval_point_w(e, EF(qacc));
    cost = wsum(constraint_update_w(e) + gauss_cost_w(e, EF(qacc)));
    WSYNC();
  }
  #pragma unroll 1
  for (int iter = 0; iter < MD.iterations; iter++) {
    real gn = 0;
    #pragma unroll 1
    PFOR(d, nv) {
      acc_grad_t tt = (acc_grad_t)EF(Ma)[d] - (acc_grad_t)EF(qfrc_smooth)[d];
      MGS_UNROLL_INNER
      for (int i = 0; i < EH.nefc; i++) tt -= (acc_grad_t)EF(J)[i * nv + d] * (acc_grad_t)EF(efc_force)[i];
      const real t = (real)tt;
      EF(grad)[d] = t;
      EF(search)[d] = t;
      gn += t * t;
    }
    gn = wsum(gn);
    // fp32: the gradient cannot be resolved below ~eps * |force terms|; floor the tolerance accordingly.  (A per-step
    // rounding-noise estimate of the gradient, K * eps * |terms| with K = 2..8, was tried as a further floor: it never
    // binds - the warm starts that iterate are far above the noise - so it is not kept.)
#ifndef MGS_TOL_FLOOR_K
#define MGS_TOL_FLOOR_K 20.0
#endif
    real tol_eff = fmax(MD.tolerance, (real)MGS_TOL_FLOOR_K * (real)REAL_EPS * scale * fabs(cost));
    // (MuJoCo tests the gradient only after an iteration; a warm start that already meets the tolerance skips
    // the Hessian here - the two answers differ by less than the solver tolerance.)
#ifdef MGS_NO_EARLY_GRAD_EXIT
    if (iter > 0 && scale * sqrt(gn) < tol_eff) break;
#else
    if (scale * sqrt(gn) < tol_eff) break;
#endif
    MGS_CLK(3);
#ifdef MGS_WIDE
    if (iter == 0 && MD.ntree <= MGS_SPARSE_H_MAXTREE) bucket_rows_w(e);
#endif
    newton_hessian_w(e);
    MGS_CLK(5);
    chol_solve_w(EF(H), EF(search), nv, 0);
    MGS_CLK(6);
    #pragma unroll 1
    PFOR(d, nv) EF(search)[d] = -EF(search)[d];
    WSYNC();
    matvec_w(EF(Mv), EF(M), EF(search), nv);
    real g1 = 0, g2 = 0, sn = 0;
    #pragma unroll 1
    PFOR(d, nv) {
      g1 += EF(search)[d] * (EF(Ma)[d] - EF(qfrc_smooth)[d]);
      g2 += EF(search)[d] * EF(Mv)[d];
      sn += EF(search)[d] * EF(search)[d];
    }
    #pragma unroll 1
    PFOR(i, EH.nefc) {
      acc_ls_t t = 0;
      MGS_UNROLL_INNER
      for (int d = 0; d < nv; d++) t += (acc_ls_t)EF(J)[i * nv + d] * (acc_ls_t)EF(search)[d];
      EF(efc_jv)[i] = (real)t;
    }
    g1 = wsum(g1); g2 = wsum(g2); sn = wsum(sn);
    WSYNC();
    real d1, d2;
    ls_eval_w(e, 0, &d1, &d2);
    const real a0 = wsum(d1);  // constraint part of the derivative at alpha = 0
    d1 = a0 + g1; d2 = wsum(d2) + g2;
    if (!(d1 < 0) || sn < R_(1e-30)) break;
    const real d1_0 = d1;
    // Stopping threshold on |d1|: MuJoCo's gtol, floored by what the arithmetic can resolve.  Near the optimum
    // (warm-started steps) the Gauss part g1 and the constraint part a0 cancel, so the rounding noise of d1 is
    // ~eps * (|g1| + |a0|) >> eps * |d1_0|; without this term the search ran to bracket collapse on the GPU
    // (ncu r1_h: 16 derivative evaluations per Newton iteration, 11 % of all issued instructions).
    real gtol = fmax(MD.tolerance * MD.ls_tolerance * sqrt(sn) / scale, R_(50.0) * (real)REAL_EPS * fabs(d1_0));
#ifndef MGS_NO_LS_NOISE_FLOOR
    gtol = fmax(gtol, R_(16.0) * (real)REAL_EPS * (fabs(g1) + fabs(a0)));
#endif
    // Exact line search on the convex, piecewise-smooth 1-D cost: zero of its monotone derivative.
    // Newton steps while they stay inside the bracket [lo, hi]; otherwise the secant of the end
    // derivatives; bisection when the same end moved twice in a row or two iterations did not halve the
    // bracket (kinks where the second derivative jumps - cone zone changes - make plain Newton bounce).
    // Without convergence the answer is `lo` (negative derivative: by convexity the cost went down).
    real alpha = -d1 / d2, lo = 0, hi = -1, dlo = d1, dhi = 0, w1 = -1, w2 = -1;
    int last_side = 0, same_side = 0, converged = 0;
    #pragma unroll 1
    for (int it = 0; it < MD.ls_iterations; it++) {
      ls_eval_w(e, alpha, &d1, &d2);
      d1 = wsum(d1) + g1 + alpha * g2; d2 = wsum(d2) + g2;
      if (fabs(d1) < gtol) { converged = 1; break; }
      const int side = d1 < 0 ? -1 : 1;
      same_side = (side == last_side) ? same_side + 1 : 0;
      last_side = side;
      if (d1 < 0) { lo = alpha; dlo = d1; } else { hi = alpha; dhi = d1; }
      real an = alpha - d1 / d2;
      if (hi < 0) { if (!(an > alpha)) an = 2 * alpha; }
      else {
        const real w = hi - lo;
        if (!(an > lo && an < hi)) an = (same_side >= 1) ? R_(0.5) * (lo + hi) : lo - dlo * (hi - lo) / (dhi - dlo);
        if (!(an > lo && an < hi) || (w2 > 0 && w > R_(0.5) * w2)) an = R_(0.5) * (lo + hi);
        w2 = w1; w1 = w;
        if (w <= R_(8.0) * (real)REAL_EPS * hi) break;
      }
      alpha = an;
    }
    if (!converged) alpha = lo > 0 ? lo : alpha;
    if (!(alpha > 0)) break;
    #pragma unroll 1
    PFOR(d, nv) { EF(qacc)[d] += alpha * EF(search)[d]; EF(Ma)[d] += alpha * EF(Mv)[d]; }
    #pragma unroll 1
    PFOR(i, EH.nefc) EF(efc_jar)[i] += alpha * EF(efc_jv)[i];
    WSYNC();
    real oldcost = cost;
    cost = wsum(constraint_update_w(e) + gauss_cost_w(e, EF(qacc)));
    WSYNC();
    if (cost > oldcost) {
      // never accept an uphill step: back to the previous iterate (forces consistent with it) and stop
      #pragma unroll 1
      PFOR(d, nv) { EF(qacc)[d] -= alpha * EF(search)[d]; EF(Ma)[d] -= alpha * EF(Mv)[d]; }
      #pragma unroll 1
      PFOR(i, EH.nefc) EF(efc_jar)[i] -= alpha * EF(efc_jv)[i];
      WSYNC();
      cost = wsum(constraint_update_w(e) + gauss_cost_w(e, EF(qacc)));
      WSYNC();
      break;
    }
    EH.niter = iter + 1;
    tol_eff = fmax(MD.tolerance, (real)MGS_TOL_FLOOR_K * (real)REAL_EPS * scale * fabs(cost));
#ifndef MGS_NO_IMPROVEMENT_EXIT
    if (scale * (oldcost - cost) < tol_eff) break;
#endif
  }
  WSYNC();
}

// min 0.5 x'Ax + b'x  s.t. sum (x_j/d_j)^2 <= r^2 , N in {2,3}: Newton on the multiplier.
// Fully unrolled for a compile-time N and inlined into its caller: every array below is indexed with
// compile-time constants and lives in registers (the pointer-argument version kept A, b, d and the result
// in local memory: ncu r1_h showed 3.6 % of all stall samples as long-scoreboard waits on its first loads).
template <int N>
MGS_DEV void qcqp_small(real *res, const real *A, const real *b, const real *d, real r) {
  real As[N * N], bs[N], v[N], la = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    bs[i] = b[i] * d[i];
    v[i] = 0;
#pragma unroll
    for (int j = 0; j < N; j++) As[i * N + j] = A[i * N + j] * d[i] * d[j];
  }
  #pragma unroll 1
  for (int iter = 0; iter < 20; iter++) {
    real P[N * N];
    if (N == 2) {
      real a00 = As[0] + la, a01 = As[1], a11 = As[3] + la, det = a00 * a11 - a01 * a01;
      if (det < R_(1e-10)) {
#pragma unroll
        for (int i = 0; i < N; i++) res[i] = 0;
        return;
      }
      real id = R_(1.0) / det;
      P[0] = a11 * id; P[1] = -a01 * id; P[2] = -a01 * id; P[3] = a00 * id;
    } else {
      real a00 = As[0] + la, a01 = As[1], a02 = As[2], a11 = As[N + 1] + la, a12 = As[N + 2], a22 = As[2 * N + 2] + la;
      real c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
      real det = a00 * c00 + a01 * c01 + a02 * c02;
      if (det < R_(1e-10)) {
#pragma unroll
        for (int i = 0; i < N; i++) res[i] = 0;
        return;
      }
      real id = R_(1.0) / det;
      P[0] = c00 * id; P[1] = c01 * id; P[2] = c02 * id;
      P[N] = P[1]; P[N + 1] = (a00 * a22 - a02 * a02) * id; P[N + 2] = (a01 * a02 - a00 * a12) * id;
      P[2 * N] = P[2]; P[2 * N + 1] = P[N + 2]; P[2 * N + 2] = (a00 * a11 - a01 * a01) * id;
    }
    real val = -r * r;
#pragma unroll
    for (int i = 0; i < N; i++) {
      real s = 0;
#pragma unroll
      for (int j = 0; j < N; j++) s -= P[i * N + j] * bs[j];
      v[i] = s; val += s * s;
    }
    if (val < R_(1e-10)) break;
    real deriv = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      real s = 0;
#pragma unroll
      for (int j = 0; j < N; j++) s += P[i * N + j] * v[j];
      deriv -= 2 * v[i] * s;
    }
    real delta = -val / deriv;
    if (delta < R_(1e-10)) break;
    la += delta;
  }
#pragma unroll
  for (int i = 0; i < N; i++) res[i] = v[i] * d[i];
}

// One Gauss-Seidel update of the N friction rows (N = 2: condim 3; N = 3: condim >= 4, torsion included) of
// contact c (first row i).  Residuals: lanes split the dofs, butterfly sums; the small QCQP runs replicated in
// registers on every lane; then wvec += M^-1 J_c' delta (M^-1 is block diagonal per kinematic tree).
// Returns the cost decrease contribution (`change`, <= 0 when accepted).
template <int N>
MGS_DEV real noslip_contact_w(Env &e, int c, int i, int p, const real *AC, real *T, const real *Bc) {
  const int nv = MD.nv;
  real res[N], Ac[N * N], old[N], bc[N], v[N], delta[N], fr[N];
#pragma unroll
  for (int j = 0; j < N; j++) res[j] = 0;
  #pragma unroll 1
  PFOR(d, nv) {
    const real u = EF(qacc_smooth)[d] + EF(wvec)[d];
#pragma unroll
    for (int j = 0; j < N; j++) res[j] += EF(J)[(i + 1 + j) * nv + d] * u;
  }
  {
    // (one combined reduction: in the env-per-CTA variant every reduction costs two CTA barriers)
    real r0 = res[0], r1 = res[1], r2 = N > 2 ? res[N - 1] : R_(0.0);
    wsum3(r0, r1, r2);
    res[0] = r0; res[1] = r1;
    if (N > 2) res[N - 1] = r2;
    int q = 0;
#pragma unroll
    for (int j = 0; j < N; j++) {
      res[j] = res[j] - EF(efc_aref)[i + 1 + j];
      old[j] = EF(efc_force)[i + 1 + j];
      fr[j] = LDG(MD.pair_friction + 5 * p + j);
#pragma unroll
      for (int k = j; k < N; k++) { Ac[j * N + k] = Ac[k * N + j] = AC[6 * c + q]; q++; }
    }
  }
#pragma unroll
  for (int j = 0; j < N; j++) {
    bc[j] = res[j];
#pragma unroll
    for (int k = 0; k < N; k++) bc[j] -= Ac[j * N + k] * old[k];
  }
  const real fnorm = EF(efc_force)[i];
  if (fnorm < MGS_MINVAL) {
#pragma unroll
    for (int j = 0; j < N; j++) v[j] = 0;
  } else qcqp_small<N>(v, Ac, bc, fr, fnorm);
  real change = 0;
#pragma unroll
  for (int j = 0; j < N; j++) delta[j] = v[j] - old[j];
#pragma unroll
  for (int j = 0; j < N; j++) {
    change += delta[j] * res[j];
#pragma unroll
    for (int k = 0; k < N; k++) change += R_(0.5) * delta[j] * Ac[j * N + k] * delta[k];
  }
  if (change > R_(1e-10)) {
#pragma unroll
    for (int j = 0; j < N; j++) { v[j] = old[j]; delta[j] = 0; }
    change = 0;
  }
  // w += M^-1 (J_c' delta): with the precomputed B_c = M^-1 J_c' (Bc != 0) three FMAs per dof
#ifdef MGS_WIDE
  WSYNC();  // every warp has read `old` / the residual inputs before the forces are overwritten
#endif
  if (MGS_LANE == 0) {
#pragma unroll
    for (int j = 0; j < N; j++) EF(efc_force)[i + 1 + j] = v[j];
  }
  if (Bc) {
    #pragma unroll 1
    PFOR(d, nv) {
      real t = 0;
#pragma unroll
      for (int j = 0; j < N; j++) t += Bc[j * nv + d] * delta[j];
      EF(wvec)[d] += t;
    }
    WSYNC();
    return change;
  }
  #pragma unroll 1
  PFOR(d, nv) {
    real t = 0;
#pragma unroll
    for (int j = 0; j < N; j++) t += EF(J)[(i + 1 + j) * nv + d] * delta[j];
    T[d] = t;
  }
  WSYNC();
  #pragma unroll 1
  PFOR(d, nv) {
    int ro, lo, tn;
    blk_row(d, ro, lo, tn);
    const int hi = lo + tn;
    real t = 0;
    const real *Mrow = EF(Minv) + ro;
    MGS_UNROLL_INNER
    for (int k = lo; k < hi; k++) t += Mrow[k] * T[k];
    EF(wvec)[d] += t;
  }
  WSYNC();
  return change;
}

#ifdef MGS_WIDE
// ENV-PER-CTA variant: the same Gauss-Seidel update of one contact, executed by ONE WARP (shuffle reductions, __syncwarp, no CTA
// barrier) and restricted to the dofs of the contact's own (at most two) kinematic trees S = [lo1, lo1 + n1) u [lo2, lo2 + n2) - its
// Jacobian rows and B_c = M^-1 J_c' are zero elsewhere.  Contacts that share no tree commute exactly, so solve_noslip_w runs them
// concurrently on different warps (level schedule below); this routine therefore touches nothing outside S.
template <int N>
MGS_DEV real noslip_contact_warp(Env &e, int c, int i, int p, const real *AC, real *T, const real *Bc, int lo1, int n1, int lo2, int n2) {
  const int nv = MD.nv, l32 = threadIdx.x & 31, m = n1 + n2;
  real res[N], Ac[N * N], old[N], bc[N], v[N], delta[N], fr[N];
#pragma unroll
  for (int j = 0; j < N; j++) res[j] = 0;
  #pragma unroll 1
  for (int u = l32; u < m; u += 32) {
    const int d = u < n1 ? lo1 + u : lo2 + (u - n1);
    const real w = EF(qacc_smooth)[d] + EF(wvec)[d];
#pragma unroll
    for (int j = 0; j < N; j++) res[j] += EF(J)[(i + 1 + j) * nv + d] * w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int j = 0; j < N; j++) res[j] += __shfl_xor_sync(0xffffffffu, res[j], o);
  {
    int q = 0;
#pragma unroll
    for (int j = 0; j < N; j++) {
      res[j] = res[j] - EF(efc_aref)[i + 1 + j];
      old[j] = EF(efc_force)[i + 1 + j];
      fr[j] = LDG(MD.pair_friction + 5 * p + j);
#pragma unroll
      for (int k = j; k < N; k++) { Ac[j * N + k] = Ac[k * N + j] = AC[6 * c + q]; q++; }
    }
  }
#pragma unroll
  for (int j = 0; j < N; j++) {
    bc[j] = res[j];
#pragma unroll
    for (int k = 0; k < N; k++) bc[j] -= Ac[j * N + k] * old[k];
  }
  const real fnorm = EF(efc_force)[i];
  if (fnorm < MGS_MINVAL) {
#pragma unroll
    for (int j = 0; j < N; j++) v[j] = 0;
  } else qcqp_small<N>(v, Ac, bc, fr, fnorm);
  real change = 0;
#pragma unroll
  for (int j = 0; j < N; j++) delta[j] = v[j] - old[j];
#pragma unroll
  for (int j = 0; j < N; j++) {
    change += delta[j] * res[j];
#pragma unroll
    for (int k = 0; k < N; k++) change += R_(0.5) * delta[j] * Ac[j * N + k] * delta[k];
  }
  if (change > R_(1e-10)) {
#pragma unroll
    for (int j = 0; j < N; j++) { v[j] = old[j]; delta[j] = 0; }
    change = 0;
  }
  __syncwarp();  // every lane has read the old forces
  if (l32 == 0) {
#pragma unroll
    for (int j = 0; j < N; j++) EF(efc_force)[i + 1 + j] = v[j];
  }
  if (Bc) {
    #pragma unroll 1
    for (int u = l32; u < m; u += 32) {
      const int d = u < n1 ? lo1 + u : lo2 + (u - n1);
      real t = 0;
#pragma unroll
      for (int j = 0; j < N; j++) t += Bc[j * nv + d] * delta[j];
      EF(wvec)[d] += t;
    }
    __syncwarp();
    return change;
  }
  #pragma unroll 1
  for (int u = l32; u < m; u += 32) {
    const int d = u < n1 ? lo1 + u : lo2 + (u - n1);
    real t = 0;
#pragma unroll
    for (int j = 0; j < N; j++) t += EF(J)[(i + 1 + j) * nv + d] * delta[j];
    T[d] = t;
  }
  __syncwarp();
  #pragma unroll 1
  for (int u = l32; u < m; u += 32) {
    const int d = u < n1 ? lo1 + u : lo2 + (u - n1);
    int ro, lo, tn;
    blk_row(d, ro, lo, tn);
    const real *Mrow = EF(Minv) + ro;
    real t = 0;
    MGS_UNROLL_INNER
    for (int k = lo; k < lo + tn; k++) t += Mrow[k] * T[k];
    EF(wvec)[d] += t;
  }
  __syncwarp();
  return change;
}
// trees of contact c: (first dof, dof count) of the trees of its two bodies (count 0: static body); returns 0 for a contact without dofs
MGS_DEV int contact_trees(const Env &e, int c, int &lo1, int &n1, int &lo2, int &n2, int &t1, int &t2) {
  const int p = IARR(EF(con_pair))[c];
  const int b1 = LDG(MD.cgeom_bodyid + LDG(MD.pair_geom1 + p)), b2 = LDG(MD.cgeom_bodyid + LDG(MD.pair_geom2 + p));
  t1 = LDG(MD.body_treeid + b1); t2 = LDG(MD.body_treeid + b2);
  if (t1 < 0) { t1 = t2; t2 = -1; }
  if (t2 == t1) t2 = -1;
  lo1 = n1 = lo2 = n2 = 0;
  if (t1 < 0) return 0;
  if (t2 >= 0 && t2 < t1) { const int t = t1; t1 = t2; t2 = t; }
  #pragma unroll 1
  for (int d = 0; d < MD.nv; d += LDG(MD.dof_treenum + d)) {
    const int t = LDG(MD.dof_treeid + d);
    if (t == t1) { lo1 = d; n1 = LDG(MD.dof_treenum + d); }
    if (t == t2) { lo2 = d; n2 = LDG(MD.dof_treenum + d); }
  }
  return 1;
}
#endif

// mj_solNoSlip: Gauss-Seidel on the friction rows with the UNREGULARISED A = J M^-1 J'.
// wvec tracks qacc - qacc_smooth = M^-1 J' f, so no nefc x nefc matrix is formed: a row's residual is
// J_i (qacc_smooth + w) - aref_i and a force change dF moves w by M^-1 J' dF.  The small diagonal blocks
// A_c = J_c M^-1 J_c' (one per contact, <= 3x3) do not change between sweeps: they are computed once, one
// contact per lane, into the (now free) jar/jv scratch.
MGS_DEVN void solve_noslip_w(Env &e) {
  const int nv = MD.nv;
  const real scale = R_(1.0) / (MD.meaninertia * (nv > 1 ? nv : 1));
  real *AC = EF(efc_jar);  // 6 reals per contact: upper triangle of A_c (jar and jv are dead after Newton)
  real *T = EF(nsB);       // nv-vector scratch
  // Starting point w of the Gauss-Seidel sweeps.  mj_solNoSlip works in force space (residual_i = sum_j AR_ij f_j + b_i) and
  // ends with qacc = qacc_smooth + M^-1 J' f: the acceleration it reasons about is w = M^-1 J' f, what the current FORCES
  // produce.  The primal iterate qacc - qacc_smooth equals it only when the Newton solve reached M (qacc - qacc_smooth) = J' f;
  // on a cone-transition step of a VX300 grasp the two differed by 49 rad/s^2 on the 1e-5 kg m^2 object (found by the fp64
  // host build vs the oracle, which restates the force-space form).
  //   fp64 build: w = M^-1 J' f, exactly the oracle / MuJoCo formulation.
  //   fp32 build: w = qacc - qacc_smooth.  In fp32 the Newton solve stops at a gradient floor of ~1e-6 |J' f|, and that residual
  //   divided by such inertias is tens of rad/s^2 on EVERY step: with the force-space start 23 % of VX300 rollouts blew up
  //   (1-lane fp32 host build vs oracle, 48 candidates), with the primal start the labels agree 48/48.  The primal iterate is
  //   the better-conditioned value of the same quantity; the two coincide for a converged solve.
#if defined(MGS_REAL_DOUBLE) || defined(MGS_NOSLIP_FORCE_START)
  #pragma unroll 1
  PFOR(d, nv) {
    real t = 0;
    MGS_UNROLL_INNER
    for (int i = 0; i < EH.nefc; i++) t += EF(J)[i * nv + d] * EF(efc_force)[i];
    T[d] = t;
  }
  WSYNC();
  #pragma unroll 1
  PFOR(d, nv) {
    int ro, lo, tn;
    blk_row(d, ro, lo, tn);
    const int hi = lo + tn;
    const real *Mrow = EF(Minv) + ro;
    real t = 0;
    MGS_UNROLL_INNER
    for (int k = lo; k < hi; k++) t += Mrow[k] * T[k];
    EF(wvec)[d] = t;
  }
#else
  #pragma unroll 1
  PFOR(d, nv) EF(wvec)[d] = EF(qacc)[d] - EF(qacc_smooth)[d];
#endif
  WSYNC();
  // B_c = M^-1 J_c' (3 x nv per contact) for the first `ncov` contacts, kept in H (nv*nv reals, free between the
  // Newton solve and the integrator): one (row, dof) per lane, M^-1 block diagonal per kinematic tree.  With B_c a
  // force change costs 3 FMAs per dof instead of a J' product plus an M^-1 product (ncu r1_i: the two were 3.6 k of
  // the 8.4 k noslip instructions per step), and A_c = J_c B_c needs nv MACs per entry instead of nv * tree.
  const int ncon = EH.ncon;
  const int ncov = (nv / 3 < ncon) ? nv / 3 : ncon;
  real *B = EF(H);
  #pragma unroll 1
  for (int c = 0; c < ncov; c++) {
    const int i = IARR(EF(con_efc))[c];
    if (i < 0) continue;
    #pragma unroll 1
    PFOR(r, 3 * nv) {
      const int j = (r >= 2 * nv) ? 2 : (r >= nv ? 1 : 0), d = r - j * nv;
      // rows beyond the contact's friction dims are never read (dim 3: j < 2)
      int ro, lo, tn;
      blk_row(d, ro, lo, tn);
      const int hi = lo + tn;
      const real *Mrow = EF(Minv) + ro, *Jrow = EF(J) + (i + 1 + j) * nv;
      real t = 0;
      MGS_UNROLL_INNER
      for (int k = lo; k < hi; k++) t += Mrow[k] * Jrow[k];
      B[c * 3 * nv + r] = t;
    }
  }
  WSYNC();
  // A_c: one (contact, upper-triangle entry) per lane
  #pragma unroll 1
  PFOR(idx, 6 * ncon) {
    const int c = idx / 6, q = idx - 6 * c;
    const int i = IARR(EF(con_efc))[c];
    if (i < 0) continue;
    const int dim = LDG(MD.pair_condim + IARR(EF(con_pair))[c]);
    int n = dim - 1;
    if (n > 3) n = 3;
    // q -> (j,k), j <= k, packed row-major over the n x n upper triangle
    int j = 0, k = q;
    if (k >= n) { k -= n; j = 1; k += 1; if (k >= n) { k -= n; j = 2; k += 2; } }
    if (j >= n || k >= n || q >= n * (n + 1) / 2) continue;
    const real *Jj = EF(J) + (i + 1 + j) * nv, *Jk = EF(J) + (i + 1 + k) * nv;
    real acc = 0;
    if (c < ncov) {
      const real *Bk = B + c * 3 * nv + k * nv;
      MGS_UNROLL_INNER
      for (int a = 0; a < nv; a++) acc += Jj[a] * Bk[a];
    } else {
      #pragma unroll 1
      for (int a = 0; a < nv; a++) {
        const real ja = Jj[a];
        if (ja == 0) continue;
        int ro, lo, tn;
        blk_row(a, ro, lo, tn);
        const int hi = lo + tn;
        real t = 0;
        const real *Mrow = EF(Minv) + ro;
        MGS_UNROLL_INNER
        for (int b2 = lo; b2 < hi; b2++) t += Mrow[b2] * Jk[b2];
        acc += ja * t;
      }
    }
    AC[6 * c + q] = acc;
  }
  WSYNC();
#ifdef MGS_WIDE
  // LEVEL SCHEDULE of the contact sweep.  Two contacts that share no kinematic tree commute exactly (disjoint rows of w, zero
  // coupling), so the sequential Gauss-Seidel order only matters along chains of contacts that share a tree: level(c) = 1 + the
  // largest level of an earlier contact that shares a tree with c.  Contacts of one level run concurrently, one warp each; the levels
  // run in order.  Same result as the sequential sweep; in the config-5 pile ~40 contacts form ~20 levels.
  int *lvl_list = IARR(EF(efc_list)), *lvl_start = IARR(EF(nsS));
  real *chg = EF(efc_list) + LY.ncon_max;  // per-contact cost change of the current sweep (efc_list holds nefc_max >= 3 ncon_max words)
  // NOT ENABLED by default (-DMGS_NOSLIP_LEVELS turns it on).  Measured on config 5: the noslip stage itself drops from 317 k to 195 k
  // cycles per step and one step agrees with the sequential sweep to rounding (tools/ab_wide.py: qacc 3e-6 after one step), but the
  // full-schedule bench got SLOWER (147 k -> 128 k env-steps/s: more Newton iterations downstream) and the 10-step trajectory test
  // against the oracle went from 2e-5 to 6e-4.  Unexplained within the round's GPU budget, so the sequential sweep stays.
#ifdef MGS_NOSLIP_LEVELS
  const int use_levels = MD.ntree <= MGS_SPARSE_H_MAXTREE && ncon <= LANES;
#else
  const int use_levels = 0;
#endif
  int nlevel = 0;
  if (use_levels) {
    int *lvl = IARR(EF(efc_key));
    #pragma unroll 1
    PFOR(c, ncon) {  // the trees of every contact, in parallel: t1 | t2 << 8 (255: none), -1: nothing to do for this contact
      int lo1, n1, lo2, n2, t1, t2, w = -1;
      const int i = IARR(EF(con_efc))[c];
      if (i >= 0 && LDG(MD.pair_condim + IARR(EF(con_pair))[c]) >= 3 && contact_trees(e, c, lo1, n1, lo2, n2, t1, t2)) w = t1 | ((t2 >= 0 ? t2 : 255) << 8);
      lvl[c] = w;
    }
    WSYNC();
    if (threadIdx.x == 0) {  // the recurrence itself is sequential (a few instructions per contact)
      int last[MGS_SPARSE_H_MAXTREE];
      for (int t = 0; t < MGS_SPARSE_H_MAXTREE; t++) last[t] = 0;
      int mx = 0;
      for (int c = 0; c < ncon; c++) {
        const int w = lvl[c];
        int L = 0;
        if (w >= 0) {
          const int t1 = w & 255, t2 = w >> 8;
          L = last[t1];
          if (t2 != 255 && last[t2] > L) L = last[t2];
          L += 1;
          last[t1] = L;
          if (t2 != 255) last[t2] = L;
          if (L > mx) mx = L;
        }
        lvl[c] = L;  // 0: nothing to do for this contact
      }
      // counting sort by level (contact order kept inside a level)
      for (int L = 0; L <= mx + 1; L++) lvl_start[L] = 0;
      for (int c = 0; c < ncon; c++) if (lvl[c] > 0) lvl_start[lvl[c] + 1]++;
      for (int L = 1; L <= mx + 1; L++) lvl_start[L] += lvl_start[L - 1];
      for (int c = 0; c < ncon; c++) if (lvl[c] > 0) { lvl_list[lvl_start[lvl[c]]] = c; lvl_start[lvl[c]]++; }
      for (int L = mx + 1; L >= 1; L--) lvl_start[L] = lvl_start[L - 1];  // undo the running pointers: lvl_start[L] = first slot of level L
      lvl_start[0] = mx;
    }
    WSYNC();
    nlevel = lvl_start[0];
  }
#endif
  #pragma unroll 1
  for (int iter = 0; iter < MD.noslip_iterations; iter++) {
    real improvement = 0;
    if (iter == 0) {
      real t = 0;
      #pragma unroll 1
      PFOR(i, EH.nefc) {
        int type = EFC_TYPE(i);
        int fr = type == CT_FRICTION_DOF || (type == CT_CONTACT && IARR(EF(con_efc))[EFC_ID(i)] != i);
        if (fr) t += R_(0.5) * EF(efc_force)[i] * EF(efc_force)[i] * EF(efc_R)[i];
      }
      improvement += wsum(t);
    }
    // dry-friction rows: J_i is the unit vector of dof d, so everything is a table lookup
    #pragma unroll 1
    for (int i = EH.ne; i < EH.ne + EH.nf; i++) {
      const int d = EFC_ID(i);
      int ro, lo, tn;
      blk_row(d, ro, lo, tn);
      const int hi = lo + tn;
      const real *Mrow = EF(Minv) + ro;  // row d of M^-1 (= column d: symmetric), nonzero inside d's tree
      const real res = EF(qacc_smooth)[d] + EF(wvec)[d] - EF(efc_aref)[i], Aii = Mrow[d];
      const real old = EF(efc_force)[i], fl = EF(efc_aux)[i];
      real fn = old - res / fmax(MGS_MINVAL, Aii);
      fn = fmax(-fl, fmin(fl, fn));
      real delta = fn - old, change = R_(0.5) * delta * delta * Aii + delta * res;
      if (change > R_(1e-10)) { fn = old; delta = 0; change = 0; }
      WSYNC();
      #pragma unroll 1
      PFOR(k, nv) if (k >= lo && k < hi) EF(wvec)[k] += Mrow[k] * delta;
      PFOR(k, 1) EF(efc_force)[i] = fn;
      improvement -= change;
      WSYNC();
    }
    // contact friction dims
#ifdef MGS_WIDE
    if (use_levels) {
      #pragma unroll 1
      PFOR(c, ncon) chg[c] = 0;
      WSYNC();
      #pragma unroll 1
      for (int L = 1; L <= nlevel; L++) {
        const int q0 = L == 1 ? 0 : lvl_start[L], q1 = lvl_start[L + 1];
        #pragma unroll 1
#ifdef MGS_NOSLIP_LEVELS_ONE_WARP  // (debug: the level order, but one warp does every contact)
        for (int q = q0; q < q1 && threadIdx.x < 32; q++) {
#else
        for (int q = q0 + (int)(threadIdx.x >> 5); q < q1; q += MGS_NWARP) {
#endif
          const int c = lvl_list[q], i = IARR(EF(con_efc))[c], p = IARR(EF(con_pair))[c], dim = LDG(MD.pair_condim + p);
          int lo1, n1, lo2, n2, t1, t2;
          contact_trees(e, c, lo1, n1, lo2, n2, t1, t2);
          const real *Bc = (c < ncov) ? B + c * 3 * nv : (const real *)0;
          const real ch = (dim == 3) ? noslip_contact_warp<2>(e, c, i, p, AC, T, Bc, lo1, n1, lo2, n2)
                                     : noslip_contact_warp<3>(e, c, i, p, AC, T, Bc, lo1, n1, lo2, n2);
          if ((threadIdx.x & 31) == 0) chg[c] = ch;
        }
        WSYNC();
      }
      #pragma unroll 1
      for (int c = 0; c < ncon; c++) improvement -= chg[c];  // summed in contact order, like the sequential sweep
    } else
#endif
    #pragma unroll 1
    for (int c = 0; c < EH.ncon; c++) {
      const int i = IARR(EF(con_efc))[c];
      if (i < 0) continue;
      const int p = IARR(EF(con_pair))[c], dim = LDG(MD.pair_condim + p);
      if (dim < 3) continue;
      const real *Bc = (c < ncov) ? B + c * 3 * nv : (const real *)0;
      improvement -= (dim == 3) ? noslip_contact_w<2>(e, c, i, p, AC, T, Bc) : noslip_contact_w<3>(e, c, i, p, AC, T, Bc);
    }
    if (improvement * scale < MD.noslip_tolerance) break;
  }
  // qacc = qacc_smooth + w ; qfrc_constraint = M w
  matvec_w(EF(qfrc_constraint), EF(M), EF(wvec), nv);
  #pragma unroll 1
  PFOR(d, nv) EF(qacc)[d] = EF(qacc_smooth)[d] + EF(wvec)[d];
  WSYNC();
}

// ---------------------------------------------------------------------------------- forward / integrate
// mj_forward.  `with_integrate_slot`: callers that do not integrate afterwards (the initial forward of a
// candidate) still take the integrate stage's barrier so that all warps of the CTA stay stage-aligned.
MGS_DEVN void forward_w(Env &e) {
  const int nv = MD.nv;
  MGS_STAGE_BARRIER(0);
  MGS_CLK(11);
  kinematics_w(e);
  inertia_w(e);
  transmission_w(e);
  MGS_CLK(0);
  MGS_STAGE_BARRIER(1);
  collision_w(e);
  MGS_CLK(1);
  MGS_STAGE_BARRIER(2);
  smooth_forces_w(e);
  make_constraint_w(e);
  MGS_CLK(2);
  MGS_STAGE_BARRIER(3);
  if (EH.nefc == 0) {
    #pragma unroll 1
    PFOR(d, nv) { EF(qacc)[d] = EF(qacc_smooth)[d]; EF(qacc_ws)[d] = EF(qacc_smooth)[d]; EF(qfrc_constraint)[d] = 0; }
    WSYNC();
  } else {
    solve_newton_w(e);
    #pragma unroll 1
    PFOR(d, nv) EF(qacc_ws)[d] = EF(qacc)[d];
    WSYNC();
  }
  MGS_CLK(3);
  MGS_STAGE_BARRIER(4);
  if (EH.nefc != 0) {
    if (MD.noslip_iterations > 0) solve_noslip_w(e);
    else {
      #pragma unroll 1
      PFOR(d, nv) {
        real t = 0;
        #pragma unroll 1
        for (int i = 0; i < EH.nefc; i++) t += EF(J)[i * nv + d] * EF(efc_force)[i];
        EF(qfrc_constraint)[d] = t;
      }
      WSYNC();
    }
  }
  MGS_CLK(7);
}

MGS_DEVN int bad_state_w(const Env &e, int check_acc) {
  int bad = 0;
  #pragma unroll 1
  PFOR(i, MD.nq) bad |= !(fabs(EF(qpos)[i]) < R_(1e10));
  #pragma unroll 1
  PFOR(i, MD.nv) bad |= !(fabs(EF(qvel)[i]) < R_(1e10));
  if (check_acc) PFOR(i, MD.nv) bad |= !(fabs(EF(qacc)[i]) < R_(1e10));
  return wany(bad);
}

// implicitfast: (M - h dF/dv) a = qfrc_smooth + qfrc_constraint; v += h a; q integrates with new v
MGS_DEVN void integrate_w(Env &e) {
  const int nv = MD.nv;
  const real h = MD.timestep;
  #pragma unroll 1
  PFOR(idx, MD.nM) {  // block storage: one entry of one tree's tile per lane
    const int ij = LDG(MD.blk_ij + idx), i = ij >> 8, j = ij & 255;
    real a = EF(M)[idx];
    if (i == j) a += h * LDG(MD.dof_damping + i);
    #pragma unroll 1
    for (int u = 0; u < MD.nu; u++) {
      real bv = LDG(MD.actuator_biasprm + 3 * u + 2);
      if (bv == 0) continue;
      if (LDG(MD.actuator_forcelimited + u)) {
        real f = EF(act_force)[u];
        if (f <= LDG(MD.actuator_forcerange + 2 * u) || f >= LDG(MD.actuator_forcerange + 2 * u + 1)) continue;
      }
      a -= h * bv * EF(act_moment)[u * nv + i] * EF(act_moment)[u * nv + j];
    }
    EF(H)[idx] = a;
  }
  #pragma unroll 1
  PFOR(d, nv) EF(search)[d] = EF(qfrc_smooth)[d] + EF(qfrc_constraint)[d];
  WSYNC();
  chol_factor_w(EF(H), nv, 1);
  chol_solve_w(EF(H), EF(search), nv, 1);
  #pragma unroll 1
  PFOR(d, nv) EF(qvel)[d] += h * EF(search)[d];
  WSYNC();
#ifdef MGS_QPOS_COMP
  // position update on qpos = hi + lo in double (a handful of fp64 operations per joint and step)
  #pragma unroll 1
  PFOR(j, MD.njnt) {
    const int qa = LDG(MD.jnt_qposadr + j), da = LDG(MD.jnt_dofadr + j);
    const double hd = (double)h;
    if (LDG(MD.jnt_type + j) == JNT_FREE) {
      for (int k = 0; k < 3; k++) {
        const double Q = (double)EF(qpos)[qa + k] + (double)EF(qpos_lo)[qa + k] + hd * (double)EF(qvel)[da + k];
        const float hi = (float)Q;
        EF(qpos)[qa + k] = hi; EF(qpos_lo)[qa + k] = (float)(Q - (double)hi);
      }
      const double w0 = EF(qvel)[da + 3], w1 = EF(qvel)[da + 4], w2 = EF(qvel)[da + 5];
      const double wn = sqrt(w0 * w0 + w1 * w1 + w2 * w2), ang = wn * hd;
      if (ang > 0) {
        // half-angle sine / cosine: series below 1e-2 rad (error < 1e-17), library calls above
        const double ha = 0.5 * ang, ha2 = ha * ha;
        const double sn = (ha < 1e-2 ? ha * (1.0 - ha2 * (1.0 / 6.0) * (1.0 - ha2 * (1.0 / 20.0) * (1.0 - ha2 * (1.0 / 42.0)))) : sin(ha)) / wn;
        const double cs = ha < 1e-2 ? 1.0 - ha2 * 0.5 * (1.0 - ha2 * (1.0 / 12.0) * (1.0 - ha2 * (1.0 / 30.0))) : cos(ha);
        const double b0 = cs, b1 = sn * w0, b2 = sn * w1, b3 = sn * w2;
        const double a0 = (double)EF(qpos)[qa + 3] + (double)EF(qpos_lo)[qa + 3], a1 = (double)EF(qpos)[qa + 4] + (double)EF(qpos_lo)[qa + 4],
                     a2 = (double)EF(qpos)[qa + 5] + (double)EF(qpos_lo)[qa + 5], a3 = (double)EF(qpos)[qa + 6] + (double)EF(qpos_lo)[qa + 6];
        double r[4] = {a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3, a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                       a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1, a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0};
        const double inv = 1.0 / sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
        for (int k = 0; k < 4; k++) {
          const double Q = r[k] * inv;
          const float hi = (float)Q;
          EF(qpos)[qa + 3 + k] = hi; EF(qpos_lo)[qa + 3 + k] = (float)(Q - (double)hi);
        }
      }
    } else {
      const double Q = (double)EF(qpos)[qa] + (double)EF(qpos_lo)[qa] + hd * (double)EF(qvel)[da];
      const float hi = (float)Q;
      EF(qpos)[qa] = hi; EF(qpos_lo)[qa] = (float)(Q - (double)hi);
    }
  }
  WSYNC();
}
#else
  #pragma unroll 1
  PFOR(j, MD.njnt) {
    int qa = LDG(MD.jnt_qposadr + j), da = LDG(MD.jnt_dofadr + j);
    if (LDG(MD.jnt_type + j) == JNT_FREE) {
      for (int k = 0; k < 3; k++) EF(qpos)[qa + k] += h * EF(qvel)[da + k];
      real w[3] = {EF(qvel)[da + 3], EF(qvel)[da + 4], EF(qvel)[da + 5]};
      real ang = sqrt(dot3(w, w)) * h;
      if (ang > 0) {
        normalize3(w);
        real sn = sin(R_(0.5) * ang), q[4] = {cos(R_(0.5) * ang), sn * w[0], sn * w[1], sn * w[2]}, r[4];
        mulquat(r, EF(qpos) + qa + 3, q);
        normquat(r);
        EF(qpos)[qa + 3] = r[0]; EF(qpos)[qa + 4] = r[1]; EF(qpos)[qa + 5] = r[2]; EF(qpos)[qa + 6] = r[3];
      }
    } else EF(qpos)[qa] += h * EF(qvel)[da];
  }
  WSYNC();
}
#endif

// mj_step x nstep; returns nonzero if the state blew up (the env is then labelled failed)
MGS_DEVN int step_w(Env &e, int nstep, int *steps_done) {
  #pragma unroll 1
  for (int k = 0; k < nstep; k++) {
    if (PRM.qvel_clip > 0) {  // scene generation clamps qvel before every step (clutter_table.py:215-221); MGS_MODE_STEP only
      #pragma unroll 1
      PFOR(i, MD.nv) EF(qvel)[i] = fmax(-PRM.qvel_clip, fmin(PRM.qvel_clip, EF(qvel)[i]));
      WSYNC();
    }
    if (EH.bad || bad_state_w(e, 0)) { EH.bad = 1; return 1; }
    forward_w(e);
    MGS_STAGE_BARRIER(5);
    if (bad_state_w(e, 1)) { EH.bad = 1; return 1; }
    integrate_w(e);
    MGS_CLK(8);
    (*steps_done)++;
  }
  return 0;
}
