// mgs_solver.cuh - constraint rows, primal Newton solver, noslip, implicitfast (warp per env).
//
// Row order and arithmetic follow SURVEY.md 8(a-MJ): equality -> dof friction -> limits ->
// contacts; soft constraints from solref/solimp (refsafe), R from body/dof invweight0, elliptic
// cones with impratio; Newton on qacc with H = M + J'DJ + cone blocks factored by a
// warp-cooperative Cholesky in shared memory; exact 1-D line search on the convex cost; noslip
// Gauss-Seidel on friction rows with the unregularised A = J M^-1 J'.
#pragma once
#include "mgs_collide.cuh"

MGS_DEV real impedance_f(const real *solimp, real pos) {
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

// accumulate sign * axis . (d point / d qdot) into row(s): walks the dof chain of `body`
MGS_DEV void jac_rows_point(const DevModel &m, const Env &e, int body, const real *point, real sign, const real *axes, int naxis_t,
                            int naxis_r, real *Jrows) {
  // axes: [naxis_t translational axes (3 each)] followed by [naxis_r rotational axes]; rows are consecutive in Jrows
  const int nv = m.nv;
  real off[3];
  sub3(off, point, e.rootcom + 3 * LDG(m.body_rootid + body));
  while (body > 0 && LDG(m.body_dofnum + body) == 0) body = LDG(m.body_parentid + body);
  if (body == 0) return;
  for (int d = LDG(m.body_dofadr + body) + LDG(m.body_dofnum + body) - 1; d >= 0; d = LDG(m.dof_parentid + d)) {
    const real *c = e.cdof + 6 * d;
    real lin[3];
    cross3(lin, c, off);
    lin[0] += c[3]; lin[1] += c[4]; lin[2] += c[5];
    for (int k = 0; k < naxis_t; k++) Jrows[k * nv + d] += sign * dot3(axes + 3 * k, lin);
    for (int k = 0; k < naxis_r; k++) Jrows[(naxis_t + k) * nv + d] += sign * dot3(axes + 3 * (naxis_t + k), c);
  }
}

MGS_DEV void kbi_from_solref(const DevModel &m, const real *solref, const real *solimp, real pos, int friction_row, real *k, real *b, real *imp) {
  *imp = impedance_f(solimp, pos);
  real dmax = fmin(R_(0.9999), fmax(R_(0.0001), solimp[1]));
  if (solref[0] > 0) {
    real tc = fmax(solref[0], 2 * m.timestep), dr = solref[1];
    *k = R_(1.0) / fmax(MGS_MINVAL, dmax * dmax * tc * tc * dr * dr);
    *b = R_(2.0) / fmax(MGS_MINVAL, dmax * tc);
  } else {
    *k = -solref[0] / fmax(MGS_MINVAL, dmax * dmax);
    *b = -solref[1] / fmax(MGS_MINVAL, dmax);
  }
  if (friction_row) *k = 0;
}

MGS_DEV void finish_row(const DevModel &m, Env &e, int r, int type, int id, real pos, real floss, real diagApprox, const real *solref,
                        const real *solimp, int friction_row) {
  const int nv = m.nv;
  real k, b, imp, vel = 0;
  kbi_from_solref(m, solref, solimp, pos, friction_row, &k, &b, &imp);
  for (int d = 0; d < nv; d++) vel += e.J[r * nv + d] * e.qvel[d];
  IARR(e.efc_type)[r] = type;
  IARR(e.efc_id)[r] = id;
  e.efc_pos[r] = pos;
  e.efc_floss[r] = floss;
  e.efc_imp[r] = imp;
  real R = fmax(MGS_MINVAL, (1 - imp) * diagApprox / imp);
  e.efc_R[r] = R;
  e.efc_D[r] = R_(1.0) / R;
  e.efc_aref[r] = -b * vel - k * imp * (friction_row ? R_(0.0) : pos);
}

// mj_makeConstraint + mj_makeImpedance + mj_referenceConstraint
MGS_DEVN void make_constraint_w(const DevModel &m, Env &e) {
  const int nv = m.nv;
  // --- row bookkeeping (warp-uniform)
  int ne = m.ne_rows, nf = 0, nl = 0;
  for (int d = 0; d < nv; d++) nf += LDG(m.dof_frictionloss + d) > 0;
  // limits: count with a scan so that rows come out in joint order
  int nlim_total = 0;
  int row_con0;
  {
    int cnt_base = 0;
    for (int j0 = 0; j0 < m.njnt; j0 += LANES) {
      int j = j0 + MGS_LANE, cnt = 0;
      real dist0 = 0, dist1 = 0;
      if (j < m.njnt && LDG(m.jnt_limited + j) && LDG(m.jnt_type + j) != JNT_FREE) {
        real q = e.qpos[LDG(m.jnt_qposadr + j)], mg = LDG(m.jnt_margin + j);
        dist0 = q - LDG(m.jnt_range + 2 * j);
        dist1 = LDG(m.jnt_range + 2 * j + 1) - q;
        cnt = (dist0 < mg) + (dist1 < mg);
      }
      int total, off = wscan_excl(cnt, &total);
      if (cnt) {
        real mg = LDG(m.jnt_margin + j);
        int r = ne + nf + cnt_base + off, dof = LDG(m.jnt_dofadr + j);
        real sr[2], si[5];
        sr[0] = LDG(m.jnt_solref + 2 * j); sr[1] = LDG(m.jnt_solref + 2 * j + 1);
        for (int k = 0; k < 5; k++) si[k] = LDG(m.jnt_solimp + 5 * j + k);
        if (dist0 < mg && r < e.nefc_max) {
          for (int d = 0; d < nv; d++) e.J[r * nv + d] = 0;
          e.J[r * nv + dof] = 1;
          finish_row(m, e, r, CT_LIMIT, j, dist0 - mg, 0, LDG(m.dof_invweight0 + dof), sr, si, 0);
          r++;
        }
        if (dist1 < mg && r < e.nefc_max) {
          for (int d = 0; d < nv; d++) e.J[r * nv + d] = 0;
          e.J[r * nv + dof] = -1;
          finish_row(m, e, r, CT_LIMIT, j, dist1 - mg, 0, LDG(m.dof_invweight0 + dof), sr, si, 0);
        }
      }
      cnt_base += total;
    }
    nlim_total = cnt_base;
  }
  nl = nlim_total;
  row_con0 = ne + nf + nl;
  // contact rows: prefix sum of condim over contacts (capacity-limited)
  int nefc;
  {
    int base = row_con0;
    for (int c0 = 0; c0 < e.ncon; c0 += LANES) {
      int c = c0 + MGS_LANE, dim = 0;
      if (c < e.ncon) dim = LDG(m.pair_condim + IARR(e.con_pair)[c]);
      int total, off = wscan_excl(dim, &total);
      if (c < e.ncon) IARR(e.con_efc)[c] = (base + off + dim <= e.nefc_max) ? base + off : -1;
      base += total;
    }
    if (base > e.nefc_max) {
      // drop the contacts that do not fit (flagged); rows of the kept ones stay contiguous
      e.overflow += 1;
      int keep = row_con0;
      for (int c = 0; c < e.ncon; c++) {
        int r = IARR(e.con_efc)[c];
        if (r >= 0) keep = r + LDG(m.pair_condim + IARR(e.con_pair)[c]);
      }
      base = keep;
    }
    nefc = base;
  }
  WSYNC();
  // --- equality rows (lane per equality)
  PFOR(q, m.neq) {
    if (!LDG(m.eq_active + q)) continue;
    int r0 = LDG(m.eq_rowadr + q), type = LDG(m.eq_type + q);
    real data[11], sr[2], si[5];
    for (int k = 0; k < 11; k++) data[k] = LDG(m.eq_data + 11 * q + k);
    sr[0] = LDG(m.eq_solref + 2 * q); sr[1] = LDG(m.eq_solref + 2 * q + 1);
    for (int k = 0; k < 5; k++) si[k] = LDG(m.eq_solimp + 5 * q + k);
    if (type == EQ_JOINT) {
      int j1 = LDG(m.eq_obj1id + q), j2 = LDG(m.eq_obj2id + q);
      int qa1 = LDG(m.jnt_qposadr + j1), d1 = LDG(m.jnt_dofadr + j1);
      real q1 = e.qpos[qa1] - LDG(m.qpos0 + qa1), pos, deriv = 0, da = LDG(m.dof_invweight0 + d1);
      for (int d = 0; d < nv; d++) e.J[r0 * nv + d] = 0;
      e.J[r0 * nv + d1] = 1;
      if (j2 >= 0) {
        int qa2 = LDG(m.jnt_qposadr + j2), d2 = LDG(m.jnt_dofadr + j2);
        real dif = e.qpos[qa2] - LDG(m.qpos0 + qa2);
        real poly = data[0] + dif * (data[1] + dif * (data[2] + dif * (data[3] + dif * data[4])));
        deriv = data[1] + dif * (2 * data[2] + dif * (3 * data[3] + dif * 4 * data[4]));
        pos = q1 - poly;
        e.J[r0 * nv + d2] = -deriv;
        da += LDG(m.dof_invweight0 + d2);
      } else pos = q1 - data[0];
      finish_row(m, e, r0, CT_EQUALITY, q, pos, 0, da, sr, si, 0);
      continue;
    }
    int b1 = LDG(m.eq_obj1id + q), b2 = LDG(m.eq_obj2id + q);
    const real *a1 = (type == EQ_WELD) ? data + 3 : data, *a2 = (type == EQ_WELD) ? data : data + 3;
    real p1[3], p2[3], cpos[6];
    mulmatvec3(p1, e.xmat + 9 * b1, a1); add3(p1, p1, e.xpos + 3 * b1);
    mulmatvec3(p2, e.xmat + 9 * b2, a2); add3(p2, p2, e.xpos + 3 * b2);
    sub3(cpos, p1, p2);
    int nrow = (type == EQ_WELD) ? 6 : 3;
    for (int k = 0; k < nrow * nv; k++) e.J[r0 * nv + k] = 0;
    const real eye[18] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 0, 0, 0, 1, 0, 0, 0, 1};
    int nr_rot = (type == EQ_WELD) ? 3 : 0;
    jac_rows_point(m, e, b1, p1, R_(1.0), eye, 3, nr_rot, e.J + r0 * nv);
    jac_rows_point(m, e, b2, p2, R_(-1.0), eye, 3, nr_rot, e.J + r0 * nv);
    real tran = LDG(m.body_invweight0 + 2 * b1) + LDG(m.body_invweight0 + 2 * b2);
    real rot = LDG(m.body_invweight0 + 2 * b1 + 1) + LDG(m.body_invweight0 + 2 * b2 + 1);
    if (type == EQ_WELD) {
      real ts = data[10], quat[4], quat1[4], quat2[4], quat3[4];
      mulquat(quat, e.xquat + 4 * b1, data + 6);
      quat1[0] = e.xquat[4 * b2]; quat1[1] = -e.xquat[4 * b2 + 1]; quat1[2] = -e.xquat[4 * b2 + 2]; quat1[3] = -e.xquat[4 * b2 + 3];
      mulquat(quat2, quat1, quat);
      cpos[3] = ts * quat2[1]; cpos[4] = ts * quat2[2]; cpos[5] = ts * quat2[3];
      for (int d = 0; d < nv; d++) {
        real ax[4] = {0, e.J[(r0 + 3) * nv + d], e.J[(r0 + 4) * nv + d], e.J[(r0 + 5) * nv + d]};
        if (ax[1] == 0 && ax[2] == 0 && ax[3] == 0) continue;
        mulquat(quat2, quat1, ax);
        mulquat(quat3, quat2, quat);
        e.J[(r0 + 3) * nv + d] = R_(0.5) * ts * quat3[1];
        e.J[(r0 + 4) * nv + d] = R_(0.5) * ts * quat3[2];
        e.J[(r0 + 5) * nv + d] = R_(0.5) * ts * quat3[3];
      }
    }
    for (int k = 0; k < nrow; k++) finish_row(m, e, r0 + k, CT_EQUALITY, q, cpos[k], 0, k < 3 ? tran : rot, sr, si, 0);
  }
  // --- dof friction rows (lane per dof; row index = rank among friction dofs)
  PFOR(d, nv) {
    real fl = LDG(m.dof_frictionloss + d);
    if (fl <= 0) continue;
    int r = ne;
    for (int k = 0; k < d; k++) r += LDG(m.dof_frictionloss + k) > 0;
    for (int k = 0; k < nv; k++) e.J[r * nv + k] = 0;
    e.J[r * nv + d] = 1;
    real sr[2], si[5];
    sr[0] = LDG(m.dof_solref + 2 * d); sr[1] = LDG(m.dof_solref + 2 * d + 1);
    for (int k = 0; k < 5; k++) si[k] = LDG(m.dof_solimp + 5 * d + k);
    finish_row(m, e, r, CT_FRICTION_DOF, d, 0, fl, LDG(m.dof_invweight0 + d), sr, si, 1);
  }
  // --- contact rows (lane per contact)
  PFOR(c, e.ncon) {
    int r0 = IARR(e.con_efc)[c];
    if (r0 < 0) continue;
    int p = IARR(e.con_pair)[c], dim = LDG(m.pair_condim + p);
    int b1 = LDG(m.cgeom_bodyid + LDG(m.pair_geom1 + p)), b2 = LDG(m.cgeom_bodyid + LDG(m.pair_geom2 + p));
    for (int k = 0; k < dim * nv; k++) e.J[r0 * nv + k] = 0;
    real axes[18];
    for (int k = 0; k < 9; k++) axes[k] = e.con_frame[9 * c + k];
    for (int k = 0; k < 9; k++) axes[9 + k] = e.con_frame[9 * c + k];
    int nt = dim < 3 ? dim : 3, nr = dim > 3 ? dim - 3 : 0;
    jac_rows_point(m, e, b2, e.con_pos + 3 * c, R_(1.0), nt == 3 ? axes : axes, nt, nr, e.J + r0 * nv);
    jac_rows_point(m, e, b1, e.con_pos + 3 * c, R_(-1.0), axes, nt, nr, e.J + r0 * nv);
    real tran = LDG(m.body_invweight0 + 2 * b1) + LDG(m.body_invweight0 + 2 * b2);
    real rot = LDG(m.body_invweight0 + 2 * b1 + 1) + LDG(m.body_invweight0 + 2 * b2 + 1);
    real sr[2], si[5], fr[5];
    sr[0] = LDG(m.pair_solref + 2 * p); sr[1] = LDG(m.pair_solref + 2 * p + 1);
    for (int k = 0; k < 5; k++) { si[k] = LDG(m.pair_solimp + 5 * p + k); fr[k] = LDG(m.pair_friction + 5 * p + k); }
    real dist = e.con_dist[c];
    for (int k = 0; k < dim; k++) {
      finish_row(m, e, r0 + k, CT_CONTACT, c, dist, 0, k < 3 ? tran : rot, sr, si, k > 0);
      if (k > 0) e.efc_pos[r0 + k] = 0;
    }
    real mu = fr[0];
    if (m.cone_elliptic && dim >= 3) {
      real R0 = e.efc_R[r0], R1 = R0 / fmax(MGS_MINVAL, m.impratio);
      e.efc_R[r0 + 1] = R1;
      mu = fr[0] * sqrt(R1 / R0);
      for (int j = 2; j < dim; j++) e.efc_R[r0 + j] = R1 * fr[0] * fr[0] / fmax(MGS_MINVAL, fr[j - 1] * fr[j - 1]);
      for (int j = 1; j < dim; j++) e.efc_D[r0 + j] = R_(1.0) / e.efc_R[r0 + j];
    }
    e.con_mu[c] = mu;
  }
  e.ne = ne; e.nf = nf; e.nl = nl; e.nefc = nefc;
  WSYNC();
}

// ---------------------------------------------------------------------------------- Newton
// per-row (or per-contact) cost/force/state at jar; returns this lane's partial cost
MGS_DEV real constraint_update_w(const DevModel &m, Env &e, int want_hess) {
  real cost = 0;
  PFOR(i, e.nefc) {
    int type = IARR(e.efc_type)[i];
    real D = e.efc_D[i], jar = e.efc_jar[i];
    if (type == CT_EQUALITY) {
      e.efc_force[i] = -D * jar; cost += R_(0.5) * D * jar * jar; IARR(e.efc_state)[i] = ST_QUADRATIC;
    } else if (type == CT_FRICTION_DOF) {
      real f = e.efc_floss[i], R = e.efc_R[i];
      if (jar <= -R * f) { e.efc_force[i] = f; cost += R_(-0.5) * R * f * f - f * jar; IARR(e.efc_state)[i] = ST_LINEARNEG; }
      else if (jar >= R * f) { e.efc_force[i] = -f; cost += R_(-0.5) * R * f * f + f * jar; IARR(e.efc_state)[i] = ST_LINEARPOS; }
      else { e.efc_force[i] = -D * jar; cost += R_(0.5) * D * jar * jar; IARR(e.efc_state)[i] = ST_QUADRATIC; }
    } else if (type == CT_LIMIT) {
      if (jar < 0) { e.efc_force[i] = -D * jar; cost += R_(0.5) * D * jar * jar; IARR(e.efc_state)[i] = ST_QUADRATIC; }
      else { e.efc_force[i] = 0; IARR(e.efc_state)[i] = ST_SATISFIED; }
    } else {
      int c = IARR(e.efc_id)[i];
      if (IARR(e.con_efc)[c] != i) continue;  // the contact's first row does the whole cone
      int p = IARR(e.con_pair)[c], dim = LDG(m.pair_condim + p);
      if (dim < 3 || !m.cone_elliptic) {
        if (jar < 0) { e.efc_force[i] = -D * jar; cost += R_(0.5) * D * jar * jar; IARR(e.efc_state)[i] = ST_QUADRATIC; }
        else { e.efc_force[i] = 0; IARR(e.efc_state)[i] = ST_SATISFIED; }
        for (int j = 1; j < dim; j++) { e.efc_force[i + j] = 0; IARR(e.efc_state)[i + j] = ST_SATISFIED; }
        continue;
      }
      real mu = e.con_mu[c], fr[5], U[6], T = 0;
      for (int k = 0; k < 5; k++) fr[k] = LDG(m.pair_friction + 5 * p + k);
      U[0] = jar * mu;
      for (int j = 1; j < dim; j++) { U[j] = e.efc_jar[i + j] * fr[j - 1]; T += U[j] * U[j]; }
      real N = U[0];
      T = sqrt(T);
      int st;
      if (N >= mu * T || (T <= 0 && N >= 0)) {
        for (int j = 0; j < dim; j++) e.efc_force[i + j] = 0;
        st = ST_SATISFIED;
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int j = 0; j < dim; j++) {
          real x = e.efc_jar[i + j], Dj = e.efc_D[i + j];
          e.efc_force[i + j] = -Dj * x;
          cost += R_(0.5) * Dj * x * x;
        }
        st = ST_QUADRATIC;
      } else {
        real Dm = D / fmax(MGS_MINVAL, mu * mu * (1 + mu * mu)), NmT = N - mu * T;
        cost += R_(0.5) * Dm * NmT * NmT;
        real f0 = -Dm * NmT * mu;
        e.efc_force[i] = f0;
        for (int j = 1; j < dim; j++) e.efc_force[i + j] = -f0 / T * U[j] * fr[j - 1];
        st = ST_CONE;
        if (want_hess) {
          real *h = e.hcone + 16 * c, scl[4];
          scl[0] = mu;
          for (int j = 1; j < dim; j++) scl[j] = fr[j - 1];
          h[0] = 1;
          for (int j = 1; j < dim; j++) h[j] = h[j * dim] = -mu * U[j] / T;
          real muNT3 = mu * N / (T * T * T), dg = mu * mu - mu * N / T;
          for (int j = 1; j < dim; j++)
            for (int k = 1; k < dim; k++) h[j * dim + k] = muNT3 * U[j] * U[k] + (j == k ? dg : R_(0.0));
          for (int j = 0; j < dim; j++)
            for (int k = 0; k < dim; k++) h[j * dim + k] *= Dm * scl[j] * scl[k];
        }
      }
      for (int j = 0; j < dim; j++) IARR(e.efc_state)[i + j] = st;
    }
  }
  return cost;
}

// derivative pair of the cost along the search direction at alpha (this lane's partial sums)
MGS_DEV void ls_eval_w(const DevModel &m, const Env &e, real alpha, real *d1, real *d2) {
  real a = 0, h = 0;
  PFOR(i, e.nefc) {
    int type = IARR(e.efc_type)[i];
    real D = e.efc_D[i], jv = e.efc_jv[i], x = e.efc_jar[i] + alpha * jv;
    if (type == CT_EQUALITY) { a += D * x * jv; h += D * jv * jv; }
    else if (type == CT_FRICTION_DOF) {
      real f = e.efc_floss[i], R = e.efc_R[i];
      if (x <= -R * f) a += -f * jv;
      else if (x >= R * f) a += f * jv;
      else { a += D * x * jv; h += D * jv * jv; }
    } else if (type == CT_LIMIT) {
      if (x < 0) { a += D * x * jv; h += D * jv * jv; }
    } else {
      int c = IARR(e.efc_id)[i];
      if (IARR(e.con_efc)[c] != i) continue;
      int p = IARR(e.con_pair)[c], dim = LDG(m.pair_condim + p);
      if (dim < 3 || !m.cone_elliptic) { if (x < 0) { a += D * x * jv; h += D * jv * jv; } continue; }
      real mu = e.con_mu[c], T = 0, UV = 0, VV = 0;
      real N = x * mu, N1 = jv * mu;
      for (int j = 1; j < dim; j++) {
        real fj = LDG(m.pair_friction + 5 * p + j - 1);
        real Uj = (e.efc_jar[i + j] + alpha * e.efc_jv[i + j]) * fj, Vj = e.efc_jv[i + j] * fj;
        T += Uj * Uj; UV += Uj * Vj; VV += Vj * Vj;
      }
      T = sqrt(T);
      if (N >= mu * T || (T <= 0 && N >= 0)) {
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int j = 0; j < dim; j++) {
          real xj = e.efc_jar[i + j] + alpha * e.efc_jv[i + j], vj = e.efc_jv[i + j], Dj = e.efc_D[i + j];
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
MGS_DEV void eval_point_w(const DevModel &m, Env &e, const real *qacc) {
  const int nv = m.nv;
  PFOR(i, e.nefc) {
    real t = -e.efc_aref[i];
    for (int d = 0; d < nv; d++) t += e.J[i * nv + d] * qacc[d];
    e.efc_jar[i] = t;
  }
  matvec_w(e.Ma, e.M, qacc, nv);  // ends with WSYNC
}
MGS_DEV real gauss_cost_w(const DevModel &m, const Env &e, const real *qacc) {
  real g = 0;
  PFOR(d, m.nv) g += (e.Ma[d] - e.qfrc_smooth[d]) * (qacc[d] - e.qacc_smooth[d]);
  return R_(0.5) * g;
}

MGS_DEVN void newton_hessian_w(const DevModel &m, Env &e) {
  const int nv = m.nv, npairs = nv * (nv + 1) / 2;
  PFOR(idx, npairs) {
    // decode (a >= b) from the triangular index
    int a = (int)((sqrt(R_(8.0) * idx + 1) - 1) * R_(0.5));
    while ((a + 1) * (a + 2) / 2 <= idx) a++;
    while (a * (a + 1) / 2 > idx) a--;
    int b = idx - a * (a + 1) / 2;
    real s = e.M[a * nv + b];
    for (int i = 0; i < e.nefc; i++) {
      int st = IARR(e.efc_state)[i];
      if (st == ST_QUADRATIC) s += e.efc_D[i] * e.J[i * nv + a] * e.J[i * nv + b];
      else if (st == ST_CONE) {
        int c = IARR(e.efc_id)[i];
        if (IARR(e.con_efc)[c] != i) continue;
        int dim = LDG(m.pair_condim + IARR(e.con_pair)[c]);
        const real *h = e.hcone + 16 * c;
        for (int j = 0; j < dim; j++) {
          real ja = e.J[(i + j) * nv + a];
          if (ja == 0) continue;
          real t = 0;
          for (int k = 0; k < dim; k++) t += h[j * dim + k] * e.J[(i + k) * nv + b];
          s += ja * t;
        }
      }
    }
    e.H[a * nv + b] = s;
  }
  WSYNC();
  chol_factor_w(e.H, nv);
}

MGS_DEVN void solve_newton_w(const DevModel &m, Env &e) {
  const int nv = m.nv;
  const real scale = R_(1.0) / (m.meaninertia * (nv > 1 ? nv : 1));
  e.niter = 0;
  // warm start: cheaper of qacc_warmstart and qacc_smooth
  eval_point_w(m, e, e.qacc_ws);
  real cw = wsum(constraint_update_w(m, e, 0) + gauss_cost_w(m, e, e.qacc_ws));
  WSYNC();
  eval_point_w(m, e, e.qacc_smooth);
  real cs = wsum(constraint_update_w(m, e, 0) + gauss_cost_w(m, e, e.qacc_smooth));
  WSYNC();
  const int use_ws = cw < cs;
  PFOR(d, nv) e.qacc[d] = use_ws ? e.qacc_ws[d] : e.qacc_smooth[d];
  WSYNC();
  eval_point_w(m, e, e.qacc);
  real cost = wsum(constraint_update_w(m, e, 1) + gauss_cost_w(m, e, e.qacc));
  WSYNC();
  for (int iter = 0; iter < m.iterations; iter++) {
    real gn = 0;
    PFOR(d, nv) {
      real t = e.Ma[d] - e.qfrc_smooth[d];
      for (int i = 0; i < e.nefc; i++) t -= e.J[i * nv + d] * e.efc_force[i];
      e.grad[d] = t;
      e.search[d] = t;
      gn += t * t;
    }
    gn = wsum(gn);
    // fp32: the gradient cannot be resolved below ~eps * |force terms|; floor the tolerance accordingly
    real tol_eff = fmax(m.tolerance, R_(20.0) * (real)REAL_EPS * scale * fabs(cost));
    if (iter > 0 && scale * sqrt(gn) < tol_eff) break;
    newton_hessian_w(m, e);
    chol_solve_w(e.H, e.search, nv);
    PFOR(d, nv) e.search[d] = -e.search[d];
    WSYNC();
    matvec_w(e.Mv, e.M, e.search, nv);
    real g1 = 0, g2 = 0, sn = 0;
    PFOR(d, nv) {
      g1 += e.search[d] * (e.Ma[d] - e.qfrc_smooth[d]);
      g2 += e.search[d] * e.Mv[d];
      sn += e.search[d] * e.search[d];
    }
    PFOR(i, e.nefc) {
      real t = 0;
      for (int d = 0; d < nv; d++) t += e.J[i * nv + d] * e.search[d];
      e.efc_jv[i] = t;
    }
    g1 = wsum(g1); g2 = wsum(g2); sn = wsum(sn);
    WSYNC();
    real d1, d2;
    ls_eval_w(m, e, 0, &d1, &d2);
    d1 = wsum(d1) + g1; d2 = wsum(d2) + g2;
    if (!(d1 < 0) || sn < R_(1e-30)) break;
    const real d1_0 = d1;
    real gtol = fmax(m.tolerance * m.ls_tolerance * sqrt(sn) / scale, R_(50.0) * (real)REAL_EPS * fabs(d1_0));
    real alpha = -d1 / d2, lo = 0, hi = -1;
    for (int it = 0; it < m.ls_iterations; it++) {
      ls_eval_w(m, e, alpha, &d1, &d2);
      d1 = wsum(d1) + g1 + alpha * g2; d2 = wsum(d2) + g2;
      if (fabs(d1) < gtol) break;
      if (d1 < 0) lo = alpha; else hi = alpha;
      real an = alpha - d1 / d2;
      if (hi > 0 && (an <= lo || an >= hi)) an = R_(0.5) * (lo + hi);
      if (fabs(an - alpha) <= R_(4.0) * (real)REAL_EPS * fabs(alpha)) { alpha = an; break; }
      alpha = an;
    }
    if (!(alpha > 0)) break;
    PFOR(d, nv) { e.qacc[d] += alpha * e.search[d]; e.Ma[d] += alpha * e.Mv[d]; }
    PFOR(i, e.nefc) e.efc_jar[i] += alpha * e.efc_jv[i];
    WSYNC();
    real oldcost = cost;
    cost = wsum(constraint_update_w(m, e, 1) + gauss_cost_w(m, e, e.qacc));
    WSYNC();
    e.niter = iter + 1;
    tol_eff = fmax(m.tolerance, R_(20.0) * (real)REAL_EPS * scale * fabs(cost));
    if (scale * (oldcost - cost) < tol_eff) break;
  }
  WSYNC();
}

// min 0.5 x'Ax + b'x  s.t. sum (x_j/d_j)^2 <= r^2 , n in {2,3}: Newton on the multiplier
MGS_DEV void qcqp_small(real *res, const real *A, const real *b, const real *d, real r, int n) {
  real As[9], bs[3], v[3] = {0, 0, 0}, la = 0;
  for (int i = 0; i < n; i++) { bs[i] = b[i] * d[i]; for (int j = 0; j < n; j++) As[i * n + j] = A[i * n + j] * d[i] * d[j]; }
  for (int iter = 0; iter < 20; iter++) {
    real P[9], t[3];
    if (n == 2) {
      real a00 = As[0] + la, a01 = As[1], a11 = As[3] + la, det = a00 * a11 - a01 * a01;
      if (det < R_(1e-10)) { res[0] = res[1] = 0; return; }
      real id = R_(1.0) / det;
      P[0] = a11 * id; P[1] = -a01 * id; P[2] = -a01 * id; P[3] = a00 * id;
    } else {
      real a00 = As[0] + la, a01 = As[1], a02 = As[2], a11 = As[4] + la, a12 = As[5], a22 = As[8] + la;
      real c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
      real det = a00 * c00 + a01 * c01 + a02 * c02;
      if (det < R_(1e-10)) { res[0] = res[1] = res[2] = 0; return; }
      real id = R_(1.0) / det;
      P[0] = c00 * id; P[1] = c01 * id; P[2] = c02 * id;
      P[3] = P[1]; P[4] = (a00 * a22 - a02 * a02) * id; P[5] = (a01 * a02 - a00 * a12) * id;
      P[6] = P[2]; P[7] = P[5]; P[8] = (a00 * a11 - a01 * a01) * id;
    }
    real val = -r * r;
    for (int i = 0; i < n; i++) { real s = 0; for (int j = 0; j < n; j++) s -= P[i * n + j] * bs[j]; v[i] = s; val += s * s; }
    if (val < R_(1e-10)) break;
    real deriv = 0;
    for (int i = 0; i < n; i++) { real s = 0; for (int j = 0; j < n; j++) s += P[i * n + j] * v[j]; t[i] = s; deriv -= 2 * v[i] * s; }
    real delta = -val / deriv;
    if (delta < R_(1e-10)) break;
    la += delta;
  }
  for (int i = 0; i < n; i++) res[i] = v[i] * d[i];
}

// mj_solNoSlip.  wvec tracks qacc - qacc_smooth = M^-1 J' f so that no nefc x nefc matrix is formed.
MGS_DEVN void solve_noslip_w(const DevModel &m, Env &e) {
  const int nv = m.nv;
  const real scale = R_(1.0) / (m.meaninertia * (nv > 1 ? nv : 1));
  PFOR(d, nv) e.wvec[d] = e.qacc[d] - e.qacc_smooth[d];
  WSYNC();
  real *B = e.nsB, *S = e.nsS;
  for (int iter = 0; iter < m.noslip_iterations; iter++) {
    real improvement = 0;
    if (iter == 0) {
      real t = 0;
      PFOR(i, e.nefc) {
        int type = IARR(e.efc_type)[i];
        int fr = type == CT_FRICTION_DOF || (type == CT_CONTACT && IARR(e.con_efc)[IARR(e.efc_id)[i]] != i);
        if (fr) t += R_(0.5) * e.efc_force[i] * e.efc_force[i] * e.efc_R[i];
      }
      improvement += wsum(t);
    }
    // dry friction rows
    for (int i = e.ne; i < e.ne + e.nf; i++) {
      const real *Ji = e.J + i * nv;
      PFOR(d, nv) {
        real t = 0;
        for (int k = 0; k < nv; k++) t += e.Minv[d * nv + k] * Ji[k];
        B[d] = t;
      }
      WSYNC();
      real res = 0, Aii = 0;
      PFOR(d, nv) { res += Ji[d] * (e.qacc_smooth[d] + e.wvec[d]); Aii += Ji[d] * B[d]; }
      res = wsum(res) - e.efc_aref[i]; Aii = wsum(Aii);
      real old = e.efc_force[i], fl = e.efc_floss[i];
      real fn = old - res / fmax(MGS_MINVAL, Aii);
      fn = fmax(-fl, fmin(fl, fn));
      real delta = fn - old, change = R_(0.5) * delta * delta * Aii + delta * res;
      if (change > R_(1e-10)) { fn = old; delta = 0; change = 0; }
      WSYNC();
      PFOR(d, nv) e.wvec[d] += B[d] * delta;
      PFOR(k, 1) e.efc_force[i] = fn;
      improvement -= change;
      WSYNC();
    }
    // contact friction dims
    for (int c = 0; c < e.ncon; c++) {
      int i = IARR(e.con_efc)[c];
      if (i < 0) continue;
      int p = IARR(e.con_pair)[c], dim = LDG(m.pair_condim + p), n = dim - 1;
      if (dim < 3) continue;
      if (n > 3) n = 3;
      // B_j = M^-1 J_j'
      PFOR(idx, n * nv) {
        int j = idx / nv, d = idx - j * nv;
        const real *Jj = e.J + (i + 1 + j) * nv;
        real t = 0;
        for (int k = 0; k < nv; k++) t += e.Minv[d * nv + k] * Jj[k];
        B[j * nv + d] = t;
      }
      WSYNC();
      // Ac (n x n) and res (n): one small dot product per lane
      PFOR(idx, n * n + n) {
        if (idx < n * n) {
          int j = idx / n, k = idx - j * n;
          const real *Jj = e.J + (i + 1 + j) * nv;
          real t = 0;
          for (int d = 0; d < nv; d++) t += Jj[d] * B[k * nv + d];
          S[idx] = t;
        } else {
          int j = idx - n * n;
          const real *Jj = e.J + (i + 1 + j) * nv;
          real t = -e.efc_aref[i + 1 + j];
          for (int d = 0; d < nv; d++) t += Jj[d] * (e.qacc_smooth[d] + e.wvec[d]);
          S[idx] = t;
        }
      }
      WSYNC();
      // tiny QCQP, computed redundantly by every lane (no divergence, no extra sync)
      real Ac[9], res[3], old[3], bc[3], v[3], delta[3], fr[3];
      for (int k = 0; k < n * n; k++) Ac[k] = S[k];
      for (int j = 0; j < n; j++) { res[j] = S[n * n + j]; old[j] = e.efc_force[i + 1 + j]; fr[j] = LDG(m.pair_friction + 5 * p + j); }
      for (int j = 0; j < n; j++) { bc[j] = res[j]; for (int k = 0; k < n; k++) bc[j] -= Ac[j * n + k] * old[k]; }
      real fnorm = e.efc_force[i];
      if (fnorm < MGS_MINVAL) { for (int j = 0; j < n; j++) v[j] = 0; }
      else qcqp_small(v, Ac, bc, fr, fnorm, n);
      real change = 0;
      for (int j = 0; j < n; j++) delta[j] = v[j] - old[j];
      for (int j = 0; j < n; j++) { change += delta[j] * res[j]; for (int k = 0; k < n; k++) change += R_(0.5) * delta[j] * Ac[j * n + k] * delta[k]; }
      if (change > R_(1e-10)) { for (int j = 0; j < n; j++) { v[j] = old[j]; delta[j] = 0; } change = 0; }
      WSYNC();
      PFOR(d, nv) { real t = 0; for (int j = 0; j < n; j++) t += B[j * nv + d] * delta[j]; e.wvec[d] += t; }
      PFOR(j, n) e.efc_force[i + 1 + j] = v[j];
      improvement -= change;
      WSYNC();
    }
    if (improvement * scale < m.noslip_tolerance) break;
  }
  // qacc = qacc_smooth + w ; qfrc_constraint = M w
  matvec_w(e.qfrc_constraint, e.M, e.wvec, nv);
  PFOR(d, nv) e.qacc[d] = e.qacc_smooth[d] + e.wvec[d];
  WSYNC();
}

// ---------------------------------------------------------------------------------- forward / integrate
MGS_DEVN void forward_w(const DevModel &m, Env &e) {
  const int nv = m.nv;
  kinematics_w(m, e);
  inertia_w(m, e);
  transmission_w(m, e);
  collision_w(m, e);
  smooth_forces_w(m, e);
  make_constraint_w(m, e);
  if (e.nefc == 0) {
    PFOR(d, nv) { e.qacc[d] = e.qacc_smooth[d]; e.qacc_ws[d] = e.qacc_smooth[d]; e.qfrc_constraint[d] = 0; }
    WSYNC();
    return;
  }
  solve_newton_w(m, e);
  PFOR(d, nv) e.qacc_ws[d] = e.qacc[d];
  WSYNC();
  if (m.noslip_iterations > 0) solve_noslip_w(m, e);
  else {
    PFOR(d, nv) {
      real t = 0;
      for (int i = 0; i < e.nefc; i++) t += e.J[i * nv + d] * e.efc_force[i];
      e.qfrc_constraint[d] = t;
    }
    WSYNC();
  }
}

MGS_DEV int bad_state_w(const DevModel &m, const Env &e, int check_acc) {
  int bad = 0;
  PFOR(i, m.nq) bad |= !(fabs(e.qpos[i]) < R_(1e10));
  PFOR(i, m.nv) bad |= !(fabs(e.qvel[i]) < R_(1e10));
  if (check_acc) PFOR(i, m.nv) bad |= !(fabs(e.qacc[i]) < R_(1e10));
  return wany(bad);
}

// implicitfast: (M - h dF/dv) a = qfrc_smooth + qfrc_constraint; v += h a; q integrates with new v
MGS_DEVN void integrate_w(const DevModel &m, Env &e) {
  const int nv = m.nv;
  const real h = m.timestep;
  PFOR(idx, nv * nv) {
    int i = idx / nv, j = idx - i * nv;
    real a = e.M[idx];
    if (i == j) a += h * LDG(m.dof_damping + i);
    for (int u = 0; u < m.nu; u++) {
      real bv = LDG(m.actuator_biasprm + 3 * u + 2);
      if (bv == 0) continue;
      if (LDG(m.actuator_forcelimited + u)) {
        real f = e.act_force[u];
        if (f <= LDG(m.actuator_forcerange + 2 * u) || f >= LDG(m.actuator_forcerange + 2 * u + 1)) continue;
      }
      a -= h * bv * e.act_moment[u * nv + i] * e.act_moment[u * nv + j];
    }
    e.H[idx] = a;
  }
  PFOR(d, nv) e.search[d] = e.qfrc_smooth[d] + e.qfrc_constraint[d];
  WSYNC();
  chol_factor_w(e.H, nv);
  chol_solve_w(e.H, e.search, nv);
  PFOR(d, nv) e.qvel[d] += h * e.search[d];
  WSYNC();
  PFOR(j, m.njnt) {
    int qa = LDG(m.jnt_qposadr + j), da = LDG(m.jnt_dofadr + j);
    if (LDG(m.jnt_type + j) == JNT_FREE) {
      for (int k = 0; k < 3; k++) e.qpos[qa + k] += h * e.qvel[da + k];
      real w[3] = {e.qvel[da + 3], e.qvel[da + 4], e.qvel[da + 5]};
      real ang = sqrt(dot3(w, w)) * h;
      if (ang > 0) {
        normalize3(w);
        real sn = sin(R_(0.5) * ang), q[4] = {cos(R_(0.5) * ang), sn * w[0], sn * w[1], sn * w[2]}, r[4];
        mulquat(r, e.qpos + qa + 3, q);
        normquat(r);
        e.qpos[qa + 3] = r[0]; e.qpos[qa + 4] = r[1]; e.qpos[qa + 5] = r[2]; e.qpos[qa + 6] = r[3];
      }
    } else e.qpos[qa] += h * e.qvel[da];
  }
  WSYNC();
}

// mj_step x nstep; returns nonzero if the state blew up (the env is then labelled failed)
MGS_DEV int step_w(const DevModel &m, Env &e, int nstep, int *steps_done) {
  for (int k = 0; k < nstep; k++) {
    if (e.bad || bad_state_w(m, e, 0)) { e.bad = 1; return 1; }
    forward_w(m, e);
    if (bad_state_w(m, e, 1)) { e.bad = 1; return 1; }
    integrate_w(m, e);
    (*steps_done)++;
  }
  return 0;
}
