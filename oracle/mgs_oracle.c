/* mgs_oracle.c - CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, fp64, one-environment-at-a-time restatement of the physics the reference obtains
 * from MuJoCo 3.2.2 (`mujoco==3.2.2`, /root/reference/requirements.txt:1, not vendored and not
 * installable here) on the grasp-evaluation rollout:
 *   mj_step / mj_forward call sites: /root/reference/mgs/gripper/panda.py:241,
 *   /root/reference/mgs/env/gravityless_object_grasping.py:159-165,214,244,258,273
 *   rollout logic: /root/reference/mgs/env/gravityless_object_grasping.py:90-125 (collision mask),
 *   :127-295 (close -> lift -> shake), :306-321 (contact tests).
 * The stage-by-stage specification followed is SURVEY.md section 8(a-MJ) (MuJoCo's published
 * algorithm: CRBA/RNE in CoM-based spatial algebra, soft constraints with solref/solimp,
 * elliptic cones, primal Newton solver, noslip, implicitfast).
 *
 * PARITY STATUS: "parity unpinned" against real MuJoCo - the reference ships no golden vectors at
 * the mj_step boundary and MuJoCo cannot be run here.  The oracle is pinned instead by analytic
 * invariants (tests/test_oracle_invariants.py) and by the one MuJoCo-recorded vector the reference holds:
 * the Robotiq closed-state mjSTATE_INTEGRATION record of mgs/cli/config/gripper/robotiq_2f_85.yaml:11
 * (tests/golden/robotiq_2f85_state_close.json, tests/test_golden_robotiq.py).  Known deliberate difference: convex narrowphase builds
 * its multi-point manifold by MPR + face clipping instead of libccd MPR + multiccd perturbation / the analytic box-box routine; the
 * sphere / capsule pairs of MuJoCo's primitive table have closed forms here as well (prim_pair, tests/test_primitive_colliders.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libmgs_b200.so) never links or calls it.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mgs_model_desc.h"

#define MINVAL 1e-15
#define NCON_MAX 128
#define NEFC_MAX 640
#define MAXPOLY 8
#define MAXCLIP 16

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_CYLINDER = 5, GEOM_BOX = 6, GEOM_MESH = 7 };
enum { EQ_CONNECT = 0, EQ_WELD = 1, EQ_JOINT = 2 };
enum { CT_EQUALITY = 0, CT_FRICTION_DOF = 1, CT_LIMIT = 2, CT_CONTACT = 3 };

typedef struct {
  double pos[3], frame[9], dist;
  int pair, geom1, geom2, body1, body2, dim, efc;
  double friction[5], solref[2], solimp[5], mu;
} Contact;

typedef struct OrcSim {
  MgsModelDesc m; /* deep copy */
  void *blocks[256];
  int nblocks;
  double *body_subtreemass;
  /* state */
  double time, *qpos, *qvel, *ctrl, *mocap_pos, *mocap_quat, *qacc_warmstart;
  /* position-dependent */
  double *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *gxpos, *gxmat;
  double *subtree_com, *cinert, *crb, *cdof, *cdof_dot, *cvel, *cacc, *cfrc;
  double *M, *L, *H, *ten_length, *ten_J, *act_moment, *act_force, *act_length, *act_velocity;
  double *qfrc_passive, *qfrc_bias, *qfrc_actuator, *qfrc_smooth, *qacc_smooth, *qacc, *qfrc_constraint;
  /* contacts and constraints */
  int ncon, nefc, ne, nf, nl;
  Contact con[NCON_MAX];
  double *J; /* NEFC_MAX x nv */
  double efc_pos[NEFC_MAX], efc_margin[NEFC_MAX], efc_D[NEFC_MAX], efc_R[NEFC_MAX], efc_aref[NEFC_MAX];
  double efc_floss[NEFC_MAX], efc_force[NEFC_MAX], efc_jar[NEFC_MAX], efc_b[NEFC_MAX], efc_vel[NEFC_MAX];
  double efc_diagApprox[NEFC_MAX], efc_KBIP[NEFC_MAX][4];
  int efc_type[NEFC_MAX], efc_id[NEFC_MAX], efc_state[NEFC_MAX];
  /* diagnostics */
  int solver_niter, bad, nstep_done, ncon_overflow, ncon_peak, nefc_peak;
  int no_analytic; /* tests only: 1 = send primitive pairs through MPR as well (orc_set_analytic) */
  double qvel_clip; /* > 0: clamp qvel to +-qvel_clip before every step (ClutterTableEnv.gen_clutter, clutter_table.py:215-221) */
  /* scratch */
  double *w1, *w2, *w3, *w4, *w5, *w6, *jtmp;
} OrcSim;

/* ------------------------------------------------------------------ small math */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double *r, const double *a, const double *b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void copy3(double *r, const double *a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline void add3(double *r, const double *a, const double *b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline void sub3(double *r, const double *a, const double *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void scl3(double *r, const double *a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
static inline void addscl3(double *r, const double *a, double s) { r[0] += a[0] * s; r[1] += a[1] * s; r[2] += a[2] * s; }
static inline double norm3(const double *a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double *a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
static void mulquat(double *r, const double *a, const double *b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void normquat(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void quat2mat(double *R, const double *q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
static inline void mulmatvec3(double *r, const double *R, const double *v) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
         z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmatTvec3(double *r, const double *R, const double *v) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2], y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2],
         z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void rotvecquat(double *r, const double *v, const double *q) {
  double R[9];
  quat2mat(R, q);
  mulmatvec3(r, R, v);
}

/* dense Cholesky (lower) of an n x n SPD matrix stored row-major; returns rank deficiency count */
static int chol_factor(double *L, const double *A, int n) {
  int bad = 0;
  memcpy(L, A, sizeof(double) * n * n);
  for (int j = 0; j < n; j++) {
    double d = L[j * n + j];
    for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
    if (d < MINVAL) { d = MINVAL; bad++; }
    d = sqrt(d);
    L[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = L[i * n + j];
      for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / d;
    }
  }
  return bad;
}
static void chol_solve(const double *L, double *x, int n) { /* in place */
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

/* ------------------------------------------------------------------ create / destroy */
static void *ALLOC(OrcSim *s, size_t bytes) {
  void *p = calloc(1, bytes ? bytes : 8);
  s->blocks[s->nblocks++] = p;
  return p;
}
#define CPD(name, cnt) do { size_t n_ = (size_t)(cnt); double *p_ = (double *)ALLOC(s, n_ * sizeof(double)); if (n_) memcpy(p_, d->name, n_ * sizeof(double)); s->m.name = p_; } while (0)
#define CPI(name, cnt) do { size_t n_ = (size_t)(cnt); int *p_ = (int *)ALLOC(s, n_ * sizeof(int)); if (n_) memcpy(p_, d->name, n_ * sizeof(int)); s->m.name = p_; } while (0)
#define NEWD(cnt) ((double *)ALLOC(s, (size_t)(cnt) * sizeof(double)))

void orc_reset(OrcSim *s);

OrcSim *orc_create(const MgsModelDesc *d) {
  OrcSim *s = (OrcSim *)calloc(1, sizeof(OrcSim));
  s->m = *d;
  int nb = d->nbody, nv = d->nv, nq = d->nq, nj = d->njnt, nu = d->nu, ng = d->ncgeom, np = d->npair, nh = d->nhull;
  CPI(body_parentid, nb); CPI(body_rootid, nb); CPI(body_weldid, nb); CPI(body_mocapid, nb); CPI(body_jntadr, nb);
  CPI(body_jntnum, nb); CPI(body_dofadr, nb); CPI(body_dofnum, nb); CPD(body_pos, 3 * nb); CPD(body_quat, 4 * nb);
  CPD(body_ipos, 3 * nb); CPD(body_iquat, 4 * nb); CPD(body_mass, nb); CPD(body_inertia, 3 * nb); CPD(body_gravcomp, nb);
  CPD(body_invweight0, 2 * nb);
  CPI(jnt_type, nj); CPI(jnt_bodyid, nj); CPI(jnt_qposadr, nj); CPI(jnt_dofadr, nj); CPI(jnt_limited, nj);
  CPD(jnt_pos, 3 * nj); CPD(jnt_axis, 3 * nj); CPD(jnt_range, 2 * nj); CPD(jnt_stiffness, nj); CPD(jnt_solref, 2 * nj);
  CPD(jnt_solimp, 5 * nj); CPD(jnt_margin, nj); CPD(qpos0, nq); CPD(qpos_spring, nq);
  CPI(dof_bodyid, nv); CPI(dof_jntid, nv); CPI(dof_parentid, nv); CPD(dof_armature, nv); CPD(dof_damping, nv);
  CPD(dof_frictionloss, nv); CPD(dof_solref, 2 * nv); CPD(dof_solimp, 5 * nv); CPD(dof_invweight0, nv);
  CPI(cgeom_geomid, ng); CPI(cgeom_type, ng); CPI(cgeom_bodyid, ng); CPI(cgeom_hullid, ng); CPD(cgeom_pos, 3 * ng);
  CPD(cgeom_quat, 4 * ng); CPD(cgeom_size, 3 * ng); CPD(cgeom_rbound, ng);
  CPI(hull_vertadr, nh); CPI(hull_vertnum, nh); CPI(hull_faceadr, nh); CPI(hull_facenum, nh);
  CPD(hull_vert, 3 * d->nhullvert); CPD(hull_facenormal, 3 * d->nhullface); CPI(hull_facevertadr, d->nhullface);
  CPI(hull_facevertnum, d->nhullface); CPI(hull_facevert, d->nhullfacevert); CPI(hull_nbradr, d->nhullvert);
  CPI(hull_nbrnum, d->nhullvert); CPI(hull_nbr, d->nhullnbr);
  CPI(pair_geom1, np); CPI(pair_geom2, np); CPI(pair_condim, np); CPD(pair_friction, 5 * np); CPD(pair_solref, 2 * np);
  CPD(pair_solimp, 5 * np); CPD(pair_margin, np); CPD(pair_gap, np);
  CPI(tendon_adr, d->ntendon); CPI(tendon_num, d->ntendon); CPI(wrap_dofadr, d->nwrap); CPI(wrap_qposadr, d->nwrap);
  CPD(wrap_coef, d->nwrap);
  CPI(actuator_trntype, nu); CPI(actuator_trnid, nu); CPI(actuator_ctrllimited, nu); CPI(actuator_forcelimited, nu);
  CPD(actuator_gainprm, 3 * nu); CPD(actuator_biasprm, 3 * nu); CPD(actuator_ctrlrange, 2 * nu);
  CPD(actuator_forcerange, 2 * nu); CPD(actuator_gear, nu);
  CPI(eq_type, d->neq); CPI(eq_obj1id, d->neq); CPI(eq_obj2id, d->neq); CPI(eq_active, d->neq); CPD(eq_data, 11 * d->neq);
  CPD(eq_solref, 2 * d->neq); CPD(eq_solimp, 5 * d->neq); CPD(mocap_pos0, 3 * d->nmocap); CPD(mocap_quat0, 4 * d->nmocap);

  s->body_subtreemass = NEWD(nb);
  for (int b = 0; b < nb; b++) s->body_subtreemass[b] = s->m.body_mass[b];
  for (int b = nb - 1; b > 0; b--) s->body_subtreemass[s->m.body_parentid[b]] += s->body_subtreemass[b];

  s->qpos = NEWD(nq); s->qvel = NEWD(nv); s->ctrl = NEWD(nu); s->mocap_pos = NEWD(3 * d->nmocap);
  s->mocap_quat = NEWD(4 * d->nmocap); s->qacc_warmstart = NEWD(nv);
  s->xpos = NEWD(3 * nb); s->xquat = NEWD(4 * nb); s->xmat = NEWD(9 * nb); s->xipos = NEWD(3 * nb); s->ximat = NEWD(9 * nb);
  s->xanchor = NEWD(3 * nj); s->xaxis = NEWD(3 * nj); s->gxpos = NEWD(3 * ng); s->gxmat = NEWD(9 * ng);
  s->subtree_com = NEWD(3 * nb); s->cinert = NEWD(10 * nb); s->crb = NEWD(10 * nb); s->cdof = NEWD(6 * nv);
  s->cdof_dot = NEWD(6 * nv); s->cvel = NEWD(6 * nb); s->cacc = NEWD(6 * nb); s->cfrc = NEWD(6 * nb);
  s->M = NEWD(nv * nv); s->L = NEWD(nv * nv); s->H = NEWD(nv * nv);
  s->ten_length = NEWD(d->ntendon); s->ten_J = NEWD(d->ntendon * nv);
  s->act_moment = NEWD(nu * nv); s->act_force = NEWD(nu); s->act_length = NEWD(nu); s->act_velocity = NEWD(nu);
  s->qfrc_passive = NEWD(nv); s->qfrc_bias = NEWD(nv); s->qfrc_actuator = NEWD(nv); s->qfrc_smooth = NEWD(nv);
  s->qacc_smooth = NEWD(nv); s->qacc = NEWD(nv); s->qfrc_constraint = NEWD(nv);
  s->J = NEWD(NEFC_MAX * nv);
  s->w1 = NEWD(nv * nv + 6 * nv); s->w2 = NEWD(nv * nv + 6 * nv); s->w3 = NEWD(nv); s->w4 = NEWD(nv); s->w5 = NEWD(nv); s->w6 = NEWD(nv);
  s->jtmp = NEWD(12 * nv);
  orc_reset(s);
  return s;
}

void orc_destroy(OrcSim *s) {
  if (!s) return;
  for (int i = 0; i < s->nblocks; i++) free(s->blocks[i]);
  free(s);
}

/* mj_resetData: qpos0, zero velocities/controls/warmstart, mocap back to the model pose */
void orc_reset(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  memcpy(s->qpos, m->qpos0, sizeof(double) * m->nq);
  memset(s->qvel, 0, sizeof(double) * m->nv);
  memset(s->qacc_warmstart, 0, sizeof(double) * m->nv);
  memset(s->ctrl, 0, sizeof(double) * m->nu);
  memcpy(s->mocap_pos, m->mocap_pos0, sizeof(double) * 3 * m->nmocap);
  memcpy(s->mocap_quat, m->mocap_quat0, sizeof(double) * 4 * m->nmocap);
  s->time = 0; s->bad = 0; s->ncon = 0; s->nefc = 0; s->nstep_done = 0;
}

/* ------------------------------------------------------------------ smooth dynamics */
static void jac_point(const OrcSim *s, int body, const double *point, double *jacp, double *jacr);
/* mj_kinematics (SURVEY 8(a-MJ) "Kinematics") */
static void kinematics(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  s->xpos[0] = s->xpos[1] = s->xpos[2] = 0;
  s->xquat[0] = 1; s->xquat[1] = s->xquat[2] = s->xquat[3] = 0;
  quat2mat(s->xmat, s->xquat);
  copy3(s->xipos, s->xpos); quat2mat(s->ximat, s->xquat);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b];
    double *xp = s->xpos + 3 * b, *xq = s->xquat + 4 * b;
    if (m->body_mocapid[b] >= 0) {
      int id = m->body_mocapid[b];
      copy3(xp, s->mocap_pos + 3 * id);
      memcpy(xq, s->mocap_quat + 4 * id, 4 * sizeof(double));
      normquat(xq);
    } else {
      double t[3];
      mulmatvec3(t, s->xmat + 9 * p, m->body_pos + 3 * b);
      add3(xp, s->xpos + 3 * p, t);
      mulquat(xq, s->xquat + 4 * p, m->body_quat + 4 * b);
    }
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      int qa = m->jnt_qposadr[j];
      double *anchor = s->xanchor + 3 * j, *axis = s->xaxis + 3 * j;
      if (m->jnt_type[j] == JNT_FREE) {
        copy3(xp, s->qpos + qa);
        normquat(s->qpos + qa + 3);
        memcpy(xq, s->qpos + qa + 3, 4 * sizeof(double));
        copy3(anchor, xp);
        axis[0] = 0; axis[1] = 0; axis[2] = 1;
        continue;
      }
      rotvecquat(axis, m->jnt_axis + 3 * j, xq);
      rotvecquat(anchor, m->jnt_pos + 3 * j, xq);
      add3(anchor, anchor, xp);
      double dq = s->qpos[qa] - m->qpos0[qa];
      if (m->jnt_type[j] == JNT_SLIDE) {
        addscl3(xp, axis, dq);
      } else { /* hinge */
        double ql[4] = {cos(0.5 * dq), sin(0.5 * dq) * m->jnt_axis[3 * j], sin(0.5 * dq) * m->jnt_axis[3 * j + 1],
                        sin(0.5 * dq) * m->jnt_axis[3 * j + 2]};
        double t[3];
        mulquat(xq, xq, ql);
        rotvecquat(t, m->jnt_pos + 3 * j, xq);
        sub3(xp, anchor, t);
      }
    }
    normquat(xq);
    quat2mat(s->xmat + 9 * b, xq);
    double t[3], qi[4];
    mulmatvec3(t, s->xmat + 9 * b, m->body_ipos + 3 * b);
    add3(s->xipos + 3 * b, xp, t);
    mulquat(qi, xq, m->body_iquat + 4 * b);
    quat2mat(s->ximat + 9 * b, qi);
  }
  for (int g = 0; g < m->ncgeom; g++) {
    int b = m->cgeom_bodyid[g];
    double t[3], q[4];
    mulmatvec3(t, s->xmat + 9 * b, m->cgeom_pos + 3 * g);
    add3(s->gxpos + 3 * g, s->xpos + 3 * b, t);
    mulquat(q, s->xquat + 4 * b, m->cgeom_quat + 4 * g);
    quat2mat(s->gxmat + 9 * g, q);
  }
}

/* spatial inertia about `origin + dif`: [Ixx Iyy Izz Ixy Ixz Iyz, m*dif, m] */
static void inert_com(double *res, const double *diag, const double *R, const double *dif, double mass) {
  double I[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      I[3 * i + j] = R[3 * i] * diag[0] * R[3 * j] + R[3 * i + 1] * diag[1] * R[3 * j + 1] + R[3 * i + 2] * diag[2] * R[3 * j + 2];
  double d2 = dot3(dif, dif);
  res[0] = I[0] + mass * (d2 - dif[0] * dif[0]);
  res[1] = I[4] + mass * (d2 - dif[1] * dif[1]);
  res[2] = I[8] + mass * (d2 - dif[2] * dif[2]);
  res[3] = I[1] - mass * dif[0] * dif[1];
  res[4] = I[2] - mass * dif[0] * dif[2];
  res[5] = I[5] - mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2];
  res[9] = mass;
}
/* res = I * v for spatial v = [ang; lin] */
static void mul_inert_vec(double *res, const double *I, const double *v) {
  double t[3];
  res[0] = I[0] * v[0] + I[3] * v[1] + I[4] * v[2];
  res[1] = I[3] * v[0] + I[1] * v[1] + I[5] * v[2];
  res[2] = I[4] * v[0] + I[5] * v[1] + I[2] * v[2];
  cross3(t, I + 6, v + 3);
  add3(res, res, t);
  cross3(t, I + 6, v);
  res[3] = I[9] * v[3] - t[0]; res[4] = I[9] * v[4] - t[1]; res[5] = I[9] * v[5] - t[2];
}
static void cross_motion(double *res, const double *vel, const double *v) {
  double t[3];
  cross3(res, vel, v);
  cross3(res + 3, vel, v + 3);
  cross3(t, vel + 3, v);
  add3(res + 3, res + 3, t);
}
static void cross_force(double *res, const double *vel, const double *f) {
  double t[3];
  cross3(res, vel, f);
  cross3(t, vel + 3, f + 3);
  add3(res, res, t);
  cross3(res + 3, vel, f + 3);
}

/* mj_comPos: subtree CoM, CoM-based inertias and motion axes */
static void com_pos(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  int nb = m->nbody;
  for (int b = 0; b < nb; b++) scl3(s->subtree_com + 3 * b, s->xipos + 3 * b, m->body_mass[b]);
  for (int b = nb - 1; b > 0; b--) add3(s->subtree_com + 3 * m->body_parentid[b], s->subtree_com + 3 * m->body_parentid[b], s->subtree_com + 3 * b);
  for (int b = 0; b < nb; b++) {
    if (s->body_subtreemass[b] < MINVAL) copy3(s->subtree_com + 3 * b, s->xipos + 3 * b);
    else scl3(s->subtree_com + 3 * b, s->subtree_com + 3 * b, 1.0 / s->body_subtreemass[b]);
  }
  memset(s->cinert, 0, 10 * sizeof(double));
  for (int b = 1; b < nb; b++) {
    double dif[3];
    sub3(dif, s->xipos + 3 * b, s->subtree_com + 3 * m->body_rootid[b]);
    inert_com(s->cinert + 10 * b, m->body_inertia + 3 * b, s->ximat + 9 * b, dif, m->body_mass[b]);
  }
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    double off[3];
    sub3(off, s->subtree_com + 3 * m->body_rootid[b], s->xanchor + 3 * j);
    if (m->jnt_type[j] == JNT_FREE) {
      memset(s->cdof + 6 * da, 0, 36 * sizeof(double));
      for (int k = 0; k < 3; k++) s->cdof[6 * (da + k) + 3 + k] = 1;
      for (int k = 0; k < 3; k++) {
        double ax[3] = {s->xmat[9 * b + k], s->xmat[9 * b + 3 + k], s->xmat[9 * b + 6 + k]};
        double *c = s->cdof + 6 * (da + 3 + k);
        copy3(c, ax);
        cross3(c + 3, ax, off);
      }
    } else if (m->jnt_type[j] == JNT_SLIDE) {
      double *c = s->cdof + 6 * da;
      c[0] = c[1] = c[2] = 0;
      copy3(c + 3, s->xaxis + 3 * j);
    } else {
      double *c = s->cdof + 6 * da;
      copy3(c, s->xaxis + 3 * j);
      cross3(c + 3, s->xaxis + 3 * j, off);
    }
  }
}

/* fixed tendons: length and Jacobian */
static void tendon(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  memset(s->ten_J, 0, sizeof(double) * m->ntendon * m->nv);
  for (int t = 0; t < m->ntendon; t++) {
    double L = 0;
    for (int w = m->tendon_adr[t]; w < m->tendon_adr[t] + m->tendon_num[t]; w++) {
      L += m->wrap_coef[w] * s->qpos[m->wrap_qposadr[w]];
      s->ten_J[t * m->nv + m->wrap_dofadr[w]] += m->wrap_coef[w];
    }
    s->ten_length[t] = L;
  }
}

/* mj_transmission: actuator length and moment arm */
static void transmission(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  int nv = m->nv;
  memset(s->act_moment, 0, sizeof(double) * m->nu * nv);
  for (int a = 0; a < m->nu; a++) {
    double gear = m->actuator_gear[a];
    if (m->actuator_trntype[a] == 0) {
      int j = m->actuator_trnid[a];
      s->act_length[a] = gear * s->qpos[m->jnt_qposadr[j]];
      s->act_moment[a * nv + m->jnt_dofadr[j]] = gear;
    } else {
      int t = m->actuator_trnid[a];
      s->act_length[a] = gear * s->ten_length[t];
      for (int d = 0; d < nv; d++) s->act_moment[a * nv + d] = gear * s->ten_J[t * nv + d];
    }
  }
}

/* mj_crb + factor: composite rigid body mass matrix (dense) and its Cholesky factor */
static void crb_factor(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  int nb = m->nbody, nv = m->nv;
  memcpy(s->crb, s->cinert, sizeof(double) * 10 * nb);
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int k = 0; k < 10; k++) s->crb[10 * p + k] += s->crb[10 * b + k];
  }
  memset(s->M, 0, sizeof(double) * nv * nv);
  for (int i = 0; i < nv; i++) {
    double buf[6];
    mul_inert_vec(buf, s->crb + 10 * m->dof_bodyid[i], s->cdof + 6 * i);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double v = 0;
      for (int k = 0; k < 6; k++) v += s->cdof[6 * j + k] * buf[k];
      s->M[i * nv + j] = s->M[j * nv + i] = v;
    }
    s->M[i * nv + i] += m->dof_armature[i];
  }
  chol_factor(s->L, s->M, nv);
}

/* mj_comVel */
static void com_vel(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  memset(s->cvel, 0, 6 * sizeof(double));
  for (int b = 1; b < m->nbody; b++) {
    double cv[6];
    memcpy(cv, s->cvel + 6 * m->body_parentid[b], 6 * sizeof(double));
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      int da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == JNT_FREE) {
        memset(s->cdof_dot + 6 * da, 0, 18 * sizeof(double));
        for (int k = 0; k < 3; k++)
          for (int c = 0; c < 6; c++) cv[c] += s->cdof[6 * (da + k) + c] * s->qvel[da + k];
        for (int k = 3; k < 6; k++) cross_motion(s->cdof_dot + 6 * (da + k), cv, s->cdof + 6 * (da + k));
        for (int k = 3; k < 6; k++)
          for (int c = 0; c < 6; c++) cv[c] += s->cdof[6 * (da + k) + c] * s->qvel[da + k];
      } else {
        cross_motion(s->cdof_dot + 6 * da, cv, s->cdof + 6 * da);
        for (int c = 0; c < 6; c++) cv[c] += s->cdof[6 * da + c] * s->qvel[da];
      }
    }
    memcpy(s->cvel + 6 * b, cv, 6 * sizeof(double));
  }
}

/* mj_passive: joint springs and dampers (gravcomp: bodies with gravcomp != 0) */
static void passive(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  for (int d = 0; d < m->nv; d++) s->qfrc_passive[d] = -m->dof_damping[d] * s->qvel[d];
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] == JNT_FREE || m->jnt_stiffness[j] == 0) continue;
    int qa = m->jnt_qposadr[j];
    s->qfrc_passive[m->jnt_dofadr[j]] -= m->jnt_stiffness[j] * (s->qpos[qa] - m->qpos_spring[qa]);
  }
  /* gravity compensation: force -gravity * mass * gravcomp at the body CoM (clutter_table.py:56, camera body) */
  for (int b = 1; b < m->nbody; b++) {
    if (m->body_gravcomp[b] == 0 || m->body_mass[b] == 0) continue;
    double *jp = s->jtmp;
    jac_point(s, b, s->xipos + 3 * b, jp, NULL);
    for (int d = 0; d < m->nv; d++)
      for (int k = 0; k < 3; k++) s->qfrc_passive[d] += jp[k * m->nv + d] * (-m->gravity[k] * m->body_mass[b] * m->body_gravcomp[b]);
  }
}

/* mj_rne(flg_acc=0): Coriolis, centrifugal and gravity forces */
static void rne(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  int nb = m->nbody;
  double *cacc = s->cacc, *cfrc = s->cfrc;
  cacc[0] = cacc[1] = cacc[2] = 0;
  cacc[3] = -m->gravity[0]; cacc[4] = -m->gravity[1]; cacc[5] = -m->gravity[2];
  memset(cfrc, 0, 6 * sizeof(double));
  for (int b = 1; b < nb; b++) {
    double *a = cacc + 6 * b, t1[6], t2[6];
    memcpy(a, cacc + 6 * m->body_parentid[b], 6 * sizeof(double));
    for (int d = m->body_dofadr[b]; d >= 0 && d < m->body_dofadr[b] + m->body_dofnum[b]; d++)
      for (int c = 0; c < 6; c++) a[c] += s->cdof_dot[6 * d + c] * s->qvel[d];
    mul_inert_vec(t1, s->cinert + 10 * b, s->cvel + 6 * b);
    cross_force(t2, s->cvel + 6 * b, t1);
    mul_inert_vec(t1, s->cinert + 10 * b, a);
    for (int c = 0; c < 6; c++) cfrc[6 * b + c] = t1[c] + t2[c];
  }
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int c = 0; c < 6; c++) cfrc[6 * p + c] += cfrc[6 * b + c];
  }
  for (int d = 0; d < m->nv; d++) {
    double v = 0;
    for (int c = 0; c < 6; c++) v += s->cdof[6 * d + c] * cfrc[6 * m->dof_bodyid[d] + c];
    s->qfrc_bias[d] = v;
  }
}

/* mj_fwdActuation: ctrl clamp, affine gain/bias, force clamp, generalized force */
static void actuation(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  int nv = m->nv;
  memset(s->qfrc_actuator, 0, sizeof(double) * nv);
  for (int a = 0; a < m->nu; a++) {
    double c = s->ctrl[a], vel = 0;
    if (m->actuator_ctrllimited[a]) c = fmax(m->actuator_ctrlrange[2 * a], fmin(m->actuator_ctrlrange[2 * a + 1], c));
    for (int d = 0; d < nv; d++) vel += s->act_moment[a * nv + d] * s->qvel[d];
    s->act_velocity[a] = vel;
    const double *g = m->actuator_gainprm + 3 * a, *bp = m->actuator_biasprm + 3 * a;
    double f = g[0] * c + bp[0] + bp[1] * s->act_length[a] + bp[2] * vel;
    if (m->actuator_forcelimited[a]) f = fmax(m->actuator_forcerange[2 * a], fmin(m->actuator_forcerange[2 * a + 1], f));
    s->act_force[a] = f;
    for (int d = 0; d < nv; d++) s->qfrc_actuator[d] += s->act_moment[a * nv + d] * f;
  }
}

/* Jacobian of a world point attached to `body`: jacp, jacr are 3 x nv (row-major), zeroed here */
static void jac_point(const OrcSim *s, int body, const double *point, double *jacp, double *jacr) {
  const MgsModelDesc *m = &s->m;
  int nv = m->nv;
  if (jacp) memset(jacp, 0, sizeof(double) * 3 * nv);
  if (jacr) memset(jacr, 0, sizeof(double) * 3 * nv);
  double off[3];
  sub3(off, point, s->subtree_com + 3 * m->body_rootid[body]);
  while (body > 0 && m->body_dofnum[body] == 0) body = m->body_parentid[body];
  if (body == 0) return;
  for (int d = m->body_dofadr[body] + m->body_dofnum[body] - 1; d >= 0; d = m->dof_parentid[d]) {
    const double *c = s->cdof + 6 * d;
    if (jacr) { jacr[d] = c[0]; jacr[nv + d] = c[1]; jacr[2 * nv + d] = c[2]; }
    if (jacp) {
      double t[3];
      cross3(t, c, off);
      jacp[d] = c[3] + t[0]; jacp[nv + d] = c[4] + t[1]; jacp[2 * nv + d] = c[5] + t[2];
    }
  }
}

/* ------------------------------------------------------------------ collision */
/* Convex narrowphase.  MuJoCo 3.2.2 routes mesh pairs through libccd's MPR (Snethen's Minkowski
 * Portal Refinement, restated here from the published algorithm) and adds extra points with the
 * multiccd perturbation; box-box has an analytic routine.  This restatement treats boxes and
 * hulls uniformly as convex polytopes: MPR supplies the penetration direction/depth, then the
 * most-aligned faces of the two polytopes are clipped against each other (Sutherland-Hodgman)
 * to produce up to 4 contact points on the reference face (a cylinder offers its caps as faces).  Pairs of
 * MuJoCo's primitive table (sphere / capsule against sphere / capsule / box, sphere-cylinder) have closed
 * forms further down (prim_pair); the remaining round pairs get the single MPR point. */
typedef struct { double v[3], v1[3], v2[3]; } SupPt;

static int is_polytope(int type) { return type == GEOM_BOX || type == GEOM_MESH; }

/* support point of collision geom g (world frame) in world direction d (unit) */
static void geom_support(const OrcSim *s, int g, const double *d, double *out) {
  const MgsModelDesc *m = &s->m;
  const double *R = s->gxmat + 9 * g, *sz = m->cgeom_size + 3 * g;
  double dl[3], p[3] = {0, 0, 0};
  mulmatTvec3(dl, R, d);
  int type = m->cgeom_type[g];
  if (type == GEOM_BOX) {
    p[0] = dl[0] >= 0 ? sz[0] : -sz[0]; p[1] = dl[1] >= 0 ? sz[1] : -sz[1]; p[2] = dl[2] >= 0 ? sz[2] : -sz[2];
  } else if (type == GEOM_MESH) {
    int h = m->cgeom_hullid[g], best = 0;
    const double *V = m->hull_vert + 3 * m->hull_vertadr[h];
    double bd = -1e300;
    for (int i = 0; i < m->hull_vertnum[h]; i++) {
      double t = dot3(V + 3 * i, dl);
      if (t > bd) { bd = t; best = i; }
    }
    copy3(p, V + 3 * best);
  } else if (type == GEOM_SPHERE) {
    scl3(p, dl, sz[0]);
  } else if (type == GEOM_CAPSULE) {
    scl3(p, dl, sz[0]);
    p[2] += dl[2] >= 0 ? sz[1] : -sz[1];
  } else if (type == GEOM_CYLINDER) {
    double n = sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
    if (n > MINVAL) { p[0] = dl[0] / n * sz[0]; p[1] = dl[1] / n * sz[0]; }
    p[2] = dl[2] >= 0 ? sz[1] : -sz[1];
  }
  mulmatvec3(out, R, p);
  add3(out, out, s->gxpos + 3 * g);
}

static void mink_support(const OrcSim *s, int g1, int g2, const double *d, SupPt *o) {
  double nd[3] = {-d[0], -d[1], -d[2]};
  geom_support(s, g1, d, o->v1);
  geom_support(s, g2, nd, o->v2);
  sub3(o->v, o->v1, o->v2);
}

static void portal_dir(const SupPt *p, double *dir) {
  double a[3], b[3];
  sub3(a, p[2].v, p[1].v);
  sub3(b, p[3].v, p[1].v);
  cross3(dir, a, b);
  normalize3(dir);
}

static void expand_portal(SupPt *p, const SupPt *v4) {
  double v4v0[3];
  cross3(v4v0, v4->v, p[0].v);
  if (dot3(p[1].v, v4v0) > 0) {
    if (dot3(p[2].v, v4v0) > 0) p[1] = *v4; else p[3] = *v4;
  } else {
    if (dot3(p[3].v, v4v0) > 0) p[2] = *v4; else p[1] = *v4;
  }
}

static int reach_tolerance(const SupPt *p, const SupPt *v4, const double *dir, double tol) {
  double d1 = dot3(p[1].v, dir), d2 = dot3(p[2].v, dir), d3 = dot3(p[3].v, dir), d4 = dot3(v4->v, dir);
  double mn = fmin(d4 - d1, fmin(d4 - d2, d4 - d3));
  return mn <= tol;
}

/* closest point to the origin on triangle a,b,c (Ericson 5.1.5) */
static void closest_on_triangle(const double *a, const double *b, const double *c, double *out) {
  double ab[3], ac[3], ap[3] = {-a[0], -a[1], -a[2]};
  sub3(ab, b, a); sub3(ac, c, a);
  double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0 && d2 <= 0) { copy3(out, a); return; }
  double bp[3] = {-b[0], -b[1], -b[2]};
  double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0 && d4 <= d3) { copy3(out, b); return; }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); copy3(out, a); addscl3(out, ab, v); return; }
  double cp[3] = {-c[0], -c[1], -c[2]};
  double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0 && d5 <= d6) { copy3(out, c); return; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); copy3(out, a); addscl3(out, ac, w); return; }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    double w = (d4 - d3) / ((d4 - d3) + (d5 - d6)), bc[3];
    sub3(bc, c, b); copy3(out, b); addscl3(out, bc, w); return;
  }
  double den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
  copy3(out, a); addscl3(out, ab, v); addscl3(out, ac, w);
}

/* returns 1 if penetrating; fills depth (>0), dir (unit, from g1 to g2) and pos */
static int mpr_penetration(const OrcSim *s, int g1, int g2, double *depth, double *dir, double *pos) {
  const MgsModelDesc *m = &s->m;
  const double tol = m->mpr_tolerance;
  SupPt p[4], v4;
  double d[3], va[3], vb[3];
  /* portal centre: interior point of the Minkowski difference */
  copy3(p[0].v1, s->gxpos + 3 * g1); copy3(p[0].v2, s->gxpos + 3 * g2);
  sub3(p[0].v, p[0].v1, p[0].v2);
  if (dot3(p[0].v, p[0].v) < 1e-28) p[0].v[0] += 1e-10;
  scl3(d, p[0].v, -1); normalize3(d);
  mink_support(s, g1, g2, d, &p[1]);
  if (dot3(p[1].v, d) <= 0) return 0;
  cross3(d, p[0].v, p[1].v);
  if (dot3(d, d) < 1e-28 * fmax(1e-30, dot3(p[0].v, p[0].v) * dot3(p[1].v, p[1].v))) {
    /* origin on the segment v0-v1: penetration along v1 */
    *depth = norm3(p[1].v);
    copy3(dir, p[1].v); normalize3(dir);
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p[1].v1[k] + p[1].v2[k]);
    return *depth > 0;
  }
  normalize3(d);
  mink_support(s, g1, g2, d, &p[2]);
  if (dot3(p[2].v, d) <= 0) return 0;
  sub3(va, p[1].v, p[0].v); sub3(vb, p[2].v, p[0].v);
  cross3(d, va, vb); normalize3(d);
  if (dot3(d, p[0].v) > 0) { SupPt t = p[1]; p[1] = p[2]; p[2] = t; scl3(d, d, -1); }
  for (int it = 0;; it++) {
    if (it > 100) return 0;
    mink_support(s, g1, g2, d, &p[3]);
    if (dot3(p[3].v, d) <= 0) return 0;
    int cont = 0;
    cross3(va, p[1].v, p[3].v);
    if (dot3(va, p[0].v) < -1e-300) { p[2] = p[3]; cont = 1; }
    if (!cont) {
      cross3(va, p[3].v, p[2].v);
      if (dot3(va, p[0].v) < -1e-300) { p[1] = p[3]; cont = 1; }
    }
    if (!cont) break;
    sub3(va, p[1].v, p[0].v); sub3(vb, p[2].v, p[0].v);
    cross3(d, va, vb); normalize3(d);
  }
  /* refine until the portal face passes the origin */
  for (int it = 0;; it++) {
    portal_dir(p, d);
    if (dot3(d, p[1].v) >= 0) break; /* origin inside the portal */
    mink_support(s, g1, g2, d, &v4);
    if (dot3(v4.v, d) < 0 || reach_tolerance(p, &v4, d, tol) || it > m->mpr_iterations) return 0;
    expand_portal(p, &v4);
  }
  /* find penetration */
  for (int it = 0;; it++) {
    portal_dir(p, d);
    mink_support(s, g1, g2, d, &v4);
    if (reach_tolerance(p, &v4, d, tol) || it > m->mpr_iterations) {
      double c[3];
      closest_on_triangle(p[1].v, p[2].v, p[3].v, c);
      *depth = norm3(c);
      if (*depth < MINVAL) copy3(dir, d); else scl3(dir, c, 1.0 / *depth);
      /* position: barycentric combination of the portal (libccd findPos) */
      double b[4], t[3], sum;
      cross3(t, p[1].v, p[2].v); b[0] = dot3(t, p[3].v);
      cross3(t, p[3].v, p[2].v); b[1] = dot3(t, p[0].v);
      cross3(t, p[0].v, p[1].v); b[2] = dot3(t, p[3].v);
      cross3(t, p[2].v, p[1].v); b[3] = dot3(t, p[0].v);
      sum = b[0] + b[1] + b[2] + b[3];
      if (sum <= 0) {
        b[0] = 0;
        cross3(t, p[2].v, p[3].v); b[1] = dot3(t, d);
        cross3(t, p[3].v, p[1].v); b[2] = dot3(t, d);
        cross3(t, p[1].v, p[2].v); b[3] = dot3(t, d);
        sum = b[1] + b[2] + b[3];
      }
      pos[0] = pos[1] = pos[2] = 0;
      for (int k = 0; k < 4; k++) { addscl3(pos, p[k].v1, b[k]); addscl3(pos, p[k].v2, b[k]); }
      scl3(pos, pos, 0.5 / sum);
      return 1;
    }
    expand_portal(p, &v4);
  }
}

/* face of polytope geom g whose outward normal is most aligned with world direction n */
static int best_face(const OrcSim *s, int g, const double *n, double *align) {
  const MgsModelDesc *m = &s->m;
  int h = m->cgeom_hullid[g], best = 0;
  double nl[3], bd = -1e300;
  mulmatTvec3(nl, s->gxmat + 9 * g, n);
  if (m->cgeom_type[g] == GEOM_CYLINDER) { /* two faces: the caps (0: +z, 1: -z) */
    *align = fabs(nl[2]);
    return nl[2] >= 0 ? 0 : 1;
  }
  const double *FN = m->hull_facenormal + 3 * m->hull_faceadr[h];
  for (int f = 0; f < m->hull_facenum[h]; f++) {
    double t = dot3(FN + 3 * f, nl);
    if (t > bd) { bd = t; best = f; }
  }
  *align = bd;
  return best;
}
/* cos, sin of k * 45 deg: a cylinder cap enters the clipping as the octagon inscribed in its rim */
static const double OCT[8][2] = {{1, 0}, {0.70710678118654752, 0.70710678118654752}, {0, 1}, {-0.70710678118654752, 0.70710678118654752},
                                 {-1, 0}, {-0.70710678118654752, -0.70710678118654752}, {0, -1}, {0.70710678118654752, -0.70710678118654752}};
static int face_polygon(const OrcSim *s, int g, int f, double poly[][3], double *nw) {
  const MgsModelDesc *m = &s->m;
  if (m->cgeom_type[g] == GEOM_CYLINDER) {
    const double *sz = m->cgeom_size + 3 * g, z = f == 0 ? sz[1] : -sz[1];
    for (int i = 0; i < 8; i++) { /* counter-clockwise seen from outside, like the hull faces */
      const double *cs = OCT[f == 0 ? i : 7 - i];
      double v[3] = {sz[0] * cs[0], sz[0] * cs[1], z};
      mulmatvec3(poly[i], s->gxmat + 9 * g, v);
      add3(poly[i], poly[i], s->gxpos + 3 * g);
    }
    double nl[3] = {0, 0, f == 0 ? 1.0 : -1.0};
    mulmatvec3(nw, s->gxmat + 9 * g, nl);
    return 8;
  }
  int h = m->cgeom_hullid[g], gf = m->hull_faceadr[h] + f;
  int n = m->hull_facevertnum[gf];
  const double *V = m->hull_vert + 3 * m->hull_vertadr[h];
  for (int i = 0; i < n; i++) {
    mulmatvec3(poly[i], s->gxmat + 9 * g, V + 3 * m->hull_facevert[m->hull_facevertadr[gf] + i]);
    add3(poly[i], poly[i], s->gxpos + 3 * g);
  }
  mulmatvec3(nw, s->gxmat + 9 * g, m->hull_facenormal + 3 * gf);
  return n;
}

static void make_frame(double *frame) { /* frame[0:3] = normal given; fill tangents (mju_makeFrame) */
  double *x = frame, *y = frame + 3, *z = frame + 6;
  if (fabs(x[1]) < 0.5) { y[0] = 0; y[1] = 1; y[2] = 0; } else { y[0] = 0; y[1] = 0; y[2] = 1; }
  double d = dot3(x, y);
  addscl3(y, x, -d);
  normalize3(y);
  cross3(z, x, y);
}

static void add_contact(OrcSim *s, int pair, const double *pos, const double *normal, double dist) {
  const MgsModelDesc *m = &s->m;
  if (s->ncon >= NCON_MAX) { s->ncon_overflow++; return; }
  Contact *c = &s->con[s->ncon++];
  int g1 = m->pair_geom1[pair], g2 = m->pair_geom2[pair];
  copy3(c->pos, pos); copy3(c->frame, normal); make_frame(c->frame);
  c->dist = dist; c->pair = pair;
  c->geom1 = m->cgeom_geomid[g1]; c->geom2 = m->cgeom_geomid[g2];
  c->body1 = m->cgeom_bodyid[g1]; c->body2 = m->cgeom_bodyid[g2];
  c->dim = m->pair_condim[pair];
  memcpy(c->friction, m->pair_friction + 5 * pair, 5 * sizeof(double));
  memcpy(c->solref, m->pair_solref + 2 * pair, 2 * sizeof(double));
  memcpy(c->solimp, m->pair_solimp + 5 * pair, 5 * sizeof(double));
  c->mu = c->friction[0]; c->efc = -1;
}

/* ---- analytic primitive pairs.  MuJoCo's collision table (engine_collision_driver.c, mjCOLLISIONFUNC) sends sphere-sphere,
 * sphere-capsule, sphere-cylinder, sphere-box, capsule-capsule and capsule-box to closed-form routines of
 * engine_collision_primitive.c / engine_collision_box.c and everything else that is convex to the ccd path.  The routines below
 * restate the published closed forms: every case reduces to the sphere-sphere (or plane-sphere) primitive at the closest feature;
 * capsule-capsule gives two points only for parallel axes, capsule-box up to two (the deepest point of the axis segment plus the
 * segment end that also touches: the reference implementation's choice of the second point is not reproducible without its source,
 * the COUNT is).  The Allegro (4 capsules) and Shadow (3 spheres, 10 capsules, 7 cylinders) hands are the users, mostly in
 * self-collision and against the ground / table box; their contacts with mesh objects stay on the convex path like in MuJoCo.
 * Normals point from the pair's first geom to its second; dist < 0 is penetration; pos is the midpoint of the overlap. */
typedef struct { int n; double pos[2][3], normal[2][3], dist[2]; } PrimCon;

static int raw_sphere_sphere(const double *p1, double r1, const double *p2, double r2, double margin, double *pos, double *normal, double *dist) {
  double dif[3];
  sub3(dif, p2, p1);
  double cd = norm3(dif);
  if (cd > margin + r1 + r2) return 0;
  *dist = cd - r1 - r2;
  if (cd < MINVAL) { normal[0] = 1; normal[1] = 0; normal[2] = 0; } else scl3(normal, dif, 1.0 / cd);
  copy3(pos, p1);
  addscl3(pos, normal, r1 + 0.5 * *dist);
  return 1;
}
static void geom_axis(const OrcSim *s, int g, double *axis) { const double *R = s->gxmat + 9 * g; axis[0] = R[2]; axis[1] = R[5]; axis[2] = R[8]; }
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* sphere centre ps (world), radius rs, against box geom gb; normal from the sphere to the box */
static int raw_sphere_box(const OrcSim *s, const double *ps, double rs, int gb, double margin, double *pos, double *normal, double *dist) {
  const double *R = s->gxmat + 9 * gb, *sz = s->m.cgeom_size + 3 * gb;
  double t[3], c[3], cl[3], d[3], nl[3];
  sub3(t, ps, s->gxpos + 3 * gb);
  mulmatTvec3(c, R, t);
  for (int k = 0; k < 3; k++) cl[k] = clampd(c[k], -sz[k], sz[k]);
  sub3(d, cl, c);
  double dn = norm3(d);
  if (dn > MINVAL) { /* centre outside the box: closest point of the box */
    if (dn > rs + margin) return 0;
    scl3(nl, d, 1.0 / dn);
    *dist = dn - rs;
  } else { /* centre inside: leave through the nearest face */
    int k = 0;
    double fd = sz[0] - fabs(c[0]);
    for (int i = 1; i < 3; i++) if (sz[i] - fabs(c[i]) < fd) { fd = sz[i] - fabs(c[i]); k = i; }
    nl[0] = nl[1] = nl[2] = 0;
    nl[k] = c[k] >= 0 ? -1.0 : 1.0;
    *dist = -(rs + fd);
  }
  double pl[3];
  copy3(pl, c);
  addscl3(pl, nl, rs + 0.5 * *dist);
  mulmatvec3(pos, R, pl);
  add3(pos, pos, s->gxpos + 3 * gb);
  mulmatvec3(normal, R, nl);
  return 1;
}

/* derivative of the squared distance between the box and the point c + t * a (box frame) */
static double seg_box_dfdt(const double *c, const double *a, const double *sz, double t) {
  double g = 0;
  for (int k = 0; k < 3; k++) {
    double p = c[k] + t * a[k], ex = fabs(p) - sz[k];
    if (ex > 0) g += 2 * (p > 0 ? ex : -ex) * a[k];
  }
  return g;
}

static void prim_push(PrimCon *o, const double *pos, const double *normal, double dist, double sign) {
  if (o->n >= 2) return;
  copy3(o->pos[o->n], pos);
  scl3(o->normal[o->n], normal, sign);
  o->dist[o->n] = dist;
  o->n++;
}

/* returns 1 when the pair's types have a closed-form routine (contacts, possibly none, in *o) */
static int prim_pair(const OrcSim *s, int pair, PrimCon *o) {
  const MgsModelDesc *m = &s->m;
  int ga = m->pair_geom1[pair], gb = m->pair_geom2[pair];
  int ta = m->cgeom_type[ga], tb = m->cgeom_type[gb];
  double sign = 1.0; /* the routines are written for type(ga) <= type(gb) */
  if (ta > tb) { int t = ga; ga = gb; gb = t; t = ta; ta = tb; tb = t; sign = -1.0; }
  const double margin = m->pair_margin[pair];
  const double *pa = s->gxpos + 3 * ga, *pb = s->gxpos + 3 * gb, *sa = m->cgeom_size + 3 * ga, *sb = m->cgeom_size + 3 * gb;
  double pos[3], nrm[3], dist;
  o->n = 0;
  if (ta == GEOM_SPHERE && tb == GEOM_SPHERE) {
    if (raw_sphere_sphere(pa, sa[0], pb, sb[0], margin, pos, nrm, &dist)) prim_push(o, pos, nrm, dist, sign);
    return 1;
  }
  if (ta == GEOM_SPHERE && tb == GEOM_CAPSULE) { /* closest point of the capsule's axis segment */
    double ax[3], t[3], q[3];
    geom_axis(s, gb, ax);
    sub3(t, pa, pb);
    double x = clampd(dot3(ax, t), -sb[1], sb[1]);
    copy3(q, pb);
    addscl3(q, ax, x);
    if (raw_sphere_sphere(pa, sa[0], q, sb[0], margin, pos, nrm, &dist)) prim_push(o, pos, nrm, dist, sign);
    return 1;
  }
  if (ta == GEOM_SPHERE && tb == GEOM_CYLINDER) { /* side, cap or rim of the cylinder */
    double ax[3], v[3], pr[3], q[3];
    geom_axis(s, gb, ax);
    sub3(v, pa, pb);
    double x = dot3(v, ax), R = sb[0], h = sb[1];
    copy3(pr, v);
    addscl3(pr, ax, -x);
    double rho = norm3(pr);
    int side = fabs(x) < h, cap = rho < R;
    if (side && cap) { if (h - fabs(x) < R - rho) side = 0; else cap = 0; } /* centre inside: nearest surface */
    if (side) {
      copy3(q, pb);
      addscl3(q, ax, x);
      if (raw_sphere_sphere(pa, sa[0], q, R, margin, pos, nrm, &dist)) prim_push(o, pos, nrm, dist, sign);
    } else if (cap) { /* plane of the nearer cap; normal from the sphere into the cylinder */
      double sg = x >= 0 ? 1.0 : -1.0;
      dist = fabs(x) - h - sa[0];
      if (dist <= margin) {
        scl3(nrm, ax, -sg);
        copy3(pos, pa);
        addscl3(pos, nrm, sa[0] + 0.5 * dist);
        prim_push(o, pos, nrm, dist, sign);
      }
    } else { /* rim circle */
      copy3(q, pb);
      addscl3(q, ax, x >= 0 ? h : -h);
      if (rho > MINVAL) addscl3(q, pr, R / rho);
      if (raw_sphere_sphere(pa, sa[0], q, 0.0, margin, pos, nrm, &dist)) prim_push(o, pos, nrm, dist, sign);
    }
    return 1;
  }
  if (ta == GEOM_SPHERE && tb == GEOM_BOX) {
    if (raw_sphere_box(s, pa, sa[0], gb, margin, pos, nrm, &dist)) prim_push(o, pos, nrm, dist, sign);
    return 1;
  }
  if (ta == GEOM_CAPSULE && tb == GEOM_CAPSULE) {
    double a1[3], a2[3], dif[3], q1[3], q2[3];
    geom_axis(s, ga, a1);
    geom_axis(s, gb, a2);
    sub3(dif, pa, pb);
    const double l1 = sa[1], l2 = sb[1];
    double mb = -dot3(a1, a2), u = -dot3(a1, dif), v = dot3(a2, dif), det = 1.0 - mb * mb;
    if (fabs(det) >= MINVAL) { /* general position: closest points of the two segments */
      double x1 = (u - mb * v) / det, x2 = (v - mb * u) / det;
      if (x1 > l1) { x1 = l1; x2 = v - mb * x1; } else if (x1 < -l1) { x1 = -l1; x2 = v - mb * x1; }
      if (x2 > l2) { x2 = l2; x1 = clampd(u - mb * x2, -l1, l1); } else if (x2 < -l2) { x2 = -l2; x1 = clampd(u - mb * x2, -l1, l1); }
      copy3(q1, pa); addscl3(q1, a1, x1);
      copy3(q2, pb); addscl3(q2, a2, x2);
      if (raw_sphere_sphere(q1, sa[0], q2, sb[0], margin, pos, nrm, &dist)) prim_push(o, pos, nrm, dist, sign);
    } else { /* parallel axes: the ends of capsule 1 against segment 2, then the ends of capsule 2 against segment 1, two points at most */
      for (int e = 0; e < 2 && o->n < 2; e++) {
        double x1 = e ? -l1 : l1, x2 = clampd(v - mb * x1, -l2, l2);
        copy3(q1, pa); addscl3(q1, a1, x1);
        copy3(q2, pb); addscl3(q2, a2, x2);
        if (raw_sphere_sphere(q1, sa[0], q2, sb[0], margin, pos, nrm, &dist)) prim_push(o, pos, nrm, dist, sign);
      }
      for (int e = 0; e < 2 && o->n < 2; e++) {
        double x2 = e ? -l2 : l2, x1 = clampd(u - mb * x2, -l1, l1);
        if (fabs(fabs(x1) - l1) < MINVAL) continue; /* that end of capsule 1 was tested above */
        copy3(q1, pa); addscl3(q1, a1, x1);
        copy3(q2, pb); addscl3(q2, a2, x2);
        if (raw_sphere_sphere(q1, sa[0], q2, sb[0], margin, pos, nrm, &dist)) prim_push(o, pos, nrm, dist, sign);
      }
    }
    return 1;
  }
  if (ta == GEOM_CAPSULE && tb == GEOM_BOX) {
    /* axis segment p(t) = c + t a, t in [-1, 1], in the box frame.  f(t) = squared distance to the box is convex and its derivative
     * piecewise linear with kinks where a coordinate crosses a face plane: the minimiser is bracketed by the two kinks (or segment
     * ends) around the sign change of f' and found by linear interpolation between them. */
    const double *R = s->gxmat + 9 * gb;
    double ax[3], t3[3], c[3], a[3], T[8];
    geom_axis(s, ga, ax);
    sub3(t3, pa, pb);
    mulmatTvec3(c, R, t3);
    mulmatTvec3(a, R, ax);
    scl3(a, a, sa[1]);
    int nT = 0;
    T[nT++] = -1; T[nT++] = 1;
    for (int k = 0; k < 3; k++)
      if (fabs(a[k]) > MINVAL)
        for (int sg = -1; sg <= 1; sg += 2) { double t = (sg * sb[k] - c[k]) / a[k]; if (t > -1 && t < 1) T[nT++] = t; }
    double tl = -2, tr = 2, gl = 0, gr = 0;
    for (int i = 0; i < nT; i++) {
      double g = seg_box_dfdt(c, a, sb, T[i]);
      if (g <= 0 && T[i] > tl) { tl = T[i]; gl = g; }
      if (g >= 0 && T[i] < tr) { tr = T[i]; gr = g; }
    }
    double ts;
    if (tl < -1.5) ts = -1; else if (tr > 1.5) ts = 1;
    else if (gr - gl > MINVAL) ts = tl + (tr - tl) * (-gl) / (gr - gl);
    else ts = 0.5 * (tl + tr);
    /* candidates: the two ends, then the closest point; first contact = deepest (earlier candidate on ties), second = the deepest
     * other candidate in contact that is not the same point of the segment */
    double ct[3] = {-1, 1, ts}, cpos[3][3], cn[3][3], cd[3];
    int hit[3];
    for (int i = 0; i < 3; i++) {
      double q[3];
      copy3(q, pa);
      addscl3(q, ax, sa[1] * ct[i]);
      hit[i] = raw_sphere_box(s, q, sa[0], gb, margin, cpos[i], cn[i], &cd[i]);
    }
    int i1 = -1, i2 = -1;
    for (int i = 0; i < 3; i++) if (hit[i] && (i1 < 0 || cd[i] < cd[i1])) i1 = i;
    if (i1 >= 0) {
      for (int i = 0; i < 3; i++) if (i != i1 && hit[i] && fabs(ct[i] - ct[i1]) > 0.05 && (i2 < 0 || cd[i] < cd[i2])) i2 = i;
      prim_push(o, cpos[i1], cn[i1], cd[i1], sign);
      if (i2 >= 0) prim_push(o, cpos[i2], cn[i2], cd[i2], sign);
    }
    return 1;
  }
  return 0;
}

#define FACE_ALIGN_MIN 0.9990 /* below this the contact is edge/vertex-like: single MPR point */

static int has_face(const OrcSim *s, int g, const double *n) {
  int t = s->m.cgeom_type[g];
  if (is_polytope(t)) return 1;
#ifndef ORC_NO_CYLINDER_CAPS
  if (t == GEOM_CYLINDER && !s->no_analytic) {
    const double *R = s->gxmat + 9 * g;
    return fabs(R[2] * n[0] + R[5] * n[1] + R[8] * n[2]) >= FACE_ALIGN_MIN;
  }
#endif
  return 0;
}

static void collide_pair(OrcSim *s, int pair) {
  const MgsModelDesc *m = &s->m;
  int g1 = m->pair_geom1[pair], g2 = m->pair_geom2[pair];
  double dc[3], depth, n[3], pos[3];
  sub3(dc, s->gxpos + 3 * g1, s->gxpos + 3 * g2);
  double rr = m->cgeom_rbound[g1] + m->cgeom_rbound[g2] + m->pair_margin[pair];
  if (dot3(dc, dc) > rr * rr) return;
#ifndef ORC_NO_ANALYTIC_PRIMS
  if (!s->no_analytic) {
    PrimCon pc;
    if (prim_pair(s, pair, &pc)) {
      for (int k = 0; k < pc.n; k++) add_contact(s, pair, pc.pos[k], pc.normal[k], pc.dist[k]);
      return;
    }
  }
#endif
  if (!mpr_penetration(s, g1, g2, &depth, n, pos)) return;
  if (!(depth > 0)) return;
  /* multi-point manifold: both geoms must offer a flat face - polytopes always do, a cylinder when the contact normal is along its
   * axis (a cap).  MuJoCo gets the extra points of flat ccd contacts from the multiccd perturbation (enabled by the reference's scene
   * options); here the two most-aligned faces are clipped against each other (cap against cap is left a single point). */
  if (has_face(s, g1, n) && has_face(s, g2, n) && (is_polytope(m->cgeom_type[g1]) || is_polytope(m->cgeom_type[g2]))) {
    double a1, a2, nn[3] = {-n[0], -n[1], -n[2]};
    int f1 = best_face(s, g1, n, &a1), f2 = best_face(s, g2, nn, &a2);
    if (fmax(a1, a2) >= FACE_ALIGN_MIN) {
      int refg = a1 >= a2 ? g1 : g2, incg = a1 >= a2 ? g2 : g1;
      int reff = a1 >= a2 ? f1 : f2;
      double ref[MAXPOLY][3], nref[3], inc[MAXPOLY][3], ninc[3], A[MAXCLIP][3], B[MAXCLIP][3];
      int nr = face_polygon(s, refg, reff, ref, nref);
      /* incident face: most anti-parallel to the reference normal */
      double mn[3] = {-nref[0], -nref[1], -nref[2]}, al;
      int incf = best_face(s, incg, mn, &al);
      int na = face_polygon(s, incg, incf, inc, ninc);
      for (int i = 0; i < na; i++) copy3(A[i], inc[i]);
      for (int e = 0; e < nr && na > 0; e++) {
        double edge[3], sn[3];
        sub3(edge, ref[(e + 1) % nr], ref[e]);
        cross3(sn, edge, nref); /* outward side normal */
        int nb2 = 0;
        for (int i = 0; i < na; i++) {
          const double *P = A[i], *Q = A[(i + 1) % na];
          double t0[3], t1[3];
          sub3(t0, P, ref[e]); sub3(t1, Q, ref[e]);
          double dp = dot3(t0, sn), dq = dot3(t1, sn);
          if (dp <= 0 && nb2 < MAXCLIP) copy3(B[nb2++], P);
          if ((dp <= 0) != (dq <= 0) && nb2 < MAXCLIP) {
            double t = dp / (dp - dq);
            for (int k = 0; k < 3; k++) B[nb2][k] = P[k] + t * (Q[k] - P[k]);
            nb2++;
          }
        }
        na = nb2;
        for (int i = 0; i < na; i++) copy3(A[i], B[i]);
      }
      /* keep penetrating points */
      double dist[MAXCLIP];
      int np = 0;
      for (int i = 0; i < na; i++) {
        double t[3];
        sub3(t, A[i], ref[0]);
        double dd = dot3(t, nref);
        if (dd < 0) { copy3(A[np], A[i]); dist[np] = dd; np++; }
      }
      if (np > 0) {
        int sel[4], ns = 0;
        /* reduce to <=4: deepest, farthest from it, farthest from that line on either side */
        int i0 = 0;
        for (int i = 1; i < np; i++) if (dist[i] < dist[i0]) i0 = i;
        sel[ns++] = i0;
        if (np > 1) {
          int i1 = -1; double bd = -1;
          for (int i = 0; i < np; i++) { double t[3]; sub3(t, A[i], A[i0]); double d2 = dot3(t, t); if (i != i0 && d2 > bd) { bd = d2; i1 = i; } }
          if (i1 >= 0 && bd > 1e-12) {
            sel[ns++] = i1;
            double e01[3]; sub3(e01, A[i1], A[i0]);
            int i2 = -1, i3 = -1; double mx = 1e-12, mnv = -1e-12;
            for (int i = 0; i < np; i++) {
              if (i == i0 || i == i1) continue;
              double t[3], c[3]; sub3(t, A[i], A[i0]); cross3(c, e01, t);
              double sa = dot3(c, nref);
              if (sa > mx) { mx = sa; i2 = i; }
              if (sa < mnv) { mnv = sa; i3 = i; }
            }
            if (i2 >= 0) sel[ns++] = i2;
            if (i3 >= 0) sel[ns++] = i3;
          }
        }
        double nout[3];
        if (refg == g1) copy3(nout, nref); else scl3(nout, nref, -1);
        for (int k = 0; k < ns; k++) {
          double cp[3];
          copy3(cp, A[sel[k]]);
          addscl3(cp, nref, -0.5 * dist[sel[k]]);
          add_contact(s, pair, cp, nout, dist[sel[k]]);
        }
        return;
      }
    }
  }
  add_contact(s, pair, pos, n, -depth);
}

static void collision(OrcSim *s) {
  s->ncon = 0;
  for (int p = 0; p < s->m.npair; p++) collide_pair(s, p);
}

/* ------------------------------------------------------------------ constraints */
/* solimp -> impedance at violation `pos` (mj getimpedance) */
static double impedance(const double *solimp, double pos, double margin) {
  double dmin = fmin(0.9999, fmax(0.0001, solimp[0])), dmax = fmin(0.9999, fmax(0.0001, solimp[1]));
  double width = fmax(MINVAL, solimp[2]), mid = fmin(0.9999, fmax(0.0001, solimp[3])), power = fmax(1.0, solimp[4]);
  if (dmin == dmax || width <= MINVAL) return 0.5 * (dmin + dmax);
  double x = fabs(pos - margin) / width, y;
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}

static int add_row(OrcSim *s, int type, int id, double pos, double margin, double floss, double diagApprox) {
  int i = s->nefc;
  if (i >= NEFC_MAX) return -1;
  memset(s->J + (size_t)i * s->m.nv, 0, sizeof(double) * s->m.nv);
  s->efc_type[i] = type; s->efc_id[i] = id; s->efc_pos[i] = pos; s->efc_margin[i] = margin;
  s->efc_floss[i] = floss; s->efc_diagApprox[i] = diagApprox;
  s->nefc++;
  return i;
}

/* mj_makeConstraint: rows in MuJoCo's order - equality, dof friction, limits, contacts */
static void make_constraint(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  int nv = m->nv;
  double *jp1 = s->jtmp, *jr1 = s->jtmp + 3 * nv, *jp2 = s->jtmp + 6 * nv, *jr2 = s->jtmp + 9 * nv;
  s->nefc = 0;
  /* equality */
  for (int e = 0; e < m->neq; e++) {
    if (!m->eq_active[e]) continue;
    const double *data = m->eq_data + 11 * e;
    if (m->eq_type[e] == EQ_JOINT) {
      int j1 = m->eq_obj1id[e], j2 = m->eq_obj2id[e];
      double pos, deriv = 0, dif = 0;
      double q1 = s->qpos[m->jnt_qposadr[j1]] - m->qpos0[m->jnt_qposadr[j1]];
      if (j2 >= 0) {
        dif = s->qpos[m->jnt_qposadr[j2]] - m->qpos0[m->jnt_qposadr[j2]];
        double poly = data[0] + dif * (data[1] + dif * (data[2] + dif * (data[3] + dif * data[4])));
        deriv = data[1] + dif * (2 * data[2] + dif * (3 * data[3] + dif * 4 * data[4]));
        pos = q1 - poly;
      } else pos = q1 - data[0];
      double da = m->dof_invweight0[m->jnt_dofadr[j1]] + (j2 >= 0 ? m->dof_invweight0[m->jnt_dofadr[j2]] : 0);
      int r = add_row(s, CT_EQUALITY, e, pos, 0, 0, da);
      if (r < 0) continue;
      s->J[r * nv + m->jnt_dofadr[j1]] = 1;
      if (j2 >= 0) s->J[r * nv + m->jnt_dofadr[j2]] = -deriv;
      continue;
    }
    int b1 = m->eq_obj1id[e], b2 = m->eq_obj2id[e];
    double p1[3], p2[3], cpos[6];
    const double *a1 = m->eq_type[e] == EQ_WELD ? data + 3 : data, *a2 = m->eq_type[e] == EQ_WELD ? data : data + 3;
    mulmatvec3(p1, s->xmat + 9 * b1, a1); add3(p1, p1, s->xpos + 3 * b1);
    mulmatvec3(p2, s->xmat + 9 * b2, a2); add3(p2, p2, s->xpos + 3 * b2);
    sub3(cpos, p1, p2);
    jac_point(s, b1, p1, jp1, jr1);
    jac_point(s, b2, p2, jp2, jr2);
    double tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    double rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    for (int k = 0; k < 3; k++) {
      int r = add_row(s, CT_EQUALITY, e, cpos[k], 0, 0, tran);
      if (r < 0) continue;
      for (int d = 0; d < nv; d++) s->J[r * nv + d] = jp1[k * nv + d] - jp2[k * nv + d];
    }
    if (m->eq_type[e] == EQ_WELD) {
      double ts = data[10], quat[4], quat1[4], quat2[4], quat3[4];
      mulquat(quat, s->xquat + 4 * b1, data + 6);
      quat1[0] = s->xquat[4 * b2]; quat1[1] = -s->xquat[4 * b2 + 1]; quat1[2] = -s->xquat[4 * b2 + 2]; quat1[3] = -s->xquat[4 * b2 + 3];
      mulquat(quat2, quat1, quat);
      int r0 = s->nefc;
      for (int k = 0; k < 3; k++) add_row(s, CT_EQUALITY, e, ts * quat2[1 + k], 0, 0, rot);
      if (s->nefc != r0 + 3) continue;
      for (int d = 0; d < nv; d++) {
        double ax[4] = {0, jr1[d] - jr2[d], jr1[nv + d] - jr2[nv + d], jr1[2 * nv + d] - jr2[2 * nv + d]};
        mulquat(quat2, quat1, ax);
        mulquat(quat3, quat2, quat);
        for (int k = 0; k < 3; k++) s->J[(r0 + k) * nv + d] = 0.5 * ts * quat3[1 + k];
      }
    }
  }
  s->ne = s->nefc;
  /* dof friction loss */
  for (int d = 0; d < nv; d++) {
    if (m->dof_frictionloss[d] <= 0) continue;
    int r = add_row(s, CT_FRICTION_DOF, d, 0, 0, m->dof_frictionloss[d], m->dof_invweight0[d]);
    if (r >= 0) s->J[r * nv + d] = 1;
  }
  s->nf = s->nefc - s->ne;
  /* joint limits */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] == JNT_FREE) continue;
    double q = s->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - q);
      if (dist < margin) {
        int r = add_row(s, CT_LIMIT, j, dist, margin, 0, m->dof_invweight0[m->jnt_dofadr[j]]);
        if (r >= 0) s->J[r * nv + m->jnt_dofadr[j]] = -side;
      }
    }
  }
  s->nl = s->nefc - s->ne - s->nf;
  /* contacts (elliptic cones: dim rows each) */
  for (int c = 0; c < s->ncon; c++) {
    Contact *con = &s->con[c];
    int dim = con->dim;
    if (s->nefc + dim > NEFC_MAX) { con->efc = -1; continue; }
    jac_point(s, con->body1, con->pos, jp1, jr1);
    jac_point(s, con->body2, con->pos, jp2, jr2);
    double tran = m->body_invweight0[2 * con->body1] + m->body_invweight0[2 * con->body2];
    double rot = m->body_invweight0[2 * con->body1 + 1] + m->body_invweight0[2 * con->body2 + 1];
    con->efc = s->nefc;
    for (int k = 0; k < dim; k++) {
      int r = add_row(s, CT_CONTACT, c, k == 0 ? con->dist : 0, 0, 0, k < 3 ? tran : rot);
      const double *ax = con->frame + 3 * (k < 3 ? k : k - 3);
      const double *A1 = k < 3 ? jp1 : jr1, *A2 = k < 3 ? jp2 : jr2;
      for (int d = 0; d < nv; d++)
        s->J[r * nv + d] = ax[0] * (A2[d] - A1[d]) + ax[1] * (A2[nv + d] - A1[nv + d]) + ax[2] * (A2[2 * nv + d] - A1[2 * nv + d]);
    }
  }
  /* impedance, regularisation (mj_makeImpedance) */
  for (int i = 0; i < s->nefc; i++) {
    const double *solref, *solimp;
    double pos = s->efc_pos[i], margin = s->efc_margin[i];
    int friction_row = 0;
    switch (s->efc_type[i]) {
      case CT_EQUALITY: solref = m->eq_solref + 2 * s->efc_id[i]; solimp = m->eq_solimp + 5 * s->efc_id[i]; break;
      case CT_FRICTION_DOF: solref = m->dof_solref + 2 * s->efc_id[i]; solimp = m->dof_solimp + 5 * s->efc_id[i]; friction_row = 1; break;
      case CT_LIMIT: solref = m->jnt_solref + 2 * s->efc_id[i]; solimp = m->jnt_solimp + 5 * s->efc_id[i]; break;
      default: {
        Contact *con = &s->con[s->efc_id[i]];
        solref = con->solref; solimp = con->solimp;
        if (i != con->efc) { friction_row = 1; pos = s->efc_pos[con->efc]; margin = s->efc_margin[con->efc]; }
      }
    }
    double imp = impedance(solimp, pos, margin);
    double dmax = fmin(0.9999, fmax(0.0001, solimp[1]));
    double k, b;
    if (solref[0] > 0) {
      double tc = fmax(solref[0], 2 * m->timestep), dr = solref[1]; /* refsafe */
      k = 1.0 / fmax(MINVAL, dmax * dmax * tc * tc * dr * dr);
      b = 2.0 / fmax(MINVAL, dmax * tc);
    } else { k = -solref[0] / fmax(MINVAL, dmax * dmax); b = -solref[1] / fmax(MINVAL, dmax); }
    if (friction_row) k = 0;
    s->efc_KBIP[i][0] = k; s->efc_KBIP[i][1] = b; s->efc_KBIP[i][2] = imp; s->efc_KBIP[i][3] = 0;
    s->efc_R[i] = fmax(MINVAL, (1 - imp) * s->efc_diagApprox[i] / imp);
  }
  if (m->cone_elliptic) {
    for (int c = 0; c < s->ncon; c++) {
      Contact *con = &s->con[c];
      int i = con->efc;
      if (i < 0 || con->dim < 3) continue;
      s->efc_R[i + 1] = s->efc_R[i] / fmax(MINVAL, m->impratio);
      con->mu = con->friction[0] * sqrt(s->efc_R[i + 1] / s->efc_R[i]);
      for (int j = 2; j < con->dim; j++)
        s->efc_R[i + j] = s->efc_R[i + 1] * con->friction[0] * con->friction[0] / fmax(MINVAL, con->friction[j - 1] * con->friction[j - 1]);
    }
  }
  for (int i = 0; i < s->nefc; i++) s->efc_D[i] = 1.0 / s->efc_R[i];
}

/* mj_referenceConstraint: aref = -b*vel - k*imp*(pos-margin) */
static void reference_constraint(OrcSim *s) {
  int nv = s->m.nv;
  for (int i = 0; i < s->nefc; i++) {
    double v = 0;
    for (int d = 0; d < nv; d++) v += s->J[i * nv + d] * s->qvel[d];
    s->efc_vel[i] = v;
    double pm = s->efc_pos[i] - s->efc_margin[i];
    if (s->efc_type[i] == CT_CONTACT && i != s->con[s->efc_id[i]].efc) pm = 0;
    s->efc_aref[i] = -s->efc_KBIP[i][1] * v - s->efc_KBIP[i][0] * s->efc_KBIP[i][2] * pm;
  }
}

/* ------------------------------------------------------------------ solver */
enum { ST_SATISFIED = 0, ST_QUADRATIC = 1, ST_LINEARNEG = 2, ST_LINEARPOS = 3, ST_CONE = 4 };

/* constraint cost, forces and states at jar (mj_constraintUpdate); cone Hessians into hcone[ncon][36] */
static double constraint_update(OrcSim *s, const double *jar, double *force, double *hcone) {
  double cost = 0;
  for (int i = 0; i < s->nefc; i++) {
    double D = s->efc_D[i], R = s->efc_R[i];
    switch (s->efc_type[i]) {
      case CT_EQUALITY:
        force[i] = -D * jar[i]; cost += 0.5 * D * jar[i] * jar[i]; s->efc_state[i] = ST_QUADRATIC; break;
      case CT_FRICTION_DOF: {
        double f = s->efc_floss[i];
        if (jar[i] <= -R * f) { force[i] = f; cost += -0.5 * R * f * f - f * jar[i]; s->efc_state[i] = ST_LINEARNEG; }
        else if (jar[i] >= R * f) { force[i] = -f; cost += -0.5 * R * f * f + f * jar[i]; s->efc_state[i] = ST_LINEARPOS; }
        else { force[i] = -D * jar[i]; cost += 0.5 * D * jar[i] * jar[i]; s->efc_state[i] = ST_QUADRATIC; }
        break;
      }
      case CT_LIMIT:
        if (jar[i] < 0) { force[i] = -D * jar[i]; cost += 0.5 * D * jar[i] * jar[i]; s->efc_state[i] = ST_QUADRATIC; }
        else { force[i] = 0; s->efc_state[i] = ST_SATISFIED; }
        break;
      default: {
        Contact *con = &s->con[s->efc_id[i]];
        int dim = con->dim;
        if (dim < 3 || !s->m.cone_elliptic) { /* frictionless */
          if (jar[i] < 0) { force[i] = -D * jar[i]; cost += 0.5 * D * jar[i] * jar[i]; s->efc_state[i] = ST_QUADRATIC; }
          else { force[i] = 0; s->efc_state[i] = ST_SATISFIED; }
          break;
        }
        double mu = con->mu, U[6], N, T = 0;
        U[0] = jar[i] * mu;
        for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * con->friction[j - 1]; T += U[j] * U[j]; }
        N = U[0]; T = sqrt(T);
        int st;
        if (N >= mu * T || (T <= 0 && N >= 0)) {
          for (int j = 0; j < dim; j++) force[i + j] = 0;
          st = ST_SATISFIED;
        } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
          for (int j = 0; j < dim; j++) { force[i + j] = -s->efc_D[i + j] * jar[i + j]; cost += 0.5 * s->efc_D[i + j] * jar[i + j] * jar[i + j]; }
          st = ST_QUADRATIC;
        } else {
          double Dm = s->efc_D[i] / fmax(MINVAL, mu * mu * (1 + mu * mu)), NmT = N - mu * T;
          cost += 0.5 * Dm * NmT * NmT;
          force[i] = -Dm * NmT * mu;
          for (int j = 1; j < dim; j++) force[i + j] = -force[i] / T * U[j] * con->friction[j - 1];
          st = ST_CONE;
          if (hcone) {
            double *h = hcone + 36 * s->efc_id[i], scl[6];
            scl[0] = mu;
            for (int j = 1; j < dim; j++) scl[j] = con->friction[j - 1];
            h[0] = 1;
            for (int j = 1; j < dim; j++) h[j] = h[j * dim] = -mu * U[j] / T;
            for (int j = 1; j < dim; j++)
              for (int k = 1; k < dim; k++) h[j * dim + k] = mu * N / (T * T * T) * U[j] * U[k] + (j == k ? mu * mu - mu * N / T : 0);
            for (int j = 0; j < dim; j++)
              for (int k = 0; k < dim; k++) h[j * dim + k] *= Dm * scl[j] * scl[k];
          }
        }
        for (int j = 0; j < dim; j++) s->efc_state[i + j] = st;
        i += dim - 1;
      }
    }
  }
  return cost;
}

/* first and second derivative of the total cost along qacc + alpha*search */
static void ls_eval(const OrcSim *s, const double *jar, const double *jv, double alpha, double g1, double g2, double *d1, double *d2) {
  double a = g1 + alpha * g2, h = g2;
  for (int i = 0; i < s->nefc; i++) {
    double D = s->efc_D[i], R = s->efc_R[i], x = jar[i] + alpha * jv[i];
    switch (s->efc_type[i]) {
      case CT_EQUALITY: a += D * x * jv[i]; h += D * jv[i] * jv[i]; break;
      case CT_FRICTION_DOF: {
        double f = s->efc_floss[i];
        if (x <= -R * f) a += -f * jv[i];
        else if (x >= R * f) a += f * jv[i];
        else { a += D * x * jv[i]; h += D * jv[i] * jv[i]; }
        break;
      }
      case CT_LIMIT: if (x < 0) { a += D * x * jv[i]; h += D * jv[i] * jv[i]; } break;
      default: {
        const Contact *con = &s->con[s->efc_id[i]];
        int dim = con->dim;
        if (dim < 3 || !s->m.cone_elliptic) { if (x < 0) { a += D * x * jv[i]; h += D * jv[i] * jv[i]; } break; }
        double mu = con->mu, U[6], V[6], N, T = 0, UV = 0, VV = 0;
        U[0] = x * mu; V[0] = jv[i] * mu;
        for (int j = 1; j < dim; j++) {
          U[j] = (jar[i + j] + alpha * jv[i + j]) * con->friction[j - 1];
          V[j] = jv[i + j] * con->friction[j - 1];
          T += U[j] * U[j]; UV += U[j] * V[j]; VV += V[j] * V[j];
        }
        N = U[0]; T = sqrt(T);
        if (N >= mu * T || (T <= 0 && N >= 0)) {
        } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
          for (int j = 0; j < dim; j++) {
            double xj = jar[i + j] + alpha * jv[i + j];
            a += s->efc_D[i + j] * xj * jv[i + j]; h += s->efc_D[i + j] * jv[i + j] * jv[i + j];
          }
        } else {
          double Dm = s->efc_D[i] / fmax(MINVAL, mu * mu * (1 + mu * mu)), NmT = N - mu * T;
          double T1 = UV / T, T2 = VV / T - UV * UV / (T * T * T), N1 = V[0];
          a += Dm * NmT * (N1 - mu * T1);
          h += Dm * ((N1 - mu * T1) * (N1 - mu * T1) - NmT * mu * T2);
        }
        i += dim - 1;
      }
    }
  }
  *d1 = a; *d2 = h;
}

static void mul_M(const OrcSim *s, double *res, const double *v) {
  int nv = s->m.nv;
  for (int i = 0; i < nv; i++) {
    double t = 0;
    for (int j = 0; j < nv; j++) t += s->M[i * nv + j] * v[j];
    res[i] = t;
  }
}

/* H = M + J' D J (+ cone blocks) for the current states, Cholesky into s->H */
static void newton_hessian(OrcSim *s, const double *hcone) {
  int nv = s->m.nv;
  double *H = s->w1;
  memcpy(H, s->M, sizeof(double) * nv * nv);
  for (int i = 0; i < s->nefc; i++) {
    if (s->efc_type[i] == CT_CONTACT && s->efc_state[i] == ST_CONE) {
      const Contact *con = &s->con[s->efc_id[i]];
      int dim = con->dim;
      const double *h = hcone + 36 * s->efc_id[i];
      for (int j = 0; j < dim; j++)
        for (int k = 0; k < dim; k++) {
          double c = h[j * dim + k];
          if (c == 0) continue;
          const double *Jj = s->J + (i + j) * nv, *Jk = s->J + (i + k) * nv;
          for (int a = 0; a < nv; a++) {
            if (Jj[a] == 0) continue;
            for (int b = 0; b < nv; b++) H[a * nv + b] += c * Jj[a] * Jk[b];
          }
        }
      i += dim - 1;
      continue;
    }
    if (s->efc_state[i] != ST_QUADRATIC) continue;
    const double *Ji = s->J + i * nv;
    double D = s->efc_D[i];
    for (int a = 0; a < nv; a++) {
      if (Ji[a] == 0) continue;
      for (int b = 0; b < nv; b++) H[a * nv + b] += D * Ji[a] * Ji[b];
    }
  }
  chol_factor(s->H, H, nv);
}

static double total_cost(OrcSim *s, const double *qacc, double *jar, double *force, double *hcone, double *Ma) {
  int nv = s->m.nv;
  for (int i = 0; i < s->nefc; i++) {
    double t = -s->efc_aref[i];
    for (int d = 0; d < nv; d++) t += s->J[i * nv + d] * qacc[d];
    jar[i] = t;
  }
  double cost = constraint_update(s, jar, force, hcone);
  mul_M(s, Ma, qacc);
  double g = 0;
  for (int d = 0; d < nv; d++) g += (Ma[d] - s->qfrc_smooth[d]) * (qacc[d] - s->qacc_smooth[d]);
  return cost + 0.5 * g;
}

/* mj_solNewton (primal): exact Newton with exact line search */
static void solve_newton(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  int nv = m->nv, nefc = s->nefc;
  double *qacc = s->qacc, *Ma = s->w3, *grad = s->w4, *search = s->w5, *Mv = s->w6;
  static __thread double jar[NEFC_MAX], jv[NEFC_MAX], hcone[NCON_MAX * 36], ftmp[NEFC_MAX];
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  s->solver_niter = 0;
  /* warmstart */
  memcpy(qacc, s->qacc_warmstart, sizeof(double) * nv);
  double cw = total_cost(s, qacc, jar, ftmp, NULL, Ma);
  double cs = total_cost(s, s->qacc_smooth, jar, ftmp, NULL, Ma);
  if (!(cw < cs)) memcpy(qacc, s->qacc_smooth, sizeof(double) * nv);
  double cost = total_cost(s, qacc, jar, s->efc_force, hcone, Ma);
  for (int iter = 0; iter < m->iterations; iter++) {
    for (int d = 0; d < nv; d++) {
      double t = Ma[d] - s->qfrc_smooth[d];
      for (int i = 0; i < nefc; i++) t -= s->J[i * nv + d] * s->efc_force[i];
      grad[d] = t;
    }
    double gn = 0;
    for (int d = 0; d < nv; d++) gn += grad[d] * grad[d];
    if (iter > 0 && scale * sqrt(gn) < m->tolerance) break;
    newton_hessian(s, hcone);
    memcpy(search, grad, sizeof(double) * nv);
    chol_solve(s->H, search, nv);
    for (int d = 0; d < nv; d++) search[d] = -search[d];
    /* line search */
    mul_M(s, Mv, search);
    double g1 = 0, g2 = 0, sn = 0;
    for (int d = 0; d < nv; d++) { g1 += search[d] * (Ma[d] - s->qfrc_smooth[d]); g2 += search[d] * Mv[d]; sn += search[d] * search[d]; }
    for (int i = 0; i < nefc; i++) {
      double t = 0;
      for (int d = 0; d < nv; d++) t += s->J[i * nv + d] * search[d];
      jv[i] = t;
    }
    double gtol = m->tolerance * m->ls_tolerance * sqrt(sn) / scale;
    /* exact line search on the convex, piecewise-smooth 1-D cost: find the zero of its (monotone)
     * derivative.  Newton steps while they stay inside the bracket [lo, hi]; otherwise the secant of the
     * bracket's end derivatives, or bisection when the same end moved twice in a row (kinks where the
     * second derivative jumps - cone zone changes - defeat plain Newton).  If the tolerance is not met
     * the answer is `lo`, a point with negative derivative: by convexity its cost is below the start's. */
    double d1, d2, alpha = 0, lo = 0, hi = -1, dlo, dhi = 0;
    int last_side = 0, same_side = 0, converged = 0;
    double w1 = -1, w2 = -1; /* bracket widths one and two iterations ago */
    ls_eval(s, jar, jv, 0, g1, g2, &d1, &d2);
    if (d1 >= 0 || sn < 1e-300) break;
    dlo = d1;
    alpha = -d1 / d2;
    for (int it = 0; it < m->ls_iterations; it++) {
      ls_eval(s, jar, jv, alpha, g1, g2, &d1, &d2);
      if (fabs(d1) < gtol) { converged = 1; break; }
      int side = d1 < 0 ? -1 : 1;
      same_side = (side == last_side) ? same_side + 1 : 0;
      last_side = side;
      if (d1 < 0) { lo = alpha; dlo = d1; } else { hi = alpha; dhi = d1; }
      double an = alpha - d1 / d2;
      if (hi < 0) { if (!(an > alpha)) an = 2 * alpha; }
      else {
        const double w = hi - lo;
        if (!(an > lo && an < hi)) an = (same_side >= 1) ? 0.5 * (lo + hi) : lo - dlo * (hi - lo) / (dhi - dlo);
        /* Newton/secant steps that bounce across a kink shrink the bracket too slowly: bisect whenever
         * two iterations did not halve it */
        if (!(an > lo && an < hi) || (w2 > 0 && w > 0.5 * w2)) an = 0.5 * (lo + hi);
        w2 = w1; w1 = w;
        if (w <= 1e-14 * hi) break;
      }
      alpha = an;
    }
    if (!converged) alpha = lo > 0 ? lo : alpha;
#ifdef ORC_DEBUG
    fprintf(stderr, "newton iter %d cost %.10g gradnorm %.4g alpha %.6g d1 %.4g lo %.4g hi %.4g conv %d\n", iter, cost, sqrt(gn), alpha, d1, lo, hi, converged);
#endif
    if (alpha <= 0) break;
    for (int d = 0; d < nv; d++) { qacc[d] += alpha * search[d]; Ma[d] += alpha * Mv[d]; }
    for (int i = 0; i < nefc; i++) jar[i] += alpha * jv[i];
    double oldcost = cost;
    cost = constraint_update(s, jar, s->efc_force, hcone);
    double g = 0;
    for (int d = 0; d < nv; d++) g += (Ma[d] - s->qfrc_smooth[d]) * (qacc[d] - s->qacc_smooth[d]);
    cost += 0.5 * g;
    if (cost > oldcost) {
      /* never accept an uphill step (an unconverged line search that only had the first, overshooting
       * point): go back to the previous iterate, whose forces are consistent with it, and stop */
      for (int d = 0; d < nv; d++) { qacc[d] -= alpha * search[d]; Ma[d] -= alpha * Mv[d]; }
      for (int i = 0; i < nefc; i++) jar[i] -= alpha * jv[i];
      cost = constraint_update(s, jar, s->efc_force, hcone);
      break;
    }
    s->solver_niter = iter + 1;
#ifdef ORC_DEBUG
    fprintf(stderr, "   -> new cost %.10g (improvement %.4g)\n", cost, oldcost - cost);
#endif
    if (scale * (oldcost - cost) < m->tolerance) break;
  }
  memcpy(s->efc_jar, jar, sizeof(double) * nefc);
}

/* min 0.5 x'Ax + b'x  s.t. sum (x_j/d_j)^2 <= r^2  (mju_QCQP restated) */
static void qcqp(double *res, const double *A, const double *b, const double *d, double r, int n) {
  double As[25], bs[5], Al[25], Lc[25], v[5], t[5], la = 0;
  for (int i = 0; i < n; i++) { bs[i] = b[i] * d[i]; for (int j = 0; j < n; j++) As[i * n + j] = A[i * n + j] * d[i] * d[j]; }
  for (int iter = 0; iter < 20; iter++) {
    memcpy(Al, As, sizeof(double) * n * n);
    for (int i = 0; i < n; i++) Al[i * n + i] += la;
    if (chol_factor(Lc, Al, n)) { for (int i = 0; i < n; i++) res[i] = 0; return; }
    for (int i = 0; i < n; i++) v[i] = -bs[i];
    chol_solve(Lc, v, n);
    double val = -r * r;
    for (int i = 0; i < n; i++) val += v[i] * v[i];
    if (val < 1e-10) break;
    memcpy(t, v, sizeof(double) * n);
    chol_solve(Lc, t, n);
    double deriv = 0;
    for (int i = 0; i < n; i++) deriv += -2 * v[i] * t[i];
    double delta = -val / deriv;
    if (delta < 1e-10) break;
    la += delta;
  }
  for (int i = 0; i < n; i++) res[i] = v[i] * d[i];
}

/* mj_solNoSlip: Gauss-Seidel on friction rows with the UNREGULARISED A = J M^-1 J' */
static void solve_noslip(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  int nv = m->nv, nefc = s->nefc;
  if (m->noslip_iterations <= 0 || nefc == 0) return;
  double *B = (double *)malloc(sizeof(double) * (size_t)nefc * nv); /* rows: M^-1 J_i' for friction rows */
  double *w = s->w3;
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  /* w = M^-1 J' f.  mj_solNoSlip works in force space (residual_i = sum_j AR_ij f_j + b_i with b = J qacc_smooth - aref)
   * and ends with qacc = qacc_smooth + M^-1 J' f, so the acceleration it reasons about is the one the current FORCES
   * produce - not the primal iterate qacc, which differs from it whenever the Newton solve stopped short of
   * M (qacc - qacc_smooth) = J' f (on bodies with tiny inertia the gap can be tens of rad/s^2). */
  for (int d = 0; d < nv; d++) {
    double t = 0;
    for (int i = 0; i < nefc; i++) t += s->J[i * nv + d] * s->efc_force[i];
    w[d] = t;
  }
  chol_solve(s->L, w, nv);
  for (int i = 0; i < nefc; i++) {
    int fr = s->efc_type[i] == CT_FRICTION_DOF || (s->efc_type[i] == CT_CONTACT && i != s->con[s->efc_id[i]].efc);
    if (!fr) continue;
    memcpy(B + (size_t)i * nv, s->J + i * nv, sizeof(double) * nv);
    chol_solve(s->L, B + (size_t)i * nv, nv);
  }
  double *f = s->efc_force;
  for (int iter = 0; iter < m->noslip_iterations; iter++) {
    double improvement = 0;
    if (iter == 0)
      for (int i = 0; i < nefc; i++) {
        int fr = s->efc_type[i] == CT_FRICTION_DOF || (s->efc_type[i] == CT_CONTACT && i != s->con[s->efc_id[i]].efc);
        if (fr) improvement += 0.5 * f[i] * f[i] * s->efc_R[i];
      }
    for (int i = s->ne; i < s->ne + s->nf; i++) {
      const double *Ji = s->J + i * nv, *Bi = B + (size_t)i * nv;
      double res = -s->efc_aref[i], Aii = 0;
      for (int d = 0; d < nv; d++) { res += Ji[d] * (s->qacc_smooth[d] + w[d]); Aii += Ji[d] * Bi[d]; }
      double old = f[i], fn = old - res / fmax(MINVAL, Aii);
      fn = fmax(-s->efc_floss[i], fmin(s->efc_floss[i], fn));
      double delta = fn - old, change = 0.5 * delta * delta * Aii + delta * res;
      if (change > 1e-10) { fn = old; delta = 0; change = 0; }
      f[i] = fn;
      for (int d = 0; d < nv; d++) w[d] += Bi[d] * delta;
      improvement -= change;
    }
    for (int c = 0; c < s->ncon; c++) {
      Contact *con = &s->con[c];
      int i = con->efc, dim = con->dim;
      if (i < 0 || dim < 3) continue;
      int n = dim - 1;
      double Ac[25], res[5], bc[5], old[5], v[5], delta[5];
      for (int j = 0; j < n; j++) {
        const double *Jj = s->J + (i + 1 + j) * nv;
        double r = -s->efc_aref[i + 1 + j];
        for (int d = 0; d < nv; d++) r += Jj[d] * (s->qacc_smooth[d] + w[d]);
        res[j] = r; old[j] = f[i + 1 + j];
        for (int k = 0; k < n; k++) {
          const double *Bk = B + (size_t)(i + 1 + k) * nv;
          double a = 0;
          for (int d = 0; d < nv; d++) a += Jj[d] * Bk[d];
          Ac[j * n + k] = a;
        }
      }
      for (int j = 0; j < n; j++) { bc[j] = res[j]; for (int k = 0; k < n; k++) bc[j] -= Ac[j * n + k] * old[k]; }
      if (f[i] < MINVAL) for (int j = 0; j < n; j++) v[j] = 0;
      else qcqp(v, Ac, bc, con->friction, f[i], n);
      double change = 0;
      for (int j = 0; j < n; j++) delta[j] = v[j] - old[j];
      for (int j = 0; j < n; j++) { change += delta[j] * res[j]; for (int k = 0; k < n; k++) change += 0.5 * delta[j] * Ac[j * n + k] * delta[k]; }
      if (change > 1e-10) { for (int j = 0; j < n; j++) { v[j] = old[j]; delta[j] = 0; } change = 0; }
      for (int j = 0; j < n; j++) {
        f[i + 1 + j] = v[j];
        const double *Bj = B + (size_t)(i + 1 + j) * nv;
        for (int d = 0; d < nv; d++) w[d] += Bj[d] * delta[j];
      }
      improvement -= change;
    }
    if (improvement * scale < m->noslip_tolerance) break;
  }
  /* qfrc_constraint = J' f ; qacc = qacc_smooth + M^-1 qfrc_constraint */
  for (int d = 0; d < nv; d++) {
    double t = 0;
    for (int i = 0; i < nefc; i++) t += s->J[i * nv + d] * f[i];
    s->qfrc_constraint[d] = t;
  }
  memcpy(s->qacc, s->qfrc_constraint, sizeof(double) * nv);
  chol_solve(s->L, s->qacc, nv);
  for (int d = 0; d < nv; d++) s->qacc[d] += s->qacc_smooth[d];
  free(B);
}

/* ------------------------------------------------------------------ forward / step */
void orc_forward(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  int nv = m->nv;
  kinematics(s); com_pos(s); tendon(s); transmission(s); crb_factor(s);
  collision(s); make_constraint(s);
  if (s->ncon > s->ncon_peak) s->ncon_peak = s->ncon;
  if (s->nefc > s->nefc_peak) s->nefc_peak = s->nefc;
  com_vel(s); passive(s); rne(s); actuation(s);
  for (int d = 0; d < nv; d++) s->qfrc_smooth[d] = s->qfrc_passive[d] - s->qfrc_bias[d] + s->qfrc_actuator[d];
  memcpy(s->qacc_smooth, s->qfrc_smooth, sizeof(double) * nv);
  chol_solve(s->L, s->qacc_smooth, nv);
  reference_constraint(s);
  if (s->nefc == 0) {
    memcpy(s->qacc, s->qacc_smooth, sizeof(double) * nv);
    memcpy(s->qacc_warmstart, s->qacc_smooth, sizeof(double) * nv);
    memset(s->qfrc_constraint, 0, sizeof(double) * nv);
    return;
  }
  solve_newton(s);
  for (int d = 0; d < nv; d++) {
    double t = 0;
    for (int i = 0; i < s->nefc; i++) t += s->J[i * nv + d] * s->efc_force[i];
    s->qfrc_constraint[d] = t;
  }
  memcpy(s->qacc_warmstart, s->qacc, sizeof(double) * nv);
  solve_noslip(s);
}

static int bad_vec(const double *v, int n) {
  for (int i = 0; i < n; i++) if (!(fabs(v[i]) < 1e10)) return 1;
  return 0;
}

/* implicitfast: (M - h*dF/dv) a = f ; v += h a ; q integrates with the new v */
static void integrate(OrcSim *s) {
  const MgsModelDesc *m = &s->m;
  int nv = m->nv;
  double h = m->timestep, *A = s->w1, *Lc = s->w2, *a = s->w3;
  memcpy(A, s->M, sizeof(double) * nv * nv);
  for (int d = 0; d < nv; d++) A[d * nv + d] += h * m->dof_damping[d];
  for (int u = 0; u < m->nu; u++) {
    double bv = m->actuator_biasprm[3 * u + 2];
    if (bv == 0) continue;
    if (m->actuator_forcelimited[u] && (s->act_force[u] <= m->actuator_forcerange[2 * u] || s->act_force[u] >= m->actuator_forcerange[2 * u + 1])) continue;
    const double *mo = s->act_moment + u * nv;
    for (int i = 0; i < nv; i++) {
      if (mo[i] == 0) continue;
      for (int j = 0; j < nv; j++) A[i * nv + j] -= h * bv * mo[i] * mo[j];
    }
  }
  chol_factor(Lc, A, nv);
  for (int d = 0; d < nv; d++) a[d] = s->qfrc_smooth[d] + s->qfrc_constraint[d];
  chol_solve(Lc, a, nv);
  for (int d = 0; d < nv; d++) s->qvel[d] += h * a[d];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      for (int k = 0; k < 3; k++) s->qpos[qa + k] += h * s->qvel[da + k];
      double w[3] = {s->qvel[da + 3], s->qvel[da + 4], s->qvel[da + 5]};
      double ang = norm3(w) * h;
      if (ang > 0) {
        double ax[3] = {w[0], w[1], w[2]}, q[4], r[4];
        normalize3(ax);
        q[0] = cos(0.5 * ang); q[1] = sin(0.5 * ang) * ax[0]; q[2] = sin(0.5 * ang) * ax[1]; q[3] = sin(0.5 * ang) * ax[2];
        mulquat(r, s->qpos + qa + 3, q);
        memcpy(s->qpos + qa + 3, r, 4 * sizeof(double));
        normquat(s->qpos + qa + 3);
      }
    } else s->qpos[qa] += h * s->qvel[da];
  }
  s->time += h;
}

void orc_set_analytic(OrcSim *s, int on) { s->no_analytic = !on; }
void orc_set_qvel_clip(OrcSim *s, double clip) { s->qvel_clip = clip > 0 ? clip : 0; }

int orc_step(OrcSim *s, int nstep) {
  for (int k = 0; k < nstep; k++) {
    if (s->qvel_clip > 0)
      for (int i = 0; i < s->m.nv; i++) s->qvel[i] = s->qvel[i] > s->qvel_clip ? s->qvel_clip : (s->qvel[i] < -s->qvel_clip ? -s->qvel_clip : s->qvel[i]);
    if (s->bad || bad_vec(s->qpos, s->m.nq) || bad_vec(s->qvel, s->m.nv)) { s->bad = 1; return -1; }
    orc_forward(s);
    if (bad_vec(s->qacc, s->m.nv)) { s->bad = 1; return -1; }
    integrate(s);
    s->nstep_done++;
  }
  return 0;
}

/* ------------------------------------------------------------------ rollout logic (reference L4) */
typedef struct {
  int nstep_close, nstep_lift, shake_steps, repose_on_close;
  double lift_dist, shake_dist;
} OrcRolloutCfg;

/* set_qpos(joints) + set_pose(base) of /root/reference/mgs/core/simualtion.py:45-49 and
 * /root/reference/mgs/gripper/base.py:48-59 (the intermediate mj_forward calls only move the
 * Newton warm start and are folded into the single forward that follows). */
void orc_place(OrcSim *s, const double *pose7, int base_qadr, const double *joints, const int *jadr, int nj) {
  for (int k = 0; k < nj; k++) s->qpos[jadr[k]] = joints[k];
  for (int k = 0; k < 7; k++) s->qpos[base_qadr + k] = pose7[k];
  for (int k = 0; k < 3; k++) s->mocap_pos[k] = pose7[k];
  for (int k = 0; k < 4; k++) s->mocap_quat[k] = pose7[3 + k];
}

/* check_contact_with_object, gravityless_object_grasping.py:309-321 */
int orc_contact_with_object(const OrcSim *s) {
  int g = s->m.ground_geomid;
  for (int c = 0; c < s->ncon; c++) {
    int a = s->con[c].geom1, b = s->con[c].geom2;
    if ((a < g && b > g) || (a > g && b < g)) return 1;
  }
  return 0;
}

/* grasp_collision_mask body, gravityless_object_grasping.py:112-122: returns 1 if ANY contact */
int orc_grasp_collision(OrcSim *s, const double *pose7, int base_qadr, const double *joints, const int *jadr, int nj) {
  orc_reset(s);
  orc_place(s, pose7, base_qadr, joints, jadr, nj);
  orc_forward(s);
  return s->ncon != 0;
}

static void round_f32(double *v, int n) { for (int i = 0; i < n; i++) v[i] = (double)(float)v[i]; }

/* grasp_stability_evaluation_from_joints body for one candidate, :158-277 */
int orc_grasp_stability(OrcSim *s, const double *pose7, int base_qadr, const double *joints, const int *jadr, int nj,
                        const double *close_ctrl, const OrcRolloutCfg *cfg, long long *steps_out) {
  int label = 0;
  orc_reset(s);
  orc_place(s, pose7, base_qadr, joints, jadr, nj);
  orc_forward(s);
  /* close_gripper_at: mocap <- pose, ctrl <- close signal, mj_step x nstep_close (panda.py:225-241) */
  if (cfg->repose_on_close) orc_place(s, pose7, base_qadr, NULL, NULL, 0);
  for (int k = 0; k < 3; k++) s->mocap_pos[k] = pose7[k];
  for (int k = 0; k < 4; k++) s->mocap_quat[k] = pose7[3 + k];
  for (int u = 0; u < s->m.nu; u++) s->ctrl[u] = close_ctrl[u];
  if (orc_step(s, cfg->nstep_close) || !orc_contact_with_object(s)) goto done;
  /* lift, :205-226 */
  {
    double z0 = s->mocap_pos[2], zt = z0 + cfg->lift_dist;
    for (int t = 0; t < cfg->nstep_lift; t++) {
      s->mocap_pos[2] = z0 + (zt - z0) * ((double)t / cfg->nstep_lift);
      if (orc_step(s, 1)) goto done;
      if (t > 0 && t % 100 == 0 && !orc_contact_with_object(s)) goto done;
    }
    if (!orc_contact_with_object(s)) goto done;
  }
  /* shake, :229-276 (current_mocap_pose goes through SE3Pose => float32) */
  {
    double p32[3] = {s->mocap_pos[0], s->mocap_pos[1], s->mocap_pos[2]};
    double q32[4] = {s->mocap_quat[0], s->mocap_quat[1], s->mocap_quat[2], s->mocap_quat[3]};
    double R[9], back[3], right[3], left[3], tb[3], tr[3], tl[3], start[3];
    round_f32(p32, 3); round_f32(q32, 4);
    normquat(q32); quat2mat(R, q32); round_f32(R, 9);
    for (int k = 0; k < 3; k++) { back[k] = -R[3 * k + 2]; right[k] = R[3 * k + 1]; left[k] = -R[3 * k + 1]; }
    for (int k = 0; k < 3; k++) tb[k] = p32[k] + back[k] * cfg->shake_dist;
    copy3(start, s->mocap_pos);
    for (int t = 0; t < cfg->shake_steps; t++) {
      for (int k = 0; k < 3; k++) s->mocap_pos[k] = start[k] + (tb[k] - start[k]) * ((double)t / cfg->shake_steps);
      if (orc_step(s, 1)) goto done;
    }
    if (!orc_contact_with_object(s)) goto done;
    for (int k = 0; k < 3; k++) tr[k] = tb[k] + right[k] * cfg->shake_dist;
    copy3(start, s->mocap_pos);
    for (int t = 0; t < cfg->shake_steps; t++) {
      for (int k = 0; k < 3; k++) s->mocap_pos[k] = start[k] + (tr[k] - start[k]) * ((double)t / cfg->shake_steps);
      if (orc_step(s, 1)) goto done;
    }
    if (!orc_contact_with_object(s)) goto done;
    /* left: starts again from `start` (the beginning of the right move) - reference quirk 2 */
    for (int k = 0; k < 3; k++) tl[k] = start[k] + left[k] * (2 * cfg->shake_dist);
    for (int t = 0; t < 2 * cfg->shake_steps; t++) {
      for (int k = 0; k < 3; k++) s->mocap_pos[k] = start[k] + (tl[k] - start[k]) * ((double)t / (2 * cfg->shake_steps));
      if (orc_step(s, 1)) goto done;
    }
    if (!orc_contact_with_object(s)) goto done;
    label = 1;
  }
done:
  if (steps_out) *steps_out = s->nstep_done;
  return label;
}

/* ------------------------------------------------------------------ clutter table (reference clutter_table.py) */
/* check_gripper_collision (:237-252): gripper <-> table or anything numbered after the table */
int orc_gripper_collision(const OrcSim *s) {
  int g = s->m.ground_geomid;
  for (int c = 0; c < s->ncon; c++) {
    int a = s->con[c].geom1, b = s->con[c].geom2;
    if ((a < g && b > g) || (a > g && b < g) || (a == g && b < g) || (a < g && b == g)) return 1;
  }
  return 0;
}
/* check_gripper_contact (:254-270): as written, `A or (B and not C)` reduces to gripper <-> anything after the table */
int orc_gripper_contact(const OrcSim *s) { return orc_contact_with_object(s); }

/* load the scene state record: qpos[nq] qvel[nv] qacc_warmstart[nv] ctrl[nu] mocap_pos[3] mocap_quat[4] */
void orc_set_record(OrcSim *s, const double *rec) {
  const MgsModelDesc *m = &s->m;
  memcpy(s->qpos, rec, sizeof(double) * m->nq); rec += m->nq;
  memcpy(s->qvel, rec, sizeof(double) * m->nv); rec += m->nv;
  memcpy(s->qacc_warmstart, rec, sizeof(double) * m->nv); rec += m->nv;
  memcpy(s->ctrl, rec, sizeof(double) * m->nu); rec += m->nu;
  if (m->nmocap) { memcpy(s->mocap_pos, rec, sizeof(double) * 3); memcpy(s->mocap_quat, rec + 3, sizeof(double) * 4); }
  s->bad = 0; s->nstep_done = 0; s->ncon = 0; s->nefc = 0;
}
void orc_get_record(const OrcSim *s, double *rec) {
  const MgsModelDesc *m = &s->m;
  memcpy(rec, s->qpos, sizeof(double) * m->nq); rec += m->nq;
  memcpy(rec, s->qvel, sizeof(double) * m->nv); rec += m->nv;
  memcpy(rec, s->qacc_warmstart, sizeof(double) * m->nv); rec += m->nv;
  memcpy(rec, s->ctrl, sizeof(double) * m->nu); rec += m->nu;
  if (m->nmocap) { memcpy(rec, s->mocap_pos, sizeof(double) * 3); memcpy(rec + 3, s->mocap_quat, sizeof(double) * 4); }
}

/* ClutterTableEnv.grasp_collision_mask body (:356-364); the workspace bound test (:344-354) is done by the caller */
int orc_clutter_collision(OrcSim *s, const double *scene, const double *pose7, int base_qadr, const double *joints, const int *jadr, int nj) {
  orc_set_record(s, scene);
  orc_place(s, pose7, base_qadr, joints, jadr, nj);
  orc_forward(s);
  return orc_gripper_collision(s);
}

/* ClutterTableEnv.grasp_stable_mask body (:288-317): restore scene, place, close, lift with the check at (t+1) % 100 == 0 */
int orc_clutter_stable(OrcSim *s, const double *scene, const double *pose7, int base_qadr, const double *joints, const int *jadr, int nj,
                       const double *close_ctrl, const OrcRolloutCfg *cfg, long long *steps_out) {
  int label = 0;
  orc_set_record(s, scene);
  orc_place(s, pose7, base_qadr, joints, jadr, nj);
  orc_forward(s);
  if (cfg->repose_on_close) orc_place(s, pose7, base_qadr, NULL, NULL, 0);
  for (int k = 0; k < 3; k++) s->mocap_pos[k] = pose7[k];
  for (int k = 0; k < 4; k++) s->mocap_quat[k] = pose7[3 + k];
  for (int u = 0; u < s->m.nu; u++) s->ctrl[u] = close_ctrl[u];
  if (orc_step(s, cfg->nstep_close)) goto done;
  {
    double z0 = s->mocap_pos[2], zt = z0 + cfg->lift_dist;
    label = 1;
    for (int t = 0; t < cfg->nstep_lift; t++) {
      s->mocap_pos[2] = z0 + (zt - z0) * ((double)t / cfg->nstep_lift);
      if (orc_step(s, 1)) { label = 0; break; }
      if ((t + 1) % 100 == 0 && !orc_gripper_contact(s)) { label = 0; break; }
    }
  }
done:
  if (steps_out) *steps_out = s->nstep_done;
  return label;
}

/* ------------------------------------------------------------------ threaded batch (CPU baseline) */
#include <pthread.h>
typedef struct {
  const MgsModelDesc *d; int n, nthreads, tid, base_qadr, nj, mode;
  const double *poses, *joints, *close_ctrl; const int *jadr; const OrcRolloutCfg *cfg;
  unsigned char *labels; long long *steps; const double *scene;
} BatchArg;
static void *batch_worker(void *p) {
  BatchArg *a = (BatchArg *)p;
  OrcSim *s = orc_create(a->d);
  for (int i = a->tid; i < a->n; i += a->nthreads) {
    if (a->mode == 0) {
      a->labels[i] = (unsigned char)!orc_grasp_collision(s, a->poses + 7 * i, a->base_qadr, a->joints + (size_t)a->nj * i, a->jadr, a->nj);
      if (a->steps) a->steps[i] = 0;
    } else if (a->mode == 2) {
      a->labels[i] = (unsigned char)!orc_clutter_collision(s, a->scene, a->poses + 7 * i, a->base_qadr, a->joints + (size_t)a->nj * i, a->jadr, a->nj);
      if (a->steps) a->steps[i] = 0;
    } else if (a->mode == 3) {
      long long st = 0;
      a->labels[i] = (unsigned char)orc_clutter_stable(s, a->scene, a->poses + 7 * i, a->base_qadr, a->joints + (size_t)a->nj * i, a->jadr, a->nj,
                                                       a->close_ctrl, a->cfg, &st);
      if (a->steps) a->steps[i] = st;
    } else {
      long long st = 0;
      a->labels[i] = (unsigned char)orc_grasp_stability(s, a->poses + 7 * i, a->base_qadr, a->joints + (size_t)a->nj * i, a->jadr, a->nj,
                                                        a->close_ctrl, a->cfg, &st);
      if (a->steps) a->steps[i] = st;
    }
  }
  orc_destroy(s);
  return NULL;
}
/* mode 0: collision-free mask; mode 1: stability labels; 2 / 3: the clutter-table versions (need `scene`).
 * One OrcSim per thread, candidates strided. */
int orc_batch_scene(const MgsModelDesc *d, int mode, int n, const double *poses, int base_qadr, const double *joints, const int *jadr,
                    int nj, const double *close_ctrl, const OrcRolloutCfg *cfg, int nthreads, unsigned char *labels, long long *steps,
                    const double *scene);
int orc_batch(const MgsModelDesc *d, int mode, int n, const double *poses, int base_qadr, const double *joints, const int *jadr,
              int nj, const double *close_ctrl, const OrcRolloutCfg *cfg, int nthreads, unsigned char *labels, long long *steps) {
  return orc_batch_scene(d, mode, n, poses, base_qadr, joints, jadr, nj, close_ctrl, cfg, nthreads, labels, steps, NULL);
}
int orc_batch_scene(const MgsModelDesc *d, int mode, int n, const double *poses, int base_qadr, const double *joints, const int *jadr,
                    int nj, const double *close_ctrl, const OrcRolloutCfg *cfg, int nthreads, unsigned char *labels, long long *steps,
                    const double *scene) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  BatchArg args[256];
  for (int t = 0; t < nthreads; t++) {
    BatchArg a = {d, n, nthreads, t, base_qadr, nj, mode, poses, joints, close_ctrl, jadr, cfg, labels, steps, scene};
    args[t] = a;
    pthread_create(&th[t], NULL, batch_worker, &args[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  return 0;
}

/* ------------------------------------------------------------------ accessors for tests */
#define ACC(name) double *orc_##name(OrcSim *s) { return s->name; }
ACC(qpos) ACC(qvel) ACC(ctrl) ACC(mocap_pos) ACC(mocap_quat) ACC(qacc_warmstart) ACC(xpos) ACC(xquat) ACC(xmat) ACC(xipos)
ACC(gxpos) ACC(gxmat) ACC(M) ACC(qfrc_bias) ACC(qfrc_passive) ACC(qfrc_actuator) ACC(qfrc_smooth) ACC(qacc_smooth) ACC(qacc)
ACC(qfrc_constraint) ACC(J) ACC(efc_pos) ACC(efc_D) ACC(efc_R) ACC(efc_aref) ACC(efc_force) ACC(efc_jar) ACC(subtree_com) ACC(cdof)
int orc_ncon(OrcSim *s) { return s->ncon; }
int orc_ncon_peak(OrcSim *s) { return s->ncon_peak; }
int orc_nefc_peak(OrcSim *s) { return s->nefc_peak; }
int orc_nefc(OrcSim *s) { return s->nefc; }
int orc_niter(OrcSim *s) { return s->solver_niter; }
int orc_bad(OrcSim *s) { return s->bad; }
int *orc_efc_type(OrcSim *s) { return s->efc_type; }
/* contact c -> out[0:3]=pos, [3:12]=frame, [12]=dist, [13]=geom1, [14]=geom2, [15]=dim, [16]=mu, [17]=efc */
void orc_contact(OrcSim *s, int c, double *out) {
  const Contact *k = &s->con[c];
  memcpy(out, k->pos, 3 * sizeof(double)); memcpy(out + 3, k->frame, 9 * sizeof(double));
  out[12] = k->dist; out[13] = k->geom1; out[14] = k->geom2; out[15] = k->dim; out[16] = k->mu; out[17] = k->efc;
}
void orc_kinematics_only(OrcSim *s) { kinematics(s); com_pos(s); tendon(s); transmission(s); crb_factor(s); }
void orc_collision_only(OrcSim *s) { kinematics(s); collision(s); }
