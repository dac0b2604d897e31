// mgs_rollout.cuh - the fused per-candidate programs (reference L4 logic on the device).
//
//   MGS_MODE_COLLISION : grasp_collision_mask body   (gravityless_object_grasping.py:112-122)
//   MGS_MODE_STABILITY : grasp_stability_evaluation_from_joints body (:158-277): reset, place,
//                        close nstep_close steps, contact test, lift with the periodic contact
//                        test, shake back / right / left, label
//   MGS_MODE_STEP      : generic "load state, mj_step x nstep, store state" used by parity tests
// One warp runs one candidate start to finish; state never leaves shared memory in between.
#pragma once
#include "mgs_solver.cuh"

#define MGS_DIAG_HEADER 8
static inline int mgs_diag_stride(int nv, int nbody, int ncon_max, int nefc_max) {
  return MGS_DIAG_HEADER + 3 * nv + nv * nv + 7 * nbody + 5 * ncon_max + 4 * nefc_max;
}

MGS_DEV void reset_w(const DevModel &m, Env &e) {
  PFOR(i, m.nq) e.qpos[i] = LDG(m.qpos0 + i);
  PFOR(i, m.nv) { e.qvel[i] = 0; e.qacc_ws[i] = 0; }
  PFOR(i, m.nu) e.ctrl[i] = 0;
  PFOR(i, m.nmocap) {
    for (int k = 0; k < 3; k++) e.mocap[7 * i + k] = LDG(m.mocap_pos0 + 3 * i + k);
    for (int k = 0; k < 4; k++) e.mocap[7 * i + 3 + k] = LDG(m.mocap_quat0 + 4 * i + k);
  }
  e.bad = 0; e.overflow = 0; e.ncon = 0; e.nefc = 0;
  WSYNC();
}

// set_qpos(joints) + set_pose(base): simualtion.py:45-49, gripper/base.py:48-59
MGS_DEV void place_w(const RolloutParams &prm, Env &e, const float *pose7, const float *joints) {
  if (joints) PFOR(k, prm.nj) e.qpos[prm.joint_qposadr[k]] = (real)LDG(joints + k);
  PFOR(k, 7) {
    real v = (real)LDG(pose7 + k);
    e.qpos[prm.base_qposadr + k] = v;
    e.mocap[k] = v;
  }
  WSYNC();
}

MGS_DEV real round32(real x) { return (real)(float)x; }

// linear mocap ramp of n steps from start to target (targets are never quite reached: t/n, t<n)
MGS_DEV int ramp_w(const DevModel &m, Env &e, const real *start, const real *target, int n, int check_every, int *steps) {
  for (int t = 0; t < n; t++) {
    real f = (real)t / (real)n;
    PFOR(k, 3) e.mocap[k] = start[k] + (target[k] - start[k]) * f;
    WSYNC();
    if (step_w(m, e, 1, steps)) return 1;
    if (check_every > 0 && t > 0 && t % check_every == 0 && !contact_with_object_w(m, e)) return 1;
  }
  return 0;
}

MGS_DEVN int stability_program_w(const DevModel &m, Env &e, const RolloutParams &prm, const float *pose7, const float *joints, int *steps) {
  reset_w(m, e);
  place_w(prm, e, pose7, joints);
  forward_w(m, e);
  // close_gripper_at (panda.py:225-241 and the five siblings): mocap <- pose, ctrl <- close signal
  if (prm.repose_on_close) place_w(prm, e, pose7, (const float *)0);
  PFOR(k, 7) e.mocap[k] = (real)LDG(pose7 + k);
  PFOR(u, m.nu) e.ctrl[u] = prm.close_ctrl[u];
  WSYNC();
  if (step_w(m, e, prm.nstep_close, steps)) return 0;
  if (!contact_with_object_w(m, e)) return 0;
  // lift (:205-226)
  real start[3], target[3];
  copy3(start, e.mocap);
  copy3(target, start);
  target[2] = start[2] + prm.lift_dist;
  WSYNC();
  if (ramp_w(m, e, start, target, prm.nstep_lift, 100, steps)) return 0;
  if (!contact_with_object_w(m, e)) return 0;
  // shake (:229-276); current_mocap_pose passes through SE3Pose => float32 pos/quat/rotation
  real p32[3], q32[4], Rm[9], tb[3], tr[3], tl[3];
  for (int k = 0; k < 3; k++) p32[k] = round32(e.mocap[k]);
  for (int k = 0; k < 4; k++) q32[k] = round32(e.mocap[3 + k]);
  normquat(q32);
  quat2mat(Rm, q32);
  for (int k = 0; k < 9; k++) Rm[k] = round32(Rm[k]);
  for (int k = 0; k < 3; k++) tb[k] = p32[k] - Rm[3 * k + 2] * prm.shake_dist;  // back = R (0,0,-1)
  copy3(start, e.mocap);
  WSYNC();
  if (ramp_w(m, e, start, tb, prm.shake_steps, 0, steps)) return 0;
  if (!contact_with_object_w(m, e)) return 0;
  for (int k = 0; k < 3; k++) tr[k] = tb[k] + Rm[3 * k + 1] * prm.shake_dist;  // right = R (0,1,0)
  copy3(start, e.mocap);
  WSYNC();
  if (ramp_w(m, e, start, tr, prm.shake_steps, 0, steps)) return 0;
  if (!contact_with_object_w(m, e)) return 0;
  // left: restarts from the START of the right move (reference quirk: ~2 cm mocap jump at t=0)
  for (int k = 0; k < 3; k++) tl[k] = start[k] - Rm[3 * k + 1] * (2 * prm.shake_dist);
  if (ramp_w(m, e, start, tl, 2 * prm.shake_steps, 0, steps)) return 0;
  if (!contact_with_object_w(m, e)) return 0;
  return 1;
}

MGS_DEV void write_diag_w(const DevModel &m, const Env &e, real *o) {
  const int nv = m.nv;
  PFOR(k, 1) { o[0] = (real)e.ncon; o[1] = (real)e.nefc; o[2] = (real)e.niter; o[3] = (real)e.bad; o[4] = (real)e.overflow; o[5] = (real)e.ne; o[6] = (real)e.nf; o[7] = (real)e.nl; }
  real *p = o + MGS_DIAG_HEADER;
  PFOR(d, nv) { p[d] = e.qacc[d]; p[nv + d] = e.qacc_smooth[d]; p[2 * nv + d] = e.qfrc_smooth[d]; }
  p += 3 * nv;
  PFOR(i, nv * nv) p[i] = e.M[i];
  p += nv * nv;
  PFOR(i, 3 * m.nbody) p[i] = e.xpos[i];
  p += 3 * m.nbody;
  PFOR(i, 4 * m.nbody) p[i] = e.xquat[i];
  p += 4 * m.nbody;
  PFOR(c, e.ncon_max) {
    int ok = c < e.ncon;
    p[5 * c] = ok ? e.con_dist[c] : 0;
    for (int k = 0; k < 3; k++) p[5 * c + 1 + k] = ok ? e.con_pos[3 * c + k] : 0;
    p[5 * c + 4] = ok ? (real)IARR(e.con_pair)[c] : -1;
  }
  p += 5 * e.ncon_max;
  PFOR(i, e.nefc_max) {
    int ok = i < e.nefc;
    p[4 * i] = ok ? e.efc_aref[i] : 0; p[4 * i + 1] = ok ? e.efc_D[i] : 0;
    p[4 * i + 2] = ok ? e.efc_force[i] : 0; p[4 * i + 3] = ok ? e.efc_jar[i] : 0;
  }
}

// run one candidate / environment on this warp
MGS_DEVN void run_env_w(const DevModel &m, Env &e, const RolloutParams &prm, const BatchIO &io, int env) {
  int steps = 0;
  if (prm.mode == MGS_MODE_STEP) {
    const real *in = io.state_in + (size_t)env * io.state_stride;
    e.bad = 0; e.overflow = 0;
    PFOR(i, m.nq) e.qpos[i] = in[i];
    PFOR(i, m.nv) { e.qvel[i] = in[m.nq + i]; e.qacc_ws[i] = in[m.nq + m.nv + i]; }
    PFOR(i, m.nu) e.ctrl[i] = in[m.nq + 2 * m.nv + i];
    PFOR(i, 7 * m.nmocap) e.mocap[i] = in[m.nq + 2 * m.nv + m.nu + i];
    WSYNC();
    if (prm.nstep > 0) step_w(m, e, prm.nstep, &steps);
    else forward_w(m, e);
    real *out = io.state_out + (size_t)env * io.state_stride;
    PFOR(i, m.nq) out[i] = e.qpos[i];
    PFOR(i, m.nv) { out[m.nq + i] = e.qvel[i]; out[m.nq + m.nv + i] = e.qacc_ws[i]; }
    PFOR(i, m.nu) out[m.nq + 2 * m.nv + i] = e.ctrl[i];
    PFOR(i, 7 * m.nmocap) out[m.nq + 2 * m.nv + m.nu + i] = e.mocap[i];
    if (io.diag_out) write_diag_w(m, e, io.diag_out + (size_t)env * io.diag_stride);
    PFOR(k, 1) { if (io.labels) io.labels[env] = (uint8_t)(e.bad ? 0 : 1); if (io.steps) io.steps[env] = steps; }
    WSYNC();
    return;
  }
  const float *pose7 = io.pose7 + (size_t)env * 7, *joints = io.joints + (size_t)env * prm.nj;
  int label;
  if (prm.mode == MGS_MODE_COLLISION) {
    reset_w(m, e);
    place_w(prm, e, pose7, joints);
    forward_w(m, e);
    label = (e.ncon == 0);  // collision-free mask: no contact of any kind (check_contact, :306-307)
  } else {
    label = stability_program_w(m, e, prm, pose7, joints, &steps);
  }
  PFOR(k, 1) { io.labels[env] = (uint8_t)label; if (io.steps) io.steps[env] = steps; }
  if (io.state_out) {
    real *out = io.state_out + (size_t)env * io.state_stride;
    PFOR(i, m.nq) out[i] = e.qpos[i];
    PFOR(i, m.nv) { out[m.nq + i] = e.qvel[i]; out[m.nq + m.nv + i] = e.qacc_ws[i]; }
    PFOR(i, m.nu) out[m.nq + 2 * m.nv + i] = e.ctrl[i];
    PFOR(i, 7 * m.nmocap) out[m.nq + 2 * m.nv + m.nu + i] = e.mocap[i];
  }
  WSYNC();
}
