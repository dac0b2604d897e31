// lane1.cpp - TEST HARNESS ONLY.  Compiles the device source (mj_grasp_sim_b200/csrc/*.cuh) with
// -DMGS_HOST so that every warp-cooperative routine runs as a 1-lane scalar program on the CPU.
// The CPU-only test tier uses it to compare the kernel's arithmetic with the fp64 oracle before
// any GPU time is spent.  It is never linked into, loaded by, or reachable from the product
// library libmgs_b200.so (which fails loudly without a CUDA device).
#define MGS_HOST 1
#include <stdlib.h>

#include <string>

#include "../../include/mgs_b200.h"
#include "../../mj_grasp_sim_b200/csrc/mgs_model_build.h"
#include "../../mj_grasp_sim_b200/csrc/mgs_rollout.cuh"

struct L1Model {
  ModelBlob blob;
  DevModel dm;
  Layout L;
  int state_stride, diag_stride;
  std::vector<real> scratch;
  std::vector<float> aux;
  double qvel_clip = 0;
};
static std::string g_err;

extern "C" const char *l1_last_error(void) { return g_err.c_str(); }

extern "C" int l1_model_create(const MgsModelDesc *desc, L1Model **out) {
  L1Model *M = new L1Model();
  if (!build_model_blob(desc, M->blob, g_err)) { delete M; return -1; }
  M->dm = M->blob.dm;
  rebase_model(M->dm, M->blob.bytes.data());
  layout_compute(&M->L, desc->nq, desc->nv, desc->nu, desc->nbody, desc->njnt, desc->nmocap, desc->ntendon, desc->ncgeom,
                 M->blob.ncon_max, M->blob.nefc_max, desc->npair, M->blob.dm.nM);
  M->state_stride = desc->nq + 2 * desc->nv + desc->nu + 7 * desc->nmocap;
  M->diag_stride = mgs_diag_stride(desc->nv, desc->nbody, M->blob.ncon_max, M->blob.nefc_max);
  M->scratch.assign(M->L.total, 0);
  *out = M;
  return 0;
}
extern "C" void l1_model_destroy(L1Model *M) { delete M; }
extern "C" int l1_model_info(const L1Model *M, MgsModelInfo *info) {
  memset(info, 0, sizeof(*info));
  info->nq = M->dm.nq; info->nv = M->dm.nv; info->nu = M->dm.nu; info->nmocap = M->dm.nmocap;
  info->state_stride = M->state_stride; info->diag_stride = M->diag_stride;
  info->ncon_max = M->L.ncon_max; info->nefc_max = M->L.nefc_max;
  info->smem_bytes_per_env = (int)(M->L.total * sizeof(real));
  info->real_bytes = (int)sizeof(real);
  return 0;
}

static void run_all(L1Model *M, const RolloutParams &prm, const BatchIO &io_in) {
  Env e;
  BatchIO io = io_in;
  M->aux.assign((size_t)4 * (prm.n > 0 ? prm.n : 1), 0.0f);
  io.aux = M->aux.data();
  g_k.m = M->dm; g_k.L = M->L; g_k.prm = prm; g_k.io = io;
  for (int env = 0; env < prm.n; env++) {
    env_bind(e, M->scratch.data());
    run_env_w(e, env);
  }
}

extern "C" int l1_last_aux(L1Model *M, int n, float *out) {
  if ((size_t)4 * n > M->aux.size()) { g_err = "l1_last_aux: the most recent call had fewer candidates"; return -1; }
  memcpy(out, M->aux.data(), sizeof(float) * 4 * n);
  return 0;
}
extern "C" int l1_set_qvel_clip(L1Model *M, double clip) { M->qvel_clip = clip > 0 ? clip : 0; return 0; }

extern "C" int l1_step_host(L1Model *M, int n, int nstep, const void *state_in, void *state_out, void *diag_out) {
  RolloutParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.mode = MGS_MODE_STEP; prm.n = n; prm.nstep = nstep; prm.qvel_clip = (real)M->qvel_clip;
  BatchIO io;
  memset(&io, 0, sizeof(io));
  io.state_in = (const real *)state_in; io.state_out = (real *)state_out; io.diag_out = (real *)diag_out;
  io.state_stride = M->state_stride; io.diag_stride = M->diag_stride;
  run_all(M, prm, io);
  return 0;
}

extern "C" int l1_clutter_host(L1Model *M, int mode, int n, const double *scene, const float *pose7, const float *joints, int nj,
                               const int *joint_qposadr, int base_qposadr, const double *close_ctrl, const MgsRolloutCfg *cfg, uint8_t *labels,
                               int *steps) {
  RolloutParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.mode = mode; prm.n = n; prm.nj = nj; prm.base_qposadr = base_qposadr;
  for (int k = 0; k < nj; k++) prm.joint_qposadr[k] = joint_qposadr[k];
  if (cfg) {
    prm.nstep_close = cfg->nstep_close; prm.nstep_lift = cfg->nstep_lift; prm.shake_steps = cfg->shake_steps;
    prm.repose_on_close = cfg->repose_on_close; prm.lift_dist = (real)cfg->lift_dist; prm.shake_dist = (real)cfg->shake_dist;
  }
  if (close_ctrl) for (int u = 0; u < M->dm.nu; u++) prm.close_ctrl[u] = (real)close_ctrl[u];
  std::vector<real> rec(M->state_stride);
  for (int i = 0; i < M->state_stride; i++) rec[i] = (real)scene[i];
  BatchIO io;
  memset(&io, 0, sizeof(io));
  io.pose7 = pose7; io.joints = joints; io.labels = labels; io.steps = steps;
  io.state_in = rec.data(); io.state_stride = M->state_stride;
  run_all(M, prm, io);
  return 0;
}

extern "C" int l1_rollout_host(L1Model *M, int mode, int n, const float *pose7, const float *joints, int nj, const int *joint_qposadr,
                               int base_qposadr, const double *close_ctrl, const MgsRolloutCfg *cfg, uint8_t *labels, int *steps) {
  RolloutParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.mode = mode; prm.n = n; prm.nj = nj; prm.base_qposadr = base_qposadr;
  for (int k = 0; k < nj; k++) prm.joint_qposadr[k] = joint_qposadr[k];
  if (cfg) {
    prm.nstep_close = cfg->nstep_close; prm.nstep_lift = cfg->nstep_lift; prm.shake_steps = cfg->shake_steps;
    prm.repose_on_close = cfg->repose_on_close; prm.lift_dist = (real)cfg->lift_dist; prm.shake_dist = (real)cfg->shake_dist;
  }
  if (close_ctrl) for (int u = 0; u < M->dm.nu; u++) prm.close_ctrl[u] = (real)close_ctrl[u];
  BatchIO io;
  memset(&io, 0, sizeof(io));
  io.pose7 = pose7; io.joints = joints; io.labels = labels; io.steps = steps;
  run_all(M, prm, io);
  return 0;
}
