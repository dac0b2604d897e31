// mgs_model_build.h - host side: MgsModelDesc (fp64, from the MJCF compiler) -> one contiguous
// blob of `real`/int arrays + a DevModel whose pointers index into it.  The CUDA library uploads
// the blob once per model (it is KBs: read-only, shared by every warp, L1/L2 resident).
#pragma once
#include <string>
#include <vector>

#include "../../include/mgs_model_desc.h"
#include "mgs_common.cuh"

struct ModelBlob {
  std::vector<char> bytes;
  DevModel dm;  // pointers hold byte OFFSETS until rebase()
  int ncon_max, nefc_max;
  int rows_static, rows_per_contact;  // nefc_max default = rows_static + ncon_max * rows_per_contact
  int nfreeobj;                       // free-floating leaf bodies (grasped / cluttered objects)
};

namespace mgs_detail {
template <typename T, typename S>
inline const T *put(std::vector<char> &b, const S *src, size_t n) {
  size_t off = (b.size() + 15) & ~size_t(15);
  b.resize(off + (n ? n : 1) * sizeof(T));
  T *dst = reinterpret_cast<T *>(b.data() + off);
  for (size_t i = 0; i < n; i++) dst[i] = (T)src[i];
  return reinterpret_cast<const T *>(off);
}
}  // namespace mgs_detail

inline bool build_model_blob(const MgsModelDesc *d, ModelBlob &out, std::string &err) {
  using mgs_detail::put;
  DevModel &m = out.dm;
  std::vector<char> &b = out.bytes;
  b.clear();
  b.resize(16);
  memset(&m, 0, sizeof(m));
  if (d->nu > MGS_MAX_NU) { err = "too many actuators"; return false; }
  if (d->nv > 255) { err = "nv > 255 (the lower-triangle index table packs a dof index in 8 bits)"; return false; }
  if (d->nmocap > 1) { err = "at most one mocap body is supported"; return false; }
  for (int p = 0; p < d->npair; p++)
    if (d->pair_condim[p] != 1 && d->pair_condim[p] != 3 && d->pair_condim[p] != 4) { err = "condim must be 1, 3 or 4"; return false; }
  for (int g = 0; g < d->ncgeom; g++) {
    int t = d->cgeom_type[g];
    if (t != GEOM_SPHERE && t != GEOM_CAPSULE && t != GEOM_CYLINDER && t != GEOM_BOX && t != GEOM_MESH) { err = "unsupported geom type"; return false; }
  }
  m.nq = d->nq; m.nv = d->nv; m.nu = d->nu; m.nbody = d->nbody; m.njnt = d->njnt; m.neq = d->neq; m.nmocap = d->nmocap;
  m.ntendon = d->ntendon; m.nwrap = d->nwrap; m.ncgeom = d->ncgeom; m.npair = d->npair; m.nhull = d->nhull;
  m.cone_elliptic = d->cone_elliptic; m.iterations = d->iterations; m.ls_iterations = d->ls_iterations;
  m.noslip_iterations = d->noslip_iterations; m.mpr_iterations = d->mpr_iterations; m.ground_geomid = d->ground_geomid;
  m.timestep = (real)d->timestep; m.impratio = (real)d->impratio; m.tolerance = (real)d->tolerance; m.ls_tolerance = (real)d->ls_tolerance;
  m.noslip_tolerance = (real)d->noslip_tolerance; m.mpr_tolerance = (real)d->mpr_tolerance; m.meaninertia = (real)d->meaninertia;
  for (int k = 0; k < 3; k++) m.gravity[k] = (real)d->gravity[k];
  const int nb = d->nbody, nv = d->nv, nq = d->nq, nj = d->njnt, nu = d->nu, ng = d->ncgeom, np = d->npair, nh = d->nhull;
  // derived: tree depth, subtree mass, static equality row addresses
  std::vector<int> depth(nb, 0), rowadr(d->neq, 0);
  std::vector<double> stm(nb, 0.0);
  int maxdepth = 0;
  for (int i = 1; i < nb; i++) { depth[i] = depth[d->body_parentid[i]] + 1; if (depth[i] > maxdepth) maxdepth = depth[i]; }
  for (int i = 0; i < nb; i++) stm[i] = d->body_mass[i];
  for (int i = nb - 1; i > 0; i--) stm[d->body_parentid[i]] += stm[i];
  int ne = 0;
  for (int q = 0; q < d->neq; q++) {
    rowadr[q] = ne;
    if (d->eq_active[q]) ne += d->eq_type[q] == EQ_WELD ? 6 : (d->eq_type[q] == EQ_CONNECT ? 3 : 1);
  }
  m.maxdepth = maxdepth; m.ne_rows = ne;
#define PI_(f, n) m.f = put<int>(b, d->f, (size_t)(n))
#define PR_(f, n) m.f = put<real>(b, d->f, (size_t)(n))
  PI_(body_parentid, nb); PI_(body_rootid, nb); PI_(body_mocapid, nb); PI_(body_jntadr, nb); PI_(body_jntnum, nb); PI_(body_dofadr, nb);
  PI_(body_dofnum, nb);
  m.body_depth = put<int>(b, depth.data(), nb);
  PR_(body_pos, 3 * nb); PR_(body_quat, 4 * nb); PR_(body_ipos, 3 * nb); PR_(body_iquat, 4 * nb); PR_(body_mass, nb); PR_(body_inertia, 3 * nb);
  PR_(body_invweight0, 2 * nb);
  m.body_subtreemass = put<real>(b, stm.data(), nb);
  PR_(body_gravcomp, nb);
  for (int i = 0; i < nb; i++) m.ngravcomp += (d->body_gravcomp[i] != 0 && d->body_mass[i] != 0);
  PI_(jnt_type, nj); PI_(jnt_bodyid, nj); PI_(jnt_qposadr, nj); PI_(jnt_dofadr, nj); PI_(jnt_limited, nj);
  PR_(jnt_pos, 3 * nj); PR_(jnt_axis, 3 * nj); PR_(jnt_range, 2 * nj); PR_(jnt_stiffness, nj); PR_(jnt_solref, 2 * nj); PR_(jnt_solimp, 5 * nj);
  PR_(jnt_margin, nj); PR_(qpos0, nq); PR_(qpos_spring, nq);
  PI_(dof_bodyid, nv); PI_(dof_jntid, nv); PI_(dof_parentid, nv);
  {
    // kinematic trees = diagonal blocks of the mass matrix (dofs of one tree are contiguous)
    std::vector<int> tadr(nv, 0), tnum(nv, 0);
    int maxt = 0;
    for (int i = 0; i < nv;) {
      int root = d->body_rootid[d->dof_bodyid[i]], j = i;
      while (j < nv && d->body_rootid[d->dof_bodyid[j]] == root) j++;
      for (int k = i; k < j; k++) { tadr[k] = i; tnum[k] = j - i; }
      if (j - i > maxt) maxt = j - i;
      i = j;
    }
    m.dof_treeadr = put<int>(b, tadr.data(), nv);
    m.dof_treenum = put<int>(b, tnum.data(), nv);
    m.max_tree_dofs = maxt;
    // block storage of the per-tree matrices
    std::vector<int> rowoff(nv > 0 ? nv : 1, 0), ij;
    int nM = 0;
    for (int i = 0; i < nv;) {
      const int t = tnum[i];
      for (int r = 0; r < t; r++) {
        rowoff[i + r] = nM + r * t - i;
        for (int c = 0; c < t; c++) ij.push_back(((i + r) << 8) | (i + c));
      }
      nM += t * t;
      i += t;
    }
    m.nM = nM;
    // tree index of every dof and of every body (the tree of the nearest ancestor that has dofs; -1 for static bodies)
    std::vector<int> dtree(nv > 0 ? nv : 1, 0), btree(nb, -1);
    int ntree = 0;
    for (int i = 0; i < nv;) { for (int k = 0; k < tnum[i]; k++) dtree[i + k] = ntree; ntree++; i += tnum[i]; }
    for (int i = 1; i < nb; i++) {
      int bb = i;
      while (bb > 0 && d->body_dofnum[bb] == 0) bb = d->body_parentid[bb];
      if (bb > 0) btree[i] = dtree[d->body_dofadr[bb]];
    }
    m.ntree = ntree;
    m.dof_treeid = put<int>(b, dtree.data(), nv);
    m.body_treeid = put<int>(b, btree.data(), nb);
    m.dof_rowoff = put<int>(b, rowoff.data(), nv);
    std::vector<int> blkw(nv > 0 ? nv : 1, 0);
    for (int i = 0; i < nv; i++) blkw[i] = (int)((unsigned)rowoff[i] | ((unsigned)tadr[i] << 16) | ((unsigned)tnum[i] << 24));
    m.dof_blk = put<int>(b, blkw.data(), nv);
    m.blk_ij = put<int>(b, ij.data(), ij.size());
    std::vector<int> tm;
    for (int a = 0; a < nv; a++)
      for (int c = 0; c <= a; c++) tm.push_back(tadr[a] == tadr[c] ? rowoff[a] + c : -1);
    m.tri_madr = put<int>(b, tm.data(), tm.size());
  }
  PR_(dof_armature, nv); PR_(dof_damping, nv); PR_(dof_frictionloss, nv); PR_(dof_solref, 2 * nv); PR_(dof_solimp, 5 * nv); PR_(dof_invweight0, nv);
  PI_(cgeom_geomid, ng); PI_(cgeom_type, ng); PI_(cgeom_bodyid, ng); PI_(cgeom_hullid, ng);
  PR_(cgeom_pos, 3 * ng); PR_(cgeom_quat, 4 * ng); PR_(cgeom_size, 3 * ng); PR_(cgeom_rbound, ng);
  PI_(hull_vertadr, nh); PI_(hull_vertnum, nh); PI_(hull_faceadr, nh); PI_(hull_facenum, nh); PI_(hull_facevertadr, d->nhullface);
  PI_(hull_facevertnum, d->nhullface); PI_(hull_facevert, d->nhullfacevert); PI_(hull_nbradr, d->nhullvert); PI_(hull_nbrnum, d->nhullvert);
  PI_(hull_nbr, d->nhullnbr); PR_(hull_vert, 3 * d->nhullvert); PR_(hull_facenormal, 3 * d->nhullface);
  PI_(pair_geom1, np); PI_(pair_geom2, np); PI_(pair_condim, np); PR_(pair_friction, 5 * np); PR_(pair_solref, 2 * np); PR_(pair_solimp, 5 * np);
  PR_(pair_margin, np);
  PI_(tendon_adr, d->ntendon); PI_(tendon_num, d->ntendon); PI_(wrap_dofadr, d->nwrap); PI_(wrap_qposadr, d->nwrap); PR_(wrap_coef, d->nwrap);
  PI_(actuator_trntype, nu); PI_(actuator_trnid, nu); PI_(actuator_ctrllimited, nu); PI_(actuator_forcelimited, nu);
  PR_(actuator_gainprm, 3 * nu); PR_(actuator_biasprm, 3 * nu); PR_(actuator_ctrlrange, 2 * nu); PR_(actuator_forcerange, 2 * nu); PR_(actuator_gear, nu);
  PI_(eq_type, d->neq); PI_(eq_obj1id, d->neq); PI_(eq_obj2id, d->neq); PI_(eq_active, d->neq);
  m.eq_rowadr = put<int>(b, rowadr.data(), d->neq);
  std::vector<int> tri;
  for (int a = 0; a < nv; a++)
    for (int c = 0; c <= a; c++) tri.push_back((a << 8) | c);
  m.tri_ab = put<int>(b, tri.data(), tri.size());
  {
    std::vector<int> rank(nv > 0 ? nv : 1, -1);
    int nfr_rows = 0;
    for (int i = 0; i < nv; i++) if (d->dof_frictionloss[i] > 0) rank[i] = nfr_rows++;
    m.nf_rows = nfr_rows;
    m.dof_frictionrank = put<int>(b, rank.data(), nv);
  }
  {
    // ancestry masks: which dofs move a point fixed to body i (lane-per-dof Jacobian assembly)
    const int words = (nv + 31) / 32 > 0 ? (nv + 31) / 32 : 1;
    std::vector<unsigned int> mask((size_t)nb * words, 0u);
    for (int i = 1; i < nb; i++) {
      int bb = i;
      while (bb > 0 && d->body_dofnum[bb] == 0) bb = d->body_parentid[bb];
      if (bb == 0) continue;
      for (int k = d->body_dofadr[bb] + d->body_dofnum[bb] - 1; k >= 0; k = d->dof_parentid[k]) mask[(size_t)i * words + (k >> 5)] |= 1u << (k & 31);
    }
    m.dofmask_words = words;
    m.body_dofmask = put<unsigned int>(b, mask.data(), mask.size());
  }
  PR_(eq_data, 11 * d->neq); PR_(eq_solref, 2 * d->neq); PR_(eq_solimp, 5 * d->neq);
  PR_(mocap_pos0, 3 * d->nmocap); PR_(mocap_quat0, 4 * d->nmocap);
#undef PI_
#undef PR_
  // scratch capacities: static rows + limits + a contact budget
  int nfr = 0, nlim = 0;
  for (int i = 0; i < nv; i++) nfr += d->dof_frictionloss[i] > 0;
  for (int j = 0; j < nj; j++) nlim += d->jnt_limited[j] ? 1 : 0;
  int maxdim = 1;
  for (int p = 0; p < np; p++) if (d->pair_condim[p] > maxdim) maxdim = d->pair_condim[p];
  // Default capacities (contacts beyond them are dropped and flagged): twice the number of candidate
  // pairs that involve a free single-body tree (a grasped object), clamped to [32, 64] contacts.
  int nobjpair = 0;
  for (int p = 0; p < np; p++) {
    for (int side = 0; side < 2; side++) {
      const int bd = d->cgeom_bodyid[side ? d->pair_geom2[p] : d->pair_geom1[p]];
      bool leaf = true;
      for (int c = 0; c < nb; c++) if (d->body_parentid[c] == bd && c != bd) leaf = false;
      if (d->body_parentid[bd] == 0 && d->body_dofnum[bd] == 6 && leaf) { nobjpair++; break; }
    }
  }
  int nc = 2 * nobjpair;
  nc = nc < 32 ? 32 : (nc > 64 ? 64 : nc);
  // many-dof models (the 16-dof hands: nv >= 28) pay nv reals per Jacobian row: 32 contacts were never exceeded on the
  // round-1 workloads (tools/caps_sweep.py: Allegro / LEAP / Shadow, 1184 candidates each, 0 environments overflowed) and
  // double the number of environments resident per SM (Shadow 2 -> 4, LEAP 3 -> 5)
  // (single-object scenes only: every resting object of a clutter scene needs its own 3-4 table contacts)
  int nfreeobj = 0;
  for (int bd = 1; bd < nb; bd++) {
    bool leaf = true;
    for (int c = 0; c < nb; c++) if (d->body_parentid[c] == bd && c != bd) leaf = false;
    if (d->body_parentid[bd] == 0 && d->body_dofnum[bd] == 6 && leaf && d->body_mocapid[bd] < 0) nfreeobj++;
  }
  if (nv > 24 && nfreeobj <= 1) nc = 32;
  nc = (nc + 3) & ~3;
  out.ncon_max = nc;
  out.nfreeobj = nfreeobj;
  out.rows_static = ne + nfr + nlim;
  out.rows_per_contact = maxdim < 3 ? 3 : maxdim;
  out.nefc_max = out.rows_static + nc * out.rows_per_contact;  // every contact slot can hold the largest cone
  return true;
}

// turn offsets into pointers relative to `base`
inline void rebase_model(DevModel &m, const char *base) {
  const char **p = reinterpret_cast<const char **>(&m.body_parentid);
  const char **end = reinterpret_cast<const char **>(&m.mocap_quat0) + 1;
  for (; p != end; ++p) *p = base + reinterpret_cast<size_t>(*p);
}
