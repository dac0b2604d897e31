// mgs_collide.cuh - broadphase + convex narrowphase, ONE GEOM PAIR PER LANE.
//
// The candidate pair list is static (filtered on the host: contype/conaffinity, <exclude>,
// parent-child, same body) and its contact parameters are pre-mixed, so the device does a
// bounding-sphere cull, then Minkowski Portal Refinement (the algorithm MuJoCo 3.2.2 reaches
// through libccd for mesh pairs) with analytic box supports and hill-climbing hull supports, then
// clips the two most-aligned polytope faces against each other for a <=4 point manifold.
// Contact slots are assigned with a warp prefix sum so the contact order is deterministic
// (pair order), which keeps the Gauss-Seidel noslip pass reproducible.
#pragma once
#include "mgs_sim.cuh"

#define MGS_MAXPOLY 8
#define MGS_MAXCLIP 12
#define MGS_FACE_ALIGN_MIN R_(0.9990)

struct SupPt { real v[3], v1[3], v2[3]; int ia, ib; };  // ia/ib: vertex ids of the two support points (-1: not a vertex)

struct GeomRef {
  int cg, type, hull;
  const real *R, *p;
  int cur;  // hill-climbing start vertex (persists across support calls of one query)
};

MGS_DEV void geomref_init(GeomRef &g, const Env &e, int cg) {
  g.cg = cg;
  g.type = LDG(MD.cgeom_type + cg);
  g.hull = LDG(MD.cgeom_hullid + cg);
  g.R = EF(gxmat) + 9 * cg;
  g.p = EF(gxpos) + 3 * cg;
  g.cur = 0;
}

struct Sup { real x, y, z; int cur, vid; };

// support point (world) of collision geom `cg` in world direction d.  Arguments and result travel in
// registers (the function is deliberately NOT inlined: one copy of the code serves every call site,
// which keeps the kernel's instruction footprint small).
MGS_DEVN Sup geom_support(const real *Rm, const real *gp, int cg, int type, int hull, int cur, real dx, real dy, real dz) {
  real dl[3], p[3] = {0, 0, 0}, d[3] = {dx, dy, dz};
  mulmatTvec3(dl, Rm, d);
  real s0 = LDG(MD.cgeom_size + 3 * cg), s1 = LDG(MD.cgeom_size + 3 * cg + 1), s2 = LDG(MD.cgeom_size + 3 * cg + 2);
  int vid = -1;
  if (type == GEOM_BOX) {
    p[0] = dl[0] >= 0 ? s0 : -s0;
    p[1] = dl[1] >= 0 ? s1 : -s1;
    p[2] = dl[2] >= 0 ? s2 : -s2;
    vid = (dl[0] >= 0 ? 1 : 0) | (dl[1] >= 0 ? 2 : 0) | (dl[2] >= 0 ? 4 : 0);
  } else if (type == GEOM_MESH) {
    // hill climbing on the hull's vertex graph from the previous answer
    const int vadr = LDG(MD.hull_vertadr + hull), nvert = LDG(MD.hull_vertnum + hull);
    const real *V = MD.hull_vert + 3 * vadr;
    real best = LDG(V + 3 * cur) * dl[0] + LDG(V + 3 * cur + 1) * dl[1] + LDG(V + 3 * cur + 2) * dl[2];
#pragma unroll 1
    for (int guard = 0; guard < nvert; guard++) {
      int na = LDG(MD.hull_nbradr + vadr + cur), nn = LDG(MD.hull_nbrnum + vadr + cur), nxt = cur;
      MGS_UNROLL_INNER  // the neighbour -> vertex loads of several neighbours overlap (they were one dependent chain each)
      for (int k = 0; k < nn; k++) {
        int v = LDG(MD.hull_nbr + na + k);
        real t = LDG(V + 3 * v) * dl[0] + LDG(V + 3 * v + 1) * dl[1] + LDG(V + 3 * v + 2) * dl[2];
        if (t > best) { best = t; nxt = v; }
      }
      if (nxt == cur) break;
      cur = nxt;
    }
    ld3(p, V + 3 * cur);
    vid = cur;
  } else if (type == GEOM_SPHERE) {
    scl3(p, dl, s0);
  } else if (type == GEOM_CAPSULE) {
    scl3(p, dl, s0);
    p[2] += dl[2] >= 0 ? s1 : -s1;
  } else if (type == GEOM_CYLINDER) {
    real n = sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
    if (n > MGS_MINVAL) { p[0] = dl[0] / n * s0; p[1] = dl[1] / n * s0; }
    p[2] = dl[2] >= 0 ? s1 : -s1;
  }
  real o[3];
  mulmatvec3(o, Rm, p);
  Sup r;
  r.x = o[0] + gp[0]; r.y = o[1] + gp[1]; r.z = o[2] + gp[2]; r.cur = cur; r.vid = vid;
  return r;
}

MGS_DEV void mink_support(GeomRef &g1, GeomRef &g2, const real *d, SupPt &o) {
  Sup a = geom_support(g1.R, g1.p, g1.cg, g1.type, g1.hull, g1.cur, d[0], d[1], d[2]);
  Sup b = geom_support(g2.R, g2.p, g2.cg, g2.type, g2.hull, g2.cur, -d[0], -d[1], -d[2]);
  g1.cur = a.cur; g2.cur = b.cur;
  o.ia = a.vid; o.ib = b.vid;
  o.v1[0] = a.x; o.v1[1] = a.y; o.v1[2] = a.z;
  o.v2[0] = b.x; o.v2[1] = b.y; o.v2[2] = b.z;
  sub3(o.v, o.v1, o.v2);
}

// world position of vertex `vid` of a box / hull geom (warm start of the portal)
MGS_DEVN Sup geom_vertex(const real *Rm, const real *gp, int cg, int type, int hull, int vid) {
  real p[3], o[3];
  if (type == GEOM_BOX) {
    const real s0 = LDG(MD.cgeom_size + 3 * cg), s1 = LDG(MD.cgeom_size + 3 * cg + 1), s2 = LDG(MD.cgeom_size + 3 * cg + 2);
    p[0] = (vid & 1) ? s0 : -s0; p[1] = (vid & 2) ? s1 : -s1; p[2] = (vid & 4) ? s2 : -s2;
  } else {
    ld3(p, MD.hull_vert + 3 * (LDG(MD.hull_vertadr + hull) + vid));
  }
  mulmatvec3(o, Rm, p);
  Sup r;
  r.x = o[0] + gp[0]; r.y = o[1] + gp[1]; r.z = o[2] + gp[2]; r.cur = vid; r.vid = vid;
  return r;
}
MGS_DEV void portal_point(const GeomRef &g1, const GeomRef &g2, int ia, int ib, SupPt &o) {
  Sup a = geom_vertex(g1.R, g1.p, g1.cg, g1.type, g1.hull, ia);
  Sup b = geom_vertex(g2.R, g2.p, g2.cg, g2.type, g2.hull, ib);
  o.ia = ia; o.ib = ib;
  o.v1[0] = a.x; o.v1[1] = a.y; o.v1[2] = a.z;
  o.v2[0] = b.x; o.v2[1] = b.y; o.v2[2] = b.z;
  sub3(o.v, o.v1, o.v2);
}

MGS_DEV void portal_dir(const SupPt &p1, const SupPt &p2, const SupPt &p3, real *dir) {
  real a[3], b[3];
  sub3(a, p2.v, p1.v);
  sub3(b, p3.v, p1.v);
  cross3(dir, a, b);
  normalize3(dir);
}

MGS_DEV void expand_portal(const SupPt &p0, SupPt &p1, SupPt &p2, SupPt &p3, const SupPt &v4) {
  real v4v0[3];
  cross3(v4v0, v4.v, p0.v);
  if (dot3(p1.v, v4v0) > 0) {
    if (dot3(p2.v, v4v0) > 0) p1 = v4; else p3 = v4;
  } else {
    if (dot3(p3.v, v4v0) > 0) p2 = v4; else p1 = v4;
  }
}

MGS_DEV int reach_tolerance(const SupPt &p1, const SupPt &p2, const SupPt &p3, const SupPt &v4, const real *dir, real tol) {
  real d4 = dot3(v4.v, dir);
  real mn = fmin(d4 - dot3(p1.v, dir), fmin(d4 - dot3(p2.v, dir), d4 - dot3(p3.v, dir)));
  return mn <= tol;
}

#ifdef MGS_INLINE_COT
MGS_DEV
#else
MGS_DEVN
#endif
void closest_on_triangle(const real *a, const real *b, const real *c, real *out) {
  real ab[3], ac[3], ap[3] = {-a[0], -a[1], -a[2]};
  sub3(ab, b, a); sub3(ac, c, a);
  real d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0 && d2 <= 0) { copy3(out, a); return; }
  real bp[3] = {-b[0], -b[1], -b[2]};
  real d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0 && d4 <= d3) { copy3(out, b); return; }
  real vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { real v = d1 / (d1 - d3); copy3(out, a); addscl3(out, ab, v); return; }
  real cp[3] = {-c[0], -c[1], -c[2]};
  real d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0 && d5 <= d6) { copy3(out, c); return; }
  real vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { real w = d2 / (d2 - d6); copy3(out, a); addscl3(out, ac, w); return; }
  real va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    real w = (d4 - d3) / ((d4 - d3) + (d5 - d6)), bc[3];
    sub3(bc, c, b); copy3(out, b); addscl3(out, bc, w); return;
  }
  real den = R_(1.0) / (va + vb + vc), v = vb * den, w = vc * den;
  copy3(out, a); addscl3(out, ab, v); addscl3(out, ac, w);
}

// Minkowski Portal Refinement as a LOCKSTEP STATE MACHINE: every lane owns one geom pair, and every
// iteration of the loop makes exactly one support query at one call site.  Lanes therefore stay
// converged through the expensive part (the hill-climbing hull support) no matter in which phase of
// the algorithm their pair is (ncu r1_b: the straight-line version ran geom_support with 2 of 32 lanes
// active and spent 24 % of all issued instructions there).
// Phases: V1, V2 (first two portal vertices), V3 (third vertex / portal discovery), REFINE (portal must
// pass the origin), PEN (push the portal to the surface).  `active` = this lane has a pair to test.
// Returns 1 when penetrating: depth > 0, dir from g1 to g2 (unit), pos.
enum { MPR_V1 = 0, MPR_V2, MPR_V3, MPR_REFINE, MPR_PEN, MPR_DONE, MPR_AXIS };
MGS_DEVN int mpr_penetration(GeomRef &g1, GeomRef &g2, int active, int *cache, real *depth, real *dir, real *pos) {
  const real tol = MD.mpr_tolerance;
  SupPt p0, p1, p2, p3, s;
  real d[3] = {1, 0, 0}, va[3], vb[3];
  int state = MPR_DONE, hit = 0, it = 0;
  if (active) {
    copy3(p0.v1, g1.p); copy3(p0.v2, g2.p);
    sub3(p0.v, p0.v1, p0.v2);
    if (dot3(p0.v, p0.v) < R_(1e-28)) p0.v[0] += R_(1e-10);
    scl3(d, p0.v, -1); normalize3(d);
    state = MPR_V1;
    // WARM START: a pair that was penetrating at the previous step remembers the vertex pairs of its final
    // portal.  Rebuilt at the new poses, that triangle is almost always still a valid portal (the origin ray
    // from v0 passes through it), so the search resumes in the refinement phase and needs 1-3 support
    // queries instead of 20-30.  (MPR's answer is the boundary face of the Minkowski difference pierced by
    // that ray: it does not depend on the starting portal beyond mpr_tolerance.)
    if (cache && cache[2] >= 0 && (g1.type == GEOM_BOX || g1.type == GEOM_MESH) && (g2.type == GEOM_BOX || g2.type == GEOM_MESH)) {
      portal_point(g1, g2, cache[0] & 0xffff, cache[0] >> 16, p1);
      portal_point(g1, g2, cache[1] & 0xffff, cache[1] >> 16, p2);
      portal_point(g1, g2, cache[2] & 0xffff, cache[2] >> 16, p3);
      real a1[3], a2[3], a3[3], t[3];
      sub3(a1, p1.v, p0.v); sub3(a2, p2.v, p0.v); sub3(a3, p3.v, p0.v);
      cross3(t, a2, a3);
      real T = dot3(a1, t);
      if (T < 0) { SupPt sw = p2; p2 = p3; p3 = sw; T = -T; }
      real s12, s23, s31;
      cross3(t, p1.v, p2.v); s12 = -dot3(t, p0.v);
      cross3(t, p2.v, p3.v); s23 = -dot3(t, p0.v);
      cross3(t, p3.v, p1.v); s31 = -dot3(t, p0.v);
      if (T > R_(1e-20) && s12 >= 0 && s23 >= 0 && s31 >= 0) {
        portal_dir(p1, p2, p3, d);
        state = (dot3(d, p1.v) >= 0) ? MPR_PEN : MPR_REFINE;
        it = 0;
      }
    }
    // SEPARATING-AXIS CACHE: a pair that was found separated remembers the direction that proved it (16-bit
    // components).  One support query along it usually proves separation again; otherwise start cold.
    if (cache && cache[2] == -2) {
      const int w0 = cache[0], w1 = cache[1];
      d[0] = (real)(short)(w0 & 0xffff); d[1] = (real)(short)((w0 >> 16) & 0xffff); d[2] = (real)(short)(w1 & 0xffff);
      normalize3(d);
      state = MPR_AXIS;
    }
  }
  int sep = 0;  // d is a direction with negative support of the Minkowski difference (proves separation)
  #pragma unroll 1
  for (;;) {
    // single warp-collective test per iteration; no lane leaves the body early (no `continue`), so all
    // lanes re-converge at the bottom of the loop
    if (!wany(state != MPR_DONE)) break;
    if (state != MPR_DONE) mink_support(g1, g2, d, s);  // the one (converged) support call site
    if (state == MPR_AXIS) {
      if (dot3(s.v, d) <= 0) { state = MPR_DONE; sep = 1; }
      else { scl3(d, p0.v, -1); normalize3(d); state = MPR_V1; }  // the old axis no longer separates: cold start
    } else if (state == MPR_V1) {
      p1 = s;
      if (dot3(p1.v, d) <= 0) { state = MPR_DONE; sep = 1; }
      else {
        cross3(d, p0.v, p1.v);
        if (dot3(d, d) < R_(1e-28) * fmax(R_(1e-30), dot3(p0.v, p0.v) * dot3(p1.v, p1.v))) {
          // origin on the segment v0-v1: penetration along v1
          *depth = sqrt(dot3(p1.v, p1.v));
          copy3(dir, p1.v); normalize3(dir);
          for (int k = 0; k < 3; k++) pos[k] = R_(0.5) * (p1.v1[k] + p1.v2[k]);
          hit = *depth > 0; state = MPR_DONE;
        } else {
          normalize3(d);
          state = MPR_V2;
        }
      }
    } else if (state == MPR_V2) {
      p2 = s;
      if (dot3(p2.v, d) <= 0) { state = MPR_DONE; sep = 1; }
      else {
        sub3(va, p1.v, p0.v); sub3(vb, p2.v, p0.v);
        cross3(d, va, vb); normalize3(d);
        if (dot3(d, p0.v) > 0) { SupPt t = p1; p1 = p2; p2 = t; scl3(d, d, -1); }
        state = MPR_V3; it = 0;
      }
    } else if (state == MPR_V3) {
      p3 = s;
      if (dot3(p3.v, d) <= 0) { state = MPR_DONE; sep = 1; }
      else if (++it > 100) state = MPR_DONE;
      else {
        int cont = 0;
        cross3(va, p1.v, p3.v);
        if (dot3(va, p0.v) < R_(-1e-30)) { p2 = p3; cont = 1; }
        if (!cont) {
          cross3(va, p3.v, p2.v);
          if (dot3(va, p0.v) < R_(-1e-30)) { p1 = p3; cont = 1; }
        }
        if (cont) {
          sub3(va, p1.v, p0.v); sub3(vb, p2.v, p0.v);
          cross3(d, va, vb); normalize3(d);
        } else {
          portal_dir(p1, p2, p3, d);
          state = (dot3(d, p1.v) >= 0) ? MPR_PEN : MPR_REFINE;  // origin already inside the portal?
          it = 0;
        }
      }
    } else if (state == MPR_REFINE) {
      if (dot3(s.v, d) < 0) { state = MPR_DONE; sep = 1; }
      else if (reach_tolerance(p1, p2, p3, s, d, tol) || ++it > MD.mpr_iterations) state = MPR_DONE;
      else {
        expand_portal(p0, p1, p2, p3, s);
        portal_dir(p1, p2, p3, d);
        if (dot3(d, p1.v) >= 0) { state = MPR_PEN; it = 0; }
      }
    } else if (state == MPR_PEN) {
      if (reach_tolerance(p1, p2, p3, s, d, tol) || ++it > MD.mpr_iterations) {
        real c[3];
        closest_on_triangle(p1.v, p2.v, p3.v, c);
        *depth = sqrt(dot3(c, c));
        if (*depth < MGS_MINVAL) copy3(dir, d); else scl3(dir, c, R_(1.0) / *depth);
        real b0, b1, b2, b3, t[3], sum;
        cross3(t, p1.v, p2.v); b0 = dot3(t, p3.v);
        cross3(t, p3.v, p2.v); b1 = dot3(t, p0.v);
        cross3(t, p0.v, p1.v); b2 = dot3(t, p3.v);
        cross3(t, p2.v, p1.v); b3 = dot3(t, p0.v);
        sum = b0 + b1 + b2 + b3;
        if (sum <= 0) {
          b0 = 0;
          cross3(t, p2.v, p3.v); b1 = dot3(t, d);
          cross3(t, p3.v, p1.v); b2 = dot3(t, d);
          cross3(t, p1.v, p2.v); b3 = dot3(t, d);
          sum = b1 + b2 + b3;
        }
        real is = R_(0.5) / sum;
        for (int k = 0; k < 3; k++)
          pos[k] = (b0 * (p0.v1[k] + p0.v2[k]) + b1 * (p1.v1[k] + p1.v2[k]) + b2 * (p2.v1[k] + p2.v2[k]) + b3 * (p3.v1[k] + p3.v2[k])) * is;
        hit = 1; state = MPR_DONE;
      } else {
        expand_portal(p0, p1, p2, p3, s);
        portal_dir(p1, p2, p3, d);
      }
    }
  }
  if (cache && active) {
    // remember the final portal of a penetrating pair (vertex ids fit 16 bits), forget it otherwise
    if (hit && p1.ia >= 0 && p1.ib >= 0 && p2.ia >= 0 && p2.ib >= 0 && p3.ia >= 0 && p3.ib >= 0 &&
        (p1.ia | p1.ib | p2.ia | p2.ib | p3.ia | p3.ib) < 32768) {
      cache[0] = p1.ia | (p1.ib << 16); cache[1] = p2.ia | (p2.ib << 16); cache[2] = p3.ia | (p3.ib << 16);
    } else if (!hit && sep) {
      const int qx = (int)(d[0] * R_(32000.0)), qy = (int)(d[1] * R_(32000.0)), qz = (int)(d[2] * R_(32000.0));
      cache[0] = (qx & 0xffff) | ((qy & 0xffff) << 16); cache[1] = qz & 0xffff; cache[2] = -2;
    } else cache[2] = -1;
    cache[3] = (g1.cur & 0xffff) | ((g2.cur & 0xffff) << 16);
  }
  return hit;
}

// Face of hull `hull` whose normal is most aligned with the LOCAL direction (nx, ny, nz); first maximum wins.
// Warp-cooperative: every lane calls it with the same arguments and scans a strided share of the faces
// (the per-lane sequential scan ran with 2 of 32 lanes active and was 6-10 % of all issued instructions).
MGS_DEVN int best_face_w(int hull, real nx, real ny, real nz, real *align) {
  const int fa = LDG(MD.hull_faceadr + hull), fn = LDG(MD.hull_facenum + hull);
  const real *FN = MD.hull_facenormal + 3 * fa;
  real bd = R_(-1e30);
  int best = 0x7fffffff;
  #pragma unroll 2
  PFOR(f, fn) {
    const real t = LDG(FN + 3 * f) * nx + LDG(FN + 3 * f + 1) * ny + LDG(FN + 3 * f + 2) * nz;
    if (t > bd) { bd = t; best = f; }
  }
  wargmax(bd, best);
  *align = bd;
  return best;
}
// a cylinder has two faces, its caps (0: +z, 1: -z): answered by the lane itself, no hull scan
MGS_DEV void cap_face(const GeomRef &g, const real *n, int *face, real *align) {
  const real nz = g.R[2] * n[0] + g.R[5] * n[1] + g.R[8] * n[2];
  *face = nz >= 0 ? 0 : 1;
  *align = fabs(nz);
}
// Does geom g offer a flat face to a contact with normal n?  Polytopes always, a cylinder when n is along its axis (a cap).
MGS_DEV int has_face(const GeomRef &g, const real *n) {
  if (g.type == GEOM_BOX || g.type == GEOM_MESH) return 1;
#ifndef MGS_NO_CYLINDER_CAPS
  if (g.type == GEOM_CYLINDER) return fabs(g.R[2] * n[0] + g.R[5] * n[1] + g.R[8] * n[2]) >= MGS_FACE_ALIGN_MIN;
#endif
  return 0;
}
// best faces for every lane with `want` set (its geom `g`, world direction `n`)
#ifdef MGS_WIDE
// env-per-CTA variant: the requests are compacted into a list and every WARP serves requests q = warp, warp + 8, ... with a
// warp-level scan + arg-max: five CTA barriers per call.  (Serving the requesters one after the other with block-level broadcasts and
// arg-max was ~12 barriers per requester, four calls per step, ~50 requesters each on the config-5 scene: most of its collision time.)
MGS_DEV void best_face_lanes(const Env &e, int want, const GeomRef &g, const real *n, int *face, real *align) {
  int nreq;
  if (want && g.type == GEOM_CYLINDER) { cap_face(g, n, face, align); want = 0; }
  const int rank = wrank(want, &nreq);
  if (nreq == 0) return;
  real *rq = EF(req_off);
  if (want) {
    real nl[3];
    mulmatTvec3(nl, g.R, n);
    *IARR(rq + 6 * rank) = g.hull; rq[6 * rank + 1] = nl[0]; rq[6 * rank + 2] = nl[1]; rq[6 * rank + 3] = nl[2];
  }
  WSYNC();
  const int l32 = threadIdx.x & 31;
  #pragma unroll 1
  for (int q = threadIdx.x >> 5; q < nreq; q += MGS_NWARP) {
    const int hull = *IARR(rq + 6 * q);
    const real nx = rq[6 * q + 1], ny = rq[6 * q + 2], nz = rq[6 * q + 3];
    const int fa = LDG(MD.hull_faceadr + hull), fn = LDG(MD.hull_facenum + hull);
    const real *FN = MD.hull_facenormal + 3 * fa;
    real bd = R_(-1e30);
    int best = 0x7fffffff;
    #pragma unroll 2
    for (int f = l32; f < fn; f += 32) {
      const real t = LDG(FN + 3 * f) * nx + LDG(FN + 3 * f + 1) * ny + LDG(FN + 3 * f + 2) * nz;
      if (t > bd) { bd = t; best = f; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const real v2 = __shfl_xor_sync(0xffffffffu, bd, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, best, o);
      if (v2 > bd || (v2 == bd && i2 < best)) { bd = v2; best = i2; }
    }
    if (l32 == 0) { *IARR(rq + 6 * q + 4) = best; rq[6 * q + 5] = bd; }
  }
  WSYNC();
  if (want) { *face = *IARR(rq + 6 * rank + 4); *align = rq[6 * rank + 5]; }
  WSYNC();
}
#else
MGS_DEV void best_face_lanes(const Env &e, int want, const GeomRef &g, const real *n, int *face, real *align) {
  (void)e;
  real nl[3] = {0, 0, 0};
  if (want && g.type == GEOM_CYLINDER) { cap_face(g, n, face, align); want = 0; }
  if (want) mulmatTvec3(nl, g.R, n);
  int pending = want;
  #pragma unroll 1
  for (;;) {
    const int src = wfirst(pending);  // lanes with a request are served in lane order
    if (src < 0) break;
    real al;
    const int f = best_face_w(wbcasti(g.hull, src), wbcast(nl[0], src), wbcast(nl[1], src), wbcast(nl[2], src), &al);
    if (MGS_LANE == src) { *face = f; *align = al; pending = 0; }
  }
}
#endif
MGS_DEVN int face_polygon(const GeomRef &g, int f, real (*poly)[3], real *nw) {
  if (g.type == GEOM_CYLINDER) {
    // the cap enters the clipping as the octagon inscribed in its rim, counter-clockwise seen from outside like the hull faces
    const real rc = LDG(MD.cgeom_size + 3 * g.cg), z = f == 0 ? LDG(MD.cgeom_size + 3 * g.cg + 1) : -LDG(MD.cgeom_size + 3 * g.cg + 1);
    const real h = R_(0.70710678118654752);
    #pragma unroll 1
    for (int i = 0; i < 8; i++) {
      const int k = f == 0 ? i : 7 - i;
      // cos, sin of k * 45 deg
      const real c = (k == 0) ? R_(1.0) : (k == 4) ? R_(-1.0) : (k == 2 || k == 6) ? R_(0.0) : (k == 1 || k == 7) ? h : -h;
      const real sn = (k == 2) ? R_(1.0) : (k == 6) ? R_(-1.0) : (k == 0 || k == 4) ? R_(0.0) : (k == 1 || k == 3) ? h : -h;
      const real v[3] = {rc * c, rc * sn, z};
      mulmatvec3(poly[i], g.R, v);
      add3(poly[i], poly[i], g.p);
    }
    const real s = f == 0 ? R_(1.0) : R_(-1.0);
    nw[0] = g.R[2] * s; nw[1] = g.R[5] * s; nw[2] = g.R[8] * s;
    return 8;
  }
  int gf = LDG(MD.hull_faceadr + g.hull) + f;
  int n = LDG(MD.hull_facevertnum + gf), fva = LDG(MD.hull_facevertadr + gf), va = LDG(MD.hull_vertadr + g.hull);
  #pragma unroll 1
  for (int i = 0; i < n; i++) {
    real v[3];
    ld3(v, MD.hull_vert + 3 * (va + LDG(MD.hull_facevert + fva + i)));
    mulmatvec3(poly[i], g.R, v);
    add3(poly[i], poly[i], g.p);
  }
  real fnl[3];
  ld3(fnl, MD.hull_facenormal + 3 * gf);
  mulmatvec3(nw, g.R, fnl);
  return n;
}

struct PairContacts { int n, own2; real normal[3], pos[4][3], dist[4], normal2[3]; };  // own2: point 1 carries its own normal (normal2)

#ifndef MGS_NO_ANALYTIC_PRIMS
// ---- closed-form primitive pairs, one pair per lane (no collectives inside).  The pairs MuJoCo's collision table sends to
// engine_collision_primitive.c / engine_collision_box.c: sphere-sphere, sphere-capsule, sphere-cylinder, sphere-box, capsule-capsule,
// capsule-box; everything else convex stays on the MPR path.  Every case ends in the sphere-sphere primitive at the closest feature.
// Why not MPR for these as well (round 1 did): MPR reports the depth along the surface normal where the centre-to-centre ray leaves the
// Minkowski difference - off a capsule's end cap that is 15-20 % too deep with the normal tens of degrees off
// (tests/test_primitive_colliders.py) - and its portal resolution on a curved surface (chord sqrt(8 r tol) ~ 0.4 mm) is the size of
// the penetrations themselves.  Mirrors oracle/mgs_oracle.c prim_pair (written separately; same case analysis).
struct PrimHit { real pos[3], n[3], dist; };

MGS_DEVN int prim_sphere_sphere(const real *p1, real r1, const real *p2, real r2, real margin, PrimHit &h) {
  real dif[3];
  sub3(dif, p2, p1);
  const real cd = sqrt(dot3(dif, dif));
  if (cd > margin + r1 + r2) return 0;
  h.dist = cd - r1 - r2;
  if (cd < MGS_MINVAL) { h.n[0] = 1; h.n[1] = 0; h.n[2] = 0; } else scl3(h.n, dif, R_(1.0) / cd);
  copy3(h.pos, p1);
  addscl3(h.pos, h.n, r1 + R_(0.5) * h.dist);
  return 1;
}
MGS_DEV real clampr(real x, real lo, real hi) { return x < lo ? lo : (x > hi ? hi : x); }

// sphere (centre ps, radius rs) against the box with pose (Rb, pb) and half sizes sz; normal from the sphere to the box
MGS_DEVN int prim_sphere_box(const real *ps, real rs, const real *Rb, const real *pb, const real *sz, real margin, PrimHit &h) {
  real t[3], c[3], d[3], nl[3];
  sub3(t, ps, pb);
  mulmatTvec3(c, Rb, t);
  for (int k = 0; k < 3; k++) d[k] = clampr(c[k], -sz[k], sz[k]) - c[k];
  const real dn = sqrt(dot3(d, d));
  if (dn > MGS_MINVAL) {
    if (dn > rs + margin) return 0;
    scl3(nl, d, R_(1.0) / dn);
    h.dist = dn - rs;
  } else {  // centre inside the box: out through the nearest face
    int k = 0;
    real fd = sz[0] - fabs(c[0]);
    if (sz[1] - fabs(c[1]) < fd) { fd = sz[1] - fabs(c[1]); k = 1; }
    if (sz[2] - fabs(c[2]) < fd) { fd = sz[2] - fabs(c[2]); k = 2; }
    const real sg = (k == 0 ? c[0] : (k == 1 ? c[1] : c[2])) >= 0 ? R_(-1.0) : R_(1.0);
    nl[0] = k == 0 ? sg : 0; nl[1] = k == 1 ? sg : 0; nl[2] = k == 2 ? sg : 0;
    h.dist = -(rs + fd);
  }
  addscl3(c, nl, rs + R_(0.5) * h.dist);
  mulmatvec3(h.pos, Rb, c);
  add3(h.pos, h.pos, pb);
  mulmatvec3(h.n, Rb, nl);
  return 1;
}

// d/dt of the squared distance between the box and the point c + t a (box frame): piecewise linear in t
MGS_DEV real seg_box_slope(const real *c, const real *a, const real *sz, real t) {
  real g = 0;
  for (int k = 0; k < 3; k++) {
    const real p = c[k] + t * a[k], ex = fabs(p) - sz[k];
    if (ex > 0) g += 2 * (p > 0 ? ex : -ex) * a[k];
  }
  return g;
}

MGS_DEV int is_prim_pair(int t1, int t2) {
  const int ta = t1 < t2 ? t1 : t2, tb = t1 < t2 ? t2 : t1;
  if (ta == GEOM_SPHERE) return tb == GEOM_SPHERE || tb == GEOM_CAPSULE || tb == GEOM_CYLINDER || tb == GEOM_BOX;
  if (ta == GEOM_CAPSULE) return tb == GEOM_CAPSULE || tb == GEOM_BOX;
  return 0;
}

MGS_DEV void prim_emit(PairContacts &out, const PrimHit &h, real sign) {
  if (out.n == 0) {
    scl3(out.normal, h.n, sign);
    copy3(out.pos[0], h.pos);
    out.dist[0] = h.dist;
    out.n = 1;
  } else if (out.n == 1) {
    scl3(out.normal2, h.n, sign);
    copy3(out.pos[1], h.pos);
    out.dist[1] = h.dist;
    out.n = 2;
    out.own2 = 1;
  }
}

MGS_DEVN void prim_pair(const GeomRef &g1, const GeomRef &g2, real margin, PairContacts &out) {
  // written for type(a) <= type(b); the pair's normal points from g1 to g2
  const int swap = g1.type > g2.type;
  const GeomRef &a = swap ? g2 : g1, &b = swap ? g1 : g2;
  const real sign = swap ? R_(-1.0) : R_(1.0);
  real sa[3], sb[3];
  ld3(sa, MD.cgeom_size + 3 * a.cg);
  ld3(sb, MD.cgeom_size + 3 * b.cg);
  PrimHit h;
  if (a.type == GEOM_SPHERE) {
    if (b.type == GEOM_SPHERE) {
      if (prim_sphere_sphere(a.p, sa[0], b.p, sb[0], margin, h)) prim_emit(out, h, sign);
    } else if (b.type == GEOM_BOX) {
      if (prim_sphere_box(a.p, sa[0], b.R, b.p, sb, margin, h)) prim_emit(out, h, sign);
    } else {  // capsule or cylinder: position along the axis, distance from it
      const real ax[3] = {b.R[2], b.R[5], b.R[8]};
      real v[3], q[3];
      sub3(v, a.p, b.p);
      const real x = dot3(v, ax);
      if (b.type == GEOM_CAPSULE) {
        copy3(q, b.p);
        addscl3(q, ax, clampr(x, -sb[1], sb[1]));
        if (prim_sphere_sphere(a.p, sa[0], q, sb[0], margin, h)) prim_emit(out, h, sign);
      } else {
        const real Rc = sb[0], hc = sb[1];
        addscl3(v, ax, -x);  // v: radial part
        const real rho = sqrt(dot3(v, v));
        int side = fabs(x) < hc, cap = rho < Rc;
        if (side && cap) { if (hc - fabs(x) < Rc - rho) side = 0; else cap = 0; }
        if (side) {
          copy3(q, b.p);
          addscl3(q, ax, x);
          if (prim_sphere_sphere(a.p, sa[0], q, Rc, margin, h)) prim_emit(out, h, sign);
        } else if (cap) {
          h.dist = fabs(x) - hc - sa[0];
          if (h.dist <= margin) {
            scl3(h.n, ax, x >= 0 ? R_(-1.0) : R_(1.0));
            copy3(h.pos, a.p);
            addscl3(h.pos, h.n, sa[0] + R_(0.5) * h.dist);
            prim_emit(out, h, sign);
          }
        } else {
          copy3(q, b.p);
          addscl3(q, ax, x >= 0 ? hc : -hc);
          if (rho > MGS_MINVAL) addscl3(q, v, Rc / rho);
          if (prim_sphere_sphere(a.p, sa[0], q, 0, margin, h)) prim_emit(out, h, sign);
        }
      }
    }
    return;
  }
  // a is a capsule
  const real a1[3] = {a.R[2], a.R[5], a.R[8]};
  if (b.type == GEOM_CAPSULE) {
    const real a2[3] = {b.R[2], b.R[5], b.R[8]};
    real dif[3], q1[3], q2[3];
    sub3(dif, a.p, b.p);
    const real l1 = sa[1], l2 = sb[1], mb = -dot3(a1, a2), u = -dot3(a1, dif), v = dot3(a2, dif), det = 1 - mb * mb;
    if (fabs(det) >= MGS_MINVAL) {
      real x1 = (u - mb * v) / det, x2 = (v - mb * u) / det;
      if (x1 > l1) { x1 = l1; x2 = v - mb * x1; } else if (x1 < -l1) { x1 = -l1; x2 = v - mb * x1; }
      if (x2 > l2) { x2 = l2; x1 = clampr(u - mb * x2, -l1, l1); } else if (x2 < -l2) { x2 = -l2; x1 = clampr(u - mb * x2, -l1, l1); }
      copy3(q1, a.p); addscl3(q1, a1, x1);
      copy3(q2, b.p); addscl3(q2, a2, x2);
      if (prim_sphere_sphere(q1, sa[0], q2, sb[0], margin, h)) prim_emit(out, h, sign);
    } else {  // parallel axes: ends of capsule 1 against segment 2, then ends of capsule 2 against segment 1; two points at most
      #pragma unroll 1
      for (int k = 0; k < 4 && out.n < 2; k++) {
        real x1, x2;
        if (k < 2) { x1 = k ? -l1 : l1; x2 = clampr(v - mb * x1, -l2, l2); }
        else {
          x2 = (k & 1) ? -l2 : l2; x1 = clampr(u - mb * x2, -l1, l1);
          if (fabs(fabs(x1) - l1) < MGS_MINVAL) continue;
        }
        copy3(q1, a.p); addscl3(q1, a1, x1);
        copy3(q2, b.p); addscl3(q2, a2, x2);
        if (prim_sphere_sphere(q1, sa[0], q2, sb[0], margin, h)) prim_emit(out, h, sign);
      }
    }
    return;
  }
  // capsule-box: closest point of the axis segment c + t a, t in [-1, 1] (box frame).  f(t) = squared distance to the box is convex,
  // f' piecewise linear with kinks where a coordinate crosses a face plane: bracket the sign change of f' between two kinks (or
  // segment ends) and interpolate linearly - exact, no iteration.
  real t3[3], c[3], av[3];
  sub3(t3, a.p, b.p);
  mulmatTvec3(c, b.R, t3);
  mulmatTvec3(av, b.R, a1);
  scl3(av, av, sa[1]);
  real tl = -2, tr = 2, gl = 0, gr = 0;
  #pragma unroll 1
  for (int i = 0; i < 8; i++) {
    real t;
    if (i < 2) t = i ? R_(1.0) : R_(-1.0);
    else {
      const int k = (i - 2) >> 1;
      const real ak = k == 0 ? av[0] : (k == 1 ? av[1] : av[2]), ck = k == 0 ? c[0] : (k == 1 ? c[1] : c[2]), sk = k == 0 ? sb[0] : (k == 1 ? sb[1] : sb[2]);
      if (!(fabs(ak) > MGS_MINVAL)) continue;
      t = (((i & 1) ? sk : -sk) - ck) / ak;
      if (!(t > -1 && t < 1)) continue;
    }
    const real g = seg_box_slope(c, av, sb, t);
    if (g <= 0 && t > tl) { tl = t; gl = g; }
    if (g >= 0 && t < tr) { tr = t; gr = g; }
  }
  real ts;
  if (tl < R_(-1.5)) ts = -1; else if (tr > R_(1.5)) ts = 1;
  else if (gr - gl > MGS_MINVAL) ts = tl + (tr - tl) * (-gl) / (gr - gl);
  else ts = R_(0.5) * (tl + tr);
  // candidates: the two ends, then the closest point.  First contact: the deepest (earlier candidate on ties); second: the deepest
  // other candidate in contact that is a different point of the segment.
  PrimHit hc[3];
  int hit[3];
  #pragma unroll
  for (int i = 0; i < 3; i++) {
    const real ti = i == 0 ? R_(-1.0) : (i == 1 ? R_(1.0) : ts);
    real q[3];
    copy3(q, a.p);
    addscl3(q, a1, sa[1] * ti);
    hit[i] = prim_sphere_box(q, sa[0], b.R, b.p, sb, margin, hc[i]);
  }
  int i1 = -1, i2 = -1;
  #pragma unroll
  for (int i = 0; i < 3; i++) if (hit[i] && (i1 < 0 || hc[i].dist < (i1 == 0 ? hc[0].dist : hc[1].dist))) i1 = i;
  if (i1 < 0) return;
  const real t1 = i1 == 0 ? R_(-1.0) : (i1 == 1 ? R_(1.0) : ts);
  #pragma unroll
  for (int i = 0; i < 3; i++) {
    const real ti = i == 0 ? R_(-1.0) : (i == 1 ? R_(1.0) : ts);
    if (i != i1 && hit[i] && fabs(ti - t1) > R_(0.05) && (i2 < 0 || hc[i].dist < (i2 == 0 ? hc[0].dist : hc[1].dist))) i2 = i;
  }
  #pragma unroll
  for (int i = 0; i < 3; i++) if (i == i1) prim_emit(out, hc[i], sign);
  #pragma unroll
  for (int i = 0; i < 3; i++) if (i == i2) prim_emit(out, hc[i], sign);
}
#endif

// narrowphase for candidate pair `pair`, which already passed the bounding-sphere test (pair < 0: lane idle); fills up
// to 4 contacts.  Called by all lanes of the warp together: the MPR stage inside is warp-converged.
MGS_DEVN void collide_pair(Env &e, int pair, PairContacts &out) {
  out.n = 0;
  out.own2 = 0;
  const int active = pair >= 0;
  int c1 = 0, c2 = 0;
  if (active) { c1 = LDG(MD.pair_geom1 + pair); c2 = LDG(MD.pair_geom2 + pair); }
  GeomRef g1, g2;
  geomref_init(g1, e, c1);
  geomref_init(g2, e, c2);
  real depth = 0, n[3], pos[3];
  int *cache = (active && pair < LY.ncache) ? IARR(EF(mpr_cache)) + 4 * pair : (int *)0;
#ifndef MGS_HOST
  // config 5 lists 1,198 pairs (LEAP 1,764), the shared-memory cache holds 128: the others keep their warm start in a global, L2-resident
  // slab of this environment slot (stage clocks of config 5: cold-started MPR was 410 k of the step's 1.9 M cycles - every iteration of
  // the lockstep loop costs two hull supports out of L2, a cold pair needs 20-30 of them where a warm one needs 1-3 - and in fp32 a
  // cold start against the 20 m table box stops at the iteration cap with a noisy depth: lift labels 57/64 -> 64/64 equal to the oracle)
  if (active && pair >= LY.ncache && IO.mpr_cache_g) cache = IO.mpr_cache_g + ((size_t)MGS_ENV_SLOT * MD.npair + pair) * 4;
#endif
#ifdef MGS_NO_MPR_WARMSTART
  cache = (int *)0;
#endif
  // hull supports resume their hill climb from the vertices this pair ended on at the previous step (word 3:
  // two 16-bit vertex ids, 0 after reset); poses change by micrometres per step, so the climb is 0-1 moves
  if (cache && active) { g1.cur = cache[3] & 0xffff; g2.cur = (cache[3] >> 16) & 0xffff; }
  MGS_CLK(1);
  int prim = 0;
#ifndef MGS_NO_ANALYTIC_PRIMS
  // sphere / capsule pairs with a closed form leave the MPR path here (a warp vote keeps the code out of the way of models without
  // such geoms: the three parallel-jaw grippers and LEAP never enter)
  prim = active && is_prim_pair(g1.type, g2.type);
  if (wany(prim)) { if (prim) prim_pair(g1, g2, LDG(MD.pair_margin + pair), out); }
#endif
  int hit = mpr_penetration(g1, g2, active && !prim, cache, &depth, n, pos);
  MGS_CLK(9);
  hit = hit && (depth > 0);
  // polytope pairs: multi-point manifold from the two most-aligned faces.  The face searches are warp-cooperative,
  // so no lane leaves before them (the clipping itself stays per lane).
  // (a cylinder joins through its caps; cap against cap stays a single point)
  const int poly = hit && has_face(g1, n) && has_face(g2, n) && (g1.type != GEOM_CYLINDER || g2.type != GEOM_CYLINDER);
  real a1 = 0, a2 = 0, nn[3] = {-n[0], -n[1], -n[2]};
  int f1 = 0, f2 = 0;
  best_face_lanes(e, poly, g1, n, &f1, &a1);
  best_face_lanes(e, poly, g2, nn, &f2, &a2);
  const int clip = poly && fmax(a1, a2) >= MGS_FACE_ALIGN_MIN;
  const int ref_is_1 = a1 >= a2;
  const GeomRef &rg = ref_is_1 ? g1 : g2;
  const GeomRef &ig = ref_is_1 ? g2 : g1;
  // Clipping works on polygons with run-time indices; they live in per-lane slots of shared memory (the part of the
  // solver overlay that is free during collision), not in local memory (ncu r1_k: the local-memory polygon copies were
  // 4.7 % of all stall samples).  Lanes that clip take slots by rank; if there are more than slots, in rounds.
  int nclipping, rank = wrank(clip, &nclipping), done = 0;
  #pragma unroll 1
  for (int base = 0; base < nclipping; base += LY.nclip) {
    const int mine = clip && rank >= base && rank < base + LY.nclip;
    real *slot = EF(clip_off) + (mine ? rank - base : 0) * MGS_CLIP_STRIDE;
    real (*ref)[3] = reinterpret_cast<real (*)[3]>(slot);
    real (*A)[3] = reinterpret_cast<real (*)[3]>(slot + 3 * MGS_MAXPOLY);
    real (*B)[3] = reinterpret_cast<real (*)[3]>(slot + 3 * MGS_MAXPOLY + 3 * MGS_MAXCLIP);
    real *dist = slot + 3 * MGS_MAXPOLY + 6 * MGS_MAXCLIP;
    real nref[3] = {0, 0, 0}, mn[3], al;
    int nr = 0, incf = 0;
    if (mine) nr = face_polygon(rg, ref_is_1 ? f1 : f2, ref, nref);
    mn[0] = -nref[0]; mn[1] = -nref[1]; mn[2] = -nref[2];
    best_face_lanes(e, mine, ig, mn, &incf, &al);
    if (mine) {
      real ninc[3];
      int na = face_polygon(ig, incf, A, ninc);
      #pragma unroll 1
      for (int ed = 0; ed < nr && na > 0; ed++) {
        real edge[3], sn[3], r0[3];
        int e2 = (ed + 1 == nr) ? 0 : ed + 1;
        copy3(r0, ref[ed]);
        sub3(edge, ref[e2], r0);
        cross3(sn, edge, nref);
        int nb2 = 0;
        real P[3], t0[3];
        copy3(P, A[0]);
        sub3(t0, P, r0);
        real dp = dot3(t0, sn);
        #pragma unroll 1
        for (int i = 0; i < na; i++) {
          real Q[3], t1[3];
          copy3(Q, A[(i + 1 == na) ? 0 : i + 1]);
          sub3(t1, Q, r0);
          const real dq = dot3(t1, sn);
          if (dp <= 0 && nb2 < MGS_MAXCLIP) { copy3(B[nb2], P); nb2++; }
          if ((dp <= 0) != (dq <= 0) && nb2 < MGS_MAXCLIP) {
            real t = dp / (dp - dq);
            for (int k = 0; k < 3; k++) B[nb2][k] = P[k] + t * (Q[k] - P[k]);
            nb2++;
          }
          copy3(P, Q);
          dp = dq;
        }
        na = nb2;
        real (*T)[3] = A; A = B; B = T;  // ping-pong instead of copying back
      }
      int np = 0;
      #pragma unroll 1
      for (int i = 0; i < na; i++) {
        real t[3], a[3];
        copy3(a, A[i]);
        sub3(t, a, ref[0]);
        real dd = dot3(t, nref);
        if (dd < 0) { copy3(A[np], a); dist[np] = dd; np++; }
      }
      if (np > 0) {
        int i0 = 0, i1 = -1, i2 = -1, i3 = -1;
        #pragma unroll 1
        for (int i = 1; i < np; i++) if (dist[i] < dist[i0]) i0 = i;
        if (np > 1) {
          real bd = -1, a0[3];
          copy3(a0, A[i0]);
          #pragma unroll 1
          for (int i = 0; i < np; i++) { real t[3]; sub3(t, A[i], a0); real d2 = dot3(t, t); if (i != i0 && d2 > bd) { bd = d2; i1 = i; } }
          if (i1 >= 0 && bd > R_(1e-12)) {
            real e01[3]; sub3(e01, A[i1], a0);
            real mx = R_(1e-12), mnv = R_(-1e-12);
            #pragma unroll 1
            for (int i = 0; i < np; i++) {
              if (i == i0 || i == i1) continue;
              real t[3], c[3]; sub3(t, A[i], a0); cross3(c, e01, t);
              real sa = dot3(c, nref);
              if (sa > mx) { mx = sa; i2 = i; }
              if (sa < mnv) { mnv = sa; i3 = i; }
            }
          } else i1 = -1;
        }
        if (ref_is_1) copy3(out.normal, nref); else scl3(out.normal, nref, -1);
        int ns = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int s = k == 0 ? i0 : (k == 1 ? i1 : (k == 2 ? i2 : i3));
          if (s >= 0) {
            const real ds = dist[s];
            real a[3];
            copy3(a, A[s]);
            addscl3(a, nref, R_(-0.5) * ds);
            // out.pos / out.dist are indexed with a run-time count: write through the unrolled slot index
#pragma unroll
            for (int q = 0; q < 4; q++) if (q == ns) { copy3(out.pos[q], a); out.dist[q] = ds; }
            ns++;
          }
        }
        out.n = ns;
        done = 1;
      }
    }
    WSYNC();
  }
  MGS_CLK(10);
  if (!hit || done) return;
  out.n = 1;
  copy3(out.normal, n);
  copy3(out.pos[0], pos);
  out.dist[0] = -depth;
}

MGS_DEV void make_frame(real *frame) {
  real *x = frame, *y = frame + 3, *z = frame + 6;
  if (fabs(x[1]) < R_(0.5)) { y[0] = 0; y[1] = 1; y[2] = 0; } else { y[0] = 0; y[1] = 0; y[2] = 1; }
  real d = dot3(x, y);
  addscl3(y, x, -d);
  normalize3(y);
  cross3(z, x, y);
}

// store this lane's contacts of pair p at the next free slots (pair order = deterministic contact order)
MGS_DEV void emit_contacts(Env &e, const PairContacts &pc, int p, int &base) {
  int total, off = wscan_excl(pc.n, &total);
  #pragma unroll 1
  for (int k = 0; k < pc.n; k++) {
    int c = base + off + k;
    if (c >= LY.ncon_max) break;
    copy3(EF(con_pos) + 3 * c, pc.pos[k]);
    copy3(EF(con_normal) + 3 * c, (k == 1 && pc.own2) ? pc.normal2 : pc.normal);
    EF(con_dist)[c] = pc.dist[k];
    IARR(EF(con_pair))[c] = p;
  }
  base += total;
}

// Broadphase + narrowphase.  Pass 1 runs the bounding-sphere test one pair per lane and COMPACTS the survivors
// (in pair order) into a 32-entry buffer; whenever the buffer cannot take the next group it is flushed through the
// narrowphase.  The MPR state machine therefore runs once per 32 surviving pairs with all their lanes busy, instead
// of once per 32 listed pairs with the few survivors of that group (Panda 63 pairs, Robotiq 95, LEAP 1764).
MGS_DEVN void collision_w(Env &e) {
  int base = 0, nbuf = 0;
  int *buf = IARR(EF(nsS));
  #pragma unroll 1
  for (int p0 = 0; p0 < MD.npair; p0 += LANES) {
    const int p = p0 + MGS_LANE;
    int active = 0;
    if (p < MD.npair) {
      const int c1 = LDG(MD.pair_geom1 + p), c2 = LDG(MD.pair_geom2 + p);
      real dc[3];
      sub3(dc, EF(gxpos) + 3 * c1, EF(gxpos) + 3 * c2);
      const real rr = LDG(MD.cgeom_rbound + c1) + LDG(MD.cgeom_rbound + c2) + LDG(MD.pair_margin + p);
      active = dot3(dc, dc) <= rr * rr;
      if (!active && p < LY.ncache) IARR(EF(mpr_cache))[4 * p + 2] = -1;  // a culled pair forgets its portal / axis
#ifndef MGS_HOST
      if (!active && p >= LY.ncache && IO.mpr_cache_g) {
        int *cg = IO.mpr_cache_g + ((size_t)MGS_ENV_SLOT * MD.npair + p) * 4;
        if (cg[2] != -1) cg[2] = -1;
      }
#endif
    }
    int k, rank = wrank(active, &k);
    if (k == 0) continue;
    if (nbuf + k > LANES) {
      PairContacts pc;
      const int q = MGS_LANE < nbuf ? buf[MGS_LANE] : -1;
      collide_pair(e, q, pc);
      emit_contacts(e, pc, q, base);
      nbuf = 0;
      WSYNC();
    }
    if (active) buf[nbuf + rank] = p;
    nbuf += k;
    WSYNC();
  }
  if (nbuf > 0) {
    PairContacts pc;
    const int q = MGS_LANE < nbuf ? buf[MGS_LANE] : -1;
    collide_pair(e, q, pc);
    emit_contacts(e, pc, q, base);
  }
  if (base > LY.ncon_max) { if (MGS_LANE == 0) EH.overflow += base - LY.ncon_max; base = LY.ncon_max; }
  EH.ncon = base;
  WSYNC();
}

// gripper <-> object contact test (reference check_contact_with_object,
// gravityless_object_grasping.py:309-321): a contact whose geom ids straddle the ground geom's id
// include_ground != 0: also count gripper <-> ground/table contacts (check_gripper_collision of the clutter
// table, clutter_table.py:237-252)
MGS_DEVN int contact_with_object_w(const Env &e, int include_ground) {
  int hit = 0, g = MD.ground_geomid;
  #pragma unroll 1
  PFOR(c, EH.ncon) {
    int p = IARR(EF(con_pair))[c];
    int a = LDG(MD.cgeom_geomid + LDG(MD.pair_geom1 + p)), b = LDG(MD.cgeom_geomid + LDG(MD.pair_geom2 + p));
    if ((a < g && b > g) || (a > g && b < g)) hit = 1;
    if (include_ground && ((a == g && b < g) || (a < g && b == g))) hit = 1;
  }
  return wany(hit);
}
