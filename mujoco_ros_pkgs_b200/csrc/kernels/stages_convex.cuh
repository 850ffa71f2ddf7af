// stages_convex.cuh — narrowphase for cylinders, ellipsoids and convex meshes, one candidate pair per lane.
//
// Completes the geom-type table of mj_collision (row M5 of SURVEY 8a; MuJoCo's mjCOLLISIONFUNC, the table the
// reference lets plugins override at mujoco_ros/src/mujoco_env.cpp:163-176): plane-cylinder (analytic, up to four
// points), plane-convex (support point; meshes add hull vertices inside the margin band), and every other convex pair
// by Minkowski Portal Refinement on the support mappings -- the algorithm of libccd's ccdMPRPenetration, which MuJoCo's
// mjc_Convex calls with opt.mpr_iterations / opt.mpr_tolerance.  The portal lives in registers / local memory of the
// lane that owns the pair; nothing is shared between lanes, so divergent iteration counts only cost lane time.
#pragma once
#include <float.h>

#include "env_ctx.cuh"
#include "pair_con.cuh"

namespace b2k {

struct CvxGeom {
  int type, nvert;
  const double *pos, *mat, *size, *vert;
};

__device__ __forceinline__ double sgnd(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

__device__ __forceinline__ CvxGeom cvxGeom(const double* gxpos, const double* gxmat, int g) {
  const DevModel& m = c_dm;
  CvxGeom c;
  c.type = m.geom_type[g];
  c.pos = gxpos + 3 * g;
  c.mat = gxmat + 9 * g;
  c.size = m.geom_size + 3 * g;
  c.vert = nullptr;
  c.nvert = 0;
  if (c.type == B2MJ_GEOM_MESH) {
    const int id = m.geom_dataid[g];
    if (id >= 0) { c.vert = m.mesh_vert + 3 * m.mesh_vertadr[id]; c.nvert = m.mesh_vertnum[id]; }
  }
  return c;
}

// support point of a convex geom along a world direction (mjccd_support without inflation)
__device__ __noinline__ void cvxSupport(const CvxGeom& g, const double* dir, double* res) {
  double ld[3], lr[3] = {0, 0, 0};
  rotVecMatT(ld, dir, g.mat);
  if (g.type == B2MJ_GEOM_SPHERE) {
    scl3(lr, ld, g.size[0]);
  } else if (g.type == B2MJ_GEOM_CAPSULE) {
    scl3(lr, ld, g.size[0]);
    lr[2] += sgnd(ld[2]) * g.size[1];
  } else if (g.type == B2MJ_GEOM_ELLIPSOID) {
    double t[3] = {ld[0] * g.size[0], ld[1] * g.size[1], ld[2] * g.size[2]};
    normalize3(t);
    lr[0] = t[0] * g.size[0]; lr[1] = t[1] * g.size[1]; lr[2] = t[2] * g.size[2];
  } else if (g.type == B2MJ_GEOM_CYLINDER) {
    const double t = sqrt(ld[0] * ld[0] + ld[1] * ld[1]);
    if (t > B2K_MINVAL) { lr[0] = ld[0] / t * g.size[0]; lr[1] = ld[1] / t * g.size[0]; }
    lr[2] = sgnd(ld[2]) * g.size[1];
  } else if (g.type == B2MJ_GEOM_BOX) {
    lr[0] = sgnd(ld[0]) * g.size[0]; lr[1] = sgnd(ld[1]) * g.size[1]; lr[2] = sgnd(ld[2]) * g.size[2];
  } else if (g.type == B2MJ_GEOM_MESH) {
    double best = -DBL_MAX;
    int arg = 0;
    B2K_NOUNROLL for (int v = 0; v < g.nvert; v++) {
      const double s = dot3(g.vert + 3 * v, ld);
      if (s > best) { best = s; arg = v; }
    }
    if (g.nvert) copy3(lr, g.vert + 3 * arg);
  }
  rotVecMat(res, lr, g.mat);
  addTo3(res, g.pos);
}

__device__ __forceinline__ void cvxEmitPlane(PairCon& o, int n, const double* normal, double dist, const double* p) {
  o.dist[n] = dist;
  copy3(o.pos + 3 * n, p);
  addToScl3(o.pos + 3 * n, normal, -dist * 0.5);
}

// plane vs cylinder: lowest rim point of each cap plus two more points of the near cap at +-120 degrees
__device__ __noinline__ int c_planeCylinder(PairCon& o, double margin, const double* pos1, const double* mat1, const double* pos2,
                                            const double* mat2, const double* size2) {
  double normal[3] = {mat1[2], mat1[5], mat1[8]}, axis[3] = {mat2[2], mat2[5], mat2[8]}, vec[3], tmp[3], p[3];
  double prjaxis = dot3(normal, axis);
  if (prjaxis > 0) { scl3(axis, axis, -1); prjaxis = -prjaxis; }
  sub3(tmp, pos2, pos1);
  const double dist0 = dot3(tmp, normal);
  scl3(vec, axis, prjaxis);
  sub3(vec, vec, normal);
  const double len2 = dot3(vec, vec);
  if (len2 >= B2K_MINVAL * B2K_MINVAL) scl3(vec, vec, size2[0] / sqrt(len2));
  else { vec[0] = mat2[0] * size2[0]; vec[1] = mat2[3] * size2[0]; vec[2] = mat2[6] * size2[0]; }
  const double prjvec = dot3(vec, normal);
  scl3(axis, axis, size2[1]);
  prjaxis *= size2[1];
  if (dist0 + prjaxis + prjvec > margin) return 0;
  o.shared_frame = 1;
  copy3(o.frame, normal);
  zero3(o.frame + 3);
  int cnt = 0;
  add3(p, pos2, vec); addTo3(p, axis);
  cvxEmitPlane(o, cnt++, normal, dist0 + prjaxis + prjvec, p);
  if (dist0 - prjaxis + prjvec <= margin) {
    add3(p, pos2, vec); sub3(p, p, axis);
    cvxEmitPlane(o, cnt++, normal, dist0 - prjaxis + prjvec, p);
  }
  const double prjvec1 = -prjvec * 0.5;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    double vec1[3];
    cross(vec1, vec, axis);
    normalize3(vec1);
    scl3(vec1, vec1, size2[0] * sqrt(3.0) / 2);
    for (int s = 1; s >= -1; s -= 2) {
      add3(p, pos2, axis);
      addToScl3(p, vec1, (double)s);
      addToScl3(p, vec, -0.5);
      cvxEmitPlane(o, cnt++, normal, dist0 + prjaxis + prjvec1, p);
    }
  }
  return cnt;
}

// plane vs convex: support point against the normal; meshes add up to three more hull vertices inside the margin band
__device__ __noinline__ int c_planeConvex(PairCon& o, double margin, const double* pos1, const double* mat1, const CvxGeom& g) {
  const double normal[3] = {mat1[2], mat1[5], mat1[8]};
  double dir[3] = {-normal[0], -normal[1], -normal[2]}, sp[3], tmp[3];
  cvxSupport(g, dir, sp);
  sub3(tmp, sp, pos1);
  const double dist = dot3(tmp, normal);
  if (dist > margin) return 0;
  o.shared_frame = 1;
  copy3(o.frame, normal);
  zero3(o.frame + 3);
  int cnt = 0;
  cvxEmitPlane(o, cnt++, normal, dist, sp);
  if (g.type == B2MJ_GEOM_MESH) {
    B2K_NOUNROLL for (int v = 0; v < g.nvert && cnt < 4; v++) {
      double w[3];
      rotVecMat(w, g.vert + 3 * v, g.mat);
      addTo3(w, g.pos);
      sub3(tmp, w, sp);
      if (dot3(tmp, tmp) < 1e-20) continue;
      sub3(tmp, w, pos1);
      const double dv = dot3(tmp, normal);
      if (dv <= margin && dv <= 0) cvxEmitPlane(o, cnt++, normal, dv, w);
    }
  }
  return cnt;
}

// ---- Minkowski Portal Refinement ----
struct MprSupp {
  double v[3], v1[3], v2[3];
};

__device__ __forceinline__ bool mprZero(double x) { return fabs(x) < DBL_EPSILON; }
__device__ __forceinline__ bool mprEq(double a, double b) {
  const double ab = fabs(a - b);
  if (ab < DBL_EPSILON) return true;
  return ab < DBL_EPSILON * fmax(fabs(a), fabs(b));
}

__device__ __noinline__ void mprSupport(const CvxGeom& a, const CvxGeom& b, double inflate, const double* dir, MprSupp& s) {
  double nd[3] = {-dir[0], -dir[1], -dir[2]};
  cvxSupport(a, dir, s.v1);
  addToScl3(s.v1, dir, inflate);
  cvxSupport(b, nd, s.v2);
  addToScl3(s.v2, nd, inflate);
  sub3(s.v, s.v1, s.v2);
}

__device__ __forceinline__ void mprPortalDir(const MprSupp* p, double* dir) {
  double a[3], b[3];
  sub3(a, p[2].v, p[1].v);
  sub3(b, p[3].v, p[1].v);
  cross(dir, a, b);
  normalize3(dir);
}

__device__ __forceinline__ bool mprReachTol(const MprSupp* p, const MprSupp& v4, const double* dir, double tol) {
  const double dv4 = dot3(v4.v, dir);
  double d = dv4 - dot3(p[1].v, dir);
  d = fmin(d, dv4 - dot3(p[2].v, dir));
  d = fmin(d, dv4 - dot3(p[3].v, dir));
  return mprEq(d, tol) || d < tol;
}

__device__ __forceinline__ void mprExpand(MprSupp* p, const MprSupp& v4) {
  double c[3];
  cross(c, v4.v, p[0].v);
  int slot;
  if (dot3(p[1].v, c) > 0) slot = dot3(p[2].v, c) > 0 ? 1 : 3;
  else slot = dot3(p[3].v, c) > 0 ? 2 : 1;
  p[slot] = v4;
}

__device__ __forceinline__ double mprOriginSeg2(const double* a, const double* b, double* w) {
  double d[3];
  sub3(d, b, a);
  const double t = -dot3(a, d) / dot3(d, d);
  if (t < 0 || mprZero(t)) { copy3(w, a); return dot3(a, a); }
  if (t > 1 || mprEq(t, 1)) { copy3(w, b); return dot3(b, b); }
  w[0] = a[0] + t * d[0]; w[1] = a[1] + t * d[1]; w[2] = a[2] + t * d[2];
  return dot3(w, w);
}

__device__ __noinline__ double mprOriginTri2(const double* x0, const double* B, const double* C, double* w) {
  double d1[3], d2[3];
  sub3(d1, B, x0);
  sub3(d2, C, x0);
  const double v = dot3(d1, d1), ww = dot3(d2, d2), p = dot3(x0, d1), q = dot3(x0, d2), r = dot3(d1, d2);
  const double s = (q * r - ww * p) / (ww * v - r * r), t = (-s * r - q) / ww;
  if ((mprZero(s) || s > 0) && (mprEq(s, 1) || s < 1) && (mprZero(t) || t > 0) && (mprEq(t, 1) || t < 1) &&
      (mprEq(t + s, 1) || t + s < 1)) {
    for (int i = 0; i < 3; i++) w[i] = x0[i] + s * d1[i] + t * d2[i];
    return dot3(w, w);
  }
  double best = mprOriginSeg2(x0, B, w), w2[3];
  double d = mprOriginSeg2(x0, C, w2);
  if (d < best) { best = d; copy3(w, w2); }
  d = mprOriginSeg2(B, C, w2);
  if (d < best) { best = d; copy3(w, w2); }
  return best;
}

__device__ __noinline__ void mprFindPos(const MprSupp* p, double* pos) {
  double dir[3], vec[3], b[4];
  mprPortalDir(p, dir);
  cross(vec, p[1].v, p[2].v); b[0] = dot3(vec, p[3].v);
  cross(vec, p[3].v, p[2].v); b[1] = dot3(vec, p[0].v);
  cross(vec, p[0].v, p[1].v); b[2] = dot3(vec, p[3].v);
  cross(vec, p[2].v, p[1].v); b[3] = dot3(vec, p[0].v);
  double sum = b[0] + b[1] + b[2] + b[3];
  if (mprZero(sum) || sum < 0) {
    b[0] = 0;
    cross(vec, p[2].v, p[3].v); b[1] = dot3(vec, dir);
    cross(vec, p[3].v, p[1].v); b[2] = dot3(vec, dir);
    cross(vec, p[1].v, p[2].v); b[3] = dot3(vec, dir);
    sum = b[1] + b[2] + b[3];
  }
  const double inv = 1 / sum;
  double p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  B2K_NOUNROLL for (int i = 0; i < 4; i++) {
    addToScl3(p1, p[i].v1, b[i]);
    addToScl3(p2, p[i].v2, b[i]);
  }
  for (int k = 0; k < 3; k++) pos[k] = 0.5 * inv * (p1[k] + p2[k]);
}

// one contact from the penetration query of the two geoms inflated by margin / 2 each (mjc_Convex)
__device__ __noinline__ int c_convexConvexAt(PairCon& o, int n, double margin, const CvxGeom& ga, const CvxGeom& gb) {
  const DevModel& m = c_dm;
  const double inflate = 0.5 * margin, tol = m.opt.mpr_tolerance;
  const int maxit = m.opt.mpr_iterations;
  MprSupp p[4], v4;
  double d[3], va[3], vb[3], depth, dir[3], pos[3];
  bool done = false;
  // discoverPortal
  sub3(p[0].v, ga.pos, gb.pos);
  copy3(p[0].v1, ga.pos);
  copy3(p[0].v2, gb.pos);
  if (mprEq(p[0].v[0], 0) && mprEq(p[0].v[1], 0) && mprEq(p[0].v[2], 0)) p[0].v[0] += DBL_EPSILON * 10;
  scl3(d, p[0].v, -1);
  normalize3(d);
  mprSupport(ga, gb, inflate, d, p[1]);
  double dt = dot3(p[1].v, d);
  if (mprZero(dt) || dt < 0) return 0;
  cross(d, p[0].v, p[1].v);
  if (mprZero(dot3(d, d))) {
    if (mprEq(p[1].v[0], 0) && mprEq(p[1].v[1], 0) && mprEq(p[1].v[2], 0)) {
      depth = 0;
      zero3(dir);
    } else {
      depth = norm3(p[1].v);
      copy3(dir, p[1].v);
      normalize3(dir);
    }
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p[1].v1[k] + p[1].v2[k]);
    done = true;
  }
  if (!done) {
    normalize3(d);
    mprSupport(ga, gb, inflate, d, p[2]);
    dt = dot3(p[2].v, d);
    if (mprZero(dt) || dt < 0) return 0;
    sub3(va, p[1].v, p[0].v);
    sub3(vb, p[2].v, p[0].v);
    cross(d, va, vb);
    normalize3(d);
    if (dot3(d, p[0].v) > 0) {
      const MprSupp t = p[1]; p[1] = p[2]; p[2] = t;
      scl3(d, d, -1);
    }
    B2K_NOUNROLL for (int guard = 0;; guard++) {
      if (guard > 1000) return 0;
      mprSupport(ga, gb, inflate, d, p[3]);
      dt = dot3(p[3].v, d);
      if (mprZero(dt) || dt < 0) return 0;
      bool cont = false;
      cross(va, p[1].v, p[3].v);
      dt = dot3(va, p[0].v);
      if (dt < 0 && !mprZero(dt)) { p[2] = p[3]; cont = true; }
      if (!cont) {
        cross(va, p[3].v, p[2].v);
        dt = dot3(va, p[0].v);
        if (dt < 0 && !mprZero(dt)) { p[1] = p[3]; cont = true; }
      }
      if (!cont) break;
      sub3(va, p[1].v, p[0].v);
      sub3(vb, p[2].v, p[0].v);
      cross(d, va, vb);
      normalize3(d);
    }
    // refinePortal
    B2K_NOUNROLL for (int guard = 0;; guard++) {
      if (guard > 1000) return 0;
      mprPortalDir(p, d);
      dt = dot3(d, p[1].v);
      if (mprZero(dt) || dt > 0) break;
      mprSupport(ga, gb, inflate, d, v4);
      dt = dot3(v4.v, d);
      if (!(mprZero(dt) || dt > 0) || mprReachTol(p, v4, d, tol)) return 0;
      mprExpand(p, v4);
    }
    // findPenetr
    B2K_NOUNROLL for (int it = 0;; it++) {
      mprPortalDir(p, d);
      mprSupport(ga, gb, inflate, d, v4);
      if (mprReachTol(p, v4, d, tol) || it > maxit) {
        double w[3];
        depth = sqrt(mprOriginTri2(p[1].v, p[2].v, p[3].v, w));
        if (mprZero(w[0]) && mprZero(w[1]) && mprZero(w[2])) copy3(w, d);
        copy3(dir, w);
        normalize3(dir);
        mprFindPos(p, pos);
        break;
      }
      mprExpand(p, v4);
    }
  }
  if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) return 0;
  o.dist[n] = margin - depth;
  copy3(o.frame + 6 * n, dir);
  zero3(o.frame + 6 * n + 3);
  copy3(o.pos + 3 * n, pos);
  return 1;
}
__device__ __forceinline__ int c_convexConvex(PairCon& o, double margin, const CvxGeom& ga, const CvxGeom& gb) {
  return c_convexConvexAt(o, 0, margin, ga, gb);
}

// height field (geom 1) vs convex geom (mjc_ConvexHField): the geom's bounding box in the hfield frame selects a
// sub-grid; every grid triangle under it, extruded down to -base, is a six-vertex prism tested with the MPR query -- one
// contact per penetrating prism, up to the B2K_MAXPAIRCON a lane holds (MuJoCo: mjMAXCONPAIR = 50).  The prism is handed
// to the support mapping as a six-vertex hull around its own centroid (the MPR's interior point).  All prisms of a pair
// are walked by the lane that owns the pair.
__device__ __noinline__ int c_hfieldConvex(PairCon& o, double margin, const double* pos1, const double* mat1, int hid,
                                           const CvxGeom& gb, double rbound2) {
  const DevModel& m = c_dm;
  const double* hsize = m.hfield_size + 4 * hid;
  const int nrow = m.hfield_nrow[hid], ncol = m.hfield_ncol[hid];
  const double* data = m.hfield_data + m.hfield_adr[hid];
  double dif[3], pos[3];
  sub3(dif, gb.pos, pos1);
  rotVecMatT(pos, dif, mat1);
  const double r2 = rbound2 + margin;
  if (pos[0] > hsize[0] + r2 || pos[0] < -hsize[0] - r2 || pos[1] > hsize[1] + r2 || pos[1] < -hsize[1] - r2 ||
      pos[2] > hsize[2] + r2 || pos[2] < -hsize[3] - r2)
    return 0;
  double xmin[3], xmax[3];
  B2K_NOUNROLL for (int i = 0; i < 3; i++) {
    double dirw[3], sp[3], loc[3];
    const double ax[3] = {mat1[i], mat1[3 + i], mat1[6 + i]};
    copy3(dirw, ax);
    cvxSupport(gb, dirw, sp);
    sub3(dif, sp, pos1);
    rotVecMatT(loc, dif, mat1);
    xmax[i] = loc[i] + margin;
    scl3(dirw, ax, -1.0);
    cvxSupport(gb, dirw, sp);
    sub3(dif, sp, pos1);
    rotVecMatT(loc, dif, mat1);
    xmin[i] = loc[i] - margin;
  }
  if (xmin[0] > hsize[0] || xmax[0] < -hsize[0] || xmin[1] > hsize[1] || xmax[1] < -hsize[1] || xmin[2] > hsize[2] ||
      xmax[2] < -hsize[3])
    return 0;
  int cmin = (int)floor((xmin[0] + hsize[0]) / (2 * hsize[0]) * (ncol - 1));
  int cmax = (int)ceil((xmax[0] + hsize[0]) / (2 * hsize[0]) * (ncol - 1));
  int rmin = (int)floor((xmin[1] + hsize[1]) / (2 * hsize[1]) * (nrow - 1));
  int rmax = (int)ceil((xmax[1] + hsize[1]) / (2 * hsize[1]) * (nrow - 1));
  cmin = max(cmin, 0);
  rmin = max(rmin, 0);
  cmax = min(cmax, ncol - 1);
  rmax = min(rmax, nrow - 1);
  const double dx = 2 * hsize[0] / (ncol - 1), dy = 2 * hsize[1] / (nrow - 1);
  double prism[18], vrel[18], cl[3], cw[3];
  B2K_NOUNROLL for (int k = 0; k < 18; k++) prism[k] = 0;
  CvxGeom ga;
  ga.type = B2MJ_GEOM_MESH;
  ga.nvert = 6;
  ga.pos = cw;
  ga.mat = mat1;
  ga.size = hsize;  // unused by the hull support
  ga.vert = vrel;
  int cnt = 0;
  B2K_NOUNROLL for (int r = rmin; r < rmax; r++) {
    int nvert = 0;
    B2K_NOUNROLL for (int c = cmin; c <= cmax; c++) {
      B2K_NOUNROLL for (int i = 0; i < 2; i++) {
        const int rr = r + (i == 0 ? 1 : 0);
        B2K_NOUNROLL for (int k = 0; k < 3; k++) {
          prism[k] = prism[3 + k]; prism[3 + k] = prism[6 + k];
          prism[9 + k] = prism[12 + k]; prism[12 + k] = prism[15 + k];
        }
        prism[6] = prism[15] = dx * c - hsize[0];
        prism[7] = prism[16] = dy * rr - hsize[1];
        prism[8] = -hsize[3];
        prism[17] = data[rr * ncol + c] * hsize[2];
        if (++nvert <= 2) continue;
        if (prism[11] < xmin[2] && prism[14] < xmin[2] && prism[17] < xmin[2]) continue;
        B2K_NOUNROLL for (int k = 0; k < 3; k++) {
          double s = 0;
          B2K_NOUNROLL for (int v = 0; v < 6; v++) s += prism[3 * v + k];
          s /= 6;
          cl[k] = s;
          B2K_NOUNROLL for (int v = 0; v < 6; v++) vrel[3 * v + k] = prism[3 * v + k] - s;
        }
        rotVecMat(cw, cl, mat1);
        addTo3(cw, pos1);
        if (c_convexConvexAt(o, cnt, margin, ga, gb) == 1) {
          if (++cnt >= B2K_MAXPAIRCON) return cnt;
        }
      }
    }
  }
  return cnt;
}

}  // namespace b2k
