// orc_convex.cpp — convex narrowphase of the CPU oracle: plane-cylinder, plane-convex (ellipsoid / mesh) and the
// general convex-convex test by Minkowski Portal Refinement for every pair without a dedicated primitive function.
// TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// Restates MuJoCo 2.3.7 engine_collision_primitive.c::mjc_PlaneCylinder, engine_collision_convex.c::mjc_PlaneConvex /
// mjc_Convex with its support functions, and the MPR of libccd 2.x (mpr.c: discoverPortal, refinePortal, findPenetr,
// findPos) that mjc_Convex calls -- libccd is a third-party dependency of MuJoCo, itself an un-vendored binary
// dependency of the reference (mujoco_ros/CMakeLists.txt:61); neither source is in /root/reference.  The reference
// reaches this code through mj_step (mujoco_env.cpp:498) and lets plugins override table entries
// (registerCollisionFunction, mujoco_env.cpp:163-176).  "parity unpinned".
#include <cfloat>
#include <cmath>

#include "orc_convex.h"
#include "orc_math.h"

namespace orc {

static inline double sgn(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

// ---- plane vs cylinder: lowest rim point of each cap, plus two more points of the near cap at +-120 degrees ----
int planeCylinder(Con* con, double margin, const double* pos1, const double* mat1, const double* pos2, const double* mat2,
                  const double* size2) {
  double normal[3] = {mat1[2], mat1[5], mat1[8]}, axis[3] = {mat2[2], mat2[5], mat2[8]}, vec[3], tmp[3];
  double prjaxis = dot3(normal, axis);
  if (prjaxis > 0) { scl3(axis, axis, -1); prjaxis = -prjaxis; }  // axis points towards the plane
  sub3(tmp, pos2, pos1);
  const double dist0 = dot3(tmp, normal);
  // rim direction: the part of -normal perpendicular to the axis, scaled to the radius
  scl3(vec, axis, prjaxis);
  sub3(vec, vec, normal);
  const double len2 = dot3(vec, vec);
  if (len2 >= MINVAL * MINVAL) scl3(vec, vec, size2[0] / std::sqrt(len2));
  else { vec[0] = mat2[0] * size2[0]; vec[1] = mat2[3] * size2[0]; vec[2] = mat2[6] * size2[0]; }  // disk parallel to plane
  const double prjvec = dot3(vec, normal);
  scl3(axis, axis, size2[1]);
  prjaxis *= size2[1];
  int cnt = 0;
  auto emit = [&](double dist, const double* p) {
    con[cnt].dist = dist;
    zero(con[cnt].frame, 9);
    copy3(con[cnt].frame, normal);
    copy3(con[cnt].pos, p);
    addToScl3(con[cnt].pos, normal, -dist * 0.5);
    cnt++;
  };
  double p[3];
  if (dist0 + prjaxis + prjvec > margin) return 0;
  add3(p, pos2, vec); addTo3(p, axis);
  emit(dist0 + prjaxis + prjvec, p);
  if (dist0 - prjaxis + prjvec <= margin) {
    add3(p, pos2, vec); sub3(p, p, axis);
    emit(dist0 - prjaxis + prjvec, p);
  }
  const double prjvec1 = -prjvec * 0.5;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    double vec1[3];
    cross(vec1, vec, axis);
    normalize3(vec1);
    scl3(vec1, vec1, size2[0] * std::sqrt(3.0) / 2);
    for (int s = 1; s >= -1; s -= 2) {
      add3(p, pos2, axis);
      addToScl3(p, vec1, (double)s);
      addToScl3(p, vec, -0.5);
      emit(dist0 + prjaxis + prjvec1, p);
    }
  }
  return cnt;
}

// ---- support mapping of a convex geom in world coordinates (mjccd_support without the margin inflation) ----
void convexSupport(const ConvexGeom& g, const double* dir, double* res) {
  double ld[3], lr[3] = {0, 0, 0};
  rotVecMatT(ld, dir, g.mat);
  switch (g.type) {
    case B2MJ_GEOM_SPHERE: scl3(lr, ld, g.size[0]); break;
    case B2MJ_GEOM_CAPSULE:
      scl3(lr, ld, g.size[0]);
      lr[2] += sgn(ld[2]) * g.size[1];
      break;
    case B2MJ_GEOM_ELLIPSOID: {
      double t[3] = {ld[0] * g.size[0], ld[1] * g.size[1], ld[2] * g.size[2]};
      normalize3(t);
      lr[0] = t[0] * g.size[0]; lr[1] = t[1] * g.size[1]; lr[2] = t[2] * g.size[2];
      break;
    }
    case B2MJ_GEOM_CYLINDER: {
      const double t = std::sqrt(ld[0] * ld[0] + ld[1] * ld[1]);
      if (t > MINVAL) { lr[0] = ld[0] / t * g.size[0]; lr[1] = ld[1] / t * g.size[0]; }
      lr[2] = sgn(ld[2]) * g.size[1];
      break;
    }
    case B2MJ_GEOM_BOX:
      for (int i = 0; i < 3; i++) lr[i] = sgn(ld[i]) * g.size[i];
      break;
    case B2MJ_GEOM_MESH: {  // exhaustive search over the hull vertices, first maximum wins
      double best = -DBL_MAX;
      int arg = 0;
      for (int v = 0; v < g.nvert; v++) {
        const double s = dot3(g.vert + 3 * v, ld);
        if (s > best) { best = s; arg = v; }
      }
      if (g.nvert) copy3(lr, g.vert + 3 * arg);
      break;
    }
    default: break;
  }
  rotVecMat(res, lr, g.mat);
  addTo3(res, g.pos);
}

// ---- plane vs convex (mjc_PlaneConvex): the support point against the plane normal; meshes add up to three more
// hull vertices that lie within the margin band, visited in vertex order ----
int planeConvex(Con* con, double margin, const double* pos1, const double* mat1, const ConvexGeom& g) {
  const double normal[3] = {mat1[2], mat1[5], mat1[8]};
  double dir[3] = {-normal[0], -normal[1], -normal[2]}, sp[3], tmp[3];
  convexSupport(g, dir, sp);
  sub3(tmp, sp, pos1);
  const double dist = dot3(tmp, normal);
  if (dist > margin) return 0;
  int cnt = 0;
  auto emit = [&](double dd, const double* p) {
    con[cnt].dist = dd;
    zero(con[cnt].frame, 9);
    copy3(con[cnt].frame, normal);
    copy3(con[cnt].pos, p);
    addToScl3(con[cnt].pos, normal, -dd * 0.5);
    cnt++;
  };
  emit(dist, sp);
  if (g.type == B2MJ_GEOM_MESH) {
    for (int v = 0; v < g.nvert && cnt < 4; v++) {
      double w[3];
      rotVecMat(w, g.vert + 3 * v, g.mat);
      addTo3(w, g.pos);
      sub3(tmp, w, sp);
      if (dot3(tmp, tmp) < 1e-20) continue;  // the support vertex itself
      sub3(tmp, w, pos1);
      const double dv = dot3(tmp, normal);
      if (dv <= margin && dv <= 0) emit(dv, w);
    }
  }
  return cnt;
}

// ------------------------------------------------------------------------------------------------
// Minkowski Portal Refinement (libccd mpr.c)
// ------------------------------------------------------------------------------------------------
namespace {

const double kEps = DBL_EPSILON;
inline bool isZero(double x) { return std::fabs(x) < kEps; }
inline bool eq(double a, double b) {
  const double ab = std::fabs(a - b);
  if (ab < kEps) return true;
  const double aa = std::fabs(a), bb = std::fabs(b);
  return ab < kEps * (bb > aa ? bb : aa);
}
inline bool vecEq(const double* a, const double* b) { return eq(a[0], b[0]) && eq(a[1], b[1]) && eq(a[2], b[2]); }

struct Supp {
  double v[3], v1[3], v2[3];  // point of the Minkowski difference and its witnesses on the two geoms
};

struct Pair {
  const ConvexGeom *a, *b;
  double inflate;  // each geom is inflated by this much along the query direction (half of the contact margin)
};

void support(const Pair& P, const double* dir, Supp* s) {
  double nd[3] = {-dir[0], -dir[1], -dir[2]};
  convexSupport(*P.a, dir, s->v1);
  addToScl3(s->v1, dir, P.inflate);
  convexSupport(*P.b, nd, s->v2);
  addToScl3(s->v2, nd, P.inflate);
  sub3(s->v, s->v1, s->v2);
}

void portalDir(const Supp* p, double* dir) {
  double a[3], b[3];
  sub3(a, p[2].v, p[1].v);
  sub3(b, p[3].v, p[1].v);
  cross(dir, a, b);
  normalize3(dir);
}

bool portalReachTolerance(const Supp* p, const Supp& v4, const double* dir, double tol) {
  const double dv4 = dot3(v4.v, dir);
  double d = dv4 - dot3(p[1].v, dir);
  d = std::fmin(d, dv4 - dot3(p[2].v, dir));
  d = std::fmin(d, dv4 - dot3(p[3].v, dir));
  return eq(d, tol) || d < tol;
}

void expandPortal(Supp* p, const Supp& v4) {
  double v4v0[3];
  cross(v4v0, v4.v, p[0].v);
  if (dot3(p[1].v, v4v0) > 0) {
    if (dot3(p[2].v, v4v0) > 0) p[1] = v4;
    else p[3] = v4;
  } else {
    if (dot3(p[3].v, v4v0) > 0) p[2] = v4;
    else p[1] = v4;
  }
}

// squared distance from the origin to segment (a, b), closest point in w
double originSegmentDist2(const double* a, const double* b, double* w) {
  double d[3], na[3] = {-a[0], -a[1], -a[2]};
  sub3(d, b, a);
  double t = dot3(na, d) / dot3(d, d);
  if (t < 0 || isZero(t)) { copy3(w, a); return dot3(a, a); }
  if (t > 1 || eq(t, 1)) { copy3(w, b); return dot3(b, b); }
  w[0] = a[0] + t * d[0]; w[1] = a[1] + t * d[1]; w[2] = a[2] + t * d[2];
  return dot3(w, w);
}

// squared distance from the origin to triangle (x0, B, C), closest point in w (ccdVec3PointTriDist2 with P = 0)
double originTriDist2(const double* x0, const double* B, const double* C, double* w) {
  double d1[3], d2[3], a[3];
  sub3(d1, B, x0);
  sub3(d2, C, x0);
  copy3(a, x0);  // a = x0 - P with P the origin
  const double u = dot3(a, a), v = dot3(d1, d1), ww = dot3(d2, d2), p = dot3(a, d1), q = dot3(a, d2), r = dot3(d1, d2);
  const double den = ww * v - r * r;
  double s = (q * r - ww * p) / den, t = (-s * r - q) / ww;
  if ((isZero(s) || s > 0) && (eq(s, 1) || s < 1) && (isZero(t) || t > 0) && (eq(t, 1) || t < 1) && (eq(t + s, 1) || t + s < 1)) {
    for (int i = 0; i < 3; i++) w[i] = x0[i] + s * d1[i] + t * d2[i];
    (void)u;
    return dot3(w, w);
  }
  double best = originSegmentDist2(x0, B, w), w2[3];
  double d = originSegmentDist2(x0, C, w2);
  if (d < best) { best = d; copy3(w, w2); }
  d = originSegmentDist2(B, C, w2);
  if (d < best) { best = d; copy3(w, w2); }
  return best;
}

void findPos(const Supp* p, double* pos) {
  double dir[3], vec[3], b[4];
  portalDir(p, dir);
  cross(vec, p[1].v, p[2].v); b[0] = dot3(vec, p[3].v);
  cross(vec, p[3].v, p[2].v); b[1] = dot3(vec, p[0].v);
  cross(vec, p[0].v, p[1].v); b[2] = dot3(vec, p[3].v);
  cross(vec, p[2].v, p[1].v); b[3] = dot3(vec, p[0].v);
  double sum = b[0] + b[1] + b[2] + b[3];
  if (isZero(sum) || sum < 0) {
    b[0] = 0;
    cross(vec, p[2].v, p[3].v); b[1] = dot3(vec, dir);
    cross(vec, p[3].v, p[1].v); b[2] = dot3(vec, dir);
    cross(vec, p[1].v, p[2].v); b[3] = dot3(vec, dir);
    sum = b[1] + b[2] + b[3];
  }
  const double inv = 1 / sum;
  double p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  for (int i = 0; i < 4; i++) {
    addToScl3(p1, p[i].v1, b[i]);
    addToScl3(p2, p[i].v2, b[i]);
  }
  for (int k = 0; k < 3; k++) pos[k] = 0.5 * inv * (p1[k] + p2[k]);
}

// returns 0 and fills depth / dir / pos when the inflated geoms intersect, -1 otherwise
int mprPenetration(const Pair& P, int max_iterations, double tol, double* depth, double* dir, double* pos) {
  Supp p[4], v4;
  double d[3], va[3], vb[3];
  // ---- discoverPortal ----
  sub3(p[0].v, P.a->pos, P.b->pos);
  copy3(p[0].v1, P.a->pos);
  copy3(p[0].v2, P.b->pos);
  const double origin[3] = {0, 0, 0};
  if (vecEq(p[0].v, origin)) p[0].v[0] += kEps * 10;
  scl3(d, p[0].v, -1);
  normalize3(d);
  support(P, d, &p[1]);
  double dt = dot3(p[1].v, d);
  if (isZero(dt) || dt < 0) return -1;
  cross(d, p[0].v, p[1].v);
  if (isZero(dot3(d, d))) {
    if (vecEq(p[1].v, origin)) {  // touching contact on v1
      *depth = 0;
      zero3(dir);
    } else {                      // origin on the segment v0 - v1
      *depth = norm3(p[1].v);
      copy3(dir, p[1].v);
      normalize3(dir);
    }
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p[1].v1[k] + p[1].v2[k]);
    return 0;
  }
  normalize3(d);
  support(P, d, &p[2]);
  dt = dot3(p[2].v, d);
  if (isZero(dt) || dt < 0) return -1;
  sub3(va, p[1].v, p[0].v);
  sub3(vb, p[2].v, p[0].v);
  cross(d, va, vb);
  normalize3(d);
  if (dot3(d, p[0].v) > 0) {  // portal faces oriented away from v0
    Supp t = p[1]; p[1] = p[2]; p[2] = t;
    scl3(d, d, -1);
  }
  for (int guard = 0;; guard++) {
    if (guard > 1000) return -1;
    support(P, d, &p[3]);
    dt = dot3(p[3].v, d);
    if (isZero(dt) || dt < 0) return -1;
    bool cont = false;
    cross(va, p[1].v, p[3].v);
    dt = dot3(va, p[0].v);
    if (dt < 0 && !isZero(dt)) { p[2] = p[3]; cont = true; }
    if (!cont) {
      cross(va, p[3].v, p[2].v);
      dt = dot3(va, p[0].v);
      if (dt < 0 && !isZero(dt)) { p[1] = p[3]; cont = true; }
    }
    if (!cont) break;
    sub3(va, p[1].v, p[0].v);
    sub3(vb, p[2].v, p[0].v);
    cross(d, va, vb);
    normalize3(d);
  }
  // ---- refinePortal ----
  for (int guard = 0;; guard++) {
    if (guard > 1000) return -1;
    portalDir(p, d);
    dt = dot3(d, p[1].v);
    if (isZero(dt) || dt > 0) break;  // the portal encapsulates the origin
    support(P, d, &v4);
    dt = dot3(v4.v, d);
    if (!(isZero(dt) || dt > 0) || portalReachTolerance(p, v4, d, tol)) return -1;
    expandPortal(p, v4);
  }
  // ---- findPenetr ----
  for (int it = 0;; it++) {
    portalDir(p, d);
    support(P, d, &v4);
    if (portalReachTolerance(p, v4, d, tol) || it > max_iterations) {
      double w[3];
      *depth = std::sqrt(originTriDist2(p[1].v, p[2].v, p[3].v, w));
      if (isZero(w[0]) && isZero(w[1]) && isZero(w[2])) copy3(w, d);
      copy3(dir, w);
      normalize3(dir);
      findPos(p, pos);
      return 0;
    }
    expandPortal(p, v4);
  }
}

}  // namespace

// mjc_Convex: one contact from the MPR penetration query of the geoms inflated by margin / 2 each
int convexConvex(Con* con, double margin, const ConvexGeom& g1, const ConvexGeom& g2, int mpr_iterations, double mpr_tolerance) {
  Pair P{&g1, &g2, 0.5 * margin};
  double depth, dir[3], pos[3];
  if (mprPenetration(P, mpr_iterations, mpr_tolerance, &depth, dir, pos) != 0) return 0;
  if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) return 0;  // contact found but the normal is undefined
  con->dist = margin - depth;
  zero(con->frame, 9);
  copy3(con->frame, dir);
  copy3(con->pos, pos);
  return 1;
}

// ---- height field vs convex geom (mjc_ConvexHField).  The geom's bounding box in the hfield frame (support points
// along the six axis directions) selects a sub-grid; every grid triangle under it is extruded down to -base into a
// six-vertex prism (vertices arrive strip-wise: row r + 1 then row r for every column, each new vertex closing a
// triangle with the previous two) and tested against the geom with the MPR penetration query, one contact per
// penetrating prism, the prism being object 1.  Differences from MuJoCo 2.3.7, stated: (i) the margin is handled as
// in mjc_Convex (both objects inflated by margin / 2, dist = margin - depth) where MuJoCo raises the prism tops by the
// margin; identical at the default margin 0; (ii) at most `maxcon` (8) contacts per pair where MuJoCo stops at
// mjMAXCONPAIR = 50.  Recalled from memory, neither MuJoCo nor libccd is in /root/reference: "parity unpinned".
int hfieldConvex(Con* con, int maxcon, double margin, const double* pos1, const double* mat1, const double* hsize, int nrow,
                 int ncol, const double* data, const ConvexGeom& g2, double rbound2, int mpr_iterations, double mpr_tolerance) {
  // geom centre in the hfield frame: slab test against the bounding sphere
  double dif[3], pos[3];
  sub3(dif, g2.pos, pos1);
  rotVecMatT(pos, dif, mat1);
  const double r2 = rbound2 + margin;
  for (int i = 0; i < 2; i++)
    if (pos[i] > hsize[i] + r2 || pos[i] < -hsize[i] - r2) return 0;
  if (pos[2] > hsize[2] + r2 || pos[2] < -hsize[3] - r2) return 0;
  // bounding box of the geom in the hfield frame from its support mapping
  double xmin[3], xmax[3];
  for (int i = 0; i < 3; i++) {
    double dirw[3], sp[3], loc[3];
    const double ax[3] = {mat1[i], mat1[3 + i], mat1[6 + i]};  // hfield axis i in world coordinates
    copy3(dirw, ax);
    convexSupport(g2, dirw, sp);
    sub3(dif, sp, pos1);
    rotVecMatT(loc, dif, mat1);
    xmax[i] = loc[i] + margin;
    scl3(dirw, ax, -1.0);
    convexSupport(g2, dirw, sp);
    sub3(dif, sp, pos1);
    rotVecMatT(loc, dif, mat1);
    xmin[i] = loc[i] - margin;
  }
  if (xmin[0] > hsize[0] || xmax[0] < -hsize[0] || xmin[1] > hsize[1] || xmax[1] < -hsize[1] || xmin[2] > hsize[2] ||
      xmax[2] < -hsize[3])
    return 0;
  // sub-grid under the box
  int cmin = (int)std::floor((xmin[0] + hsize[0]) / (2 * hsize[0]) * (ncol - 1));
  int cmax = (int)std::ceil((xmax[0] + hsize[0]) / (2 * hsize[0]) * (ncol - 1));
  int rmin = (int)std::floor((xmin[1] + hsize[1]) / (2 * hsize[1]) * (nrow - 1));
  int rmax = (int)std::ceil((xmax[1] + hsize[1]) / (2 * hsize[1]) * (nrow - 1));
  cmin = cmin < 0 ? 0 : cmin;
  rmin = rmin < 0 ? 0 : rmin;
  cmax = cmax > ncol - 1 ? ncol - 1 : cmax;
  rmax = rmax > nrow - 1 ? nrow - 1 : rmax;
  const double dx = 2 * hsize[0] / (ncol - 1), dy = 2 * hsize[1] / (nrow - 1);
  double prism[18] = {0};  // vertices 0-2 bottom, 3-5 top, hfield frame
  double vrel[18], cl[3], cw[3];  // the prism as a six-vertex hull around its own centroid (the MPR's interior point)
  const double zero_size[3] = {0, 0, 0};
  ConvexGeom g1{B2MJ_GEOM_MESH, cw, mat1, zero_size, vrel, 6};
  int cnt = 0;
  for (int r = rmin; r < rmax; r++) {
    int nvert = 0;
    for (int c = cmin; c <= cmax; c++) {
      for (int i = 0; i < 2; i++) {
        const int rr = r + (i == 0 ? 1 : 0);
        // shift the strip: the two newest columns of the prism move down, the new vertex enters at slots 2 / 5
        for (int k = 0; k < 3; k++) {
          prism[k] = prism[3 + k]; prism[3 + k] = prism[6 + k];
          prism[9 + k] = prism[12 + k]; prism[12 + k] = prism[15 + k];
        }
        prism[6] = prism[15] = dx * c - hsize[0];
        prism[7] = prism[16] = dy * rr - hsize[1];
        prism[8] = -hsize[3];
        prism[17] = data[rr * ncol + c] * hsize[2];
        if (++nvert <= 2) continue;
        // prism entirely below the geom's lowest point: nothing to test
        if (prism[11] < xmin[2] && prism[14] < xmin[2] && prism[17] < xmin[2]) continue;
        for (int k = 0; k < 3; k++) {
          cl[k] = 0;
          for (int v = 0; v < 6; v++) cl[k] += prism[3 * v + k];
          cl[k] /= 6;
          for (int v = 0; v < 6; v++) vrel[3 * v + k] = prism[3 * v + k] - cl[k];
        }
        rotVecMat(cw, cl, mat1);
        addTo3(cw, pos1);
        if (convexConvex(con + cnt, margin, g1, g2, mpr_iterations, mpr_tolerance) == 1) {
          if (++cnt >= maxcon) return cnt;
        }
      }
    }
  }
  return cnt;
}

}  // namespace orc
