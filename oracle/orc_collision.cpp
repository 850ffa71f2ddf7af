// orc_collision.cpp — collision detection of the CPU oracle.
// TEST INFRASTRUCTURE ONLY (see orc_math.h).  Restates MuJoCo 2.3.7 engine_collision_driver.c
// (pair filtering, parameter mixing, contact ordering) and engine_collision_primitive.c (plane-*,
// sphere-*, capsule-capsule) — the mj_collision stage inside the `mj_step` call at reference
// mujoco_env.cpp:498; the narrowphase table is what MujocoEnv::registerCollisionFunction overrides
// (mujoco_env.cpp:163-176).  sphere-box / capsule-box follow closest-feature constructions of our own
// (MuJoCo's box routines are not reproducible from memory; parity unpinned).
//
// Candidate pairs come pre-filtered and pre-ordered from the model (collpair_*): body-pair signature
// ascending, then geom ids — the order the driver emits contacts in.  Only the data-dependent tests
// (bounding sphere, narrowphase distance vs margin) run here.
#include <algorithm>
#include <cmath>

#include "orc_convex.h"
#include "orc_math.h"
#include "orc_types.h"

namespace orc {

static int planeSphere(Con* con, double margin, const double* pos1, const double* mat1, const double* pos2, double radius) {
  double normal[3] = {mat1[2], mat1[5], mat1[8]}, tmp[3];
  sub3(tmp, pos2, pos1);
  double cdist = dot3(tmp, normal);
  if (cdist > margin + radius) return 0;
  con->dist = cdist - radius;
  zero(con->frame, 9);
  copy3(con->frame, normal);
  scl3(tmp, normal, -con->dist / 2 - radius);
  add3(con->pos, pos2, tmp);
  return 1;
}

static int planeCapsule(Con* con, double margin, const double* pos1, const double* mat1, const double* pos2,
                        const double* mat2, const double* size2) {
  double axis[3] = {mat2[2], mat2[5], mat2[8]}, seg[3], p[3];
  scl3(seg, axis, size2[1]);
  int n = 0;
  add3(p, pos2, seg);
  n += planeSphere(con + n, margin, pos1, mat1, p, size2[0]);
  sub3(p, pos2, seg);
  n += planeSphere(con + n, margin, pos1, mat1, p, size2[0]);
  for (int i = 0; i < n; i++) copy3(con[i].frame + 3, axis);  // align the tangent with the capsule axis
  return n;
}

static int planeBox(Con* con, double margin, const double* pos1, const double* mat1, const double* pos2,
                    const double* mat2, const double* size2) {
  double normal[3] = {mat1[2], mat1[5], mat1[8]}, dif[3];
  sub3(dif, pos2, pos1);
  double dist = dot3(dif, normal);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    double vec[3] = {(i & 1 ? size2[0] : -size2[0]), (i & 2 ? size2[1] : -size2[1]), (i & 4 ? size2[2] : -size2[2])};
    double corner[3];
    rotVecMat(corner, vec, mat2);
    double ldist = dot3(normal, corner);
    if (dist + ldist > margin || ldist > 0) continue;
    con[cnt].dist = dist + ldist;
    zero(con[cnt].frame, 9);
    copy3(con[cnt].frame, normal);
    addTo3(corner, pos2);
    scl3(vec, normal, -con[cnt].dist / 2);
    add3(con[cnt].pos, corner, vec);
    if (++cnt >= 4) return 4;
  }
  return cnt;
}

static int sphereSphere(Con* con, double margin, const double* pos1, double r1, const double* pos2, double r2) {
  double dif[3];
  sub3(dif, pos2, pos1);
  double cdist2 = dot3(dif, dif), bound = margin + r1 + r2;
  if (cdist2 > bound * bound) return 0;
  double cdist = normalize3(dif);
  con->dist = cdist - r1 - r2;
  zero(con->frame, 9);
  copy3(con->frame, dif);
  scl3(con->pos, dif, r1 + con->dist / 2);
  addTo3(con->pos, pos1);
  return 1;
}

static int sphereCapsule(Con* con, double margin, const double* pos1, double r1, const double* pos2, const double* mat2,
                         const double* size2) {
  double axis[3] = {mat2[2], mat2[5], mat2[8]}, vec[3], p[3];
  sub3(vec, pos1, pos2);
  double x = clampd(dot3(axis, vec), -size2[1], size2[1]);
  scl3(p, axis, x);
  addTo3(p, pos2);
  return sphereSphere(con, margin, pos1, r1, p, size2[0]);
}

static int capsuleCapsule(Con* con, double margin, const double* pos1, const double* mat1, const double* size1,
                          const double* pos2, const double* mat2, const double* size2) {
  double axis1[3] = {mat1[2], mat1[5], mat1[8]}, axis2[3] = {mat2[2], mat2[5], mat2[8]}, dif[3];
  sub3(dif, pos1, pos2);
  double ma = dot3(axis1, axis1), mb = -dot3(axis1, axis2), mc = dot3(axis2, axis2);
  double u = -dot3(axis1, dif), v = dot3(axis2, dif);
  double det = ma * mc - mb * mb;
  double vec1[3], vec2[3];
  if (std::fabs(det) >= MINVAL) {
    double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > size1[1]) { x1 = size1[1]; x2 = (v - mb * size1[1]) / mc; }
    else if (x1 < -size1[1]) { x1 = -size1[1]; x2 = (v + mb * size1[1]) / mc; }
    if (x2 > size2[1]) { x2 = size2[1]; x1 = clampd((u - mb * size2[1]) / ma, -size1[1], size1[1]); }
    else if (x2 < -size2[1]) { x2 = -size2[1]; x1 = clampd((u + mb * size2[1]) / ma, -size1[1], size1[1]); }
    scl3(vec1, axis1, x1); addTo3(vec1, pos1);
    scl3(vec2, axis2, x2); addTo3(vec2, pos2);
    return sphereSphere(con, margin, vec1, size1[0], vec2, size2[0]);
  }
  // parallel axes: up to two contacts at the segment ends
  int n = 0;
  double x1, x2;
  x1 = size1[1]; x2 = clampd((v - mb * size1[1]) / mc, -size2[1], size2[1]);
  scl3(vec1, axis1, x1); addTo3(vec1, pos1); scl3(vec2, axis2, x2); addTo3(vec2, pos2);
  n += sphereSphere(con + n, margin, vec1, size1[0], vec2, size2[0]);
  x1 = -size1[1]; x2 = clampd((v + mb * size1[1]) / mc, -size2[1], size2[1]);
  scl3(vec1, axis1, x1); addTo3(vec1, pos1); scl3(vec2, axis2, x2); addTo3(vec2, pos2);
  n += sphereSphere(con + n, margin, vec1, size1[0], vec2, size2[0]);
  if (n >= 2) return n;
  x2 = size2[1]; x1 = clampd((u - mb * size2[1]) / ma, -size1[1], size1[1]);
  scl3(vec1, axis1, x1); addTo3(vec1, pos1); scl3(vec2, axis2, x2); addTo3(vec2, pos2);
  n += sphereSphere(con + n, margin, vec1, size1[0], vec2, size2[0]);
  if (n >= 2) return n;
  x2 = -size2[1]; x1 = clampd((u + mb * size2[1]) / ma, -size1[1], size1[1]);
  scl3(vec1, axis1, x1); addTo3(vec1, pos1); scl3(vec2, axis2, x2); addTo3(vec2, pos2);
  n += sphereSphere(con + n, margin, vec1, size1[0], vec2, size2[0]);
  return n;
}

// sphere (geom1) vs box (geom2): closest feature of the box to the sphere centre
static int sphereBox(Con* con, double margin, const double* pos1, double radius, const double* pos2, const double* mat2,
                     const double* size2) {
  double tmp[3], center[3], clamped[3], dif[3];
  sub3(tmp, pos1, pos2);
  rotVecMatT(center, tmp, mat2);
  for (int i = 0; i < 3; i++) clamped[i] = clampd(center[i], -size2[i], size2[i]);
  sub3(dif, center, clamped);
  double dist = norm3(dif);
  if (dist - radius > margin) return 0;
  double nloc[3], ploc[3];
  if (dist <= MINVAL) {
    // centre inside the box: push out through the nearest face
    int k = 0;
    double depth = size2[0] - std::fabs(center[0]);
    for (int i = 1; i < 3; i++) {
      double di = size2[i] - std::fabs(center[i]);
      if (di < depth) { depth = di; k = i; }
    }
    double e = center[k] >= 0 ? 1.0 : -1.0;
    zero3(nloc);
    nloc[k] = -e;
    con->dist = -(depth + radius);
    copy3(ploc, center);
    ploc[k] += e * (depth - radius) / 2;
  } else {
    scl3(nloc, dif, -1.0 / dist);
    con->dist = dist - radius;
    // midpoint between sphere surface (centre + n*r) and box surface (clamped)
    for (int i = 0; i < 3; i++) ploc[i] = 0.5 * (center[i] + nloc[i] * radius + clamped[i]);
  }
  zero(con->frame, 9);
  rotVecMat(con->frame, nloc, mat2);
  rotVecMat(con->pos, ploc, mat2);
  addTo3(con->pos, pos2);
  return 1;
}

// squared distance from point c + t*a (box frame) to the box, minimised exactly over t in [-h, h]:
// the function is convex piecewise quadratic with breakpoints where a coordinate crosses +-size.
static double segmentBoxClosest(const double* c, const double* a, double h, const double* size) {
  double bp[8];
  int nbp = 0;
  bp[nbp++] = -h;
  bp[nbp++] = h;
  for (int i = 0; i < 3; i++) {
    if (std::fabs(a[i]) < MINVAL) continue;
    double t1 = (size[i] - c[i]) / a[i], t2 = (-size[i] - c[i]) / a[i];
    if (t1 > -h && t1 < h) bp[nbp++] = t1;
    if (t2 > -h && t2 < h) bp[nbp++] = t2;
  }
  std::sort(bp, bp + nbp);
  auto d2 = [&](double t) {
    double s = 0;
    for (int i = 0; i < 3; i++) {
      double p = c[i] + t * a[i];
      double e = std::fabs(p) - size[i];
      if (e > 0) s += e * e;
    }
    return s;
  };
  double best_t = bp[0], best = d2(bp[0]);
  for (int k = 0; k + 1 < nbp; k++) {
    double lo = bp[k], hi = bp[k + 1];
    if (hi - lo < MINVAL) continue;
    // quadratic on this interval: active coordinates decided at the midpoint
    double mid = 0.5 * (lo + hi), A = 0, B = 0;
    for (int i = 0; i < 3; i++) {
      double p = c[i] + mid * a[i];
      if (std::fabs(p) > size[i]) {
        double sgn = p > 0 ? 1.0 : -1.0;
        // (sgn*(c+t a) - size)^2 -> derivative terms
        A += a[i] * a[i];
        B += a[i] * (c[i] - sgn * size[i]);
      }
    }
    double cand[2] = {hi, hi};
    int nc = 1;
    if (A > MINVAL) {
      double t = clampd(-B / A, lo, hi);
      cand[0] = t;
      cand[1] = hi;
      nc = 2;
    }
    for (int q = 0; q < nc; q++) {
      double v = d2(cand[q]);
      if (v < best) { best = v; best_t = cand[q]; }
    }
  }
  return best_t;
}

// capsule (geom1) vs box (geom2): sphere-box at the segment point closest to the box, plus the far end
static int capsuleBox(Con* con, double margin, const double* pos1, const double* mat1, const double* size1,
                      const double* pos2, const double* mat2, const double* size2) {
  double axis[3] = {mat1[2], mat1[5], mat1[8]}, tmp[3], c[3], a[3], p[3];
  sub3(tmp, pos1, pos2);
  rotVecMatT(c, tmp, mat2);
  rotVecMatT(a, axis, mat2);
  const double h = size1[1];
  double t = segmentBoxClosest(c, a, h, size2);
  int n = 0;
  scl3(p, axis, t); addTo3(p, pos1);
  n += sphereBox(con + n, margin, p, size1[0], pos2, mat2, size2);
  double t2 = t >= 0 ? -h : h;
  if (std::fabs(t2 - t) > 1e-3 * h) {
    scl3(p, axis, t2); addTo3(p, pos1);
    n += sphereBox(con + n, margin, p, size1[0], pos2, mat2, size2);
  }
  return n;
}

// narrowphase dispatch; geoms already ordered so that type1 <= type2.  Returns -1 if unsupported.
// box-box: separating-axis search over the 15 candidate axes, then either a face contact (incident
// face clipped against the side planes of the reference face, up to 8 points) or a single edge-edge
// contact.  A construction of our own, like the other box routines (MuJoCo's mjc_BoxBox is not
// reproducible from memory; parity unpinned) — the CUDA narrowphase mirrors it step by step.
static int boxBox(Con* con, double margin, const double* pos1, const double* mat1, const double* size1,
                  const double* pos2, const double* mat2, const double* size2) {
  double A[3][3], B[3][3], d[3], dA[3], dB[3], R[3][3], aR[3][3];
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) { A[i][k] = mat1[3 * k + i]; B[i][k] = mat2[3 * k + i]; }  // axes = matrix columns
  sub3(d, pos2, pos1);
  for (int i = 0; i < 3; i++) {
    dA[i] = dot3(d, A[i]);
    dB[i] = dot3(d, B[i]);
    for (int j = 0; j < 3; j++) { R[i][j] = dot3(A[i], B[j]); aR[i][j] = std::fabs(R[i][j]) + 1e-12; }
  }
  double best = -1e30;
  int code = -1;
  for (int i = 0; i < 3; i++) {
    const double s = std::fabs(dA[i]) - (size1[i] + size2[0] * aR[i][0] + size2[1] * aR[i][1] + size2[2] * aR[i][2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = i; }
  }
  for (int j = 0; j < 3; j++) {
    const double s = std::fabs(dB[j]) - (size2[j] + size1[0] * aR[0][j] + size1[1] * aR[1][j] + size1[2] * aR[2][j]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = 3 + j; }
  }
  double edgeN[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double L[3];
      cross(L, A[i], B[j]);
      const double l = norm3(L);
      if (l < 1e-6) continue;
      scl3(L, L, 1 / l);
      double rA = 0, rB = 0;
      for (int k = 0; k < 3; k++) {
        if (k != i) rA += size1[k] * std::fabs(dot3(A[k], L));
        if (k != j) rB += size2[k] * std::fabs(dot3(B[k], L));
      }
      const double dl = dot3(d, L);
      const double s = std::fabs(dl) - (rA + rB);
      if (s > margin) return 0;
      if (s > best + 0.05 * std::fabs(best) + 1e-9) {
        best = s;
        code = 6 + 3 * i + j;
        scl3(edgeN, L, dl < 0 ? -1.0 : 1.0);
      }
    }
  if (code < 6) {
    // face contact: reference box r (geom1 if code < 3), incident box c
    const bool refA = code < 3;
    const int r = refA ? code : code - 3;
    const double (*Rf)[3] = refA ? A : B;
    const double (*In)[3] = refA ? B : A;
    const double* pr = refA ? pos1 : pos2;
    const double* pi = refA ? pos2 : pos1;
    const double* sr = refA ? size1 : size2;
    const double* si = refA ? size2 : size1;
    double n[3];
    const double sgn = (refA ? dA[r] : -dB[r]) < 0 ? -1.0 : 1.0;
    scl3(n, Rf[r], sgn);
    int k = 0;
    double bestdot = -1;
    for (int q = 0; q < 3; q++) {
      const double a = std::fabs(dot3(n, In[q]));
      if (a > bestdot) { bestdot = a; k = q; }
    }
    const double fs = dot3(n, In[k]) > 0 ? -1.0 : 1.0;
    const int ku = (k + 1) % 3, kv = (k + 2) % 3, ra = (r + 1) % 3, rb = (r + 2) % 3;
    double poly[8][3], tmp[8][3];  // (x, y, h) in the reference-face frame
    int np = 4;
    const double su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
    for (int v = 0; v < 4; v++) {
      double w[3];
      for (int c = 0; c < 3; c++)
        w[c] = pi[c] + fs * si[k] * In[k][c] + su[v] * si[ku] * In[ku][c] + sv[v] * si[kv] * In[kv][c] - pr[c];
      poly[v][0] = dot3(w, Rf[ra]);
      poly[v][1] = dot3(w, Rf[rb]);
      poly[v][2] = dot3(w, n) - sr[r];
    }
    for (int pl = 0; pl < 4; pl++) {  // clip against x <= sa, -x <= sa, y <= sb, -y <= sb
      const int ax = pl >> 1;
      const double sg = (pl & 1) ? -1.0 : 1.0, lim = ax == 0 ? sr[ra] : sr[rb];
      int nq = 0;
      for (int v = 0; v < np && nq < 8; v++) {
        const double* p0 = poly[v];
        const double* p1 = poly[(v + 1) % np];
        const double e0 = sg * p0[ax] - lim, e1 = sg * p1[ax] - lim;
        if (e0 <= 0) { for (int c = 0; c < 3; c++) tmp[nq][c] = p0[c]; nq++; }
        if ((e0 < 0 && e1 > 0) || (e0 > 0 && e1 < 0)) {
          if (nq < 8) {
            const double t = e0 / (e0 - e1);
            for (int c = 0; c < 3; c++) tmp[nq][c] = p0[c] + t * (p1[c] - p0[c]);
            nq++;
          }
        }
      }
      np = nq;
      for (int v = 0; v < np; v++) for (int c = 0; c < 3; c++) poly[v][c] = tmp[v][c];
      if (np == 0) return 0;
    }
    int num = 0;
    for (int v = 0; v < np && num < 8; v++) {
      const double h = poly[v][2];
      if (h >= margin) continue;
      Con& cn = con[num++];
      cn.dist = h;
      for (int c = 0; c < 3; c++) {
        cn.pos[c] = pr[c] + poly[v][0] * Rf[ra][c] + poly[v][1] * Rf[rb][c] + (sr[r] + 0.5 * h) * n[c];
        cn.frame[c] = refA ? n[c] : -n[c];
      }
      for (int c = 3; c < 9; c++) cn.frame[c] = 0;
    }
    return num;
  }
  // edge-edge contact
  const int i = (code - 6) / 3, j = (code - 6) % 3;
  double pa[3], pb[3];
  copy3(pa, pos1);
  copy3(pb, pos2);
  for (int k = 0; k < 3; k++) {
    if (k != i) addToScl3(pa, A[k], (dot3(A[k], edgeN) > 0 ? 1.0 : -1.0) * size1[k]);
    if (k != j) addToScl3(pb, B[k], (dot3(B[k], edgeN) > 0 ? -1.0 : 1.0) * size2[k]);
  }
  // closest points of the lines pa + al*A_i and pb + be*B_j
  double w[3];
  sub3(w, pb, pa);
  const double uu = dot3(A[i], B[j]), wa = dot3(w, A[i]), wb = dot3(w, B[j]);
  const double den = 1 - uu * uu;
  double al = 0, be = 0;
  if (den > 1e-12) { al = (wa - uu * wb) / den; be = (uu * wa - wb) / den; }
  al = std::fmax(-size1[i], std::fmin(size1[i], al));
  be = std::fmax(-size2[j], std::fmin(size2[j], be));
  addToScl3(pa, A[i], al);
  addToScl3(pb, B[j], be);
  Con& cn = con[0];
  cn.dist = best;
  for (int c = 0; c < 3; c++) { cn.pos[c] = 0.5 * (pa[c] + pb[c]); cn.frame[c] = edgeN[c]; }
  for (int c = 3; c < 9; c++) cn.frame[c] = 0;
  return 1;
}

static int narrowphase(const b2mjModel* m, const OrcData* d, Con* con, int g1, int g2, double margin) {
  const int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  const double *pos1 = d->geom_xpos + 3 * g1, *mat1 = d->geom_xmat + 9 * g1, *size1 = m->geom_size + 3 * g1;
  const double *pos2 = d->geom_xpos + 3 * g2, *mat2 = d->geom_xmat + 9 * g2, *size2 = m->geom_size + 3 * g2;
  auto convex = [&](int g, int t, const double* pos, const double* mat, const double* size) {
    ConvexGeom c{t, pos, mat, size, nullptr, 0};
    if (t == B2MJ_GEOM_MESH && m->geom_dataid[g] >= 0) {
      const int id = m->geom_dataid[g];
      c.vert = m->mesh_vert + 3 * m->mesh_vertadr[id];
      c.nvert = m->mesh_vertnum[id];
    }
    return c;
  };
  if (t1 == B2MJ_GEOM_PLANE) {
    if (t2 == B2MJ_GEOM_SPHERE) return planeSphere(con, margin, pos1, mat1, pos2, size2[0]);
    if (t2 == B2MJ_GEOM_CAPSULE) return planeCapsule(con, margin, pos1, mat1, pos2, mat2, size2);
    if (t2 == B2MJ_GEOM_BOX) return planeBox(con, margin, pos1, mat1, pos2, mat2, size2);
    if (t2 == B2MJ_GEOM_CYLINDER) return planeCylinder(con, margin, pos1, mat1, pos2, mat2, size2);
    if (t2 == B2MJ_GEOM_ELLIPSOID || t2 == B2MJ_GEOM_MESH) return planeConvex(con, margin, pos1, mat1, convex(g2, t2, pos2, mat2, size2));
    return -1;
  }
  if (t1 == B2MJ_GEOM_HFIELD) {  // mjc_ConvexHField for every convex geom type
    if (t2 < B2MJ_GEOM_SPHERE || t2 > B2MJ_GEOM_MESH || m->geom_dataid[g1] < 0) return -1;
    const int h = m->geom_dataid[g1];
    return hfieldConvex(con, 8, margin, pos1, mat1, m->hfield_size + 4 * h, m->hfield_nrow[h], m->hfield_ncol[h],
                        m->hfield_data + m->hfield_adr[h], convex(g2, t2, pos2, mat2, size2), m->geom_rbound[g2],
                        m->opt.mpr_iterations, m->opt.mpr_tolerance);
  }
  // pairs with a dedicated primitive function (mjCOLLISIONFUNC); every other convex pair goes to the general MPR test
  if (t1 == B2MJ_GEOM_SPHERE && t2 == B2MJ_GEOM_SPHERE) return sphereSphere(con, margin, pos1, size1[0], pos2, size2[0]);
  if (t1 == B2MJ_GEOM_SPHERE && t2 == B2MJ_GEOM_CAPSULE) return sphereCapsule(con, margin, pos1, size1[0], pos2, mat2, size2);
  if (t1 == B2MJ_GEOM_SPHERE && t2 == B2MJ_GEOM_BOX) return sphereBox(con, margin, pos1, size1[0], pos2, mat2, size2);
  if (t1 == B2MJ_GEOM_CAPSULE && t2 == B2MJ_GEOM_CAPSULE) return capsuleCapsule(con, margin, pos1, mat1, size1, pos2, mat2, size2);
  if (t1 == B2MJ_GEOM_CAPSULE && t2 == B2MJ_GEOM_BOX) return capsuleBox(con, margin, pos1, mat1, size1, pos2, mat2, size2);
  if (t1 == B2MJ_GEOM_BOX && t2 == B2MJ_GEOM_BOX) return boxBox(con, margin, pos1, mat1, size1, pos2, mat2, size2);
  if (t1 >= B2MJ_GEOM_SPHERE && t1 <= B2MJ_GEOM_MESH && t2 >= B2MJ_GEOM_SPHERE && t2 <= B2MJ_GEOM_MESH)
    return convexConvex(con, margin, convex(g1, t1, pos1, mat1, size1), convex(g2, t2, pos2, mat2, size2),
                        m->opt.mpr_iterations, m->opt.mpr_tolerance);
  return -1;
}

// mj_collision
void collision(const b2mjModel* m, OrcData* d) {
  d->ncon() = 0;
  if (m->opt.disableflags & (B2MJ_DSBL_CONSTRAINT | B2MJ_DSBL_CONTACT)) return;
  if (m->nconmax == 0) return;
  for (int p = 0; p < m->ncollpair; p++) {
    const int g1 = m->collpair_geom1[p], g2 = m->collpair_geom2[p];
    const int pid = m->collpair_pairid[p];  // explicit <contact><pair>: its own margin / gap / contact parameters
    const double margin = pid >= 0 ? m->pair_margin[pid] : std::fmax(m->geom_margin[g1], m->geom_margin[g2]);
    const double gap = pid >= 0 ? m->pair_gap[pid] : std::fmax(m->geom_gap[g1], m->geom_gap[g2]);
    // bounding-sphere filter (planes: signed distance to the plane)
    const double r1 = m->geom_rbound[g1], r2 = m->geom_rbound[g2];
    if (r1 > 0 && r2 > 0) {
      double dif[3];
      sub3(dif, d->geom_xpos + 3 * g1, d->geom_xpos + 3 * g2);
      double bound = r1 + r2 + margin;
      if (dot3(dif, dif) > bound * bound) continue;
    } else if (m->geom_type[g1] == B2MJ_GEOM_PLANE && r2 > 0) {
      const double* mat1 = d->geom_xmat + 9 * g1;
      double normal[3] = {mat1[2], mat1[5], mat1[8]}, dif[3];
      sub3(dif, d->geom_xpos + 3 * g2, d->geom_xpos + 3 * g1);
      if (dot3(dif, normal) > margin + r2) continue;
    }
    Con con[8];
    int num = narrowphase(m, d, con, g1, g2, margin);
    if (num <= 0) continue;
    // contact parameters (mixing rules of mj_collideGeoms)
    int condim;
    double solref[2], solimp[5], fri[3], fri5[5];
    if (pid >= 0) {
      condim = m->pair_dim[pid];
      copy(solref, m->pair_solref + 2 * pid, 2);
      copy(solimp, m->pair_solimp + 5 * pid, 5);
      copy(fri5, m->pair_friction + 5 * pid, 5);
    } else if (m->geom_priority[g1] != m->geom_priority[g2]) {
      int gi = m->geom_priority[g1] > m->geom_priority[g2] ? g1 : g2;
      condim = m->geom_condim[gi];
      copy(solref, m->geom_solref + 2 * gi, 2);
      copy(solimp, m->geom_solimp + 5 * gi, 5);
      copy3(fri, m->geom_friction + 3 * gi);
    } else {
      condim = std::max(m->geom_condim[g1], m->geom_condim[g2]);
      double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2], mix;
      if (s1 >= MINVAL && s2 >= MINVAL) mix = s1 / (s1 + s2);
      else if (s1 < MINVAL && s2 < MINVAL) mix = 0.5;
      else if (s1 < MINVAL) mix = 0.0;
      else mix = 1.0;
      const double *a = m->geom_solref + 2 * g1, *b = m->geom_solref + 2 * g2;
      if (a[0] > 0 && b[0] > 0) for (int i = 0; i < 2; i++) solref[i] = mix * a[i] + (1 - mix) * b[i];
      else for (int i = 0; i < 2; i++) solref[i] = std::fmin(a[i], b[i]);
      for (int i = 0; i < 5; i++) solimp[i] = mix * m->geom_solimp[5 * g1 + i] + (1 - mix) * m->geom_solimp[5 * g2 + i];
      for (int i = 0; i < 3; i++) fri[i] = std::fmax(m->geom_friction[3 * g1 + i], m->geom_friction[3 * g2 + i]);
    }
    for (int i = 0; i < num; i++) {
      if (d->ncon() >= m->nconmax) {
        d->warning[B2MJ_WARN_CONTACTFULL]++;
        return;
      }
      const int c = d->ncon()++;
      d->contact_dist[c] = con[i].dist;
      copy3(d->contact_pos + 3 * c, con[i].pos);
      copy(d->contact_frame + 9 * c, con[i].frame, 9);
      makeFrame(d->contact_frame + 9 * c);
      d->contact_includemargin[c] = margin - gap;
      double* f = d->contact_friction + 5 * c;
      if (pid >= 0) {
        for (int k = 0; k < 5; k++) f[k] = std::fmax(B2MJ_MINMU, fri5[k]);
      } else {
        f[0] = f[1] = std::fmax(B2MJ_MINMU, fri[0]);
        f[2] = std::fmax(B2MJ_MINMU, fri[1]);
        f[3] = f[4] = std::fmax(B2MJ_MINMU, fri[2]);
      }
      copy(d->contact_solref + 2 * c, solref, 2);
      copy(d->contact_solimp + 5 * c, solimp, 5);
      d->contact_mu[c] = 0;
      d->contact_dim[c] = condim;
      d->contact_geom1[c] = g1;
      d->contact_geom2[c] = g2;
      d->contact_exclude[c] = con[i].dist >= d->contact_includemargin[c];
      d->contact_efc_address[c] = -1;
    }
  }
}

}  // namespace orc
