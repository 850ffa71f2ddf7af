// stages_collision.cuh — broad/narrow-phase collision, one candidate geom pair per lane.
//
// Replaces mj_collision inside the reference's mj_step call (mujoco_env.cpp:498; row M5 of SURVEY 8a).
// The candidate list (model collpair_*) is static, pre-filtered and pre-ordered on the host (weld
// groups, parent filter, contype/conaffinity, excludes — the same rules MuJoCo's driver applies), so
// the device only runs the data-dependent tests.  Determinism ("bit-exact contact-pair indexing"):
// each lane keeps the contacts of its pair in registers/local memory, an ordered warp prefix sum gives
// every contact its final index — no atomics, the contact order is a pure function of the geometry.
#pragma once
#include "env_ctx.cuh"
#include "pair_con.cuh"
#include "stages_convex.cuh"

namespace b2k {


__device__ __forceinline__ int c_planeSphere(PairCon& o, int n, double margin, const double* pos1, const double* mat1,
                                             const double* pos2, double radius) {
  double normal[3] = {mat1[2], mat1[5], mat1[8]}, tmp[3];
  sub3(tmp, pos2, pos1);
  const double cdist = dot3(tmp, normal);
  if (cdist > margin + radius) return 0;
  const double dist = cdist - radius;
  o.dist[n] = dist;
  copy3(o.frame + 6 * n, normal);
  zero3(o.frame + 6 * n + 3);
  scl3(tmp, normal, -dist / 2 - radius);
  add3(o.pos + 3 * n, pos2, tmp);
  return 1;
}

__device__ __forceinline__ int c_sphereSphere(PairCon& o, int n, double margin, const double* pos1, double r1,
                                              const double* pos2, double r2) {
  double dif[3];
  sub3(dif, pos2, pos1);
  const double cdist2 = dot3(dif, dif), bound = margin + r1 + r2;
  if (cdist2 > bound * bound) return 0;
  const double cdist = normalize3(dif);
  const double dist = cdist - r1 - r2;
  o.dist[n] = dist;
  copy3(o.frame + 6 * n, dif);
  zero3(o.frame + 6 * n + 3);
  scl3(o.pos + 3 * n, dif, r1 + dist / 2);
  addTo3(o.pos + 3 * n, pos1);
  return 1;
}

__device__ __forceinline__ int c_sphereBox(PairCon& o, int n, double margin, const double* pos1, double radius,
                                           const double* pos2, const double* mat2, const double* size2) {
  double tmp[3], center[3], clamped[3], dif[3];
  sub3(tmp, pos1, pos2);
  rotVecMatT(center, tmp, mat2);
  for (int i = 0; i < 3; i++) clamped[i] = clampd(center[i], -size2[i], size2[i]);
  sub3(dif, center, clamped);
  const double dist = norm3(dif);
  if (dist - radius > margin) return 0;
  double nloc[3], ploc[3];
  if (dist <= B2K_MINVAL) {
    int k = 0;
    double depth = size2[0] - fabs(center[0]);
    for (int i = 1; i < 3; i++) {
      const double di = size2[i] - fabs(center[i]);
      if (di < depth) { depth = di; k = i; }
    }
    const double s = center[k] >= 0 ? 1.0 : -1.0;
    zero3(nloc);
    nloc[k] = -s;
    o.dist[n] = -(depth + radius);
    copy3(ploc, center);
    ploc[k] += s * (depth - radius) / 2;
  } else {
    scl3(nloc, dif, -1.0 / dist);
    o.dist[n] = dist - radius;
    for (int i = 0; i < 3; i++) ploc[i] = 0.5 * (center[i] + nloc[i] * radius + clamped[i]);
  }
  rotVecMat(o.frame + 6 * n, nloc, mat2);
  zero3(o.frame + 6 * n + 3);
  rotVecMat(o.pos + 3 * n, ploc, mat2);
  addTo3(o.pos + 3 * n, pos2);
  return 1;
}


// box-box: separating-axis search over the 15 candidate axes, then a face contact (incident face
// clipped against the side planes of the reference face, up to 8 points, one shared normal) or a single
// edge-edge contact.
__device__ __noinline__ int c_boxBox(PairCon& o, double margin, const double* pos1, const double* mat1, const double* size1,
                                     const double* pos2, const double* mat2, const double* size2) {
  double A[3][3], B[3][3], d[3], dA[3], dB[3], aR[3][3];
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) { A[i][k] = mat1[3 * k + i]; B[i][k] = mat2[3 * k + i]; }
  sub3(d, pos2, pos1);
  for (int i = 0; i < 3; i++) {
    dA[i] = dot3(d, A[i]);
    dB[i] = dot3(d, B[i]);
    for (int j = 0; j < 3; j++) aR[i][j] = fabs(dot3(A[i], B[j])) + 1e-12;
  }
  double best = -1e30;
  int code = -1;
  for (int i = 0; i < 3; i++) {
    const double s = fabs(dA[i]) - (size1[i] + size2[0] * aR[i][0] + size2[1] * aR[i][1] + size2[2] * aR[i][2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = i; }
  }
  for (int j = 0; j < 3; j++) {
    const double s = fabs(dB[j]) - (size2[j] + size1[0] * aR[0][j] + size1[1] * aR[1][j] + size1[2] * aR[2][j]);
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
        if (k != i) rA += size1[k] * fabs(dot3(A[k], L));
        if (k != j) rB += size2[k] * fabs(dot3(B[k], L));
      }
      const double dl = dot3(d, L);
      const double s = fabs(dl) - (rA + rB);
      if (s > margin) return 0;
      if (s > best + 0.05 * fabs(best) + 1e-9) {
        best = s;
        code = 6 + 3 * i + j;
        scl3(edgeN, L, dl < 0 ? -1.0 : 1.0);
      }
    }
  o.shared_frame = 1;
  zero3(o.frame + 3);
  if (code < 6) {
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
      const double a = fabs(dot3(n, In[q]));
      if (a > bestdot) { bestdot = a; k = q; }
    }
    const double fs = dot3(n, In[k]) > 0 ? -1.0 : 1.0;
    const int ku = (k + 1) % 3, kv = (k + 2) % 3, ra = (r + 1) % 3, rb = (r + 2) % 3;
    double poly[8][3], tmp[8][3];
    int np = 4;
    for (int v = 0; v < 4; v++) {
      const double su = (v == 0 || v == 3) ? 1.0 : -1.0, sv = v < 2 ? 1.0 : -1.0;
      double w[3];
      for (int c = 0; c < 3; c++)
        w[c] = pi[c] + fs * si[k] * In[k][c] + su * si[ku] * In[ku][c] + sv * si[kv] * In[kv][c] - pr[c];
      poly[v][0] = dot3(w, Rf[ra]);
      poly[v][1] = dot3(w, Rf[rb]);
      poly[v][2] = dot3(w, n) - sr[r];
    }
    for (int pl = 0; pl < 4; pl++) {
      const int ax = pl >> 1;
      const double sg = (pl & 1) ? -1.0 : 1.0, lim = ax == 0 ? sr[ra] : sr[rb];
      int nq = 0;
      B2K_NOUNROLL for (int v = 0; v < np && nq < 8; v++) {
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
      B2K_NOUNROLL for (int v = 0; v < np; v++) for (int c = 0; c < 3; c++) poly[v][c] = tmp[v][c];
      if (np == 0) return 0;
    }
    int num = 0;
    for (int c = 0; c < 3; c++) o.frame[c] = refA ? n[c] : -n[c];
    B2K_NOUNROLL for (int v = 0; v < np && num < 8; v++) {
      const double h = poly[v][2];
      if (h >= margin) continue;
      o.dist[num] = h;
      for (int c = 0; c < 3; c++)
        o.pos[3 * num + c] = pr[c] + poly[v][0] * Rf[ra][c] + poly[v][1] * Rf[rb][c] + (sr[r] + 0.5 * h) * n[c];
      num++;
    }
    return num;
  }
  const int i = (code - 6) / 3, j = (code - 6) % 3;
  double pa[3], pb[3];
  copy3(pa, pos1);
  copy3(pb, pos2);
  for (int k = 0; k < 3; k++) {
    if (k != i) addToScl3(pa, A[k], (dot3(A[k], edgeN) > 0 ? 1.0 : -1.0) * size1[k]);
    if (k != j) addToScl3(pb, B[k], (dot3(B[k], edgeN) > 0 ? -1.0 : 1.0) * size2[k]);
  }
  double w[3];
  sub3(w, pb, pa);
  const double uu = dot3(A[i], B[j]), wa = dot3(w, A[i]), wb = dot3(w, B[j]);
  const double den = 1 - uu * uu;
  double al = 0, be = 0;
  if (den > 1e-12) { al = (wa - uu * wb) / den; be = (uu * wa - wb) / den; }
  al = fmax(-size1[i], fmin(size1[i], al));
  be = fmax(-size2[j], fmin(size2[j], be));
  addToScl3(pa, A[i], al);
  addToScl3(pb, B[j], be);
  o.dist[0] = best;
  for (int c = 0; c < 3; c++) { o.pos[c] = 0.5 * (pa[c] + pb[c]); o.frame[c] = edgeN[c]; }
  return 1;
}


// exact minimiser over t in [-h,h] of the squared distance from c + t*a to an origin-centred box
__device__ __forceinline__ double segmentBoxClosest(const double* c, const double* a, double h, const double* size) {
  double bp[8];
  int nbp = 0;
  bp[nbp++] = -h;
  bp[nbp++] = h;
  for (int i = 0; i < 3; i++) {
    if (fabs(a[i]) < B2K_MINVAL) continue;
    const double t1 = (size[i] - c[i]) / a[i], t2 = (-size[i] - c[i]) / a[i];
    if (t1 > -h && t1 < h) bp[nbp++] = t1;
    if (t2 > -h && t2 < h) bp[nbp++] = t2;
  }
  B2K_NOUNROLL for (int i = 1; i < nbp; i++) {  // insertion sort
    const double v = bp[i];
    int j = i - 1;
    while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; j--; }
    bp[j + 1] = v;
  }
  auto d2 = [&](double t) {
    double s = 0;
    for (int i = 0; i < 3; i++) {
      const double p = c[i] + t * a[i];
      const double ex = fabs(p) - size[i];
      if (ex > 0) s += ex * ex;
    }
    return s;
  };
  double best_t = bp[0], best = d2(bp[0]);
  for (int k = 0; k + 1 < nbp; k++) {
    const double lo = bp[k], hi = bp[k + 1];
    if (hi - lo < B2K_MINVAL) continue;
    const double mid = 0.5 * (lo + hi);
    double A = 0, B = 0;
    for (int i = 0; i < 3; i++) {
      const double p = c[i] + mid * a[i];
      if (fabs(p) > size[i]) {
        const double sgn = p > 0 ? 1.0 : -1.0;
        A += a[i] * a[i];
        B += a[i] * (c[i] - sgn * size[i]);
      }
    }
    double cand[2] = {hi, hi};
    int nc = 1;
    if (A > B2K_MINVAL) { cand[0] = clampd(-B / A, lo, hi); cand[1] = hi; nc = 2; }
    B2K_NOUNROLL for (int q = 0; q < nc; q++) {
      const double v = d2(cand[q]);
      if (v < best) { best = v; best_t = cand[q]; }
    }
  }
  return best_t;
}

__device__ __noinline__ int narrowphase(const double* gxpos, const double* gxmat, PairCon& o, int g1, int g2,
                           double margin) {
  const DevModel& m = c_dm;
  const int t1 = m.geom_type[g1], t2 = m.geom_type[g2];
  const double *pos1 = gxpos + 3 * g1, *mat1 = gxmat + 9 * g1, *size1 = m.geom_size + 3 * g1;
  const double *pos2 = gxpos + 3 * g2, *mat2 = gxmat + 9 * g2, *size2 = m.geom_size + 3 * g2;
  // narrowphase override table: the batched counterpart of MujocoEnv::registerCollisionFunction, which overwrites
  // mjCOLLISIONFUNC[t1][t2] (mujoco_env.cpp:163-176).  Device code cannot call a host plugin's function pointer, so the
  // table selects among functions the library provides (b2mj_register_collision_function).
  const int fn = m.collfunc[(t1 & 7) * 8 + (t2 & 7)];
  if (fn == B2MJ_COLLFN_NONE) return 0;
  if (fn == B2MJ_COLLFN_BOUNDING_SPHERES) {
    if (t1 == B2MJ_GEOM_PLANE) return c_planeSphere(o, 0, margin, pos1, mat1, pos2, m.geom_rbound[g2]);
    return c_sphereSphere(o, 0, margin, pos1, m.geom_rbound[g1], pos2, m.geom_rbound[g2]);
  }
  if (t1 == B2MJ_GEOM_PLANE) {
    if (t2 == B2MJ_GEOM_SPHERE) return c_planeSphere(o, 0, margin, pos1, mat1, pos2, size2[0]);
    if (t2 == B2MJ_GEOM_CAPSULE) {
      double axis[3] = {mat2[2], mat2[5], mat2[8]}, seg[3], p[3];
      scl3(seg, axis, size2[1]);
      int n = 0;
      add3(p, pos2, seg);
      n += c_planeSphere(o, n, margin, pos1, mat1, p, size2[0]);
      sub3(p, pos2, seg);
      n += c_planeSphere(o, n, margin, pos1, mat1, p, size2[0]);
      B2K_NOUNROLL for (int i = 0; i < n; i++) copy3(o.frame + 6 * i + 3, axis);
      return n;
    }
    if (t2 == B2MJ_GEOM_CYLINDER) return c_planeCylinder(o, margin, pos1, mat1, pos2, mat2, size2);
    if (t2 == B2MJ_GEOM_ELLIPSOID || t2 == B2MJ_GEOM_MESH) return c_planeConvex(o, margin, pos1, mat1, cvxGeom(gxpos, gxmat, g2));
    if (t2 == B2MJ_GEOM_BOX) {
      double normal[3] = {mat1[2], mat1[5], mat1[8]}, dif[3];
      sub3(dif, pos2, pos1);
      const double dist = dot3(dif, normal);
      int cnt = 0;
      for (int i = 0; i < 8; i++) {
        double vec[3] = {(i & 1 ? size2[0] : -size2[0]), (i & 2 ? size2[1] : -size2[1]), (i & 4 ? size2[2] : -size2[2])};
        double corner[3];
        rotVecMat(corner, vec, mat2);
        const double ldist = dot3(normal, corner);
        if (dist + ldist > margin || ldist > 0) continue;
        o.dist[cnt] = dist + ldist;
        copy3(o.frame + 6 * cnt, normal);
        zero3(o.frame + 6 * cnt + 3);
        addTo3(corner, pos2);
        scl3(vec, normal, -o.dist[cnt] / 2);
        add3(o.pos + 3 * cnt, corner, vec);
        if (++cnt >= 4) return 4;
      }
      return cnt;
    }
    return 0;
  }
  if (t1 == B2MJ_GEOM_SPHERE) {
    if (t2 == B2MJ_GEOM_SPHERE) return c_sphereSphere(o, 0, margin, pos1, size1[0], pos2, size2[0]);
    if (t2 == B2MJ_GEOM_CAPSULE) {
      double axis[3] = {mat2[2], mat2[5], mat2[8]}, vec[3], p[3];
      sub3(vec, pos1, pos2);
      const double x = clampd(dot3(axis, vec), -size2[1], size2[1]);
      scl3(p, axis, x);
      addTo3(p, pos2);
      return c_sphereSphere(o, 0, margin, pos1, size1[0], p, size2[0]);
    }
    if (t2 == B2MJ_GEOM_BOX) return c_sphereBox(o, 0, margin, pos1, size1[0], pos2, mat2, size2);
  }
  if (t1 == B2MJ_GEOM_CAPSULE) {
    if (t2 == B2MJ_GEOM_CAPSULE) {
      double axis1[3] = {mat1[2], mat1[5], mat1[8]}, axis2[3] = {mat2[2], mat2[5], mat2[8]}, dif[3];
      sub3(dif, pos1, pos2);
      const double ma = dot3(axis1, axis1), mb = -dot3(axis1, axis2), mc = dot3(axis2, axis2);
      const double u = -dot3(axis1, dif), v = dot3(axis2, dif);
      const double det = ma * mc - mb * mb;
      double vec1[3], vec2[3];
      if (fabs(det) >= B2K_MINVAL) {
        double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
        if (x1 > size1[1]) { x1 = size1[1]; x2 = (v - mb * size1[1]) / mc; }
        else if (x1 < -size1[1]) { x1 = -size1[1]; x2 = (v + mb * size1[1]) / mc; }
        if (x2 > size2[1]) { x2 = size2[1]; x1 = clampd((u - mb * size2[1]) / ma, -size1[1], size1[1]); }
        else if (x2 < -size2[1]) { x2 = -size2[1]; x1 = clampd((u + mb * size2[1]) / ma, -size1[1], size1[1]); }
        scl3(vec1, axis1, x1); addTo3(vec1, pos1);
        scl3(vec2, axis2, x2); addTo3(vec2, pos2);
        return c_sphereSphere(o, 0, margin, vec1, size1[0], vec2, size2[0]);
      }
      int n = 0;
      double x1, x2;
      x1 = size1[1]; x2 = clampd((v - mb * size1[1]) / mc, -size2[1], size2[1]);
      scl3(vec1, axis1, x1); addTo3(vec1, pos1); scl3(vec2, axis2, x2); addTo3(vec2, pos2);
      n += c_sphereSphere(o, n, margin, vec1, size1[0], vec2, size2[0]);
      x1 = -size1[1]; x2 = clampd((v + mb * size1[1]) / mc, -size2[1], size2[1]);
      scl3(vec1, axis1, x1); addTo3(vec1, pos1); scl3(vec2, axis2, x2); addTo3(vec2, pos2);
      n += c_sphereSphere(o, n, margin, vec1, size1[0], vec2, size2[0]);
      if (n >= 2) return n;
      x2 = size2[1]; x1 = clampd((u - mb * size2[1]) / ma, -size1[1], size1[1]);
      scl3(vec1, axis1, x1); addTo3(vec1, pos1); scl3(vec2, axis2, x2); addTo3(vec2, pos2);
      n += c_sphereSphere(o, n, margin, vec1, size1[0], vec2, size2[0]);
      if (n >= 2) return n;
      x2 = -size2[1]; x1 = clampd((u + mb * size2[1]) / ma, -size1[1], size1[1]);
      scl3(vec1, axis1, x1); addTo3(vec1, pos1); scl3(vec2, axis2, x2); addTo3(vec2, pos2);
      n += c_sphereSphere(o, n, margin, vec1, size1[0], vec2, size2[0]);
      return n;
    }
    if (t2 == B2MJ_GEOM_BOX) {
      double axis[3] = {mat1[2], mat1[5], mat1[8]}, tmp[3], c[3], a[3], p[3];
      sub3(tmp, pos1, pos2);
      rotVecMatT(c, tmp, mat2);
      rotVecMatT(a, axis, mat2);
      const double h = size1[1];
      const double t = segmentBoxClosest(c, a, h, size2);
      int n = 0;
      scl3(p, axis, t); addTo3(p, pos1);
      n += c_sphereBox(o, n, margin, p, size1[0], pos2, mat2, size2);
      const double tb = t >= 0 ? -h : h;
      if (fabs(tb - t) > 1e-3 * h) {
        scl3(p, axis, tb); addTo3(p, pos1);
        n += c_sphereBox(o, n, margin, p, size1[0], pos2, mat2, size2);
      }
      return n;
    }
  }
  if (t1 == B2MJ_GEOM_BOX && t2 == B2MJ_GEOM_BOX) return c_boxBox(o, margin, pos1, mat1, size1, pos2, mat2, size2);
  if (t1 == B2MJ_GEOM_HFIELD) {
    if (t2 < B2MJ_GEOM_SPHERE || t2 > B2MJ_GEOM_MESH || m.geom_dataid[g1] < 0) return 0;
    return c_hfieldConvex(o, margin, pos1, mat1, m.geom_dataid[g1], cvxGeom(gxpos, gxmat, g2), m.geom_rbound[g2]);
  }
  // every other pair of convex geoms (anything with an ellipsoid, a cylinder or a mesh): the general MPR test
  if (t1 >= B2MJ_GEOM_SPHERE && t1 <= B2MJ_GEOM_MESH && t2 >= B2MJ_GEOM_SPHERE && t2 <= B2MJ_GEOM_MESH)
    return c_convexConvex(o, margin, cvxGeom(gxpos, gxmat, g1), cvxGeom(gxpos, gxmat, g2));
  return 0;
}

// mj_collision; returns ncon (also stored)
__device__ __noinline__ int stage_collision(const Env e, int* warning) {
  const DevModel& m = c_dm;
  int* ncon_p = e.I(B2MJ_F_NCON);
  if ((m.opt.disableflags & (B2MJ_DSBL_CONSTRAINT | B2MJ_DSBL_CONTACT)) || m.nconmax == 0 || m.ncollpair == 0) {
    if (e.lane == 0) ncon_p[0] = 0;
    WSYNC();
    return 0;
  }
  const double* gxpos = e.D(B2MJ_F_GEOM_XPOS);
  const double* gxmat = e.D(B2MJ_F_GEOM_XMAT);
  double* c_dist = e.DG(B2MJ_F_CONTACT_DIST);
  double* c_pos = e.DG(B2MJ_F_CONTACT_POS);
  double* c_frame = e.DG(B2MJ_F_CONTACT_FRAME);
  double* c_inc = e.DG(B2MJ_F_CONTACT_INCLUDEMARGIN);
  double* c_fri = e.DG(B2MJ_F_CONTACT_FRICTION);
  double* c_solref = e.DG(B2MJ_F_CONTACT_SOLREF);
  double* c_solimp = e.DG(B2MJ_F_CONTACT_SOLIMP);
  double* c_mu = e.DG(B2MJ_F_CONTACT_MU);
  int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
  int* c_g1 = e.IG(B2MJ_F_CONTACT_GEOM1);
  int* c_g2 = e.IG(B2MJ_F_CONTACT_GEOM2);
  int* c_excl = e.IG(B2MJ_F_CONTACT_EXCLUDE);
  int* c_adr = e.IG(B2MJ_F_CONTACT_EFC_ADDRESS);
  int carry = 0, overflow = 0;
  B2K_NOUNROLL for (int base = 0; base < m.ncollpair; base += B2K_G) {
    const int p = base + e.lane;
    PairCon pc;
    pc.shared_frame = 0;
    int num = 0, g1 = 0, g2 = 0, pid = -1;
    double margin = 0, gap = 0;
    if (p < m.ncollpair) {
      g1 = m.collpair_geom1[p];
      g2 = m.collpair_geom2[p];
      pid = m.collpair_pairid[p];  // explicit <contact><pair>: its own margin / gap / contact parameters
      margin = pid >= 0 ? m.pair_margin[pid] : fmax(m.geom_margin[g1], m.geom_margin[g2]);
      gap = pid >= 0 ? m.pair_gap[pid] : fmax(m.geom_gap[g1], m.geom_gap[g2]);
      const double r1 = m.geom_rbound[g1], r2 = m.geom_rbound[g2];
      bool cull = false;
      if (r1 > 0 && r2 > 0) {
        double dif[3];
        sub3(dif, gxpos + 3 * g1, gxpos + 3 * g2);
        const double bound = r1 + r2 + margin;
        cull = dot3(dif, dif) > bound * bound;
      } else if (m.geom_type[g1] == B2MJ_GEOM_PLANE && r2 > 0) {
        const double* mat1 = gxmat + 9 * g1;
        double normal[3] = {mat1[2], mat1[5], mat1[8]}, dif[3];
        sub3(dif, gxpos + 3 * g2, gxpos + 3 * g1);
        cull = dot3(dif, normal) > margin + r2;
      }
      if (!cull) num = narrowphase(gxpos, gxmat, pc, g1, g2, margin);
    }
    const int incl = warpInclusiveScan(e.mask, num, e.lane);
    const int total = __shfl_sync(e.mask, incl, B2K_G - 1, B2K_G);
    if (num > 0) {
      int condim;
      double solref[2], solimp[5], fri[3], fri5[5];
      if (pid >= 0) {
        condim = m.pair_dim[pid];
        for (int i = 0; i < 2; i++) solref[i] = m.pair_solref[2 * pid + i];
        for (int i = 0; i < 5; i++) solimp[i] = m.pair_solimp[5 * pid + i];
        for (int i = 0; i < 5; i++) fri5[i] = m.pair_friction[5 * pid + i];
      } else if (m.geom_priority[g1] != m.geom_priority[g2]) {
        const int gi = m.geom_priority[g1] > m.geom_priority[g2] ? g1 : g2;
        condim = m.geom_condim[gi];
        for (int i = 0; i < 2; i++) solref[i] = m.geom_solref[2 * gi + i];
        for (int i = 0; i < 5; i++) solimp[i] = m.geom_solimp[5 * gi + i];
        for (int i = 0; i < 3; i++) fri[i] = m.geom_friction[3 * gi + i];
      } else {
        condim = max(m.geom_condim[g1], m.geom_condim[g2]);
        const double s1 = m.geom_solmix[g1], s2 = m.geom_solmix[g2];
        double mix;
        if (s1 >= B2K_MINVAL && s2 >= B2K_MINVAL) mix = s1 / (s1 + s2);
        else if (s1 < B2K_MINVAL && s2 < B2K_MINVAL) mix = 0.5;
        else if (s1 < B2K_MINVAL) mix = 0.0;
        else mix = 1.0;
        const double *a = m.geom_solref + 2 * g1, *b = m.geom_solref + 2 * g2;
        if (a[0] > 0 && b[0] > 0) for (int i = 0; i < 2; i++) solref[i] = mix * a[i] + (1 - mix) * b[i];
        else for (int i = 0; i < 2; i++) solref[i] = fmin(a[i], b[i]);
        for (int i = 0; i < 5; i++) solimp[i] = mix * m.geom_solimp[5 * g1 + i] + (1 - mix) * m.geom_solimp[5 * g2 + i];
        for (int i = 0; i < 3; i++) fri[i] = fmax(m.geom_friction[3 * g1 + i], m.geom_friction[3 * g2 + i]);
      }
      const int first = carry + incl - num;
      B2K_NOUNROLL for (int i = 0; i < num; i++) {
        const int c = first + i;
        if (c >= m.nconmax) { overflow = 1; break; }
        c_dist[c] = pc.dist[i];
        copy3(c_pos + 3 * c, pc.pos + 3 * i);
        double fr[9];
        for (int k = 0; k < 6; k++) fr[k] = pc.frame[6 * (pc.shared_frame ? 0 : i) + k];
        fr[6] = 0; fr[7] = 0; fr[8] = 0;
        makeFrame(fr);
        for (int k = 0; k < 9; k++) c_frame[9 * c + k] = fr[k];
        const double inc = margin - gap;
        c_inc[c] = inc;
        double* f = c_fri + 5 * c;
        if (pid >= 0) {
          for (int k = 0; k < 5; k++) f[k] = fmax(B2MJ_MINMU, fri5[k]);
        } else {
          f[0] = f[1] = fmax(B2MJ_MINMU, fri[0]);
          f[2] = fmax(B2MJ_MINMU, fri[1]);
          f[3] = f[4] = fmax(B2MJ_MINMU, fri[2]);
        }
        c_solref[2 * c] = solref[0]; c_solref[2 * c + 1] = solref[1];
        for (int k = 0; k < 5; k++) c_solimp[5 * c + k] = solimp[k];
        c_mu[c] = 0;
        c_dim[c] = condim;
        c_g1[c] = g1;
        c_g2[c] = g2;
        c_excl[c] = pc.dist[i] >= inc;
        c_adr[c] = -1;
      }
    }
    carry += total;
  }
  if (__any_sync(e.mask, overflow) || carry > m.nconmax) {
    if (e.lane == 0) warning[B2MJ_WARN_CONTACTFULL]++;
    carry = min(carry, m.nconmax);
  }
  if (e.lane == 0) ncon_p[0] = carry;
  WSYNC();
  return carry;
}

}  // namespace b2k
