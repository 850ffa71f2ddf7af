// orc_ray.cpp — ray casting of the CPU oracle (mj_ray / mju_rayGeom) for the rangefinder sensor.
// TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// Restates MuJoCo 2.3.7 engine_ray.c for the primitive geom types; the reference's sensor plugin publishes
// mjSENS_RANGEFINDER readings as scalars (mujoco_ros_sensors/src/mujoco_sensor_handler_plugin.cpp:77,331,569).
// "parity unpinned": MuJoCo's source is not in /root/reference.
#include <cmath>

#include "orc_math.h"
#include "orc_types.h"

namespace orc {

// smallest non-negative root of a x^2 + 2 b x + c = 0 (both roots in xx), -1 if none
static double rayQuad(double a, double b, double c, double* xx) {
  const double det0 = b * b - a * c;
  if (det0 < MINVAL) {
    xx[0] = xx[1] = -1;
    return -1;
  }
  const double det = std::sqrt(det0);
  xx[0] = (-b - det) / a;
  xx[1] = (-b + det) / a;
  if (xx[0] >= 0) return xx[0];
  if (xx[1] >= 0) return xx[1];
  return -1;
}

// mju_rayGeom: distance along vec from pnt to the surface of one primitive geom, -1 if missed
double rayGeom(const double* pos, const double* mat, const double* size, const double* pnt, const double* vec, int type) {
  double dif[3], lpnt[3], lvec[3], xx[2], x = -1, sol;
  sub3(dif, pnt, pos);
  rotVecMatT(lpnt, dif, mat);
  rotVecMatT(lvec, vec, mat);
  auto better = [&](double s) { if (s >= 0 && (x < 0 || s < x)) x = s; };
  switch (type) {
    case B2MJ_GEOM_PLANE: {
      if (lvec[2] > -MINVAL) return -1;  // not pointing at the front face
      sol = -lpnt[2] / lvec[2];
      if (sol < 0) return -1;
      const double p0 = lpnt[0] + sol * lvec[0], p1 = lpnt[1] + sol * lvec[1];
      if ((size[0] <= 0 || std::fabs(p0) <= size[0]) && (size[1] <= 0 || std::fabs(p1) <= size[1])) return sol;
      return -1;
    }
    case B2MJ_GEOM_SPHERE:
      return rayQuad(dot3(lvec, lvec), dot3(lvec, lpnt), dot3(lpnt, lpnt) - size[0] * size[0], xx);
    case B2MJ_GEOM_CAPSULE: {
      const double ssz = size[0] + size[1];
      if (rayQuad(dot3(lvec, lvec), dot3(lvec, lpnt), dot3(lpnt, lpnt) - ssz * ssz, xx) < 0) return -1;
      // round side, between the flat ends
      sol = rayQuad(lvec[0] * lvec[0] + lvec[1] * lvec[1], lvec[0] * lpnt[0] + lvec[1] * lpnt[1],
                    lpnt[0] * lpnt[0] + lpnt[1] * lpnt[1] - size[0] * size[0], xx);
      if (sol >= 0 && std::fabs(lpnt[2] + sol * lvec[2]) <= size[1]) better(sol);
      // top cap: upper half of the sphere at +size[1]
      double ldif[3] = {lpnt[0], lpnt[1], lpnt[2] - size[1]};
      rayQuad(dot3(lvec, lvec), dot3(lvec, ldif), dot3(ldif, ldif) - size[0] * size[0], xx);
      for (int i = 0; i < 2; i++)
        if (xx[i] >= 0 && lpnt[2] + xx[i] * lvec[2] >= size[1]) better(xx[i]);
      // bottom cap
      ldif[2] = lpnt[2] + size[1];
      rayQuad(dot3(lvec, lvec), dot3(lvec, ldif), dot3(ldif, ldif) - size[0] * size[0], xx);
      for (int i = 0; i < 2; i++)
        if (xx[i] >= 0 && lpnt[2] + xx[i] * lvec[2] <= -size[1]) better(xx[i]);
      return x;
    }
    case B2MJ_GEOM_ELLIPSOID: {
      const double s[3] = {1 / (size[0] * size[0]), 1 / (size[1] * size[1]), 1 / (size[2] * size[2])};
      return rayQuad(s[0] * lvec[0] * lvec[0] + s[1] * lvec[1] * lvec[1] + s[2] * lvec[2] * lvec[2],
                     s[0] * lvec[0] * lpnt[0] + s[1] * lvec[1] * lpnt[1] + s[2] * lvec[2] * lpnt[2],
                     s[0] * lpnt[0] * lpnt[0] + s[1] * lpnt[1] * lpnt[1] + s[2] * lpnt[2] * lpnt[2] - 1, xx);
    }
    case B2MJ_GEOM_CYLINDER: {
      const double ssz = size[0] * size[0] + size[1] * size[1];
      if (rayQuad(dot3(lvec, lvec), dot3(lvec, lpnt), dot3(lpnt, lpnt) - ssz, xx) < 0) return -1;
      if (std::fabs(lvec[2]) > MINVAL)
        for (int side = -1; side <= 1; side += 2) {
          sol = (side * size[1] - lpnt[2]) / lvec[2];
          if (sol >= 0) {
            const double p0 = lpnt[0] + sol * lvec[0], p1 = lpnt[1] + sol * lvec[1];
            if (p0 * p0 + p1 * p1 <= size[0] * size[0]) better(sol);
          }
        }
      sol = rayQuad(lvec[0] * lvec[0] + lvec[1] * lvec[1], lvec[0] * lpnt[0] + lvec[1] * lpnt[1],
                    lpnt[0] * lpnt[0] + lpnt[1] * lpnt[1] - size[0] * size[0], xx);
      if (sol >= 0 && std::fabs(lpnt[2] + sol * lvec[2]) <= size[1]) better(sol);
      return x;
    }
    case B2MJ_GEOM_BOX: {
      const double ssz = dot3(size, size);
      if (rayQuad(dot3(lvec, lvec), dot3(lvec, lpnt), dot3(lpnt, lpnt) - ssz, xx) < 0) return -1;
      for (int i = 0; i < 3; i++) {
        if (std::fabs(lvec[i]) <= MINVAL) continue;
        const int i0 = (i + 1) % 3, i1 = (i + 2) % 3;
        for (int side = -1; side <= 1; side += 2) {
          sol = (side * size[i] - lpnt[i]) / lvec[i];
          if (sol >= 0) {
            const double p0 = lpnt[i0] + sol * lvec[i0], p1 = lpnt[i1] + sol * lvec[i1];
            if (std::fabs(p0) <= size[i0] && std::fabs(p1) <= size[i1]) better(sol);
          }
        }
      }
      return x;
    }
    default:
      return -1;
  }
}

// mj_ray with geomgroup = NULL, flg_static = 1: nearest hit over all geoms except those of bodyexclude and fully
// transparent ones (rgba alpha == 0); -1 if nothing is hit
double ray(const b2mjModel* m, const OrcData* d, const double* pnt, const double* vec, int bodyexclude, int* geomid) {
  double best = -1;
  if (geomid) *geomid = -1;
  for (int g = 0; g < m->ngeom; g++) {
    if (m->geom_bodyid[g] == bodyexclude) continue;
    if (m->geom_rgba[4 * g + 3] == 0) continue;
    const double x = rayGeom(d->geom_xpos + 3 * g, d->geom_xmat + 9 * g, m->geom_size + 3 * g, pnt, vec, m->geom_type[g]);
    if (x >= 0 && (best < 0 || x < best)) {
      best = x;
      if (geomid) *geomid = g;
    }
  }
  return best;
}

}  // namespace orc

extern "C" double orc_ray(const b2mjModel* m, const OrcData* d, const double* pnt, const double* vec, int bodyexclude,
                          int* geomid) {
  return orc::ray(m, d, pnt, vec, bodyexclude, geomid);
}
