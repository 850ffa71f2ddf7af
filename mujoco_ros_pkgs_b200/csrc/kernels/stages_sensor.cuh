// stages_sensor.cuh — sensor evaluation (mj_sensorPos / Vel / Acc, mj_rnePostConstraint), one sensor per lane.
//
// Replaces row M10 of SURVEY 8(a): the values the reference's sensor plugin reads from d->sensordata
// (mujoco_ros_sensors/src/mujoco_sensor_handler_plugin.cpp:183-226).  Types: the 36 names that plugin
// knows (:70-105), rangefinder included (ray cast against the primitive geoms, mj_ray / mju_rayGeom).
#pragma once
#include "env_ctx.cuh"
#include "stages_smooth.cuh"

namespace b2k {

__device__ __forceinline__ void sens_cutoff(int i, double* out) {
  const DevModel& m = c_dm;
  const double cutoff = m.sensor_cutoff[i];
  if (cutoff <= 0) return;
  const int dt = m.sensor_datatype[i];
  B2K_NOUNROLL for (int k = 0; k < m.sensor_dim[i]; k++) {
    if (dt == B2MJ_DATATYPE_REAL) out[k] = clampd(out[k], -cutoff, cutoff);
    else if (dt == B2MJ_DATATYPE_POSITIVE) out[k] = fmin(cutoff, out[k]);
  }
}

__device__ __forceinline__ void sens_frame(const Env e, int type, int id, const double** pos, const double** mat,
                                           double* quat) {
  const DevModel& m = c_dm;
  const double* xquat = e.D(B2MJ_F_XQUAT);
  switch (type) {
    case B2MJ_OBJ_BODY:
      *pos = e.D(B2MJ_F_XIPOS) + 3 * id; *mat = e.DG(B2MJ_F_XIMAT) + 9 * id;
      mulQuat(quat, xquat + 4 * id, m.body_iquat + 4 * id);
      break;
    case B2MJ_OBJ_GEOM:
      *pos = e.D(B2MJ_F_GEOM_XPOS) + 3 * id; *mat = e.D(B2MJ_F_GEOM_XMAT) + 9 * id;
      mulQuat(quat, xquat + 4 * m.geom_bodyid[id], m.geom_quat + 4 * id);
      break;
    case B2MJ_OBJ_SITE:
      *pos = e.D(B2MJ_F_SITE_XPOS) + 3 * id; *mat = e.D(B2MJ_F_SITE_XMAT) + 9 * id;
      mulQuat(quat, xquat + 4 * m.site_bodyid[id], m.site_quat + 4 * id);
      break;
    default:
      *pos = e.D(B2MJ_F_XPOS) + 3 * id; *mat = e.D(B2MJ_F_XMAT) + 9 * id;
      copy4(quat, xquat + 4 * id);
  }
}

// ---- mj_ray / mju_rayGeom for the rangefinder (plane, sphere, capsule, ellipsoid, cylinder, box) ----
// smallest non-negative root of a x^2 + 2 b x + c = 0 (both roots in xx), -1 if none
__device__ __forceinline__ double rayQuad(double a, double b, double c, double* xx) {
  const double det0 = b * b - a * c;
  if (det0 < B2K_MINVAL) { xx[0] = -1; xx[1] = -1; return -1; }
  const double det = sqrt(det0);
  xx[0] = (-b - det) / a;
  xx[1] = (-b + det) / a;
  return xx[0] >= 0 ? xx[0] : (xx[1] >= 0 ? xx[1] : -1.0);
}
__device__ __noinline__ double rayGeom(const double* pos, const double* mat, const double* size, const double* pnt,
                                       const double* vec, int type) {
  double dif[3], lpnt[3], lvec[3], xx[2], x = -1, sol;
  sub3(dif, pnt, pos);
  rotVecMatT(lpnt, dif, mat);
  rotVecMatT(lvec, vec, mat);
#define B2K_RAY_BETTER(S) { const double _s = (S); if (_s >= 0 && (x < 0 || _s < x)) x = _s; }
  const double vv = dot3(lvec, lvec), vp = dot3(lvec, lpnt), pp = dot3(lpnt, lpnt);
  if (type == B2MJ_GEOM_PLANE) {
    if (lvec[2] > -B2K_MINVAL) return -1;
    sol = -lpnt[2] / lvec[2];
    if (sol < 0) return -1;
    const double p0 = lpnt[0] + sol * lvec[0], p1 = lpnt[1] + sol * lvec[1];
    return ((size[0] <= 0 || fabs(p0) <= size[0]) && (size[1] <= 0 || fabs(p1) <= size[1])) ? sol : -1.0;
  }
  if (type == B2MJ_GEOM_SPHERE) return rayQuad(vv, vp, pp - size[0] * size[0], xx);
  if (type == B2MJ_GEOM_ELLIPSOID) {
    const double s0 = 1 / (size[0] * size[0]), s1 = 1 / (size[1] * size[1]), s2 = 1 / (size[2] * size[2]);
    return rayQuad(s0 * lvec[0] * lvec[0] + s1 * lvec[1] * lvec[1] + s2 * lvec[2] * lvec[2],
                   s0 * lvec[0] * lpnt[0] + s1 * lvec[1] * lpnt[1] + s2 * lvec[2] * lpnt[2],
                   s0 * lpnt[0] * lpnt[0] + s1 * lpnt[1] * lpnt[1] + s2 * lpnt[2] * lpnt[2] - 1, xx);
  }
  const double a2 = lvec[0] * lvec[0] + lvec[1] * lvec[1], b2 = lvec[0] * lpnt[0] + lvec[1] * lpnt[1],
               c2 = lpnt[0] * lpnt[0] + lpnt[1] * lpnt[1] - size[0] * size[0];
  if (type == B2MJ_GEOM_CAPSULE) {
    const double ssz = size[0] + size[1];
    if (rayQuad(vv, vp, pp - ssz * ssz, xx) < 0) return -1;
    sol = rayQuad(a2, b2, c2, xx);
    if (sol >= 0 && fabs(lpnt[2] + sol * lvec[2]) <= size[1]) B2K_RAY_BETTER(sol)
    double ldif[3] = {lpnt[0], lpnt[1], lpnt[2] - size[1]};
    rayQuad(vv, dot3(lvec, ldif), dot3(ldif, ldif) - size[0] * size[0], xx);
    for (int i = 0; i < 2; i++)
      if (xx[i] >= 0 && lpnt[2] + xx[i] * lvec[2] >= size[1]) B2K_RAY_BETTER(xx[i])
    ldif[2] = lpnt[2] + size[1];
    rayQuad(vv, dot3(lvec, ldif), dot3(ldif, ldif) - size[0] * size[0], xx);
    for (int i = 0; i < 2; i++)
      if (xx[i] >= 0 && lpnt[2] + xx[i] * lvec[2] <= -size[1]) B2K_RAY_BETTER(xx[i])
    return x;
  }
  if (type == B2MJ_GEOM_CYLINDER) {
    if (rayQuad(vv, vp, pp - (size[0] * size[0] + size[1] * size[1]), xx) < 0) return -1;
    if (fabs(lvec[2]) > B2K_MINVAL)
      for (int side = -1; side <= 1; side += 2) {
        sol = (side * size[1] - lpnt[2]) / lvec[2];
        if (sol >= 0) {
          const double p0 = lpnt[0] + sol * lvec[0], p1 = lpnt[1] + sol * lvec[1];
          if (p0 * p0 + p1 * p1 <= size[0] * size[0]) B2K_RAY_BETTER(sol)
        }
      }
    sol = rayQuad(a2, b2, c2, xx);
    if (sol >= 0 && fabs(lpnt[2] + sol * lvec[2]) <= size[1]) B2K_RAY_BETTER(sol)
    return x;
  }
  if (type == B2MJ_GEOM_BOX) {
    if (rayQuad(vv, vp, pp - dot3(size, size), xx) < 0) return -1;
    B2K_NOUNROLL for (int i = 0; i < 3; i++) {
      if (fabs(lvec[i]) <= B2K_MINVAL) continue;
      const int i0 = (i + 1) % 3, i1 = (i + 2) % 3;
      for (int side = -1; side <= 1; side += 2) {
        sol = (side * size[i] - lpnt[i]) / lvec[i];
        if (sol >= 0) {
          const double p0 = lpnt[i0] + sol * lvec[i0], p1 = lpnt[i1] + sol * lvec[i1];
          if (fabs(p0) <= size[i0] && fabs(p1) <= size[i1]) B2K_RAY_BETTER(sol)
        }
      }
    }
    return x;
  }
#undef B2K_RAY_BETTER
  return -1;
}
// nearest hit over all geoms except those of bodyexclude and fully transparent ones; -1 if none (one lane, serial)
__device__ __noinline__ double rayScene(const Env e, const double* pnt, const double* vec, int bodyexclude) {
  const DevModel& m = c_dm;
  const double* gx = e.D(B2MJ_F_GEOM_XPOS);
  const double* gm = e.D(B2MJ_F_GEOM_XMAT);
  double best = -1;
  B2K_NOUNROLL for (int g = 0; g < m.ngeom; g++) {
    if (m.geom_bodyid[g] == bodyexclude || m.geom_rgba[4 * g + 3] == 0) continue;
    const double x = rayGeom(gx + 3 * g, gm + 9 * g, m.geom_size + 3 * g, pnt, vec, m.geom_type[g]);
    if (x >= 0 && (best < 0 || x < best)) best = x;
  }
  return best;
}

__device__ __forceinline__ int sens_body(int type, int id) {
  const DevModel& m = c_dm;
  return type == B2MJ_OBJ_GEOM ? m.geom_bodyid[id] : type == B2MJ_OBJ_SITE ? m.site_bodyid[id] : id;
}

// mj_objectVelocity / mj_objectAcceleration
__device__ __forceinline__ void objVelocity(const Env e, int type, int id, double* res, int local) {
  const DevModel& m = c_dm;
  const double *pos, *mat;
  double q[4];
  sens_frame(e, type, id, &pos, &mat, q);
  const int b = sens_body(type, id);
  transformSpatial(res, e.D(B2MJ_F_CVEL) + 6 * b, 0, pos, e.D(B2MJ_F_SUBTREE_COM) + 3 * m.body_rootid[b], local ? mat : nullptr);
}
__device__ __forceinline__ void objAcceleration(const Env e, int type, int id, double* res, int local) {
  const DevModel& m = c_dm;
  const double *pos, *mat;
  double q[4], vel[6], corr[3];
  sens_frame(e, type, id, &pos, &mat, q);
  const int b = sens_body(type, id);
  const double* com = e.D(B2MJ_F_SUBTREE_COM) + 3 * m.body_rootid[b];
  transformSpatial(res, e.DG(B2MJ_F_CACC) + 6 * b, 0, pos, com, local ? mat : nullptr);
  transformSpatial(vel, e.D(B2MJ_F_CVEL) + 6 * b, 0, pos, com, local ? mat : nullptr);
  cross(corr, vel, vel + 3);
  addTo3(res + 3, corr);
}

__device__ __forceinline__ int findLimitRow(const Env e, int nefc, int want, int id) {
  const int* type = e.IG(B2MJ_F_EFC_TYPE);
  const int* eid = e.IG(B2MJ_F_EFC_ID);
  B2K_NOUNROLL for (int r = 0; r < nefc; r++)
    if (type[r] == want && eid[r] == id) return r;
  return -1;
}

__device__ __noinline__ void stage_sensorPos(const Env e, int nefc) {
  const DevModel& m = c_dm;
  if (!m.nsensor || (m.opt.disableflags & B2MJ_DSBL_SENSOR)) return;
  double* sd = e.D(B2MJ_F_SENSORDATA);
  const double* qpos = e.D(B2MJ_F_QPOS);
  FORL(i, m.nsensor) {
    if (m.sensor_needstage[i] != B2MJ_STAGE_POS) continue;
    const int type = m.sensor_type[i], objid = m.sensor_objid[i];
    double* out = sd + m.sensor_adr[i];
    switch (type) {
      case B2MJ_SENS_MAGNETOMETER: rotVecMatT(out, m.opt.magnetic, e.D(B2MJ_F_SITE_XMAT) + 9 * objid); break;
      case B2MJ_SENS_RANGEFINDER: {  // ray along the site's z axis, the site's own body excluded
        const double* sm = e.D(B2MJ_F_SITE_XMAT) + 9 * objid;
        const double rvec[3] = {sm[2], sm[5], sm[8]};
        out[0] = rayScene(e, e.D(B2MJ_F_SITE_XPOS) + 3 * objid, rvec, m.site_bodyid[objid]);
        break;
      }
      case B2MJ_SENS_JOINTPOS: out[0] = qpos[m.jnt_qposadr[objid]]; break;
      case B2MJ_SENS_TENDONPOS: out[0] = e.D(B2MJ_F_TEN_LENGTH)[objid]; break;
      case B2MJ_SENS_ACTUATORPOS: out[0] = e.D(B2MJ_F_ACTUATOR_LENGTH)[objid]; break;
      case B2MJ_SENS_BALLQUAT: copy4(out, qpos + m.jnt_qposadr[objid]); normalize4(out); break;
      case B2MJ_SENS_JOINTLIMITPOS:
      case B2MJ_SENS_TENDONLIMITPOS: {
        const int r = findLimitRow(e, nefc, type == B2MJ_SENS_JOINTLIMITPOS ? B2MJ_CNSTR_LIMIT_JOINT : B2MJ_CNSTR_LIMIT_TENDON, objid);
        out[0] = r < 0 ? 0.0 : e.DG(B2MJ_F_EFC_POS)[r] - e.DG(B2MJ_F_EFC_MARGIN)[r];
        break;
      }
      case B2MJ_SENS_FRAMEPOS:
      case B2MJ_SENS_FRAMEQUAT:
      case B2MJ_SENS_FRAMEXAXIS:
      case B2MJ_SENS_FRAMEYAXIS:
      case B2MJ_SENS_FRAMEZAXIS: {
        const double *xp, *xm, *rp = nullptr, *rm = nullptr;
        double xq[4], rq[4];
        sens_frame(e, m.sensor_objtype[i], objid, &xp, &xm, xq);
        const int refid = m.sensor_refid[i];
        if (refid >= 0) sens_frame(e, m.sensor_reftype[i], refid, &rp, &rm, rq);
        if (type == B2MJ_SENS_FRAMEPOS) {
          if (refid < 0) copy3(out, xp);
          else { double dif[3]; sub3(dif, xp, rp); rotVecMatT(out, dif, rm); }
        } else if (type == B2MJ_SENS_FRAMEQUAT) {
          if (refid < 0) copy4(out, xq);
          else { double neg[4]; negQuat(neg, rq); mulQuat(out, neg, xq); }
          normalize4(out);
        } else {
          const int k = type - B2MJ_SENS_FRAMEXAXIS;
          double axis[3] = {xm[k], xm[k + 3], xm[k + 6]};
          if (refid < 0) copy3(out, axis);
          else rotVecMatT(out, axis, rm);
        }
        break;
      }
      case B2MJ_SENS_SUBTREECOM: copy3(out, e.D(B2MJ_F_SUBTREE_COM) + 3 * objid); break;
      case B2MJ_SENS_CLOCK: out[0] = e.D(B2MJ_F_TIME)[0]; break;
      default: break;
    }
    sens_cutoff(i, out);
  }
  WSYNC();
}

// mj_subtreeVel (serial over bodies on lane 0; only runs when a subtree sensor exists)
__device__ __noinline__ void subtreeVel_lane0(const Env e) {
  const DevModel& m = c_dm;
  if (e.lane == 0) {
    double* linvel = e.X(XF_SUBTREE_LINVEL);
    double* angmom = e.X(XF_SUBTREE_ANGMOM);
    double* bodyvel = e.X(XF_BODYVEL);
    const double* ximat = e.DG(B2MJ_F_XIMAT);
    const double* xipos = e.D(B2MJ_F_XIPOS);
    const double* com = e.D(B2MJ_F_SUBTREE_COM);
    const int nb = m.nbody;
    B2K_NOUNROLL for (int i = 0; i < nb; i++) {
      objVelocity(e, B2MJ_OBJ_BODY, i, bodyvel + 6 * i, 0);
      scl3(linvel + 3 * i, bodyvel + 6 * i + 3, m.body_mass[i]);
      double dv[3];
      rotVecMatT(dv, bodyvel + 6 * i, ximat + 9 * i);
      dv[0] *= m.body_inertia[3 * i]; dv[1] *= m.body_inertia[3 * i + 1]; dv[2] *= m.body_inertia[3 * i + 2];
      rotVecMat(angmom + 3 * i, dv, ximat + 9 * i);
    }
    for (int i = nb - 1; i >= 0; i--) {
      if (i) addTo3(linvel + 3 * m.body_parentid[i], linvel + 3 * i);
      scl3(linvel + 3 * i, linvel + 3 * i, 1 / fmax(B2K_MINVAL, m.body_subtreemass[i]));
    }
    for (int i = nb - 1; i > 0; i--) {
      const int p = m.body_parentid[i];
      double dx[3], dv[3], dp[3], dL[3];
      sub3(dx, xipos + 3 * i, com + 3 * i);
      sub3(dv, bodyvel + 6 * i + 3, linvel + 3 * i);
      scl3(dp, dv, m.body_mass[i]);
      cross(dL, dx, dp);
      addTo3(angmom + 3 * i, dL);
      addTo3(angmom + 3 * p, angmom + 3 * i);
      sub3(dx, com + 3 * i, com + 3 * p);
      sub3(dv, linvel + 3 * i, linvel + 3 * p);
      scl3(dv, dv, m.body_subtreemass[i]);
      cross(dL, dx, dv);
      addTo3(angmom + 3 * p, dL);
    }
  }
  WSYNC();
}

__device__ __noinline__ void stage_sensorVel(const Env e, int nefc) {
  const DevModel& m = c_dm;
  if (!m.nsensor || (m.opt.disableflags & B2MJ_DSBL_SENSOR)) return;
  if (m.need_subtreevel) subtreeVel_lane0(e);
  double* sd = e.D(B2MJ_F_SENSORDATA);
  const double* qvel = e.D(B2MJ_F_QVEL);
  FORL(i, m.nsensor) {
    if (m.sensor_needstage[i] != B2MJ_STAGE_VEL) continue;
    const int type = m.sensor_type[i], objid = m.sensor_objid[i];
    double* out = sd + m.sensor_adr[i];
    double tmp[6];
    switch (type) {
      case B2MJ_SENS_VELOCIMETER: objVelocity(e, B2MJ_OBJ_SITE, objid, tmp, 1); copy3(out, tmp + 3); break;
      case B2MJ_SENS_GYRO: objVelocity(e, B2MJ_OBJ_SITE, objid, tmp, 1); copy3(out, tmp); break;
      case B2MJ_SENS_JOINTVEL: out[0] = qvel[m.jnt_dofadr[objid]]; break;
      case B2MJ_SENS_TENDONVEL: out[0] = e.D(B2MJ_F_TEN_VELOCITY)[objid]; break;
      case B2MJ_SENS_ACTUATORVEL: out[0] = e.D(B2MJ_F_ACTUATOR_VELOCITY)[objid]; break;
      case B2MJ_SENS_BALLANGVEL: copy3(out, qvel + m.jnt_dofadr[objid]); break;
      case B2MJ_SENS_JOINTLIMITVEL:
      case B2MJ_SENS_TENDONLIMITVEL: {
        const int r = findLimitRow(e, nefc, type == B2MJ_SENS_JOINTLIMITVEL ? B2MJ_CNSTR_LIMIT_JOINT : B2MJ_CNSTR_LIMIT_TENDON, objid);
        out[0] = r < 0 ? 0.0 : e.DG(B2MJ_F_EFC_VEL)[r];
        break;
      }
      case B2MJ_SENS_FRAMELINVEL:
      case B2MJ_SENS_FRAMEANGVEL: {
        objVelocity(e, m.sensor_objtype[i], objid, tmp, 0);
        const int refid = m.sensor_refid[i];
        if (refid >= 0) {
          const double *xp, *xm, *rp, *rm;
          double q[4], rvel[6], rel[3], cr[3], dif[3];
          sens_frame(e, m.sensor_objtype[i], objid, &xp, &xm, q);
          sens_frame(e, m.sensor_reftype[i], refid, &rp, &rm, q);
          objVelocity(e, m.sensor_reftype[i], refid, rvel, 0);
          if (type == B2MJ_SENS_FRAMELINVEL) {
            sub3(rel, tmp + 3, rvel + 3);
            sub3(dif, xp, rp);
            cross(cr, rvel, dif);
            sub3(rel, rel, cr);
            rotVecMatT(out, rel, rm);
          } else {
            sub3(rel, tmp, rvel);
            rotVecMatT(out, rel, rm);
          }
        } else {
          copy3(out, type == B2MJ_SENS_FRAMELINVEL ? tmp + 3 : tmp);
        }
        break;
      }
      case B2MJ_SENS_SUBTREELINVEL: copy3(out, e.X(XF_SUBTREE_LINVEL) + 3 * objid); break;
      case B2MJ_SENS_SUBTREEANGMOM: copy3(out, e.X(XF_SUBTREE_ANGMOM) + 3 * objid); break;
      default: break;
    }
    sens_cutoff(i, out);
  }
  WSYNC();
}

// local contact force [normal, tangents..., torques...] of contact c
__device__ __forceinline__ void contactForce(const Env e, int c, double* lfrc) {
  const int adr = e.IG(B2MJ_F_CONTACT_EFC_ADDRESS)[c];
  const int dim = e.IG(B2MJ_F_CONTACT_DIM)[c];
  const double* f = e.DG(B2MJ_F_EFC_FORCE);
  for (int k = 0; k < 6; k++) lfrc[k] = 0;
  if (adr < 0) return;
  if (e.IG(B2MJ_F_EFC_TYPE)[adr] == B2MJ_CNSTR_CONTACT_PYRAMIDAL) {
    const double* mu = e.DG(B2MJ_F_CONTACT_FRICTION) + 5 * c;
    B2K_NOUNROLL for (int k = 0; k < 2 * (dim - 1); k++) lfrc[0] += f[adr + k];
    B2K_NOUNROLL for (int k = 1; k < dim; k++) lfrc[k] = (f[adr + 2 * (k - 1)] - f[adr + 2 * (k - 1) + 1]) * mu[k - 1];
  } else {
    B2K_NOUNROLL for (int k = 0; k < dim; k++) lfrc[k] = f[adr + k];
  }
}

// mj_rnePostConstraint: cacc, cfrc_int, cfrc_ext
__device__ __noinline__ void stage_rnePost(const Env e, int ncon, const double* xfrc) {
  const DevModel& m = c_dm;
  const double* cdof = e.D(B2MJ_F_CDOF);
  const double* cdof_dot = e.D(B2MJ_F_CDOF_DOT);
  const double* cvel = e.D(B2MJ_F_CVEL);
  const double* cinert = e.D(B2MJ_F_CINERT);
  const double* qvel = e.D(B2MJ_F_QVEL);
  const double* qacc = e.D(B2MJ_F_QACC);
  const double* com = e.D(B2MJ_F_SUBTREE_COM);
  double* cacc = e.DG(B2MJ_F_CACC);
  double* cint = e.DG(B2MJ_F_CFRC_INT);
  double* cext = e.DG(B2MJ_F_CFRC_EXT);
  // external forces, one lane per body (contacts scanned in order => deterministic sums)
  FORL(b, m.nbody) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (b > 0) {
      if (xfrc) {
        const double* x = xfrc + 6 * b;
        if (!(x[0] == 0 && x[1] == 0 && x[2] == 0 && x[3] == 0 && x[4] == 0 && x[5] == 0)) {
          double corr[6] = {x[3], x[4], x[5], x[0], x[1], x[2]}, f[6];
          transformSpatial(f, corr, 1, com + 3 * m.body_rootid[b], e.D(B2MJ_F_XIPOS) + 3 * b, nullptr);
          for (int k = 0; k < 6; k++) acc[k] += f[k];
        }
      }
      const int* g1 = e.IG(B2MJ_F_CONTACT_GEOM1);
      const int* g2 = e.IG(B2MJ_F_CONTACT_GEOM2);
      const int* cadr = e.IG(B2MJ_F_CONTACT_EFC_ADDRESS);
      B2K_NOUNROLL for (int c = 0; c < ncon; c++) {
        if (cadr[c] < 0) continue;
        const int b1 = m.geom_bodyid[g1[c]], b2 = m.geom_bodyid[g2[c]];
        if (b1 != b && b2 != b) continue;
        double lfrc[6], cf[6], f[6];
        contactForce(e, c, lfrc);
        const double* fr = e.DG(B2MJ_F_CONTACT_FRAME) + 9 * c;
        rotVecMatT(cf + 3, lfrc, fr);
        rotVecMatT(cf, lfrc + 3, fr);
        transformSpatial(f, cf, 1, com + 3 * m.body_rootid[b], e.DG(B2MJ_F_CONTACT_POS) + 3 * c, nullptr);
        if (b1 == b) for (int k = 0; k < 6; k++) acc[k] -= f[k];
        if (b2 == b) for (int k = 0; k < 6; k++) acc[k] += f[k];
      }
    }
    for (int k = 0; k < 6; k++) cext[6 * b + k] = acc[k];
  }
  if (e.lane < 6) {
    double g = 0;
    if (e.lane >= 3 && !(m.opt.disableflags & B2MJ_DSBL_GRAVITY)) g = -m.env_gravity[e.lane - 3];
    cacc[e.lane] = g;
    cint[e.lane] = 0;
  }
  WSYNC();
  B2K_NOUNROLL for (int l = 1; l < m.nlevel; l++) {
    const int ladr = m.level_bodyadr[l], lnum = m.level_bodynum[l];
    FORL(k, lnum) {
      const int i = m.level_body[ladr + k];
      const int bda = m.body_dofadr[i], dn = m.body_dofnum[i];
      double tmp[6], tmp1[6], acc[6], body[6];
      mulDofVec(tmp, cdof_dot + 6 * bda, qvel + bda, dn);
      for (int c = 0; c < 6; c++) acc[c] = cacc[6 * m.body_parentid[i] + c] + tmp[c];
      mulDofVec(tmp, cdof + 6 * bda, qacc + bda, dn);
      for (int c = 0; c < 6; c++) acc[c] += tmp[c];
      for (int c = 0; c < 6; c++) cacc[6 * i + c] = acc[c];
      mulInertVec(body, cinert + 10 * i, acc);
      mulInertVec(tmp, cinert + 10 * i, cvel + 6 * i);
      crossForce(tmp1, cvel + 6 * i, tmp);
      for (int c = 0; c < 6; c++) cint[6 * i + c] = body[c] + tmp1[c] - cext[6 * i + c];
    }
    WSYNC();
  }
  if (e.lane < 6) {
    const int c = e.lane;
    for (int i = m.nbody - 1; i > 0; i--) {
      const int p = m.body_parentid[i];
      if (p) cint[6 * p + c] += cint[6 * i + c];
    }
  }
  WSYNC();
}

__device__ __forceinline__ bool pointInSite(const Env e, int site, const double* p) {
  const DevModel& m = c_dm;
  double dif[3], loc[3];
  sub3(dif, p, e.D(B2MJ_F_SITE_XPOS) + 3 * site);
  rotVecMatT(loc, dif, e.D(B2MJ_F_SITE_XMAT) + 9 * site);
  const double* s = m.site_size + 3 * site;
  switch (m.site_type[site]) {
    case B2MJ_GEOM_SPHERE: return dot3(loc, loc) <= s[0] * s[0];
    case B2MJ_GEOM_BOX: return fabs(loc[0]) <= s[0] && fabs(loc[1]) <= s[1] && fabs(loc[2]) <= s[2];
    case B2MJ_GEOM_CAPSULE: {
      const double z = clampd(loc[2], -s[1], s[1]);
      return loc[0] * loc[0] + loc[1] * loc[1] + (loc[2] - z) * (loc[2] - z) <= s[0] * s[0];
    }
    case B2MJ_GEOM_CYLINDER: return loc[0] * loc[0] + loc[1] * loc[1] <= s[0] * s[0] && fabs(loc[2]) <= s[1];
    case B2MJ_GEOM_ELLIPSOID:
      return (loc[0] / s[0]) * (loc[0] / s[0]) + (loc[1] / s[1]) * (loc[1] / s[1]) + (loc[2] / s[2]) * (loc[2] / s[2]) <= 1;
    default: return false;
  }
}

__device__ __noinline__ void stage_sensorAcc(const Env e, int nefc, int ncon, const double* xfrc) {
  const DevModel& m = c_dm;
  if (!m.nsensor || (m.opt.disableflags & B2MJ_DSBL_SENSOR)) return;
  if (m.need_rnepost) stage_rnePost(e, ncon, xfrc);
  double* sd = e.D(B2MJ_F_SENSORDATA);
  FORL(i, m.nsensor) {
    if (m.sensor_needstage[i] != B2MJ_STAGE_ACC) continue;
    const int type = m.sensor_type[i], objid = m.sensor_objid[i];
    double* out = sd + m.sensor_adr[i];
    double tmp[6];
    switch (type) {
      case B2MJ_SENS_TOUCH: {
        double s = 0;
        const int body = m.site_bodyid[objid];
        const int* g1 = e.IG(B2MJ_F_CONTACT_GEOM1);
        const int* g2 = e.IG(B2MJ_F_CONTACT_GEOM2);
        B2K_NOUNROLL for (int c = 0; c < ncon; c++) {
          if (e.IG(B2MJ_F_CONTACT_EFC_ADDRESS)[c] < 0) continue;
          const int b1 = m.geom_bodyid[g1[c]], b2 = m.geom_bodyid[g2[c]];
          if (b1 != body && b2 != body) continue;
          double lfrc[6];
          contactForce(e, c, lfrc);
          if (lfrc[0] <= 0) continue;
          if (pointInSite(e, objid, e.DG(B2MJ_F_CONTACT_POS) + 3 * c)) s += lfrc[0];
        }
        out[0] = s;
        break;
      }
      case B2MJ_SENS_ACCELEROMETER: objAcceleration(e, B2MJ_OBJ_SITE, objid, tmp, 1); copy3(out, tmp + 3); break;
      case B2MJ_SENS_FORCE:
      case B2MJ_SENS_TORQUE: {
        const int body = m.site_bodyid[objid];
        double f[6], dif[3], cr[3];
        const double* w = e.DG(B2MJ_F_CFRC_INT) + 6 * body;
        sub3(dif, e.D(B2MJ_F_SITE_XPOS) + 3 * objid, e.D(B2MJ_F_SUBTREE_COM) + 3 * m.body_rootid[body]);
        cross(cr, dif, w + 3);
        sub3(f, w, cr);
        copy3(f + 3, w + 3);
        rotVecMatT(out, type == B2MJ_SENS_FORCE ? f + 3 : f, e.D(B2MJ_F_SITE_XMAT) + 9 * objid);
        break;
      }
      case B2MJ_SENS_ACTUATORFRC: out[0] = e.D(B2MJ_F_ACTUATOR_FORCE)[objid]; break;
      case B2MJ_SENS_JOINTACTFRC: out[0] = e.D(B2MJ_F_QFRC_ACTUATOR)[m.jnt_dofadr[objid]]; break;
      case B2MJ_SENS_JOINTLIMITFRC:
      case B2MJ_SENS_TENDONLIMITFRC: {
        const int r = findLimitRow(e, nefc, type == B2MJ_SENS_JOINTLIMITFRC ? B2MJ_CNSTR_LIMIT_JOINT : B2MJ_CNSTR_LIMIT_TENDON, objid);
        out[0] = r < 0 ? 0.0 : e.DG(B2MJ_F_EFC_FORCE)[r];
        break;
      }
      case B2MJ_SENS_FRAMELINACC:
      case B2MJ_SENS_FRAMEANGACC:
        objAcceleration(e, m.sensor_objtype[i], objid, tmp, 0);
        copy3(out, type == B2MJ_SENS_FRAMELINACC ? tmp + 3 : tmp);
        break;
      default: break;
    }
    sens_cutoff(i, out);
  }
  WSYNC();
}

}  // namespace b2k
