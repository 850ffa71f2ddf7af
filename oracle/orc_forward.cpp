// orc_forward.cpp — pipeline driver of the CPU oracle: mj_forward, mj_step, integrators, sensors,
// state checks, mj_resetData.
// TEST INFRASTRUCTURE ONLY (see orc_math.h).  Restates MuJoCo 2.3.7 engine_forward.c /
// engine_sensor.c / engine_io.c(mj_resetData), i.e. the calls the reference makes at
// mujoco_env.cpp:498 (mj_step), :329/:621 (mj_forward) and :252 (mj_resetData); pipeline order as in
// SURVEY Appendix A, callback placement as reference mujoco_env.h:242-251.
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include <chrono>

#include "orc_math.h"
#include "orc_types.h"

namespace orc {

// mj_integratePos
void integratePos(const b2mjModel* m, double* qpos, const double* qvel, double dt) {
  for (int j = 0; j < m->njnt; j++) {
    int padr = m->jnt_qposadr[j], vadr = m->jnt_dofadr[j];
    switch (m->jnt_type[j]) {
      case B2MJ_JNT_FREE:
        for (int i = 0; i < 3; i++) qpos[padr + i] += dt * qvel[vadr + i];
        padr += 3; vadr += 3;
        [[fallthrough]];
      case B2MJ_JNT_BALL:
        quatIntegrate(qpos + padr, qvel + vadr, dt);
        break;
      default:
        qpos[padr] += dt * qvel[vadr];
    }
  }
}

static void apply_cutoff(const b2mjModel* m, OrcData* d, int i) {
  const double cutoff = m->sensor_cutoff[i];
  if (cutoff <= 0) return;
  const int adr = m->sensor_adr[i];
  for (int k = 0; k < m->sensor_dim[i]; k++) {
    if (m->sensor_datatype[i] == B2MJ_DATATYPE_REAL) d->sensordata[adr + k] = clampd(d->sensordata[adr + k], -cutoff, cutoff);
    else if (m->sensor_datatype[i] == B2MJ_DATATYPE_POSITIVE) d->sensordata[adr + k] = std::fmin(cutoff, d->sensordata[adr + k]);
  }
}

static void get_frame(const b2mjModel* m, const OrcData* d, int type, int id, const double** pos, const double** mat,
                      double* quat) {
  switch (type) {
    case B2MJ_OBJ_BODY:
      *pos = d->xipos + 3 * id; *mat = d->ximat + 9 * id;
      mulQuat(quat, d->xquat + 4 * id, m->body_iquat + 4 * id);
      break;
    case B2MJ_OBJ_GEOM:
      *pos = d->geom_xpos + 3 * id; *mat = d->geom_xmat + 9 * id;
      mulQuat(quat, d->xquat + 4 * m->geom_bodyid[id], m->geom_quat + 4 * id);
      break;
    case B2MJ_OBJ_SITE:
      *pos = d->site_xpos + 3 * id; *mat = d->site_xmat + 9 * id;
      mulQuat(quat, d->xquat + 4 * m->site_bodyid[id], m->site_quat + 4 * id);
      break;
    default:  // XBODY
      *pos = d->xpos + 3 * id; *mat = d->xmat + 9 * id;
      copy4(quat, d->xquat + 4 * id);
  }
}

// mj_sensorPos
void sensorPos(const b2mjModel* m, OrcData* d) {
  if (m->opt.disableflags & B2MJ_DSBL_SENSOR) return;
  for (int i = 0; i < m->nsensor; i++) {
    if (m->sensor_needstage[i] != B2MJ_STAGE_POS) continue;
    const int type = m->sensor_type[i], objid = m->sensor_objid[i], adr = m->sensor_adr[i];
    double* out = d->sensordata + adr;
    switch (type) {
      case B2MJ_SENS_MAGNETOMETER: rotVecMatT(out, m->opt.magnetic, d->site_xmat + 9 * objid); break;
      case B2MJ_SENS_RANGEFINDER: {  // ray along the site's z axis, the site's own body excluded
        const double* sm = d->site_xmat + 9 * objid;
        const double rvec[3] = {sm[2], sm[5], sm[8]};
        out[0] = ray(m, d, d->site_xpos + 3 * objid, rvec, m->site_bodyid[objid], nullptr);
        break;
      }
      case B2MJ_SENS_JOINTPOS: out[0] = d->qpos[m->jnt_qposadr[objid]]; break;
      case B2MJ_SENS_TENDONPOS: out[0] = d->ten_length[objid]; break;
      case B2MJ_SENS_ACTUATORPOS: out[0] = d->actuator_length[objid]; break;
      case B2MJ_SENS_BALLQUAT:
        copy4(out, d->qpos + m->jnt_qposadr[objid]);
        normalize4(out);
        break;
      case B2MJ_SENS_JOINTLIMITPOS:
      case B2MJ_SENS_TENDONLIMITPOS: {
        out[0] = 0;
        const int want = type == B2MJ_SENS_JOINTLIMITPOS ? B2MJ_CNSTR_LIMIT_JOINT : B2MJ_CNSTR_LIMIT_TENDON;
        for (int r = 0; r < d->nefc(); r++)
          if (d->efc_type[r] == want && d->efc_id[r] == objid) { out[0] = d->efc_pos[r] - d->efc_margin[r]; break; }
        break;
      }
      case B2MJ_SENS_FRAMEPOS:
      case B2MJ_SENS_FRAMEQUAT:
      case B2MJ_SENS_FRAMEXAXIS:
      case B2MJ_SENS_FRAMEYAXIS:
      case B2MJ_SENS_FRAMEZAXIS: {
        const double *xpos, *xmat, *rpos = nullptr, *rmat = nullptr;
        double xquat[4], rquat[4];
        get_frame(m, d, m->sensor_objtype[i], objid, &xpos, &xmat, xquat);
        const int refid = m->sensor_refid[i];
        if (refid >= 0) get_frame(m, d, m->sensor_reftype[i], refid, &rpos, &rmat, rquat);
        if (type == B2MJ_SENS_FRAMEPOS) {
          if (refid < 0) copy3(out, xpos);
          else { double dif[3]; sub3(dif, xpos, rpos); rotVecMatT(out, dif, rmat); }
        } else if (type == B2MJ_SENS_FRAMEQUAT) {
          if (refid < 0) copy4(out, xquat);
          else { double neg[4]; negQuat(neg, rquat); mulQuat(out, neg, xquat); }
          normalize4(out);
        } else {
          const int k = type - B2MJ_SENS_FRAMEXAXIS;
          double axis[3] = {xmat[k], xmat[k + 3], xmat[k + 6]};
          if (refid < 0) copy3(out, axis);
          else rotVecMatT(out, axis, rmat);
        }
        break;
      }
      case B2MJ_SENS_SUBTREECOM: copy3(out, d->subtree_com + 3 * objid); break;
      case B2MJ_SENS_CLOCK: out[0] = d->time[0]; break;
      default: break;
    }
    apply_cutoff(m, d, i);
  }
}

// mj_subtreeVel quantities, computed on demand
static void subtreeVel(const b2mjModel* m, const OrcData* d, std::vector<double>& linvel, std::vector<double>& angmom) {
  const int nb = m->nbody;
  linvel.assign(3 * nb, 0.0);
  angmom.assign(3 * nb, 0.0);
  std::vector<double> bodyvel(6 * nb, 0.0);
  for (int i = 0; i < nb; i++) {
    objectVelocity(m, d, B2MJ_OBJ_BODY, i, &bodyvel[6 * i], 0);
    scl3(&linvel[3 * i], &bodyvel[6 * i + 3], m->body_mass[i]);
    double dv[3], dl[3];
    rotVecMatT(dv, &bodyvel[6 * i], d->ximat + 9 * i);
    dv[0] *= m->body_inertia[3 * i]; dv[1] *= m->body_inertia[3 * i + 1]; dv[2] *= m->body_inertia[3 * i + 2];
    rotVecMat(dl, dv, d->ximat + 9 * i);
    copy3(&angmom[3 * i], dl);
  }
  for (int i = nb - 1; i >= 0; i--) {
    if (i) addTo3(&linvel[3 * m->body_parentid[i]], &linvel[3 * i]);
    scl3(&linvel[3 * i], &linvel[3 * i], 1 / std::fmax(MINVAL, m->body_subtreemass[i]));
  }
  for (int i = nb - 1; i > 0; i--) {
    const int p = m->body_parentid[i];
    double dx[3], dv[3], dp[3], dL[3];
    sub3(dx, d->xipos + 3 * i, d->subtree_com + 3 * i);
    sub3(dv, &bodyvel[6 * i + 3], &linvel[3 * i]);
    scl3(dp, dv, m->body_mass[i]);
    cross(dL, dx, dp);
    addTo3(&angmom[3 * i], dL);
    addTo3(&angmom[3 * p], &angmom[3 * i]);
    sub3(dx, d->subtree_com + 3 * i, d->subtree_com + 3 * p);
    sub3(dv, &linvel[3 * i], &linvel[3 * p]);
    scl3(dv, dv, m->body_subtreemass[i]);
    cross(dL, dx, dv);
    addTo3(&angmom[3 * p], dL);
  }
}

// mj_sensorVel
void sensorVel(const b2mjModel* m, OrcData* d) {
  if (m->opt.disableflags & B2MJ_DSBL_SENSOR) return;
  std::vector<double> linvel, angmom;
  bool have_subtree = false;
  for (int i = 0; i < m->nsensor; i++) {
    if (m->sensor_needstage[i] != B2MJ_STAGE_VEL) continue;
    const int type = m->sensor_type[i], objid = m->sensor_objid[i], adr = m->sensor_adr[i];
    double* out = d->sensordata + adr;
    double tmp[6];
    switch (type) {
      case B2MJ_SENS_VELOCIMETER:
        objectVelocity(m, d, B2MJ_OBJ_SITE, objid, tmp, 1);
        copy3(out, tmp + 3);
        break;
      case B2MJ_SENS_GYRO:
        objectVelocity(m, d, B2MJ_OBJ_SITE, objid, tmp, 1);
        copy3(out, tmp);
        break;
      case B2MJ_SENS_JOINTVEL: out[0] = d->qvel[m->jnt_dofadr[objid]]; break;
      case B2MJ_SENS_TENDONVEL: out[0] = d->ten_velocity[objid]; break;
      case B2MJ_SENS_ACTUATORVEL: out[0] = d->actuator_velocity[objid]; break;
      case B2MJ_SENS_BALLANGVEL: copy3(out, d->qvel + m->jnt_dofadr[objid]); break;
      case B2MJ_SENS_JOINTLIMITVEL:
      case B2MJ_SENS_TENDONLIMITVEL: {
        out[0] = 0;
        const int want = type == B2MJ_SENS_JOINTLIMITVEL ? B2MJ_CNSTR_LIMIT_JOINT : B2MJ_CNSTR_LIMIT_TENDON;
        for (int r = 0; r < d->nefc(); r++)
          if (d->efc_type[r] == want && d->efc_id[r] == objid) { out[0] = d->efc_vel[r]; break; }
        break;
      }
      case B2MJ_SENS_FRAMELINVEL:
      case B2MJ_SENS_FRAMEANGVEL: {
        objectVelocity(m, d, m->sensor_objtype[i], objid, tmp, 0);
        const int refid = m->sensor_refid[i];
        if (refid >= 0) {
          // velocity relative to, and expressed in, the reference frame
          const double *xpos, *xmat, *rpos, *rmat;
          double q[4], rvel[6], rel[3], cr[3], dif[3];
          get_frame(m, d, m->sensor_objtype[i], objid, &xpos, &xmat, q);
          get_frame(m, d, m->sensor_reftype[i], refid, &rpos, &rmat, q);
          objectVelocity(m, d, m->sensor_reftype[i], refid, rvel, 0);
          if (type == B2MJ_SENS_FRAMELINVEL) {
            sub3(rel, tmp + 3, rvel + 3);
            sub3(dif, xpos, rpos);
            cross(cr, rvel, dif);
            sub3(rel, rel, cr);
            rotVecMatT(out, rel, rmat);
          } else {
            sub3(rel, tmp, rvel);
            rotVecMatT(out, rel, rmat);
          }
        } else {
          copy3(out, type == B2MJ_SENS_FRAMELINVEL ? tmp + 3 : tmp);
        }
        break;
      }
      case B2MJ_SENS_SUBTREELINVEL:
      case B2MJ_SENS_SUBTREEANGMOM:
        if (!have_subtree) { subtreeVel(m, d, linvel, angmom); have_subtree = true; }
        copy3(out, type == B2MJ_SENS_SUBTREELINVEL ? &linvel[3 * objid] : &angmom[3 * objid]);
        break;
      default: break;
    }
    apply_cutoff(m, d, i);
  }
}

static bool point_in_site(const b2mjModel* m, const OrcData* d, int site, const double* p) {
  double dif[3], loc[3];
  sub3(dif, p, d->site_xpos + 3 * site);
  rotVecMatT(loc, dif, d->site_xmat + 9 * site);
  const double* s = m->site_size + 3 * site;
  switch (m->site_type[site]) {
    case B2MJ_GEOM_SPHERE: return dot3(loc, loc) <= s[0] * s[0];
    case B2MJ_GEOM_BOX: return std::fabs(loc[0]) <= s[0] && std::fabs(loc[1]) <= s[1] && std::fabs(loc[2]) <= s[2];
    case B2MJ_GEOM_CAPSULE: {
      double z = clampd(loc[2], -s[1], s[1]);
      return loc[0] * loc[0] + loc[1] * loc[1] + (loc[2] - z) * (loc[2] - z) <= s[0] * s[0];
    }
    case B2MJ_GEOM_CYLINDER: return loc[0] * loc[0] + loc[1] * loc[1] <= s[0] * s[0] && std::fabs(loc[2]) <= s[1];
    case B2MJ_GEOM_ELLIPSOID:
      return (loc[0] / s[0]) * (loc[0] / s[0]) + (loc[1] / s[1]) * (loc[1] / s[1]) + (loc[2] / s[2]) * (loc[2] / s[2]) <= 1;
    default: return false;
  }
}

// mj_sensorAcc
void sensorAcc(const b2mjModel* m, OrcData* d) {
  if (m->opt.disableflags & B2MJ_DSBL_SENSOR) return;
  bool need_rne = false;
  for (int i = 0; i < m->nsensor; i++) {
    if (m->sensor_needstage[i] != B2MJ_STAGE_ACC) continue;
    const int t = m->sensor_type[i];
    if (t == B2MJ_SENS_ACCELEROMETER || t == B2MJ_SENS_FORCE || t == B2MJ_SENS_TORQUE || t == B2MJ_SENS_FRAMELINACC ||
        t == B2MJ_SENS_FRAMEANGACC)
      need_rne = true;
  }
  if (need_rne) rnePostConstraint(m, d);
  for (int i = 0; i < m->nsensor; i++) {
    if (m->sensor_needstage[i] != B2MJ_STAGE_ACC) continue;
    const int type = m->sensor_type[i], objid = m->sensor_objid[i], adr = m->sensor_adr[i];
    double* out = d->sensordata + adr;
    double tmp[6];
    switch (type) {
      case B2MJ_SENS_TOUCH: {
        out[0] = 0;
        const int body = m->site_bodyid[objid];
        for (int c = 0; c < d->ncon(); c++) {
          const int ea = d->contact_efc_address[c];
          if (ea < 0) continue;
          const int b1 = m->geom_bodyid[d->contact_geom1[c]], b2 = m->geom_bodyid[d->contact_geom2[c]];
          if (b1 != body && b2 != body) continue;
          double normal_force;
          if (d->efc_type[ea] == B2MJ_CNSTR_CONTACT_PYRAMIDAL) {
            normal_force = 0;
            for (int k = 0; k < 2 * (d->contact_dim[c] - 1); k++) normal_force += d->efc_force[ea + k];
          } else {
            normal_force = d->efc_force[ea];
          }
          if (normal_force <= 0) continue;
          if (point_in_site(m, d, objid, d->contact_pos + 3 * c)) out[0] += normal_force;
        }
        break;
      }
      case B2MJ_SENS_ACCELEROMETER:
        objectAcceleration(m, d, B2MJ_OBJ_SITE, objid, tmp, 1);
        copy3(out, tmp + 3);
        break;
      case B2MJ_SENS_FORCE:
      case B2MJ_SENS_TORQUE: {
        const int body = m->site_bodyid[objid];
        // interaction wrench moved from the subtree COM to the site, site orientation
        double f[6], dif[3], cr[3];
        const double* w = d->cfrc_int + 6 * body;
        sub3(dif, d->site_xpos + 3 * objid, d->subtree_com + 3 * m->body_rootid[body]);
        cross(cr, dif, w + 3);
        sub3(f, w, cr);
        copy3(f + 3, w + 3);
        rotVecMatT(out, type == B2MJ_SENS_FORCE ? f + 3 : f, d->site_xmat + 9 * objid);
        break;
      }
      case B2MJ_SENS_ACTUATORFRC: out[0] = d->actuator_force[objid]; break;
      case B2MJ_SENS_JOINTACTFRC: out[0] = d->qfrc_actuator[m->jnt_dofadr[objid]]; break;
      case B2MJ_SENS_JOINTLIMITFRC:
      case B2MJ_SENS_TENDONLIMITFRC: {
        out[0] = 0;
        const int want = type == B2MJ_SENS_JOINTLIMITFRC ? B2MJ_CNSTR_LIMIT_JOINT : B2MJ_CNSTR_LIMIT_TENDON;
        for (int r = 0; r < d->nefc(); r++)
          if (d->efc_type[r] == want && d->efc_id[r] == objid) { out[0] = d->efc_force[r]; break; }
        break;
      }
      case B2MJ_SENS_FRAMELINACC:
      case B2MJ_SENS_FRAMEANGACC:
        objectAcceleration(m, d, m->sensor_objtype[i], objid, tmp, 0);
        copy3(out, type == B2MJ_SENS_FRAMELINACC ? tmp + 3 : tmp);
        break;
      default: break;
    }
    apply_cutoff(m, d, i);
  }
}

static bool bad(double x) { return std::isnan(x) || x > B2MJ_MAXVAL || x < -B2MJ_MAXVAL; }

static void resetData(const b2mjModel* m, OrcData* d) {
#define X(e, n, t, c) if ((c) > 0) std::memset(d->n, 0, sizeof(t) * (size_t)(c));
  ORC_FIELDS(X)
#undef X
  copy(d->qpos, m->qpos0, m->nq);
  for (int i = 0; i < m->nbody; i++) {
    const int id = m->body_mocapid[i];
    if (id < 0) continue;
    copy3(d->mocap_pos + 3 * id, m->body_pos + 3 * i);
    copy4(d->mocap_quat + 4 * id, m->body_quat + 4 * i);
  }
}

static void checkPos(const b2mjModel* m, OrcData* d) {
  for (int i = 0; i < m->nq; i++)
    if (bad(d->qpos[i])) {
      int w = d->warning[B2MJ_WARN_BADQPOS] + 1;
      resetData(m, d);
      d->warning[B2MJ_WARN_BADQPOS] = w;
      return;
    }
}
static void checkVel(const b2mjModel* m, OrcData* d) {
  for (int i = 0; i < m->nv; i++)
    if (bad(d->qvel[i])) {
      int w = d->warning[B2MJ_WARN_BADQVEL] + 1;
      resetData(m, d);
      d->warning[B2MJ_WARN_BADQVEL] = w;
      return;
    }
}

static void fwdPosition(const b2mjModel* m, OrcData* d) {
  kinematics(m, d);
  comPos(m, d);
  tendon(m, d);
  transmission(m, d);
  crb(m, d);
  factorM(m, d);
  collision(m, d);
  makeConstraint(m, d);
  projectConstraint(m, d);
}

static void fwdVelocity(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv;
  for (int i = 0; i < m->ntendon; i++) d->ten_velocity[i] = dot(d->ten_J + i * nv, d->qvel, nv);
  for (int i = 0; i < m->nu; i++) d->actuator_velocity[i] = dot(d->actuator_moment + i * nv, d->qvel, nv);
  comVel(m, d);
  passive(m, d);
  referenceConstraint(m, d);
  rne(m, d, 0, d->qfrc_bias);
}

static void forwardSkip(const b2mjModel* m, OrcData* d, bool skipsensor, bool up_to_control, bool from_control) {
  if (!from_control) {
    fwdPosition(m, d);
    if (!skipsensor) sensorPos(m, d);
    fwdVelocity(m, d);
    if (!skipsensor) sensorVel(m, d);
  }
  if (up_to_control) return;
  if (d->cb_control) {
    d->n_control_calls++;
    d->cb_control(m, d, d->cb_user);
  }
  fwdActuation(m, d);
  fwdAcceleration(m, d);
  fwdConstraint(m, d);
  if (!skipsensor) sensorAcc(m, d);
}

static void checkAcc(const b2mjModel* m, OrcData* d) {
  for (int i = 0; i < m->nv; i++)
    if (bad(d->qacc[i])) {
      int w = d->warning[B2MJ_WARN_BADQACC] + 1;
      resetData(m, d);
      d->warning[B2MJ_WARN_BADQACC] = w;
      forwardSkip(m, d, false, false, false);
      return;
    }
}

// mj_advance
static void advance(const b2mjModel* m, OrcData* d, const double* act_dot, const double* qacc, const double* qvel) {
  const double h = m->opt.timestep;
  for (int i = 0; i < m->nu; i++) {
    const int a = m->actuator_actadr[i];
    if (a < 0) continue;
    d->act[a] += act_dot[a] * h;
    if (m->actuator_actlimited[i]) d->act[a] = clampd(d->act[a], m->actuator_actrange[2 * i], m->actuator_actrange[2 * i + 1]);
  }
  for (int i = 0; i < m->nv; i++) d->qvel[i] += qacc[i] * h;
  integratePos(m, d->qpos, qvel ? qvel : d->qvel, h);
  d->time[0] += h;
}

// mj_Euler: semi-implicit Euler with implicit joint damping
static void euler(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv;
  bool damping = false;
  if (!(m->opt.disableflags & B2MJ_DSBL_EULERDAMP))
    for (int i = 0; i < nv; i++)
      if (m->dof_damping[i] > 0) { damping = true; break; }
  std::vector<double> qacc(nv);
  if (!damping) {
    copy(qacc.data(), d->qacc, nv);
  } else {
    copy(d->qH, d->qM, m->nM);
    for (int i = 0; i < nv; i++) d->qH[m->dof_Madr[i]] += m->opt.timestep * m->dof_damping[i];
    factorI(m, d->qH, d->qH, d->qHDiagInv, nullptr);
    for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    solveLD(m, qacc.data(), d->qH, d->qHDiagInv);
  }
  advance(m, d, d->act_dot, qacc.data(), nullptr);
}

// mj_implicit: (M - h qDeriv) qacc = qfrc_smooth + qfrc_constraint (orc_implicit.cpp), then mj_advance
static void implicitIntegrate(const b2mjModel* m, OrcData* d) {
  std::vector<double> qacc(m->nv);
  implicitQacc(m, d, qacc.data());
  advance(m, d, d->act_dot, qacc.data(), nullptr);
}

// mj_RungeKutta(4): classical RK4; stages 2-4 skip sensors; control callback fires in every stage
static void rungeKutta4(const b2mjModel* m, OrcData* d) {
  const int nq = m->nq, nv = m->nv, na = m->na;
  const double h = m->opt.timestep, time0 = d->time[0];
  static const double A[9] = {0.5, 0, 0, 0, 0.5, 0, 0, 0, 1}, Bw[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
  std::vector<double> X0(nq + nv + na), Xf[4], F[4], dX(2 * nv + na);
  for (int k = 0; k < 4; k++) { Xf[k].resize(nv); F[k].resize(nv + na); }
  copy(X0.data(), d->qpos, nq);
  copy(X0.data() + nq, d->qvel, nv);
  if (na) copy(X0.data() + nq + nv, d->act, na);
  copy(Xf[0].data(), d->qvel, nv);
  copy(F[0].data(), d->qacc, nv);
  if (na) copy(F[0].data() + nv, d->act_dot, na);
  for (int i = 1; i < 4; i++) {
    // dX = sum_j A[i-1][j] * {Xf[j], F[j]}
    std::fill(dX.begin(), dX.end(), 0.0);
    for (int j = 0; j < 3; j++) {
      const double a = A[(i - 1) * 3 + j];
      if (a == 0) continue;
      for (int k = 0; k < nv; k++) { dX[k] += a * Xf[j][k]; dX[nv + k] += a * F[j][k]; }
      for (int k = 0; k < na; k++) dX[2 * nv + k] += a * F[j][nv + k];
    }
    copy(d->qpos, X0.data(), nq);
    integratePos(m, d->qpos, dX.data(), h);
    for (int k = 0; k < nv; k++) d->qvel[k] = X0[nq + k] + h * dX[nv + k];
    for (int k = 0; k < na; k++) d->act[k] = X0[nq + nv + k] + h * dX[2 * nv + k];
    const double c = (i == 3) ? 1.0 : 0.5;
    d->time[0] = time0 + c * h;
    forwardSkip(m, d, true, false, false);
    copy(Xf[i].data(), d->qvel, nv);
    copy(F[i].data(), d->qacc, nv);
    if (na) copy(F[i].data() + nv, d->act_dot, na);
  }
  std::fill(dX.begin(), dX.end(), 0.0);
  for (int j = 0; j < 4; j++) {
    for (int k = 0; k < nv; k++) { dX[k] += Bw[j] * Xf[j][k]; dX[nv + k] += Bw[j] * F[j][k]; }
    for (int k = 0; k < na; k++) dX[2 * nv + k] += Bw[j] * F[j][nv + k];
  }
  copy(d->qpos, X0.data(), nq);
  copy(d->qvel, X0.data() + nq, nv);
  if (na) copy(d->act, X0.data() + nq + nv, na);
  d->time[0] = time0;
  advance(m, d, dX.data() + 2 * nv, dX.data() + nv, dX.data());
}

}  // namespace orc

using namespace orc;

extern "C" {

void orc_reset_data(const b2mjModel* m, OrcData* d) {
  resetData(m, d);
  d->n_control_calls = d->n_passive_calls = 0;
}

// mj_resetDataKeyframe: mj_resetData, then the keyframe's time / qpos / qvel / act / ctrl / mocap pose
void orc_reset_keyframe(const b2mjModel* m, OrcData* d, int key) {
  resetData(m, d);
  d->n_control_calls = d->n_passive_calls = 0;
  if (key < 0 || key >= m->nkey) return;
  d->time[0] = m->key_time[key];
  copy(d->qpos, m->key_qpos + (size_t)key * m->nq, m->nq);
  copy(d->qvel, m->key_qvel + (size_t)key * m->nv, m->nv);
  if (m->na) copy(d->act, m->key_act + (size_t)key * m->na, m->na);
  if (m->nu) copy(d->ctrl, m->key_ctrl + (size_t)key * m->nu, m->nu);
  if (m->nmocap) {
    copy(d->mocap_pos, m->key_mpos + (size_t)key * 3 * m->nmocap, 3 * m->nmocap);
    copy(d->mocap_quat, m->key_mquat + (size_t)key * 4 * m->nmocap, 4 * m->nmocap);
  }
}

void orc_forward(const b2mjModel* m, OrcData* d) { forwardSkip(m, d, false, false, false); }

void orc_step(const b2mjModel* m, OrcData* d) {
  checkPos(m, d);
  checkVel(m, d);
  forwardSkip(m, d, false, false, false);
  checkAcc(m, d);
  if (m->opt.integrator == B2MJ_INT_RK4) rungeKutta4(m, d);
  else if (m->opt.integrator == B2MJ_INT_IMPLICIT || m->opt.integrator == B2MJ_INT_IMPLICITFAST) implicitIntegrate(m, d);
  else euler(m, d);
}

void orc_step1(const b2mjModel* m, OrcData* d) {
  checkPos(m, d);
  checkVel(m, d);
  forwardSkip(m, d, false, true, false);
}

void orc_step2(const b2mjModel* m, OrcData* d) {
  // like mj_step2: no control callback here (the caller sets controls between step1 and step2)
  orc_callback saved = d->cb_control;
  d->cb_control = nullptr;
  forwardSkip(m, d, false, false, true);
  d->cb_control = saved;
  checkAcc(m, d);
  // mj_step2: Euler or implicit; RK4 falls back to Euler
  if (m->opt.integrator == B2MJ_INT_IMPLICIT || m->opt.integrator == B2MJ_INT_IMPLICITFAST) implicitIntegrate(m, d);
  else euler(m, d);
}

double orc_rollout(const b2mjModel* m, int nenv, int nsteps, double* qpos, double* qvel, const double* ctrl,
                   int nthreads, float* sensor_out) {
  if (nthreads < 1) nthreads = 1;
  std::vector<OrcData*> datas(nthreads);
  for (auto& p : datas) p = orc_make_data(m);
  auto worker = [&](int tid) {
    OrcData* d = datas[tid];
    for (int e = tid; e < nenv; e += nthreads) {
      orc_reset_data(m, d);
      copy(d->qpos, qpos + (size_t)e * m->nq, m->nq);
      copy(d->qvel, qvel + (size_t)e * m->nv, m->nv);
      volatile double published_time = 0;
      for (int s = 0; s < nsteps; s++) {
        if (ctrl && m->nu) copy(d->ctrl, ctrl + ((size_t)s * nenv + e) * m->nu, m->nu);
        orc_step(m, d);
        published_time = d->time[0];  // publishSimTime stub (mujoco_env.cpp:499)
        // last-stage callback: sensor readout arithmetic (mujoco_sensor_handler_plugin.cpp:183-226)
        if (sensor_out)
          for (int n = 0; n < m->nsensor; n++) {
            const double cutoff = m->sensor_cutoff[n] > 0 ? m->sensor_cutoff[n] : 1;
            for (int k = 0; k < m->sensor_dim[n]; k++)
              sensor_out[(size_t)e * m->nsensordata + m->sensor_adr[n] + k] = (float)(d->sensordata[m->sensor_adr[n] + k] / cutoff);
          }
      }
      (void)published_time;
      copy(qpos + (size_t)e * m->nq, d->qpos, m->nq);
      copy(qvel + (size_t)e * m->nv, d->qvel, m->nv);
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (nthreads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (auto p : datas) orc_free_data(p);
  return secs;
}

}  // extern "C"
