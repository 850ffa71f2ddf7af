// orc_smooth.cpp — smooth (constraint-free) dynamics of the CPU oracle.
// TEST INFRASTRUCTURE ONLY (see orc_math.h).  Restates MuJoCo 2.3.7 engine_core_smooth.c /
// engine_passive.c / engine_forward.c (mj_fwdActuation, mj_fwdAcceleration) / engine_support.c —
// the interior of the `mj_step` call at reference mujoco_env.cpp:498; stage list in SURVEY App. A.
#include <cmath>
#include <vector>

#include "orc_math.h"
#include "orc_types.h"

namespace orc {

// mj_kinematics: body, inertial, geom and site frames from qpos (quaternions normalised in place)
void kinematics(const b2mjModel* m, OrcData* d) {
  zero3(d->xpos);
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  zero3(d->xipos);
  zero(d->xmat, 9); zero(d->ximat, 9);
  d->xmat[0] = d->xmat[4] = d->xmat[8] = 1;
  d->ximat[0] = d->ximat[4] = d->ximat[8] = 1;
  // normalise all quaternions in qpos and mocap_quat
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] == B2MJ_JNT_FREE) normalize4(d->qpos + m->jnt_qposadr[j] + 3);
    else if (m->jnt_type[j] == B2MJ_JNT_BALL) normalize4(d->qpos + m->jnt_qposadr[j]);
  }
  for (int i = 0; i < m->nmocap; i++) normalize4(d->mocap_quat + 4 * i);

  for (int i = 1; i < m->nbody; i++) {
    double xpos[3], xquat[4];
    const int jntadr = m->body_jntadr[i], jntnum = m->body_jntnum[i];
    if (jntnum == 1 && m->jnt_type[jntadr] == B2MJ_JNT_FREE) {
      const int qadr = m->jnt_qposadr[jntadr];
      copy3(xpos, d->qpos + qadr);
      copy4(xquat, d->qpos + qadr + 3);
      copy3(d->xanchor + 3 * jntadr, xpos);
      copy3(d->xaxis + 3 * jntadr, m->jnt_axis + 3 * jntadr);
    } else {
      const int pid = m->body_parentid[i];
      const double *bodypos, *bodyquat;
      if (m->body_mocapid[i] >= 0) {
        bodypos = d->mocap_pos + 3 * m->body_mocapid[i];
        bodyquat = d->mocap_quat + 4 * m->body_mocapid[i];
      } else {
        bodypos = m->body_pos + 3 * i;
        bodyquat = m->body_quat + 4 * i;
      }
      if (pid) {
        double vec[3];
        rotVecMat(vec, bodypos, d->xmat + 9 * pid);
        add3(xpos, vec, d->xpos + 3 * pid);
        mulQuat(xquat, d->xquat + 4 * pid, bodyquat);
      } else {
        copy3(xpos, bodypos);
        copy4(xquat, bodyquat);
      }
      for (int j = 0; j < jntnum; j++) {
        const int jid = jntadr + j, qadr = m->jnt_qposadr[jid], jtype = m->jnt_type[jid];
        double xanchor[3], xaxis[3], qloc[4], vec[3];
        rotVecQuat(xaxis, m->jnt_axis + 3 * jid, xquat);
        rotVecQuat(xanchor, m->jnt_pos + 3 * jid, xquat);
        addTo3(xanchor, xpos);
        switch (jtype) {
          case B2MJ_JNT_SLIDE:
            addToScl3(xpos, xaxis, d->qpos[qadr] - m->qpos0[qadr]);
            break;
          case B2MJ_JNT_BALL:
          case B2MJ_JNT_HINGE:
            if (jtype == B2MJ_JNT_BALL) copy4(qloc, d->qpos + qadr);
            else axisAngle2Quat(qloc, m->jnt_axis + 3 * jid, d->qpos[qadr] - m->qpos0[qadr]);
            mulQuat(xquat, xquat, qloc);
            rotVecQuat(vec, m->jnt_pos + 3 * jid, xquat);
            sub3(xpos, xanchor, vec);
            break;
          default: break;
        }
        copy3(d->xanchor + 3 * jid, xanchor);
        copy3(d->xaxis + 3 * jid, xaxis);
      }
    }
    normalize4(xquat);
    copy4(d->xquat + 4 * i, xquat);
    copy3(d->xpos + 3 * i, xpos);
    quat2Mat(d->xmat + 9 * i, xquat);
  }
  // inertial frames
  for (int i = 1; i < m->nbody; i++) {
    double vec[3], q[4];
    rotVecMat(vec, m->body_ipos + 3 * i, d->xmat + 9 * i);
    add3(d->xipos + 3 * i, vec, d->xpos + 3 * i);
    mulQuat(q, d->xquat + 4 * i, m->body_iquat + 4 * i);
    quat2Mat(d->ximat + 9 * i, q);
  }
  for (int i = 0; i < m->ngeom; i++) {
    const int b = m->geom_bodyid[i];
    double vec[3], q[4];
    rotVecMat(vec, m->geom_pos + 3 * i, d->xmat + 9 * b);
    add3(d->geom_xpos + 3 * i, vec, d->xpos + 3 * b);
    mulQuat(q, d->xquat + 4 * b, m->geom_quat + 4 * i);
    quat2Mat(d->geom_xmat + 9 * i, q);
  }
  for (int i = 0; i < m->nsite; i++) {
    const int b = m->site_bodyid[i];
    double vec[3], q[4];
    rotVecMat(vec, m->site_pos + 3 * i, d->xmat + 9 * b);
    add3(d->site_xpos + 3 * i, vec, d->xpos + 3 * b);
    mulQuat(q, d->xquat + 4 * b, m->site_quat + 4 * i);
    quat2Mat(d->site_xmat + 9 * i, q);
  }
}

// mj_comPos: subtree COMs, com-based inertias and motion axes
void comPos(const b2mjModel* m, OrcData* d) {
  zero(d->subtree_com, 3 * m->nbody);
  for (int i = m->nbody - 1; i >= 0; i--) {
    addToScl3(d->subtree_com + 3 * i, d->xipos + 3 * i, m->body_mass[i]);
    if (i) addTo3(d->subtree_com + 3 * m->body_parentid[i], d->subtree_com + 3 * i);
    if (m->body_subtreemass[i] < MINVAL) copy3(d->subtree_com + 3 * i, d->xipos + 3 * i);
    else scl3(d->subtree_com + 3 * i, d->subtree_com + 3 * i, 1.0 / std::fmax(MINVAL, m->body_subtreemass[i]));
  }
  zero(d->cinert, 10);
  for (int i = 1; i < m->nbody; i++) {
    double offset[3];
    sub3(offset, d->xipos + 3 * i, d->subtree_com + 3 * m->body_rootid[i]);
    inertCom(d->cinert + 10 * i, m->body_inertia + 3 * i, d->ximat + 9 * i, offset, m->body_mass[i]);
  }
  for (int j = 0; j < m->njnt; j++) {
    const int da = 6 * m->jnt_dofadr[j], bi = m->jnt_bodyid[j];
    double offset[3], axis[3];
    sub3(offset, d->subtree_com + 3 * m->body_rootid[bi], d->xanchor + 3 * j);
    int skip = 0;
    switch (m->jnt_type[j]) {
      case B2MJ_JNT_FREE:
        zero(d->cdof + da, 18);
        for (int k = 0; k < 3; k++) d->cdof[da + 3 + 7 * k] = 1;
        skip = 18;
        [[fallthrough]];
      case B2MJ_JNT_BALL:
        for (int k = 0; k < 3; k++) {
          axis[0] = d->xmat[9 * bi + k]; axis[1] = d->xmat[9 * bi + k + 3]; axis[2] = d->xmat[9 * bi + k + 6];
          double* res = d->cdof + da + skip + 6 * k;
          copy3(res, axis);
          cross(res + 3, axis, offset);
        }
        break;
      case B2MJ_JNT_SLIDE:
        zero3(d->cdof + da);
        copy3(d->cdof + da + 3, d->xaxis + 3 * j);
        break;
      case B2MJ_JNT_HINGE:
        copy3(d->cdof + da, d->xaxis + 3 * j);
        cross(d->cdof + da + 3, d->xaxis + 3 * j, offset);
        break;
    }
  }
}

// mj_tendon (fixed tendons only): lengths and Jacobians
void tendon(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv;
  if (!m->ntendon) return;
  zero(d->ten_J, m->ntendon * nv);
  std::vector<double> j0(3 * nv), j1(3 * nv);
  for (int i = 0; i < m->ntendon; i++) {
    double len = 0, divisor = 1;
    const int w0 = m->tendon_adr[i], w1 = w0 + m->tendon_num[i];
    for (int w = w0; w < w1; w++) {
      const int type = m->wrap_type[w];
      if (type == 1) {  // mjWRAP_JOINT: fixed tendon
        const int jid = m->wrap_objid[w];
        len += m->wrap_prm[w] * d->qpos[m->jnt_qposadr[jid]];
        d->ten_J[i * nv + m->jnt_dofadr[jid]] = m->wrap_prm[w];
      } else if (type == 2) {  // mjWRAP_PULLEY: scales the branches that follow
        divisor = m->wrap_prm[w];
      } else if (type == 3 && w + 1 < w1 && m->wrap_type[w + 1] == 3) {  // site - site segment of a spatial tendon
        const int s0 = m->wrap_objid[w], s1 = m->wrap_objid[w + 1], b0 = m->site_bodyid[s0], b1 = m->site_bodyid[s1];
        const double *p0 = d->site_xpos + 3 * s0, *p1 = d->site_xpos + 3 * s1;
        double dif[3];
        sub3(dif, p1, p0);
        const double seg = normalize3(dif);
        len += seg / divisor;
        if (b0 != b1) {  // mj_jacDifPair: the segment direction dotted with the difference of the point Jacobians
          jac(m, d, j0.data(), nullptr, p0, b0);
          jac(m, d, j1.data(), nullptr, p1, b1);
          for (int k = 0; k < nv; k++) {
            double s = 0;
            for (int r = 0; r < 3; r++) s += dif[r] * (j1[r * nv + k] - j0[r * nv + k]);
            d->ten_J[i * nv + k] += s / divisor;
          }
        }
      }
    }
    d->ten_length[i] = len;
  }
}

// mj_transmission: actuator lengths and moment arms (joint and tendon transmissions)
void transmission(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv;
  if (!m->nu) return;
  zero(d->actuator_moment, m->nu * nv);
  for (int i = 0; i < m->nu; i++) {
    const int id = m->actuator_trnid[2 * i];
    const double gear = m->actuator_gear[6 * i];
    if (m->actuator_trntype[i] == B2MJ_TRN_TENDON) {
      d->actuator_length[i] = d->ten_length[id] * gear;
      for (int k = 0; k < nv; k++) d->actuator_moment[i * nv + k] = d->ten_J[id * nv + k] * gear;
    } else {
      d->actuator_length[i] = d->qpos[m->jnt_qposadr[id]] * gear;
      d->actuator_moment[i * nv + m->jnt_dofadr[id]] = gear;
    }
  }
}

// mj_crb: composite rigid body inertia -> sparse joint-space inertia qM
void crb(const b2mjModel* m, OrcData* d) {
  copy(d->crb, d->cinert, 10 * m->nbody);
  for (int i = m->nbody - 1; i > 0; i--)
    if (m->body_parentid[i] > 0)
      for (int k = 0; k < 10; k++) d->crb[10 * m->body_parentid[i] + k] += d->crb[10 * i + k];
  zero(d->qM, m->nM);
  for (int i = 0; i < m->nv; i++) {
    int Madr_ij = m->dof_Madr[i];
    double buf[6];
    mulInertVec(buf, d->crb + 10 * m->dof_bodyid[i], d->cdof + 6 * i);
    d->qM[Madr_ij] = m->dof_armature[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j]) d->qM[Madr_ij++] += dot(d->cdof + 6 * j, buf, 6);
  }
}

// mj_factorI: sparse L'*D*L factorisation of an inertia-like matrix
void factorI(const b2mjModel* m, const double* M, double* LD, double* diaginv, double* sqrtdiaginv) {
  const int nv = m->nv;
  if (LD != M) copy(LD, M, m->nM);
  for (int k = nv - 1; k >= 0; k--) {
    const int Madr_kk = m->dof_Madr[k];
    int Madr_ki = Madr_kk + 1;
    int i = m->dof_parentid[k];
    while (i >= 0) {
      const double tmp = LD[Madr_ki] / LD[Madr_kk];
      const int cnt = (i < nv - 1 ? m->dof_Madr[i + 1] : m->nM) - m->dof_Madr[i];
      for (int c = 0; c < cnt; c++) LD[m->dof_Madr[i] + c] -= LD[Madr_ki + c] * tmp;
      LD[Madr_ki] = tmp;
      i = m->dof_parentid[i];
      Madr_ki++;
    }
  }
  for (int i = 0; i < nv; i++) {
    const double D = LD[m->dof_Madr[i]];
    diaginv[i] = 1.0 / D;
    if (sqrtdiaginv) sqrtdiaginv[i] = 1.0 / std::sqrt(D);
  }
}

void factorM(const b2mjModel* m, OrcData* d) { factorI(m, d->qM, d->qLD, d->qLDiagInv, d->qLDiagSqrtInv); }

// mj_solveLD: x <- inv(L'*D*L) * x
void solveLD(const b2mjModel* m, double* x, const double* LD, const double* diaginv) {
  const int nv = m->nv;
  for (int i = nv - 1; i >= 0; i--) {
    const double tmp = x[i];
    if (tmp == 0) continue;
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[j] -= LD[adr++] * tmp;
  }
  for (int i = 0; i < nv; i++) x[i] *= diaginv[i];
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[i] -= LD[adr++] * x[j];
  }
}

void solveM(const b2mjModel* m, const OrcData* d, double* x, const double* y) {
  if (x != y) copy(x, y, m->nv);
  solveLD(m, x, d->qLD, d->qLDiagInv);
}

// mj_solveM2: x = sqrt(inv(D)) * inv(L') * y
void solveM2(const b2mjModel* m, const OrcData* d, double* x, const double* y) {
  const int nv = m->nv;
  if (x != y) copy(x, y, nv);
  for (int i = nv - 1; i >= 0; i--) {
    const double tmp = x[i];
    if (tmp == 0) continue;
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[j] -= d->qLD[adr++] * tmp;
  }
  for (int i = 0; i < nv; i++) x[i] *= d->qLDiagSqrtInv[i];
}

// mj_mulM: res = M * vec (sparse)
void mulM(const b2mjModel* m, const OrcData* d, double* res, const double* vec) {
  const int nv = m->nv;
  zero(res, nv);
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    res[i] += d->qM[adr] * vec[i];
    adr++;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) {
      res[i] += d->qM[adr] * vec[j];
      res[j] += d->qM[adr] * vec[i];
      adr++;
    }
  }
}

static void mulDofVec(double* res, const double* dof, const double* vec, int n) {
  zero(res, 6);
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 6; k++) res[k] += dof[6 * i + k] * vec[i];
}

// mj_comVel: com-based body velocities and dof-axis time derivatives
void comVel(const b2mjModel* m, OrcData* d) {
  zero(d->cvel, 6);
  for (int i = 1; i < m->nbody; i++) {
    const int bda = m->body_dofadr[i];
    double cvel[6], tmp[6], cdofdot[36];
    copy(cvel, d->cvel + 6 * m->body_parentid[i], 6);
    const int dofnum = m->body_dofnum[i];
    for (int j = 0; j < dofnum; j++) {
      switch (m->jnt_type[m->dof_jntid[bda + j]]) {
        case B2MJ_JNT_FREE:
          zero(cdofdot, 18);
          mulDofVec(tmp, d->cdof + 6 * bda, d->qvel + bda, 3);
          for (int k = 0; k < 6; k++) cvel[k] += tmp[k];
          j += 3;
          [[fallthrough]];
        case B2MJ_JNT_BALL:
          for (int k = 0; k < 3; k++) crossMotion(cdofdot + 6 * (j + k), cvel, d->cdof + 6 * (bda + j + k));
          mulDofVec(tmp, d->cdof + 6 * (bda + j), d->qvel + bda + j, 3);
          for (int k = 0; k < 6; k++) cvel[k] += tmp[k];
          j += 2;
          break;
        default:
          crossMotion(cdofdot + 6 * j, cvel, d->cdof + 6 * (bda + j));
          mulDofVec(tmp, d->cdof + 6 * (bda + j), d->qvel + bda + j, 1);
          for (int k = 0; k < 6; k++) cvel[k] += tmp[k];
      }
    }
    copy(d->cvel + 6 * i, cvel, 6);
    if (dofnum) copy(d->cdof_dot + 6 * bda, cdofdot, 6 * dofnum);
  }
}

static void inertiaBoxFluid(const b2mjModel* m, OrcData* d, int i);

// mj_passive: springs, dampers, then the passive callback (mjcb_passive, reference mujoco_env.h:247-251)
void passive(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv;
  zero(d->qfrc_passive, nv);
  if (m->opt.disableflags & B2MJ_DSBL_PASSIVE) return;
  for (int j = 0; j < m->njnt; j++) {
    const double stiffness = m->jnt_stiffness[j];
    if (stiffness == 0) continue;
    int padr = m->jnt_qposadr[j], dadr = m->jnt_dofadr[j];
    switch (m->jnt_type[j]) {
      case B2MJ_JNT_FREE:
        for (int i = 0; i < 3; i++) d->qfrc_passive[dadr + i] -= stiffness * (d->qpos[padr + i] - m->qpos_spring[padr + i]);
        dadr += 3; padr += 3;
        [[fallthrough]];
      case B2MJ_JNT_BALL: {
        double quat[4], dif[3];
        copy4(quat, d->qpos + padr);
        normalize4(quat);
        subQuat(dif, quat, m->qpos_spring + padr);
        for (int i = 0; i < 3; i++) d->qfrc_passive[dadr + i] -= stiffness * dif[i];
        break;
      }
      default:
        d->qfrc_passive[dadr] -= stiffness * (d->qpos[padr] - m->qpos_spring[padr]);
    }
  }
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] -= m->dof_damping[i] * d->qvel[i];
  for (int i = 0; i < m->ntendon; i++) {
    const double stiffness = m->tendon_stiffness[i], damping = m->tendon_damping[i];
    if (stiffness == 0 && damping == 0) continue;
    double frc = 0;
    const double length = d->ten_length[i], lower = m->tendon_lengthspring[2 * i], upper = m->tendon_lengthspring[2 * i + 1];
    if (length > upper) frc = stiffness * (upper - length);
    else if (length < lower) frc = stiffness * (lower - length);
    frc -= damping * d->ten_velocity[i];
    for (int k = 0; k < nv; k++) d->qfrc_passive[k] += d->ten_J[i * nv + k] * frc;
  }
  // gravity compensation
  if (!(m->opt.disableflags & B2MJ_DSBL_GRAVITY)) {
    for (int i = 1; i < m->nbody; i++) {
      if (m->body_gravcomp[i] == 0) continue;
      double force[3], torque[3] = {0, 0, 0};
      scl3(force, m->opt.gravity, -(m->body_mass[i] * m->body_gravcomp[i]));
      applyFT(m, d, force, torque, d->xipos + 3 * i, i, d->qfrc_passive);
    }
  }
  // body-level viscosity, lift and drag (inertia-box fluid model); the reference exposes density / viscosity / wind in
  // its option panel (mujoco_ros/src/viewer.cpp:597-600)
  if (m->opt.viscosity > 0 || m->opt.density > 0)
    for (int i = 1; i < m->nbody; i++)
      if (m->body_mass[i] >= MINVAL) inertiaBoxFluid(m, d, i);
  if (d->cb_passive) {
    d->n_passive_calls++;
    d->cb_passive(m, d, d->cb_user);
  }
}

// mj_rne: recursive Newton-Euler; flg_acc=0 gives the bias force (Coriolis, centrifugal, gravity)
void rne(const b2mjModel* m, OrcData* d, int flg_acc, double* result) {
  const int nb = m->nbody;
  std::vector<double> cacc(6 * nb, 0.0), cfrc(6 * nb, 0.0);
  if (!(m->opt.disableflags & B2MJ_DSBL_GRAVITY)) scl3(cacc.data() + 3, m->opt.gravity, -1);
  for (int i = 1; i < nb; i++) {
    const int bda = m->body_dofadr[i];
    double tmp[6], tmp1[6];
    mulDofVec(tmp, d->cdof_dot + 6 * bda, d->qvel + bda, m->body_dofnum[i]);
    for (int k = 0; k < 6; k++) cacc[6 * i + k] = cacc[6 * m->body_parentid[i] + k] + tmp[k];
    if (flg_acc) {
      mulDofVec(tmp, d->cdof + 6 * bda, d->qacc + bda, m->body_dofnum[i]);
      for (int k = 0; k < 6; k++) cacc[6 * i + k] += tmp[k];
    }
    mulInertVec(&cfrc[6 * i], d->cinert + 10 * i, &cacc[6 * i]);
    mulInertVec(tmp, d->cinert + 10 * i, d->cvel + 6 * i);
    crossForce(tmp1, d->cvel + 6 * i, tmp);
    for (int k = 0; k < 6; k++) cfrc[6 * i + k] += tmp1[k];
  }
  zero(cfrc.data(), 6);
  for (int i = nb - 1; i > 0; i--)
    if (m->body_parentid[i])
      for (int k = 0; k < 6; k++) cfrc[6 * m->body_parentid[i] + k] += cfrc[6 * i + k];
  for (int i = 0; i < m->nv; i++) result[i] = dot(d->cdof + 6 * i, &cfrc[6 * m->dof_bodyid[i]], 6);
}

// mj_jac: Jacobians (3 x nv each, may be null) of a world-frame point attached to a body
void jac(const b2mjModel* m, const OrcData* d, double* jacp, double* jacr, const double* point, int body) {
  const int nv = m->nv;
  if (jacp) zero(jacp, 3 * nv);
  if (jacr) zero(jacr, 3 * nv);
  double offset[3];
  sub3(offset, point, d->subtree_com + 3 * m->body_rootid[body]);
  while (body && !m->body_dofnum[body]) body = m->body_parentid[body];
  if (!body) return;
  int i = m->body_dofadr[body] + m->body_dofnum[body] - 1;
  while (i >= 0) {
    const double* cd = d->cdof + 6 * i;
    if (jacr) { jacr[i] = cd[0]; jacr[nv + i] = cd[1]; jacr[2 * nv + i] = cd[2]; }
    if (jacp) {
      double tmp[3];
      cross(tmp, cd, offset);
      jacp[i] = cd[3] + tmp[0]; jacp[nv + i] = cd[4] + tmp[1]; jacp[2 * nv + i] = cd[5] + tmp[2];
    }
    i = m->dof_parentid[i];
  }
}

// mj_applyFT: qfrc += J' * [force; torque] applied at a point of a body
void applyFT(const b2mjModel* m, const OrcData* d, const double* force, const double* torque, const double* point,
             int body, double* qfrc) {
  const int nv = m->nv;
  if (!nv) return;
  std::vector<double> jp(3 * nv), jr(3 * nv);
  jac(m, d, jp.data(), jr.data(), point, body);
  for (int k = 0; k < nv; k++) {
    double s = 0;
    if (force) s += jp[k] * force[0] + jp[nv + k] * force[1] + jp[2 * nv + k] * force[2];
    if (torque) s += jr[k] * torque[0] + jr[nv + k] * torque[1] + jr[2 * nv + k] * torque[2];
    qfrc[k] += s;
  }
}

// mj_fwdActuation: ctrl clamp, activation dynamics, gain/bias, force clamp, qfrc_actuator
void fwdActuation(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv, nu = m->nu;
  zero(d->qfrc_actuator, nv);
  if (nu) zero(d->actuator_force, nu);
  if (!nu || (m->opt.disableflags & B2MJ_DSBL_ACTUATION)) return;
  std::vector<double> ctrl(nu);
  bool bad = false;
  for (int i = 0; i < nu; i++) {
    ctrl[i] = d->ctrl[i];
    if (std::isnan(ctrl[i]) || ctrl[i] > B2MJ_MAXVAL || ctrl[i] < -B2MJ_MAXVAL) bad = true;
  }
  if (bad) {
    d->warning[B2MJ_WARN_BADCTRL]++;
    for (int i = 0; i < nu; i++) ctrl[i] = d->ctrl[i] = 0;
  }
  for (int i = 0; i < nu; i++)
    if (m->actuator_ctrllimited[i] && !(m->opt.disableflags & B2MJ_DSBL_CLAMPCTRL))
      ctrl[i] = clampd(ctrl[i], m->actuator_ctrlrange[2 * i], m->actuator_ctrlrange[2 * i + 1]);
  for (int i = 0; i < nu; i++) {
    const int a = m->actuator_actadr[i];
    if (a < 0) continue;
    const double* prm = m->actuator_dynprm + B2MJ_NDYN * i;
    switch (m->actuator_dyntype[i]) {
      case B2MJ_DYN_INTEGRATOR: d->act_dot[a] = ctrl[i]; break;
      case B2MJ_DYN_FILTER: d->act_dot[a] = (ctrl[i] - d->act[a]) / std::fmax(MINVAL, prm[0]); break;
      default: d->act_dot[a] = 0;
    }
  }
  for (int i = 0; i < nu; i++) {
    const double* gp = m->actuator_gainprm + B2MJ_NGAIN * i;
    const double* bp = m->actuator_biasprm + B2MJ_NBIAS * i;
    double gain = gp[0];
    if (m->actuator_gaintype[i] == B2MJ_GAIN_AFFINE)
      gain = gp[0] + gp[1] * d->actuator_length[i] + gp[2] * d->actuator_velocity[i];
    const int a = m->actuator_actadr[i];
    d->actuator_force[i] = gain * (a < 0 ? ctrl[i] : d->act[a]);
    if (m->actuator_biastype[i] == B2MJ_BIAS_AFFINE)
      d->actuator_force[i] += bp[0] + bp[1] * d->actuator_length[i] + bp[2] * d->actuator_velocity[i];
  }
  for (int i = 0; i < nu; i++)
    if (m->actuator_forcelimited[i])
      d->actuator_force[i] = clampd(d->actuator_force[i], m->actuator_forcerange[2 * i], m->actuator_forcerange[2 * i + 1]);
  for (int k = 0; k < nv; k++) {
    double s = 0;
    for (int i = 0; i < nu; i++) s += d->actuator_moment[i * nv + k] * d->actuator_force[i];
    d->qfrc_actuator[k] = s;
  }
}

// mj_fwdAcceleration: qfrc_smooth and qacc_smooth = M \ qfrc_smooth
void fwdAcceleration(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv;
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i];
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] += d->qfrc_applied[i];
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] += d->qfrc_actuator[i];
  for (int i = 1; i < m->nbody; i++) {
    const double* x = d->xfrc_applied + 6 * i;
    if (x[0] == 0 && x[1] == 0 && x[2] == 0 && x[3] == 0 && x[4] == 0 && x[5] == 0) continue;
    applyFT(m, d, x, x + 3, d->xipos + 3 * i, i, d->qfrc_smooth);
  }
  solveM(m, d, d->qacc_smooth, d->qfrc_smooth);
}

// mju_transformSpatial
static void transformSpatial(double* res, const double* vec, int flg_force, const double* newpos, const double* oldpos,
                             const double* rotnew2old) {
  double dif[3], cr[3], tran[6];
  copy(tran, vec, 6);
  sub3(dif, newpos, oldpos);
  if (flg_force) {
    cross(cr, dif, vec + 3);
    sub3(tran, vec, cr);
  } else {
    cross(cr, dif, vec);
    sub3(tran + 3, vec + 3, cr);
  }
  if (rotnew2old) {
    rotVecMatT(res, tran, rotnew2old);
    rotVecMatT(res + 3, tran + 3, rotnew2old);
  } else {
    copy(res, tran, 6);
  }
}

static void object_frame(const b2mjModel* m, const OrcData* d, int objtype, int objid, int* bodyid, const double** pos,
                         const double** rot) {
  switch (objtype) {
    case B2MJ_OBJ_BODY: *bodyid = objid; *pos = d->xipos + 3 * objid; *rot = d->ximat + 9 * objid; break;
    case B2MJ_OBJ_GEOM: *bodyid = m->geom_bodyid[objid]; *pos = d->geom_xpos + 3 * objid; *rot = d->geom_xmat + 9 * objid; break;
    case B2MJ_OBJ_SITE: *bodyid = m->site_bodyid[objid]; *pos = d->site_xpos + 3 * objid; *rot = d->site_xmat + 9 * objid; break;
    default: *bodyid = objid; *pos = d->xpos + 3 * objid; *rot = d->xmat + 9 * objid; break;  // XBODY
  }
}

// mj_objectVelocity: 6D velocity [angular; linear] of an object-centred frame
void objectVelocity(const b2mjModel* m, const OrcData* d, int objtype, int objid, double* res, int flg_local) {
  int bodyid;
  const double *pos, *rot;
  object_frame(m, d, objtype, objid, &bodyid, &pos, &rot);
  transformSpatial(res, d->cvel + 6 * bodyid, 0, pos, d->subtree_com + 3 * m->body_rootid[bodyid], flg_local ? rot : nullptr);
}

// mj_inertiaBoxFluidModel: viscous drag of the sphere equivalent to the body's inertia box, quadratic lift / drag on
// the box faces, relative to the wind; applied at the body's centre of mass
static void inertiaBoxFluid(const b2mjModel* m, OrcData* d, int i) {
  const double* inertia = m->body_inertia + 3 * i;
  const double mass = m->body_mass[i], rho = m->opt.density, mu = m->opt.viscosity, kPi = 3.14159265358979323846;
  double box[3], lvel[6], wind[6] = {0, 0, 0, 0, 0, 0}, lwind[6], lfrc[6] = {0, 0, 0, 0, 0, 0}, bfrc[6];
  box[0] = std::sqrt(std::fmax(MINVAL, inertia[1] + inertia[2] - inertia[0]) / mass * 6.0);
  box[1] = std::sqrt(std::fmax(MINVAL, inertia[0] + inertia[2] - inertia[1]) / mass * 6.0);
  box[2] = std::sqrt(std::fmax(MINVAL, inertia[0] + inertia[1] - inertia[2]) / mass * 6.0);
  objectVelocity(m, d, B2MJ_OBJ_BODY, i, lvel, 1);
  copy3(wind + 3, m->opt.wind);
  transformSpatial(lwind, wind, 0, d->xipos + 3 * i, d->subtree_com + 3 * m->body_rootid[i], d->ximat + 9 * i);
  for (int k = 0; k < 3; k++) lvel[3 + k] -= lwind[3 + k];
  if (mu > 0) {
    const double diam = (box[0] + box[1] + box[2]) / 3.0;
    scl3(lfrc, lvel, -kPi * diam * diam * diam * mu);
    scl3(lfrc + 3, lvel + 3, -3.0 * kPi * diam * mu);
  }
  if (rho > 0) {
    lfrc[3] -= 0.5 * rho * box[1] * box[2] * std::fabs(lvel[3]) * lvel[3];
    lfrc[4] -= 0.5 * rho * box[0] * box[2] * std::fabs(lvel[4]) * lvel[4];
    lfrc[5] -= 0.5 * rho * box[0] * box[1] * std::fabs(lvel[5]) * lvel[5];
    const double b0 = box[0] * box[0] * box[0] * box[0], b1 = box[1] * box[1] * box[1] * box[1],
                 b2 = box[2] * box[2] * box[2] * box[2];
    lfrc[0] -= rho * box[0] * (b1 + b2) * std::fabs(lvel[0]) * lvel[0] / 64.0;
    lfrc[1] -= rho * box[1] * (b0 + b2) * std::fabs(lvel[1]) * lvel[1] / 64.0;
    lfrc[2] -= rho * box[2] * (b0 + b1) * std::fabs(lvel[2]) * lvel[2] / 64.0;
  }
  rotVecMat(bfrc, lfrc, d->ximat + 9 * i);
  rotVecMat(bfrc + 3, lfrc + 3, d->ximat + 9 * i);
  applyFT(m, d, bfrc + 3, bfrc, d->xipos + 3 * i, i, d->qfrc_passive);
}

// mj_objectAcceleration (needs cacc from rnePostConstraint)
void objectAcceleration(const b2mjModel* m, const OrcData* d, int objtype, int objid, double* res, int flg_local) {
  int bodyid;
  const double *pos, *rot;
  double vel[6], corr[3];
  object_frame(m, d, objtype, objid, &bodyid, &pos, &rot);
  const double* com = d->subtree_com + 3 * m->body_rootid[bodyid];
  transformSpatial(res, d->cacc + 6 * bodyid, 0, pos, com, flg_local ? rot : nullptr);
  transformSpatial(vel, d->cvel + 6 * bodyid, 0, pos, com, flg_local ? rot : nullptr);
  cross(corr, vel, vel + 3);
  addTo3(res + 3, corr);
}

// mj_rnePostConstraint: cacc, cfrc_int, cfrc_ext including applied and contact forces
void rnePostConstraint(const b2mjModel* m, OrcData* d) {
  const int nb = m->nbody;
  zero(d->cfrc_ext, 6 * nb);
  for (int i = 1; i < nb; i++) {
    const double* x = d->xfrc_applied + 6 * i;
    if (x[0] == 0 && x[1] == 0 && x[2] == 0 && x[3] == 0 && x[4] == 0 && x[5] == 0) continue;
    double corr[6] = {x[3], x[4], x[5], x[0], x[1], x[2]}, f[6];
    transformSpatial(f, corr, 1, d->subtree_com + 3 * m->body_rootid[i], d->xipos + 3 * i, nullptr);
    for (int k = 0; k < 6; k++) d->cfrc_ext[6 * i + k] += f[k];
  }
  for (int c = 0; c < d->ncon(); c++) {
    const int adr = d->contact_efc_address[c];
    if (adr < 0) continue;
    const int dim = d->contact_dim[c];
    double lfrc[6] = {0, 0, 0, 0, 0, 0};
    if (d->efc_type[adr] == B2MJ_CNSTR_CONTACT_PYRAMIDAL) {
      const double* mu = d->contact_friction + 5 * c;
      for (int k = 0; k < 2 * (dim - 1); k++) lfrc[0] += d->efc_force[adr + k];
      for (int k = 1; k < dim; k++) lfrc[k] = (d->efc_force[adr + 2 * (k - 1)] - d->efc_force[adr + 2 * (k - 1) + 1]) * mu[k - 1];
    } else {
      for (int k = 0; k < dim; k++) lfrc[k] = d->efc_force[adr + k];
    }
    // contact frame rows are the axes: world = frame' * local
    const double* fr = d->contact_frame + 9 * c;
    double cf[6];
    rotVecMatT(cf + 3, lfrc, fr);      // force
    rotVecMatT(cf, lfrc + 3, fr);      // torque
    const double* pos = d->contact_pos + 3 * c;
    int b1 = m->geom_bodyid[d->contact_geom1[c]], b2 = m->geom_bodyid[d->contact_geom2[c]];
    double f[6];
    if (b1) {
      transformSpatial(f, cf, 1, d->subtree_com + 3 * m->body_rootid[b1], pos, nullptr);
      for (int k = 0; k < 6; k++) d->cfrc_ext[6 * b1 + k] -= f[k];
    }
    if (b2) {
      transformSpatial(f, cf, 1, d->subtree_com + 3 * m->body_rootid[b2], pos, nullptr);
      for (int k = 0; k < 6; k++) d->cfrc_ext[6 * b2 + k] += f[k];
    }
  }
  // forward pass
  zero(d->cacc, 6);
  if (!(m->opt.disableflags & B2MJ_DSBL_GRAVITY)) scl3(d->cacc + 3, m->opt.gravity, -1);
  zero(d->cfrc_int, 6);
  for (int i = 1; i < nb; i++) {
    const int bda = m->body_dofadr[i];
    double tmp[6], tmp1[6], body[6];
    mulDofVec(tmp, d->cdof_dot + 6 * bda, d->qvel + bda, m->body_dofnum[i]);
    for (int k = 0; k < 6; k++) d->cacc[6 * i + k] = d->cacc[6 * m->body_parentid[i] + k] + tmp[k];
    mulDofVec(tmp, d->cdof + 6 * bda, d->qacc + bda, m->body_dofnum[i]);
    for (int k = 0; k < 6; k++) d->cacc[6 * i + k] += tmp[k];
    mulInertVec(body, d->cinert + 10 * i, d->cacc + 6 * i);
    mulInertVec(tmp, d->cinert + 10 * i, d->cvel + 6 * i);
    crossForce(tmp1, d->cvel + 6 * i, tmp);
    for (int k = 0; k < 6; k++) d->cfrc_int[6 * i + k] = body[k] + tmp1[k] - d->cfrc_ext[6 * i + k];
  }
  for (int i = nb - 1; i > 0; i--)
    if (m->body_parentid[i])
      for (int k = 0; k < 6; k++) d->cfrc_int[6 * m->body_parentid[i] + k] += d->cfrc_int[6 * i + k];
}

}  // namespace orc
