// stages_smooth.cuh — warp-cooperative smooth-dynamics stages (one env per warp).
//
// Replaces, for a batch of envs, the position/velocity/acceleration stages inside the reference's
// `mj_step(model_.get(), data_.get())` call (mujoco_ros/src/mujoco_env.cpp:498,552,593): rows M2-M4,
// M7, M8, M11 of SURVEY.md 8(a).  Parallelisation inside the warp:
//   * tree recursions run level by level, one lane per body of the level;
//   * backward accumulations (subtree COM, composite inertia, RNE forces) run one lane per vector
//     component, serial over bodies, which keeps the serial summation order;
//   * per-dof / per-joint / per-geom / per-actuator work is lane-strided.
#pragma once
#include "env_ctx.cuh"

namespace b2k {

// mj_kinematics
__device__ void stage_kinematics(const Env& e) {
  const DevModel& m = e.m;
  double* qpos = e.D(B2MJ_F_QPOS);
  double* xpos = e.D(B2MJ_F_XPOS);
  double* xquat = e.D(B2MJ_F_XQUAT);
  double* xmat = e.D(B2MJ_F_XMAT);
  double* xipos = e.D(B2MJ_F_XIPOS);
  double* ximat = e.D(B2MJ_F_XIMAT);
  double* xanchor = e.D(B2MJ_F_XANCHOR);
  double* xaxis = e.D(B2MJ_F_XAXIS);
  double* qloc = e.X(XF_QLOC);
  double* mocap_pos = m.nmocap ? e.D(B2MJ_F_MOCAP_POS) : nullptr;
  double* mocap_quat = m.nmocap ? e.D(B2MJ_F_MOCAP_QUAT) : nullptr;

  // joint-local quaternions, all joints in parallel (takes sincos off the serial chain);
  // free / ball quaternions are normalised in place in qpos
  FORL(j, m.njnt) {
    const int t = m.jnt_type[j], qa = m.jnt_qposadr[j];
    if (t == B2MJ_JNT_FREE) normalize4(qpos + qa + 3);
    else if (t == B2MJ_JNT_BALL) { normalize4(qpos + qa); copy4(qloc + 4 * j, qpos + qa); }
    else if (t == B2MJ_JNT_HINGE) axisAngle2Quat(qloc + 4 * j, m.jnt_axis + 3 * j, qpos[qa] - m.qpos0[qa]);
  }
  FORL(i, m.nmocap) normalize4(mocap_quat + 4 * i);
  if (e.lane == 0) {
    zero3(xpos); zero3(xipos);
    xquat[0] = 1; xquat[1] = 0; xquat[2] = 0; xquat[3] = 0;
    for (int k = 0; k < 9; k++) { xmat[k] = (k % 4 == 0) ? 1.0 : 0.0; ximat[k] = (k % 4 == 0) ? 1.0 : 0.0; }
  }
  WSYNC();
  for (int l = 1; l < m.nlevel; l++) {
    const int ladr = m.level_bodyadr[l], lnum = m.level_bodynum[l];
    FORL(k, lnum) {
      const int i = m.level_body[ladr + k];
      double p[3], q[4];
      const int jntadr = m.body_jntadr[i], jntnum = m.body_jntnum[i];
      if (jntnum == 1 && m.jnt_type[jntadr] == B2MJ_JNT_FREE) {
        const int qa = m.jnt_qposadr[jntadr];
        copy3(p, qpos + qa);
        copy4(q, qpos + qa + 3);
        copy3(xanchor + 3 * jntadr, p);
        copy3(xaxis + 3 * jntadr, m.jnt_axis + 3 * jntadr);
      } else {
        const int pid = m.body_parentid[i];
        const double *bpos, *bquat;
        const int mid = m.body_mocapid[i];
        if (mid >= 0) { bpos = mocap_pos + 3 * mid; bquat = mocap_quat + 4 * mid; }
        else { bpos = m.body_pos + 3 * i; bquat = m.body_quat + 4 * i; }
        if (pid) {
          double v[3];
          rotVecMat(v, bpos, xmat + 9 * pid);
          add3(p, v, xpos + 3 * pid);
          mulQuat(q, xquat + 4 * pid, bquat);
        } else {
          copy3(p, bpos);
          copy4(q, bquat);
        }
        for (int j = 0; j < jntnum; j++) {
          const int jid = jntadr + j, jt = m.jnt_type[jid];
          double anchor[3], axis[3], v[3];
          rotVecQuat(axis, m.jnt_axis + 3 * jid, q);
          rotVecQuat(anchor, m.jnt_pos + 3 * jid, q);
          addTo3(anchor, p);
          if (jt == B2MJ_JNT_SLIDE) {
            const int qa = m.jnt_qposadr[jid];
            addToScl3(p, axis, qpos[qa] - m.qpos0[qa]);
          } else if (jt == B2MJ_JNT_BALL || jt == B2MJ_JNT_HINGE) {
            mulQuat(q, q, qloc + 4 * jid);
            rotVecQuat(v, m.jnt_pos + 3 * jid, q);
            sub3(p, anchor, v);
          }
          copy3(xanchor + 3 * jid, anchor);
          copy3(xaxis + 3 * jid, axis);
        }
      }
      normalize4(q);
      copy4(xquat + 4 * i, q);
      copy3(xpos + 3 * i, p);
      quat2Mat(xmat + 9 * i, q);
    }
    WSYNC();
  }
  // inertial, geom and site frames
  FORL(i, m.nbody) {
    if (i == 0) continue;
    double v[3], q[4];
    rotVecMat(v, m.body_ipos + 3 * i, xmat + 9 * i);
    add3(xipos + 3 * i, v, xpos + 3 * i);
    mulQuat(q, xquat + 4 * i, m.body_iquat + 4 * i);
    quat2Mat(ximat + 9 * i, q);
  }
  double* gxpos = e.D(B2MJ_F_GEOM_XPOS);
  double* gxmat = e.D(B2MJ_F_GEOM_XMAT);
  FORL(i, m.ngeom) {
    const int b = m.geom_bodyid[i];
    double v[3], q[4];
    rotVecMat(v, m.geom_pos + 3 * i, xmat + 9 * b);
    add3(gxpos + 3 * i, v, xpos + 3 * b);
    mulQuat(q, xquat + 4 * b, m.geom_quat + 4 * i);
    quat2Mat(gxmat + 9 * i, q);
  }
  if (m.nsite) {
    double* sxpos = e.D(B2MJ_F_SITE_XPOS);
    double* sxmat = e.D(B2MJ_F_SITE_XMAT);
    FORL(i, m.nsite) {
      const int b = m.site_bodyid[i];
      double v[3], q[4];
      rotVecMat(v, m.site_pos + 3 * i, xmat + 9 * b);
      add3(sxpos + 3 * i, v, xpos + 3 * b);
      mulQuat(q, xquat + 4 * b, m.site_quat + 4 * i);
      quat2Mat(sxmat + 9 * i, q);
    }
  }
  WSYNC();
}

// mj_comPos
__device__ void stage_comPos(const Env& e) {
  const DevModel& m = e.m;
  const double* xipos = e.D(B2MJ_F_XIPOS);
  const double* ximat = e.D(B2MJ_F_XIMAT);
  const double* xmat = e.D(B2MJ_F_XMAT);
  const double* xanchor = e.D(B2MJ_F_XANCHOR);
  const double* xaxis = e.D(B2MJ_F_XAXIS);
  double* com = e.D(B2MJ_F_SUBTREE_COM);
  double* cinert = e.D(B2MJ_F_CINERT);
  double* cdof = e.D(B2MJ_F_CDOF);
  FORL(k, 3 * m.nbody) com[k] = 0;
  WSYNC();
  if (e.lane < 3) {
    const int c = e.lane;
    for (int i = m.nbody - 1; i >= 0; i--) {
      double s = com[3 * i + c] + xipos[3 * i + c] * m.body_mass[i];
      if (i) com[3 * m.body_parentid[i] + c] += s;
      const double sm = m.body_subtreemass[i];
      com[3 * i + c] = (sm < B2K_MINVAL) ? xipos[3 * i + c] : s * (1.0 / fmax(B2K_MINVAL, sm));
    }
  }
  WSYNC();
  FORL(i, m.nbody) {
    if (i == 0) { for (int k = 0; k < 10; k++) cinert[k] = 0; continue; }
    double off[3];
    sub3(off, xipos + 3 * i, com + 3 * m.body_rootid[i]);
    inertCom(cinert + 10 * i, m.body_inertia + 3 * i, ximat + 9 * i, off, m.body_mass[i]);
  }
  FORL(j, m.njnt) {
    const int da = 6 * m.jnt_dofadr[j], bi = m.jnt_bodyid[j];
    double off[3], axis[3];
    sub3(off, com + 3 * m.body_rootid[bi], xanchor + 3 * j);
    int skip = 0;
    switch (m.jnt_type[j]) {
      case B2MJ_JNT_FREE:
        for (int k = 0; k < 18; k++) cdof[da + k] = 0;
        for (int k = 0; k < 3; k++) cdof[da + 3 + 7 * k] = 1;
        skip = 18;
      case B2MJ_JNT_BALL:
        for (int k = 0; k < 3; k++) {
          axis[0] = xmat[9 * bi + k]; axis[1] = xmat[9 * bi + k + 3]; axis[2] = xmat[9 * bi + k + 6];
          double* r = cdof + da + skip + 6 * k;
          copy3(r, axis);
          cross(r + 3, axis, off);
        }
        break;
      case B2MJ_JNT_SLIDE:
        zero3(cdof + da);
        copy3(cdof + da + 3, xaxis + 3 * j);
        break;
      default:
        copy3(cdof + da, xaxis + 3 * j);
        cross(cdof + da + 3, xaxis + 3 * j, off);
    }
  }
  WSYNC();
}

// mj_tendon (fixed) + mj_transmission
__device__ void stage_tendon_transmission(const Env& e) {
  const DevModel& m = e.m;
  const int nv = m.nv;
  const double* qpos = e.D(B2MJ_F_QPOS);
  if (m.ntendon) {
    double* tl = e.D(B2MJ_F_TEN_LENGTH);
    double* tJ = e.D(B2MJ_F_TEN_J);
    FORL(k, m.ntendon * nv) tJ[k] = 0;
    WSYNC();
    FORL(i, m.ntendon) {
      double len = 0;
      for (int w = m.tendon_adr[i]; w < m.tendon_adr[i] + m.tendon_num[i]; w++) {
        const int jid = m.wrap_objid[w];
        len += m.wrap_prm[w] * qpos[m.jnt_qposadr[jid]];
        tJ[i * nv + m.jnt_dofadr[jid]] = m.wrap_prm[w];
      }
      tl[i] = len;
    }
    WSYNC();
  }
  if (m.nu) {
    double* al = e.D(B2MJ_F_ACTUATOR_LENGTH);
    double* am = e.D(B2MJ_F_ACTUATOR_MOMENT);
    FORL(k, m.nu * nv) am[k] = 0;
    WSYNC();
    FORL(i, m.nu) {
      const int id = m.actuator_trnid[2 * i];
      const double gear = m.actuator_gear[6 * i];
      if (m.actuator_trntype[i] == B2MJ_TRN_TENDON) {
        const double* tl = e.D(B2MJ_F_TEN_LENGTH);
        const double* tJ = e.D(B2MJ_F_TEN_J);
        al[i] = tl[id] * gear;
        for (int k = 0; k < nv; k++) am[i * nv + k] = tJ[id * nv + k] * gear;
      } else {
        al[i] = qpos[m.jnt_qposadr[id]] * gear;
        am[i * nv + m.jnt_dofadr[id]] = gear;
      }
    }
    WSYNC();
  }
}

// sparse L'DL factorisation in place (mj_factorI); lane 0 walks dofs, lanes share each row update
__device__ void factorLD(const Env& e, double* LD, double* diaginv, double* sqrtdiaginv) {
  const DevModel& m = e.m;
  const int nv = m.nv;
  for (int k = nv - 1; k >= 0; k--) {
    const int Mkk = m.dof_Madr[k];
    // ancestors of k: a = 0.. ; entry M(k, anc_a) at Mkk+1+a.  All ancestor rows update independently
    // from the *original* row k, so: phase 1 rows, phase 2 scale row k.
    const double dkk = LD[Mkk];
    int i = m.dof_parentid[k];
    int a = 0;
    while (i >= 0) {
      const double tmp = LD[Mkk + 1 + a] / dkk;
      const int rowadr = m.dof_Madr[i];
      const int cnt = (i < nv - 1 ? m.dof_Madr[i + 1] : m.nM) - rowadr;
      FORL(c, cnt) LD[rowadr + c] -= LD[Mkk + 1 + a + c] * tmp;
      i = m.dof_parentid[i];
      a++;
    }
    WSYNC();
    FORL(c, a) LD[Mkk + 1 + c] = LD[Mkk + 1 + c] / dkk;
    WSYNC();
  }
  FORL(i, nv) {
    const double Dv = LD[m.dof_Madr[i]];
    diaginv[i] = 1.0 / Dv;
    if (sqrtdiaginv) sqrtdiaginv[i] = 1.0 / sqrt(Dv);
  }
  WSYNC();
}

// x <- inv(L'DL) x, executed by ONE lane (callers distribute independent right-hand sides over lanes)
__device__ __forceinline__ void solveLD_lane(const DevModel& m, double* x, const double* LD, const double* diaginv) {
  const int nv = m.nv;
  for (int i = nv - 1; i >= 0; i--) {
    const double t = x[i];
    if (t == 0) continue;
    int adr = m.dof_Madr[i] + 1;
    for (int j = m.dof_parentid[i]; j >= 0; j = m.dof_parentid[j]) x[j] -= LD[adr++] * t;
  }
  for (int i = 0; i < nv; i++) x[i] *= diaginv[i];
  for (int i = 0; i < nv; i++) {
    int adr = m.dof_Madr[i] + 1;
    double xi = x[i];
    for (int j = m.dof_parentid[i]; j >= 0; j = m.dof_parentid[j]) xi -= LD[adr++] * x[j];
    x[i] = xi;
  }
}

// res = M * vec with one lane per output row (gather form of mj_mulM)
__device__ void mulM_warp(const Env& e, double* res, const double* vec) {
  const DevModel& m = e.m;
  const double* qM = e.D(B2MJ_F_QM);
  FORL(i, m.nv) res[i] = 0;
  WSYNC();
  // scatter form keeps MuJoCo's summation structure; executed per dof with atomics avoided by
  // splitting into the "row" part (lane i) and the "column" part (lane j) below
  FORL(i, m.nv) {
    int adr = m.dof_Madr[i];
    double s = qM[adr] * vec[i];
    adr++;
    for (int j = m.dof_parentid[i]; j >= 0; j = m.dof_parentid[j]) s += qM[adr++] * vec[j];
    res[i] = s;
  }
  WSYNC();
  // contributions M(i,j)*vec[i] to res[j] for descendants i of j: serial over i per lane j via mask
  FORL(j, m.nv) {
    double s = res[j];
    for (int i = j + 1; i < m.nv; i++) {
      // is j an ancestor of i?  walk is short (tree depth)
      int adr = m.dof_Madr[i] + 1;
      for (int p = m.dof_parentid[i]; p >= 0; p = m.dof_parentid[p], adr++)
        if (p == j) { s += qM[adr] * vec[i]; break; }
        else if (p < j) break;
    }
    res[j] = s;
  }
  WSYNC();
}

// mj_crb + mj_factorM
__device__ void stage_crb_factor(const Env& e) {
  const DevModel& m = e.m;
  const double* cinert = e.D(B2MJ_F_CINERT);
  const double* cdof = e.D(B2MJ_F_CDOF);
  double* crb = e.D(B2MJ_F_CRB);
  double* qM = e.D(B2MJ_F_QM);
  double* qLD = e.D(B2MJ_F_QLD);
  FORL(k, 10 * m.nbody) crb[k] = cinert[k];
  WSYNC();
  if (e.lane < 10) {
    const int c = e.lane;
    for (int i = m.nbody - 1; i > 0; i--) {
      const int p = m.body_parentid[i];
      if (p > 0) crb[10 * p + c] += crb[10 * i + c];
    }
  }
  WSYNC();
  FORL(i, m.nv) {
    int adr = m.dof_Madr[i];
    double buf[6];
    mulInertVec(buf, crb + 10 * m.dof_bodyid[i], cdof + 6 * i);
    qM[adr] = m.dof_armature[i] + dot6(cdof + 6 * i, buf);
    adr++;
    for (int j = m.dof_parentid[i]; j >= 0; j = m.dof_parentid[j]) qM[adr++] = dot6(cdof + 6 * j, buf);
  }
  WSYNC();
  FORL(k, m.nM) qLD[k] = qM[k];
  WSYNC();
  factorLD(e, qLD, e.D(B2MJ_F_QLDIAGINV), e.D(B2MJ_F_QLDIAGSQRTINV));
}

__device__ __forceinline__ void mulDofVec(double* res, const double* dof, const double* vec, int n) {
  for (int k = 0; k < 6; k++) res[k] = 0;
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 6; k++) res[k] += dof[6 * i + k] * vec[i];
}

// mj_comVel (level-parallel)
__device__ void stage_comVel(const Env& e) {
  const DevModel& m = e.m;
  const double* cdof = e.D(B2MJ_F_CDOF);
  const double* qvel = e.D(B2MJ_F_QVEL);
  double* cvelA = e.D(B2MJ_F_CVEL);
  double* cdof_dot = e.D(B2MJ_F_CDOF_DOT);
  if (e.lane < 6) cvelA[e.lane] = 0;
  WSYNC();
  for (int l = 1; l < m.nlevel; l++) {
    const int ladr = m.level_bodyadr[l], lnum = m.level_bodynum[l];
    FORL(k, lnum) {
      const int i = m.level_body[ladr + k];
      const int bda = m.body_dofadr[i], dofnum = m.body_dofnum[i];
      double cvel[6], tmp[6];
      for (int c = 0; c < 6; c++) cvel[c] = cvelA[6 * m.body_parentid[i] + c];
      for (int j = 0; j < dofnum; j++) {
        const int jt = m.jnt_type[m.dof_jntid[bda + j]];
        if (jt == B2MJ_JNT_FREE) {
          for (int c = 0; c < 18; c++) cdof_dot[6 * bda + c] = 0;
          mulDofVec(tmp, cdof + 6 * bda, qvel + bda, 3);
          for (int c = 0; c < 6; c++) cvel[c] += tmp[c];
          j += 3;
        }
        if (jt == B2MJ_JNT_FREE || jt == B2MJ_JNT_BALL) {
          for (int c = 0; c < 3; c++) crossMotion(cdof_dot + 6 * (bda + j + c), cvel, cdof + 6 * (bda + j + c));
          mulDofVec(tmp, cdof + 6 * (bda + j), qvel + bda + j, 3);
          for (int c = 0; c < 6; c++) cvel[c] += tmp[c];
          j += 2;
        } else {
          crossMotion(cdof_dot + 6 * (bda + j), cvel, cdof + 6 * (bda + j));
          mulDofVec(tmp, cdof + 6 * (bda + j), qvel + bda + j, 1);
          for (int c = 0; c < 6; c++) cvel[c] += tmp[c];
        }
      }
      for (int c = 0; c < 6; c++) cvelA[6 * i + c] = cvel[c];
    }
    WSYNC();
  }
}

// Jacobian-transpose application of a force/torque at a point of a body: one lane per dof.
// qfrc[k] += jacp[:,k].force + jacr[:,k].torque for dofs on the body's chain.
__device__ void applyFT_warp(const Env& e, const double* force, const double* torque, const double* point, int body,
                             double* qfrc) {
  const DevModel& m = e.m;
  const double* cdof = e.D(B2MJ_F_CDOF);
  const double* com = e.D(B2MJ_F_SUBTREE_COM);
  double off[3];
  sub3(off, point, com + 3 * m.body_rootid[body]);
  const unsigned* mask = m.body_dofmask + body * m.nmaskword;
  FORL(k, m.nv) {
    if (!((mask[k >> 5] >> (k & 31)) & 1u)) continue;
    const double* cd = cdof + 6 * k;
    double cr[3];
    cross(cr, cd, off);
    double s = (cd[3] + cr[0]) * force[0] + (cd[4] + cr[1]) * force[1] + (cd[5] + cr[2]) * force[2];
    s += cd[0] * torque[0] + cd[1] * torque[1] + cd[2] * torque[2];
    qfrc[k] += s;
  }
  WSYNC();
}

// mj_passive (springs, dampers, gravity compensation); the host passive hook is a split-step feature
__device__ void stage_passive(const Env& e) {
  const DevModel& m = e.m;
  const int nv = m.nv;
  const double* qpos = e.D(B2MJ_F_QPOS);
  const double* qvel = e.D(B2MJ_F_QVEL);
  double* qp = e.D(B2MJ_F_QFRC_PASSIVE);
  if (m.opt.disableflags & B2MJ_DSBL_PASSIVE) {
    FORL(i, nv) qp[i] = 0;
    WSYNC();
    return;
  }
  FORL(i, nv) qp[i] = -m.dof_damping[i] * qvel[i];
  WSYNC();
  FORL(j, m.njnt) {
    const double st = m.jnt_stiffness[j];
    if (st == 0) continue;
    int padr = m.jnt_qposadr[j], dadr = m.jnt_dofadr[j];
    const int jt = m.jnt_type[j];
    if (jt == B2MJ_JNT_FREE) {
      for (int i = 0; i < 3; i++) qp[dadr + i] -= st * (qpos[padr + i] - m.qpos_spring[padr + i]);
      dadr += 3; padr += 3;
    }
    if (jt == B2MJ_JNT_FREE || jt == B2MJ_JNT_BALL) {
      double quat[4], dif[3];
      copy4(quat, qpos + padr);
      normalize4(quat);
      subQuat(dif, quat, m.qpos_spring + padr);
      for (int i = 0; i < 3; i++) qp[dadr + i] -= st * dif[i];
    } else {
      qp[dadr] -= st * (qpos[padr] - m.qpos_spring[padr]);
    }
  }
  WSYNC();
  if (m.ntendon) {
    const double* tl = e.D(B2MJ_F_TEN_LENGTH);
    const double* tv = e.D(B2MJ_F_TEN_VELOCITY);
    const double* tJ = e.D(B2MJ_F_TEN_J);
    for (int i = 0; i < m.ntendon; i++) {
      const double st = m.tendon_stiffness[i], dm = m.tendon_damping[i];
      if (st == 0 && dm == 0) continue;
      double frc = 0;
      const double len = tl[i], lo = m.tendon_lengthspring[2 * i], hi = m.tendon_lengthspring[2 * i + 1];
      if (len > hi) frc = st * (hi - len);
      else if (len < lo) frc = st * (lo - len);
      frc -= dm * tv[i];
      FORL(k, nv) qp[k] += tJ[i * nv + k] * frc;
    }
    WSYNC();
  }
  if (!(m.opt.disableflags & B2MJ_DSBL_GRAVITY)) {
    const double* xipos = e.D(B2MJ_F_XIPOS);
    for (int i = 1; i < m.nbody; i++) {
      const double gc = m.body_gravcomp[i];
      if (gc == 0) continue;
      double force[3], torque[3] = {0, 0, 0};
      scl3(force, m.opt.gravity, -(m.body_mass[i] * gc));
      applyFT_warp(e, force, torque, xipos + 3 * i, i, qp);
    }
  }
}

// mj_rne(flg_acc = 0): bias forces.  Uses the cacc / cfrc_int fields as scratch.
__device__ void stage_rne_bias(const Env& e) {
  const DevModel& m = e.m;
  const double* cdof = e.D(B2MJ_F_CDOF);
  const double* cdof_dot = e.D(B2MJ_F_CDOF_DOT);
  const double* cvel = e.D(B2MJ_F_CVEL);
  const double* cinert = e.D(B2MJ_F_CINERT);
  const double* qvel = e.D(B2MJ_F_QVEL);
  double* cacc = e.D(B2MJ_F_CACC);
  double* cfrc = e.D(B2MJ_F_CFRC_INT);
  double* bias = e.D(B2MJ_F_QFRC_BIAS);
  if (e.lane < 6) {
    double g = 0;
    if (e.lane >= 3 && !(m.opt.disableflags & B2MJ_DSBL_GRAVITY)) g = -m.opt.gravity[e.lane - 3];
    cacc[e.lane] = g;
    cfrc[e.lane] = 0;
  }
  WSYNC();
  for (int l = 1; l < m.nlevel; l++) {
    const int ladr = m.level_bodyadr[l], lnum = m.level_bodynum[l];
    FORL(k, lnum) {
      const int i = m.level_body[ladr + k];
      const int bda = m.body_dofadr[i];
      double tmp[6], tmp1[6], acc[6], f[6];
      mulDofVec(tmp, cdof_dot + 6 * bda, qvel + bda, m.body_dofnum[i]);
      for (int c = 0; c < 6; c++) acc[c] = cacc[6 * m.body_parentid[i] + c] + tmp[c];
      for (int c = 0; c < 6; c++) cacc[6 * i + c] = acc[c];
      mulInertVec(f, cinert + 10 * i, acc);
      mulInertVec(tmp, cinert + 10 * i, cvel + 6 * i);
      crossForce(tmp1, cvel + 6 * i, tmp);
      for (int c = 0; c < 6; c++) cfrc[6 * i + c] = f[c] + tmp1[c];
    }
    WSYNC();
  }
  if (e.lane < 6) {
    const int c = e.lane;
    for (int i = m.nbody - 1; i > 0; i--) {
      const int p = m.body_parentid[i];
      if (p) cfrc[6 * p + c] += cfrc[6 * i + c];
    }
  }
  WSYNC();
  FORL(i, m.nv) bias[i] = dot6(cdof + 6 * i, cfrc + 6 * m.dof_bodyid[i]);
  WSYNC();
}

// tendon / actuator velocities (head of mj_fwdVelocity)
__device__ void stage_velocity_head(const Env& e) {
  const DevModel& m = e.m;
  const int nv = m.nv;
  const double* qvel = e.D(B2MJ_F_QVEL);
  if (m.ntendon) {
    const double* tJ = e.D(B2MJ_F_TEN_J);
    double* tv = e.D(B2MJ_F_TEN_VELOCITY);
    FORL(i, m.ntendon) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += tJ[i * nv + k] * qvel[k];
      tv[i] = s;
    }
  }
  if (m.nu) {
    const double* am = e.D(B2MJ_F_ACTUATOR_MOMENT);
    double* av = e.D(B2MJ_F_ACTUATOR_VELOCITY);
    FORL(i, m.nu) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += am[i * nv + k] * qvel[k];
      av[i] = s;
    }
  }
  WSYNC();
}

// mj_fwdActuation
__device__ void stage_actuation(const Env& e, int* warning) {
  const DevModel& m = e.m;
  const int nv = m.nv, nu = m.nu;
  double* qa = e.D(B2MJ_F_QFRC_ACTUATOR);
  if (!nu || (m.opt.disableflags & B2MJ_DSBL_ACTUATION)) {
    FORL(i, nv) qa[i] = 0;
    if (nu) { double* af = e.D(B2MJ_F_ACTUATOR_FORCE); FORL(i, nu) af[i] = 0; }
    WSYNC();
    return;
  }
  double* ctrl = e.D(B2MJ_F_CTRL);
  double* af = e.D(B2MJ_F_ACTUATOR_FORCE);
  const double* al = e.D(B2MJ_F_ACTUATOR_LENGTH);
  const double* av = e.D(B2MJ_F_ACTUATOR_VELOCITY);
  const double* am = e.D(B2MJ_F_ACTUATOR_MOMENT);
  const double* act = m.na ? e.D(B2MJ_F_ACT) : nullptr;
  double* act_dot = m.na ? e.D(B2MJ_F_ACT_DOT) : nullptr;
  // bad controls: warn and zero all of them
  int badc = 0;
  FORL(i, nu) {
    const double c = ctrl[i];
    if (isnan(c) || c > B2MJ_MAXVAL || c < -B2MJ_MAXVAL) badc = 1;
  }
  if (__any_sync(0xffffffffu, badc)) {
    FORL(i, nu) ctrl[i] = 0;
    if (e.lane == 0) warning[B2MJ_WARN_BADCTRL]++;
    WSYNC();
  }
  FORL(i, nu) {
    double c = ctrl[i];
    if (m.actuator_ctrllimited[i] && !(m.opt.disableflags & B2MJ_DSBL_CLAMPCTRL))
      c = clampd(c, m.actuator_ctrlrange[2 * i], m.actuator_ctrlrange[2 * i + 1]);
    const int a = m.actuator_actadr[i];
    if (a >= 0) {
      const double* prm = m.actuator_dynprm + B2MJ_NDYN * i;
      const int dt = m.actuator_dyntype[i];
      act_dot[a] = dt == B2MJ_DYN_INTEGRATOR ? c : dt == B2MJ_DYN_FILTER ? (c - act[a]) / fmax(B2K_MINVAL, prm[0]) : 0.0;
    }
    const double* gp = m.actuator_gainprm + B2MJ_NGAIN * i;
    const double* bp = m.actuator_biasprm + B2MJ_NBIAS * i;
    double gain = gp[0];
    if (m.actuator_gaintype[i] == B2MJ_GAIN_AFFINE) gain = gp[0] + gp[1] * al[i] + gp[2] * av[i];
    double f = gain * (a < 0 ? c : act[a]);
    if (m.actuator_biastype[i] == B2MJ_BIAS_AFFINE) f += bp[0] + bp[1] * al[i] + bp[2] * av[i];
    if (m.actuator_forcelimited[i]) f = clampd(f, m.actuator_forcerange[2 * i], m.actuator_forcerange[2 * i + 1]);
    af[i] = f;
  }
  WSYNC();
  FORL(k, nv) {
    double s = 0;
    for (int i = 0; i < nu; i++) s += am[i * nv + k] * af[i];
    qa[k] = s;
  }
  WSYNC();
}

// mj_fwdAcceleration
__device__ void stage_acceleration(const Env& e, const double* xfrc) {
  const DevModel& m = e.m;
  const int nv = m.nv;
  double* qs = e.D(B2MJ_F_QFRC_SMOOTH);
  double* qas = e.D(B2MJ_F_QACC_SMOOTH);
  const double* qp = e.D(B2MJ_F_QFRC_PASSIVE);
  const double* qb = e.D(B2MJ_F_QFRC_BIAS);
  const double* qap = e.D(B2MJ_F_QFRC_APPLIED);
  const double* qa = e.D(B2MJ_F_QFRC_ACTUATOR);
  FORL(i, nv) {
    double s = qp[i] - qb[i];
    s += qap[i];
    s += qa[i];
    qs[i] = s;
  }
  WSYNC();
  if (xfrc) {
    const double* xipos = e.D(B2MJ_F_XIPOS);
    for (int i = 1; i < m.nbody; i++) {
      const double* x = xfrc + 6 * i;
      if (x[0] == 0 && x[1] == 0 && x[2] == 0 && x[3] == 0 && x[4] == 0 && x[5] == 0) continue;
      applyFT_warp(e, x, x + 3, xipos + 3 * i, i, qs);
    }
  }
  FORL(i, nv) qas[i] = qs[i];
  WSYNC();
  if (e.lane == 0) solveLD_lane(m, qas, e.D(B2MJ_F_QLD), e.D(B2MJ_F_QLDIAGINV));
  WSYNC();
}

// mj_integratePos for the joints handled by this lane
__device__ void integratePos_warp(const Env& e, double* qpos, const double* qvel, double dt) {
  const DevModel& m = e.m;
  FORL(j, m.njnt) {
    int padr = m.jnt_qposadr[j], vadr = m.jnt_dofadr[j];
    const int jt = m.jnt_type[j];
    if (jt == B2MJ_JNT_FREE) {
      for (int i = 0; i < 3; i++) qpos[padr + i] += dt * qvel[vadr + i];
      padr += 3; vadr += 3;
    }
    if (jt == B2MJ_JNT_FREE || jt == B2MJ_JNT_BALL) quatIntegrate(qpos + padr, qvel + vadr, dt);
    else qpos[padr] += dt * qvel[vadr];
  }
  WSYNC();
}

// mj_advance
__device__ void advance_warp(const Env& e, const double* act_dot, const double* qacc, const double* qvel_for_pos) {
  const DevModel& m = e.m;
  const double h = m.opt.timestep;
  double* qvel = e.D(B2MJ_F_QVEL);
  if (m.na) {
    double* act = e.D(B2MJ_F_ACT);
    FORL(i, m.nu) {
      const int a = m.actuator_actadr[i];
      if (a < 0) continue;
      double v = act[a] + act_dot[a] * h;
      if (m.actuator_actlimited[i]) v = clampd(v, m.actuator_actrange[2 * i], m.actuator_actrange[2 * i + 1]);
      act[a] = v;
    }
  }
  FORL(i, m.nv) qvel[i] += qacc[i] * h;
  WSYNC();
  integratePos_warp(e, e.D(B2MJ_F_QPOS), qvel_for_pos ? qvel_for_pos : qvel, h);
  if (e.lane == 0) e.D(B2MJ_F_TIME)[0] += h;
  WSYNC();
}

// mj_Euler: semi-implicit Euler, implicit in joint damping
__device__ void stage_euler(const Env& e) {
  const DevModel& m = e.m;
  const int nv = m.nv;
  const double* act_dot = m.na ? e.D(B2MJ_F_ACT_DOT) : nullptr;
  if (!m.any_damping || (m.opt.disableflags & B2MJ_DSBL_EULERDAMP)) {
    advance_warp(e, act_dot, e.D(B2MJ_F_QACC), nullptr);
    return;
  }
  double* qH = e.X(XF_QH);
  double* qHd = e.X(XF_QHDIAGINV);
  double* acc = e.X(XF_VEC0);
  const double* qM = e.D(B2MJ_F_QM);
  const double* qs = e.D(B2MJ_F_QFRC_SMOOTH);
  const double* qc = e.D(B2MJ_F_QFRC_CONSTRAINT);
  FORL(k, m.nM) qH[k] = qM[k];
  WSYNC();
  FORL(i, nv) { qH[m.dof_Madr[i]] += m.opt.timestep * m.dof_damping[i]; acc[i] = qs[i] + qc[i]; }
  WSYNC();
  factorLD(e, qH, qHd, nullptr);
  if (e.lane == 0) solveLD_lane(m, acc, qH, qHd);
  WSYNC();
  advance_warp(e, act_dot, acc, nullptr);
}

}  // namespace b2k
