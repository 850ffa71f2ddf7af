// set_const.cpp — qpos0-dependent model constants (the role mj_setConst plays for the reference:
// callbacks.cpp:254,582 re-run it after mass / geom edits).  Runs once per model compile / edit on the
// host; it is model preparation, not the step path.
//
// Derived here: body_subtreemass, dof_M0, dof_invweight0, body_invweight0, tendon_length0,
// tendon_invweight0, actuator_length0, actuator_acc0, stat.{meaninertia,meanmass}.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "hostmath.h"
#include "model_core.h"

namespace b2mj {
using namespace hm;

namespace {

struct Kin0 {
  std::vector<double> xpos, xquat, xmat, xipos, ximat, xanchor, xaxis, com, cinert, cdof, qM, qLD, dinv;
};

void kinematics0(const b2mjModel* m, Kin0& k) {
  const int nb = m->nbody;
  k.xpos.assign(3 * nb, 0); k.xquat.assign(4 * nb, 0); k.xmat.assign(9 * nb, 0);
  k.xipos.assign(3 * nb, 0); k.ximat.assign(9 * nb, 0);
  k.xanchor.assign(3 * m->njnt + 1, 0); k.xaxis.assign(3 * m->njnt + 1, 0);
  k.xquat[0] = 1; k.xmat[0] = k.xmat[4] = k.xmat[8] = 1; k.ximat[0] = k.ximat[4] = k.ximat[8] = 1;
  const double* qpos = m->qpos0;
  for (int i = 1; i < nb; i++) {
    double xp[3], xq[4];
    int ja = m->body_jntadr[i], jn = m->body_jntnum[i];
    if (jn == 1 && m->jnt_type[ja] == B2MJ_JNT_FREE) {
      int qa = m->jnt_qposadr[ja];
      copy3(xp, qpos + qa);
      copy4(xq, qpos + qa + 3);
      normalize4(xq);
      copy3(&k.xanchor[3 * ja], xp);
      copy3(&k.xaxis[3 * ja], m->jnt_axis + 3 * ja);
    } else {
      int pid = m->body_parentid[i];
      double v[3];
      mulmatvec3(v, &k.xmat[9 * pid], m->body_pos + 3 * i);
      for (int c = 0; c < 3; c++) xp[c] = v[c] + k.xpos[3 * pid + c];
      mulquat(xq, &k.xquat[4 * pid], m->body_quat + 4 * i);
      for (int j = 0; j < jn; j++) {
        int jid = ja + j, qa = m->jnt_qposadr[jid];
        double ax[3], an[3];
        rotvecquat(ax, m->jnt_axis + 3 * jid, xq);
        rotvecquat(an, m->jnt_pos + 3 * jid, xq);
        for (int c = 0; c < 3; c++) an[c] += xp[c];
        int t = m->jnt_type[jid];
        if (t == B2MJ_JNT_SLIDE) {
          for (int c = 0; c < 3; c++) xp[c] += ax[c] * (qpos[qa] - m->qpos0[qa]);
        } else if (t == B2MJ_JNT_BALL || t == B2MJ_JNT_HINGE) {
          double ql[4];
          if (t == B2MJ_JNT_BALL) copy4(ql, qpos + qa);
          else axisangle2quat(ql, m->jnt_axis + 3 * jid, qpos[qa] - m->qpos0[qa]);
          mulquat(xq, xq, ql);
          double v2[3];
          rotvecquat(v2, m->jnt_pos + 3 * jid, xq);
          for (int c = 0; c < 3; c++) xp[c] = an[c] - v2[c];
        }
        copy3(&k.xanchor[3 * jid], an);
        copy3(&k.xaxis[3 * jid], ax);
      }
    }
    normalize4(xq);
    copy3(&k.xpos[3 * i], xp);
    copy4(&k.xquat[4 * i], xq);
    quat2mat(&k.xmat[9 * i], xq);
    double v[3], q[4];
    mulmatvec3(v, &k.xmat[9 * i], m->body_ipos + 3 * i);
    for (int c = 0; c < 3; c++) k.xipos[3 * i + c] = v[c] + xp[c];
    mulquat(q, xq, m->body_iquat + 4 * i);
    quat2mat(&k.ximat[9 * i], q);
  }
}

void inert_com(double* res, const double* inert, const double* mat, const double* dif, double mass) {
  double tmp[9];
  tmp[0] = mat[0] * inert[0]; tmp[3] = mat[1] * inert[1]; tmp[6] = mat[2] * inert[2];
  tmp[1] = mat[3] * inert[0]; tmp[4] = mat[4] * inert[1]; tmp[7] = mat[5] * inert[2];
  tmp[2] = mat[6] * inert[0]; tmp[5] = mat[7] * inert[1]; tmp[8] = mat[8] * inert[2];
  res[0] = mat[0] * tmp[0] + mat[1] * tmp[3] + mat[2] * tmp[6];
  res[1] = mat[3] * tmp[1] + mat[4] * tmp[4] + mat[5] * tmp[7];
  res[2] = mat[6] * tmp[2] + mat[7] * tmp[5] + mat[8] * tmp[8];
  res[3] = mat[0] * tmp[1] + mat[1] * tmp[4] + mat[2] * tmp[7];
  res[4] = mat[0] * tmp[2] + mat[1] * tmp[5] + mat[2] * tmp[8];
  res[5] = mat[3] * tmp[2] + mat[4] * tmp[5] + mat[5] * tmp[8];
  res[0] += mass * (dif[1] * dif[1] + dif[2] * dif[2]);
  res[1] += mass * (dif[0] * dif[0] + dif[2] * dif[2]);
  res[2] += mass * (dif[0] * dif[0] + dif[1] * dif[1]);
  res[3] -= mass * dif[0] * dif[1];
  res[4] -= mass * dif[0] * dif[2];
  res[5] -= mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2];
  res[9] = mass;
}

void mul_inert_vec(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}

void com_crb_factor(const b2mjModel* m, Kin0& k) {
  const int nb = m->nbody, nv = m->nv;
  k.com.assign(3 * nb, 0);
  for (int i = nb - 1; i >= 0; i--) {
    for (int c = 0; c < 3; c++) k.com[3 * i + c] += k.xipos[3 * i + c] * m->body_mass[i];
    if (i) for (int c = 0; c < 3; c++) k.com[3 * m->body_parentid[i] + c] += k.com[3 * i + c];
    if (m->body_subtreemass[i] < B2MJ_MINVAL) copy3(&k.com[3 * i], &k.xipos[3 * i]);
    else for (int c = 0; c < 3; c++) k.com[3 * i + c] /= m->body_subtreemass[i];
  }
  k.cinert.assign(10 * nb, 0);
  for (int i = 1; i < nb; i++) {
    double off[3];
    for (int c = 0; c < 3; c++) off[c] = k.xipos[3 * i + c] - k.com[3 * m->body_rootid[i] + c];
    inert_com(&k.cinert[10 * i], m->body_inertia + 3 * i, &k.ximat[9 * i], off, m->body_mass[i]);
  }
  k.cdof.assign(6 * nv + 1, 0);
  for (int j = 0; j < m->njnt; j++) {
    int da = 6 * m->jnt_dofadr[j], bi = m->jnt_bodyid[j];
    double off[3];
    for (int c = 0; c < 3; c++) off[c] = k.com[3 * m->body_rootid[bi] + c] - k.xanchor[3 * j + c];
    int skip = 0;
    switch (m->jnt_type[j]) {
      case B2MJ_JNT_FREE:
        for (int c = 0; c < 3; c++) k.cdof[da + 3 + 7 * c] = 1;
        skip = 18;
        [[fallthrough]];
      case B2MJ_JNT_BALL:
        for (int c = 0; c < 3; c++) {
          double ax[3] = {k.xmat[9 * bi + c], k.xmat[9 * bi + c + 3], k.xmat[9 * bi + c + 6]};
          double* d = &k.cdof[da + skip + 6 * c];
          copy3(d, ax);
          cross(d + 3, ax, off);
        }
        break;
      case B2MJ_JNT_SLIDE:
        copy3(&k.cdof[da + 3], &k.xaxis[3 * j]);
        break;
      case B2MJ_JNT_HINGE:
        copy3(&k.cdof[da], &k.xaxis[3 * j]);
        cross(&k.cdof[da + 3], &k.xaxis[3 * j], off);
        break;
    }
  }
  // composite rigid body + joint-space inertia (ancestor-chain sparse storage)
  std::vector<double> crb(k.cinert);
  for (int i = nb - 1; i > 0; i--)
    if (m->body_parentid[i] > 0)
      for (int c = 0; c < 10; c++) crb[10 * m->body_parentid[i] + c] += crb[10 * i + c];
  k.qM.assign(m->nM + 1, 0);
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    double buf[6];
    mul_inert_vec(buf, &crb[10 * m->dof_bodyid[i]], &k.cdof[6 * i]);
    k.qM[adr] = m->dof_armature[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double s = 0;
      for (int c = 0; c < 6; c++) s += k.cdof[6 * j + c] * buf[c];
      k.qM[adr++] += s;
    }
  }
  // L'DL factorisation
  k.qLD = k.qM;
  for (int kk = nv - 1; kk >= 0; kk--) {
    int Mkk = m->dof_Madr[kk];
    int i = m->dof_parentid[kk], Mki = Mkk + 1;
    while (i >= 0) {
      double tmp = k.qLD[Mki] / k.qLD[Mkk];
      int cnt = (i < nv - 1 ? m->dof_Madr[i + 1] : m->nM) - m->dof_Madr[i];
      for (int c = 0; c < cnt; c++) k.qLD[m->dof_Madr[i] + c] -= k.qLD[Mki + c] * tmp;
      k.qLD[Mki] = tmp;
      i = m->dof_parentid[i];
      Mki++;
    }
  }
  k.dinv.assign(nv + 1, 0);
  for (int i = 0; i < nv; i++) k.dinv[i] = 1.0 / k.qLD[m->dof_Madr[i]];
}

void solve_ld(const b2mjModel* m, const Kin0& k, double* x) {
  const int nv = m->nv;
  for (int i = nv - 1; i >= 0; i--) {
    double t = x[i];
    if (t == 0) continue;
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[j] -= k.qLD[adr++] * t;
  }
  for (int i = 0; i < nv; i++) x[i] *= k.dinv[i];
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[i] -= k.qLD[adr++] * x[j];
  }
}

// translational / rotational Jacobians of a world point attached to `body` (3 x nv each)
void jac_point(const b2mjModel* m, const Kin0& k, double* jp, double* jr, const double* point, int body) {
  const int nv = m->nv;
  std::memset(jp, 0, sizeof(double) * 3 * nv);
  std::memset(jr, 0, sizeof(double) * 3 * nv);
  double off[3];
  for (int c = 0; c < 3; c++) off[c] = point[c] - k.com[3 * m->body_rootid[body] + c];
  while (body && !m->body_dofnum[body]) body = m->body_parentid[body];
  if (!body) return;
  for (int i = m->body_dofadr[body] + m->body_dofnum[body] - 1; i >= 0; i = m->dof_parentid[i]) {
    const double* cd = &k.cdof[6 * i];
    double cr[3];
    cross(cr, cd, off);
    for (int c = 0; c < 3; c++) {
      jr[c * nv + i] = cd[c];
      jp[c * nv + i] = cd[3 + c] + cr[c];
    }
  }
}

}  // namespace

// body frames at qpos0 (the compiler needs them for default equality relposes)
void model_body_poses0(const b2mjModel* m, std::vector<double>& xpos, std::vector<double>& xquat) {
  Kin0 k;
  kinematics0(m, k);
  xpos = k.xpos;
  xquat = k.xquat;
}

int model_set_const(b2mjModel* m, std::string& err) {
  const int nb = m->nbody, nv = m->nv;
  // subtree masses
  for (int i = 0; i < nb; i++) m->body_subtreemass[i] = m->body_mass[i];
  for (int i = nb - 1; i > 0; i--) m->body_subtreemass[m->body_parentid[i]] += m->body_subtreemass[i];

  Kin0 k;
  kinematics0(m, k);
  com_crb_factor(m, k);
  for (int i = 0; i < nv; i++) {
    if (!(k.qLD[m->dof_Madr[i]] > 0)) {
      err = "set_const: joint-space inertia is not positive definite at qpos0 (dof " + std::to_string(i) +
            "); every moving body needs mass and inertia";
      return B2MJ_EINVAL;
    }
  }
  // dof_M0, dof_invweight0
  std::vector<double> x(nv + 1);
  double meaninertia = 0;
  for (int i = 0; i < nv; i++) {
    m->dof_M0[i] = k.qM[m->dof_Madr[i]];
    meaninertia += m->dof_M0[i];
    std::fill(x.begin(), x.end(), 0.0);
    x[i] = 1;
    solve_ld(m, k, x.data());
    m->dof_invweight0[i] = x[i];
  }
  m->stat.meaninertia = nv ? meaninertia / nv : 1.0;
  // average over the dofs of ball joints and each 3-block of free joints
  for (int j = 0; j < m->njnt; j++) {
    int da = m->jnt_dofadr[j];
    auto avg3 = [&](int a) {
      double s = (m->dof_invweight0[a] + m->dof_invweight0[a + 1] + m->dof_invweight0[a + 2]) / 3;
      m->dof_invweight0[a] = m->dof_invweight0[a + 1] = m->dof_invweight0[a + 2] = s;
    };
    if (m->jnt_type[j] == B2MJ_JNT_FREE) { avg3(da); avg3(da + 3); }
    else if (m->jnt_type[j] == B2MJ_JNT_BALL) avg3(da);
  }
  // body_invweight0 = trace(J inv(M) J')/3 for translational and rotational Jacobians at the body COM
  std::vector<double> jp(3 * nv + 1), jr(3 * nv + 1);
  double meanmass = 0;
  for (int i = 0; i < nb; i++) {
    m->body_invweight0[2 * i] = m->body_invweight0[2 * i + 1] = 0;
    if (i == 0) continue;
    meanmass += m->body_mass[i];
    if (nv == 0) continue;
    jac_point(m, k, jp.data(), jr.data(), &k.xipos[3 * i], i);
    double trp = 0, trr = 0;
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < nv; c++) x[c] = jp[r * nv + c];
      solve_ld(m, k, x.data());
      for (int c = 0; c < nv; c++) trp += jp[r * nv + c] * x[c];
      for (int c = 0; c < nv; c++) x[c] = jr[r * nv + c];
      solve_ld(m, k, x.data());
      for (int c = 0; c < nv; c++) trr += jr[r * nv + c] * x[c];
    }
    m->body_invweight0[2 * i] = trp / 3;
    m->body_invweight0[2 * i + 1] = trr / 3;
  }
  m->stat.meanmass = nb > 1 ? meanmass / (nb - 1) : 0;
  // tendons: length0, invweight0 (fixed: joint coefficients; spatial: site path with pulley divisors)
  std::vector<double> tenJ((size_t)m->ntendon * (nv + 1), 0.0);
  for (int t = 0; t < m->ntendon; t++) {
    std::fill(x.begin(), x.end(), 0.0);
    double* J = &tenJ[(size_t)t * (nv + 1)];
    double len = 0, divisor = 1;
    const int w0 = m->tendon_adr[t], w1 = w0 + m->tendon_num[t];
    for (int w = w0; w < w1; w++) {
      if (m->wrap_type[w] == 1) {
        int jid = m->wrap_objid[w];
        len += m->wrap_prm[w] * m->qpos0[m->jnt_qposadr[jid]];
        J[m->jnt_dofadr[jid]] = m->wrap_prm[w];
      } else if (m->wrap_type[w] == 2) {
        divisor = m->wrap_prm[w];
      } else if (m->wrap_type[w] == 3 && w + 1 < w1 && m->wrap_type[w + 1] == 3) {
        const int s0 = m->wrap_objid[w], s1 = m->wrap_objid[w + 1], b0 = m->site_bodyid[s0], b1 = m->site_bodyid[s1];
        double p0[3], p1[3], dif[3];
        mulmatvec3(p0, &k.xmat[9 * b0], m->site_pos + 3 * s0);
        mulmatvec3(p1, &k.xmat[9 * b1], m->site_pos + 3 * s1);
        for (int c = 0; c < 3; c++) { p0[c] += k.xpos[3 * b0 + c]; p1[c] += k.xpos[3 * b1 + c]; dif[c] = p1[c] - p0[c]; }
        const double seg = normalize3(dif);
        len += seg / divisor;
        if (b0 != b1) {
          std::vector<double> j0(3 * nv + 3, 0.0), j1(3 * nv + 3, 0.0), jr(3 * nv + 3);
          jac_point(m, k, j0.data(), jr.data(), p0, b0);
          jac_point(m, k, j1.data(), jr.data(), p1, b1);
          for (int c = 0; c < nv; c++) {
            double sdot = 0;
            for (int r = 0; r < 3; r++) sdot += dif[r] * (j1[r * nv + c] - j0[r * nv + c]);
            J[c] += sdot / divisor;
          }
        }
      }
    }
    m->tendon_length0[t] = len;
    for (int c = 0; c < nv; c++) x[c] = J[c];
    solve_ld(m, k, x.data());
    double s = 0;
    for (int c = 0; c < nv; c++) s += J[c] * x[c];
    m->tendon_invweight0[t] = s;
  }
  // actuators: length0, acc0 = |inv(M) moment'|
  for (int a = 0; a < m->nu; a++) {
    std::vector<double> mom(nv + 1, 0.0);
    double len = 0;
    int id = m->actuator_trnid[2 * a];
    double g = m->actuator_gear[6 * a];
    if (m->actuator_trntype[a] == B2MJ_TRN_JOINT || m->actuator_trntype[a] == B2MJ_TRN_JOINTINPARENT) {
      len = g * m->qpos0[m->jnt_qposadr[id]];
      mom[m->jnt_dofadr[id]] = g;
    } else if (m->actuator_trntype[a] == B2MJ_TRN_TENDON) {
      len = g * m->tendon_length0[id];
      for (int c = 0; c < nv; c++) mom[c] = g * tenJ[(size_t)id * (nv + 1) + c];
    }
    m->actuator_length0[a] = len;
    for (int c = 0; c < nv; c++) x[c] = mom[c];
    solve_ld(m, k, x.data());
    double s = 0;
    for (int c = 0; c < nv; c++) s += x[c] * x[c];
    m->actuator_acc0[a] = std::sqrt(s);
  }
  return 0;
}

}  // namespace b2mj
