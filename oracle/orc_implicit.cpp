// orc_implicit.cpp — implicit-in-velocity integrators of the CPU oracle (mj_implicit / mjINT_IMPLICIT, mjINT_IMPLICITFAST).
// TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// Restates MuJoCo 2.3.7 engine_forward.c::mj_implicitSkip and engine_derivative.c::mjd_smooth_vel
// (mjd_actuator_vel, mjd_passive_vel, mjd_rne_vel) -- the integrators the reference exposes through its
// option panel (mujoco_ros/src/viewer.cpp:579-582 "Euler\nRK4\nimplicit\nimplicitfast") and reaches through the
// same mj_step call (mujoco_env.cpp:498).  MuJoCo's source is not in /root/reference; "parity unpinned".
//
//   qDeriv = d(qfrc_actuator + qfrc_passive - qfrc_bias) / d qvel, kept on MuJoCo's sparsity pattern "D" only:
//            entries (i, j) with i == j, j an ancestor dof of i, or i an ancestor dof of j.
//   implicitfast: bias term dropped, qDeriv symmetric -> (M - h qDeriv) qacc = qfrc_smooth + qfrc_constraint by L'DL
//   implicit:     full qDeriv, (M - h qDeriv) factorised by the reverse-order, fill-in free LU of mju_factorLUSparse.
//
// The RNE derivative is computed in forward mode, one qvel direction (column) at a time, by differentiating exactly
// the recursions of orc::comVel and orc::rne; tests/test_integrators_cpu.py checks it against central differences.
#include <cmath>
#include <vector>

#include "orc_math.h"
#include "orc_types.h"

namespace orc {

static bool sameChain(const b2mjModel* m, int i, int j) {
  if (i == j) return true;
  int hi = i > j ? i : j;
  const int lo = i > j ? j : i;
  while (hi > lo) hi = m->dof_parentid[hi];
  return hi == lo;
}

// qDeriv += J' b J for one row J (restricted to the D pattern), mjd addJTBJ with n = 1
static void addJTBJ(const b2mjModel* m, std::vector<double>& qD, const double* J, double b) {
  const int nv = m->nv;
  for (int k = 0; k < nv; k++) {
    if (J[k] == 0) continue;
    for (int c = 0; c < nv; c++)
      if (sameChain(m, k, c)) qD[(size_t)k * nv + c] += J[c] * (J[k] * b);
  }
}

// d qfrc_bias / d qvel, dense [nv][nv] (row = force component, column = velocity component)
void rneVelDerivative(const b2mjModel* m, const OrcData* d, double* dbias) {
  const int nv = m->nv, nb = m->nbody;
  std::vector<double> dcvel(6 * nb), dcacc(6 * nb), dcfrc(6 * nb);
  for (int c = 0; c < nv; c++) {
    std::fill(dcvel.begin(), dcvel.end(), 0.0);
    std::fill(dcacc.begin(), dcacc.end(), 0.0);
    std::fill(dcfrc.begin(), dcfrc.end(), 0.0);
    for (int i = 1; i < nb; i++) {
      const int bda = m->body_dofadr[i], dofnum = m->body_dofnum[i], par = m->body_parentid[i];
      double v[6], a[6], tmp[6];
      copy(v, &dcvel[6 * par], 6);
      copy(a, &dcacc[6 * par], 6);
      // mirror of orc::comVel: cdof_dot_j = crossMotion(cvel so far, cdof_j), cvel += cdof_j qvel_j
      auto addDofs = [&](int j0, int cnt) {  // dcvel += sum cdof_j * dqvel_j over a joint's dofs
        for (int k = 0; k < cnt; k++)
          if (bda + j0 + k == c)
            for (int t = 0; t < 6; t++) v[t] += d->cdof[6 * (bda + j0 + k) + t];
      };
      auto accDot = [&](int j) {  // dcacc += d(cdof_dot_j) qvel_j + cdof_dot_j dqvel_j
        crossMotion(tmp, v, d->cdof + 6 * (bda + j));
        for (int t = 0; t < 6; t++) a[t] += tmp[t] * d->qvel[bda + j];
        if (bda + j == c)
          for (int t = 0; t < 6; t++) a[t] += d->cdof_dot[6 * (bda + j) + t];
      };
      for (int j = 0; j < dofnum; j++) {
        switch (m->jnt_type[m->dof_jntid[bda + j]]) {
          case B2MJ_JNT_FREE:
            // translational dofs: cdof_dot = 0, contributes only cdof_dot_j dqvel_j = 0
            addDofs(j, 3);
            j += 3;
            [[fallthrough]];
          case B2MJ_JNT_BALL:
            for (int k = 0; k < 3; k++) accDot(j + k);
            addDofs(j, 3);
            j += 2;
            break;
          default:
            accDot(j);
            addDofs(j, 1);
        }
      }
      copy(&dcvel[6 * i], v, 6);
      copy(&dcacc[6 * i], a, 6);
      // cfrc = I cacc + cvel x* (I cvel)
      double Iv[6], Idv[6], f1[6], f2[6];
      mulInertVec(&dcfrc[6 * i], d->cinert + 10 * i, a);
      mulInertVec(Iv, d->cinert + 10 * i, d->cvel + 6 * i);
      mulInertVec(Idv, d->cinert + 10 * i, v);
      crossForce(f1, v, Iv);
      crossForce(f2, d->cvel + 6 * i, Idv);
      for (int t = 0; t < 6; t++) dcfrc[6 * i + t] += f1[t] + f2[t];
    }
    for (int i = nb - 1; i > 0; i--)
      if (m->body_parentid[i])
        for (int t = 0; t < 6; t++) dcfrc[6 * m->body_parentid[i] + t] += dcfrc[6 * i + t];
    for (int k = 0; k < nv; k++) dbias[(size_t)k * nv + c] = dot(d->cdof + 6 * k, &dcfrc[6 * m->dof_bodyid[k]], 6);
  }
}

// mjd_smooth_vel: dense qDeriv on the D pattern
void smoothVelDerivative(const b2mjModel* m, const OrcData* d, int flg_bias, std::vector<double>& qD) {
  const int nv = m->nv;
  qD.assign((size_t)nv * nv, 0.0);
  // mjd_actuator_vel
  if (!(m->opt.disableflags & B2MJ_DSBL_ACTUATION)) {
    for (int i = 0; i < m->nu; i++) {
      double bias_vel = 0, gain_vel = 0;
      if (m->actuator_biastype[i] == B2MJ_BIAS_AFFINE) bias_vel = m->actuator_biasprm[B2MJ_NBIAS * i + 2];
      if (m->actuator_gaintype[i] == B2MJ_GAIN_AFFINE) gain_vel = m->actuator_gainprm[B2MJ_NGAIN * i + 2];
      if (gain_vel != 0) {
        const int a = m->actuator_actadr[i];
        bias_vel += gain_vel * (a < 0 ? d->ctrl[i] : d->act[a]);
      }
      if (bias_vel != 0) addJTBJ(m, qD, d->actuator_moment + (size_t)i * nv, bias_vel);
    }
  }
  // mjd_passive_vel
  if (!(m->opt.disableflags & B2MJ_DSBL_PASSIVE)) {
    for (int i = 0; i < nv; i++) qD[(size_t)i * nv + i] -= m->dof_damping[i];
    for (int i = 0; i < m->ntendon; i++)
      if (m->tendon_damping[i] > 0) addJTBJ(m, qD, d->ten_J + (size_t)i * nv, -m->tendon_damping[i]);
  }
  // mjd_rne_vel
  if (flg_bias) {
    std::vector<double> db((size_t)nv * nv);
    rneVelDerivative(m, d, db.data());
    for (int i = 0; i < nv; i++)
      for (int j = 0; j < nv; j++)
        if (sameChain(m, i, j)) qD[(size_t)i * nv + j] -= db[(size_t)i * nv + j];
  }
}

// mju_factorLUSparse / mju_solveLUSparse on the dense image of the D pattern: A = (U + I) L, eliminated from the last
// row upwards without pivoting; no fill-in arises because the pattern is a forest of ancestor chains
static void factorLUReverse(double* A, int n) {
  for (int i = n - 1; i >= 0; i--) {
    const double piv = A[(size_t)i * n + i];
    for (int j = i - 1; j >= 0; j--) {
      if (A[(size_t)j * n + i] == 0) continue;
      A[(size_t)j * n + i] /= piv;
      const double f = A[(size_t)j * n + i];
      for (int k = 0; k < i; k++) A[(size_t)j * n + k] -= A[(size_t)i * n + k] * f;
    }
  }
}
static void solveLUReverse(double* x, const double* LU, int n) {
  // (U + I) y = x
  for (int i = n - 1; i >= 0; i--)
    for (int j = 0; j < i; j++) x[j] -= LU[(size_t)j * n + i] * x[i];
  // L z = y
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= LU[(size_t)i * n + k] * x[k];
    x[i] = s / LU[(size_t)i * n + i];
  }
}

// mj_implicitSkip(m, d, 0) up to (excluding) mj_advance: qacc_out solves (M - h qDeriv) qacc = qfrc_smooth + qfrc_constraint
void implicitQacc(const b2mjModel* m, OrcData* d, double* qacc_out) {
  const int nv = m->nv;
  const double h = m->opt.timestep;
  std::vector<double> qD;
  for (int i = 0; i < nv; i++) qacc_out[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
  if (m->opt.integrator == B2MJ_INT_IMPLICIT) {
    smoothVelDerivative(m, d, 1, qD);
    std::vector<double> LU((size_t)nv * nv, 0.0);
    for (int i = 0; i < nv; i++) {
      int adr = m->dof_Madr[i];
      for (int j = i; j >= 0; j = m->dof_parentid[j]) {
        LU[(size_t)i * nv + j] = d->qM[adr];
        LU[(size_t)j * nv + i] = d->qM[adr];
        adr++;
      }
    }
    for (size_t k = 0; k < LU.size(); k++) LU[k] -= h * qD[k];
    factorLUReverse(LU.data(), nv);
    solveLUReverse(qacc_out, LU.data(), nv);
  } else {
    smoothVelDerivative(m, d, 0, qD);
    // qH = M - h * qDeriv on the entries of M (row i, columns = i and its ancestors)
    for (int i = 0; i < nv; i++) {
      int adr = m->dof_Madr[i];
      for (int j = i; j >= 0; j = m->dof_parentid[j]) {
        d->qH[adr] = d->qM[adr] - h * qD[(size_t)i * nv + j];
        adr++;
      }
    }
    factorI(m, d->qH, d->qH, d->qHDiagInv, nullptr);
    solveLD(m, qacc_out, d->qH, d->qHDiagInv);
  }
}

}  // namespace orc

extern "C" {
// test hooks: the dense derivative matrices after a forward pass
void orc_rne_vel_derivative(const b2mjModel* m, const OrcData* d, double* dbias /* [nv][nv] */) {
  orc::rneVelDerivative(m, d, dbias);
}
void orc_smooth_vel_derivative(const b2mjModel* m, const OrcData* d, int flg_bias, double* qderiv /* [nv][nv] */) {
  std::vector<double> qD;
  orc::smoothVelDerivative(m, d, flg_bias, qD);
  for (size_t k = 0; k < qD.size(); k++) qderiv[k] = qD[k];
}
}
