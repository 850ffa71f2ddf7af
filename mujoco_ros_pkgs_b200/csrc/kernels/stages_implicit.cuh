// stages_implicit.cuh — implicit-in-velocity integrators (mjINT_IMPLICIT, mjINT_IMPLICITFAST), one env per warp.
//
// Replaces mj_implicit inside the reference's mj_step call (mujoco_ros/src/mujoco_env.cpp:498) for models that select
// those integrators -- the reference exposes the choice in its option panel (mujoco_ros/src/viewer.cpp:579-582).
// Row M11 of SURVEY 8(a).  Semantics (mj_implicitSkip + mjd_smooth_vel of MuJoCo 2.3.7):
//
//   qDeriv = d(qfrc_actuator + qfrc_passive [- qfrc_bias]) / d qvel on MuJoCo's pattern "D" (dof pairs on one chain)
//   implicitfast: (M - h qDeriv) symmetric  -> the same inverse / L'DL machinery as the Euler damping matrix qH
//   implicit:     (M - h qDeriv) with the RNE term -> dense reverse-order LU without pivoting (mju_factorLUSparse order)
//
// B200 form of the RNE derivative: MuJoCo walks the tree once per column; here the chain / subtree bit masks give
// closed forms, so all (column, body) pairs are independent work items:
//   d cvel_b / d v_c      = cdof_c                                   if c on chain(b)
//   d cdof_dot_k / d v_c  = cdof_c x_m cdof_k                        if c in pre(k)   (dof_premask, see stage_comVel)
//   d cacc_b / d v_c      = sum_{k on chain(b)} [c in pre(k)] (cdof_c x_m cdof_k) v_k  +  [c on chain(b)] cdof_dot_c
//   d cfrc_b / d v_c      = I_b dcacc + dcvel x_f (I_b cvel_b) + cvel_b x_f (I_b dcvel)
//   d bias_k / d v_c      = cdof_k . sum_{b' in subtree(body(k))} d cfrc_b' / d v_c
#pragma once
#include "stages_smooth.cuh"

namespace b2k {

__device__ __forceinline__ bool maskBit(const unsigned* mask, int k) { return (mask[k >> 5] >> (k & 31)) & 1u; }

// dofs i, j lie on one ancestor chain (MuJoCo's qDeriv sparsity pattern)
__device__ __forceinline__ bool sameChain(int i, int j) {
  const DevModel& m = c_dm;
  const int hi = i > j ? i : j, lo = i > j ? j : i;
  return maskBit(m.body_dofmask + m.dof_bodyid[hi] * m.nmaskword, lo);
}

// entry (i, j) of d(qfrc_actuator + qfrc_passive)/d qvel (mjd_actuator_vel + mjd_passive_vel); (i, j) on one chain
__device__ __forceinline__ double smoothVelEntry(const Env e, int i, int j) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  double s = 0;
  if (m.nu && !(m.opt.disableflags & B2MJ_DSBL_ACTUATION)) {
    const double* mom = e.D(B2MJ_F_ACTUATOR_MOMENT);
    const double* ctrl = e.D(B2MJ_F_CTRL);
    const double* act = m.na ? e.D(B2MJ_F_ACT) : nullptr;
    B2K_NOUNROLL for (int a = 0; a < m.nu; a++) {
      double bias_vel = 0, gain_vel = 0;
      if (m.actuator_biastype[a] == B2MJ_BIAS_AFFINE) bias_vel = m.actuator_biasprm[B2MJ_NBIAS * a + 2];
      if (m.actuator_gaintype[a] == B2MJ_GAIN_AFFINE) gain_vel = m.actuator_gainprm[B2MJ_NGAIN * a + 2];
      if (gain_vel != 0) {
        const int ad = m.actuator_actadr[a];
        bias_vel += gain_vel * (ad < 0 ? ctrl[a] : act[ad]);
      }
      if (bias_vel != 0) s += mom[a * nv + j] * (mom[a * nv + i] * bias_vel);
    }
  }
  if (!(m.opt.disableflags & B2MJ_DSBL_PASSIVE)) {
    if (i == j) s -= m.dof_damping[i];
    if (m.ntendon) {
      const double* tJ = e.D(B2MJ_F_TEN_J);
      B2K_NOUNROLL for (int t = 0; t < m.ntendon; t++) {
        const double b = m.tendon_damping[t];
        if (b > 0) s += tJ[t * nv + j] * (tJ[t * nv + i] * -b);
      }
    }
  }
  return s;
}

// implicitfast: qH = M - h qDeriv (symmetric), factorised like the Euler damping matrix; acc = inv(qH) rhs
__device__ __noinline__ void implicitFastSolve(const Env e, double* acc, const double* rhs) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  const double h = m.opt.timestep;
  const double* qM = e.D(B2MJ_F_QM);
  if (m.dense_small) {
    double* Hinv = e.X(XF_HINV);
    FORL(k, nv * nv) Hinv[k] = 0;
    WSYNC();
    FORL(t, m.nM) {
      const int i = m.M_row[t], j = m.M_col[t];
      const double v = qM[t] - h * smoothVelEntry(e, i, j);
      Hinv[i * nv + j] = v;
      Hinv[j * nv + i] = v;
    }
    WSYNC();
    invertSPD2(e, Hinv, nullptr, nv);
    mulDense_warp(e, acc, Hinv, rhs, nv);
    return;
  }
  double* qH = e.XG(XF_QH);
  FORL(t, m.nM) qH[t] = qM[t] - h * smoothVelEntry(e, m.M_row[t], m.M_col[t]);
  WSYNC();
  factorLD2(e, qH, nullptr, e.XG(XF_QHDIAGINV), nullptr, nullptr);
  invL2(e, qH, e.XG(XF_QHW), nullptr, nullptr);
  FORL(i, nv) acc[i] = rhs[i];
  WSYNC();
  solveW_warp(e, acc, e.XG(XF_QHW), e.XG(XF_QHDIAGINV), e.X(XF_VEC0));
}

// implicit: A = M - h qDeriv with the RNE velocity derivative, dense in the env's L2 arena; A = (U + I) L
__device__ __noinline__ void implicitFullSolve(const Env e, double* acc, const double* rhs) {
  const DevModel& m = c_dm;
  const int nv = m.nv, nb = m.nbody;
  const double h = m.opt.timestep;
  const double* qM = e.D(B2MJ_F_QM);
  const double* cdof = e.D(B2MJ_F_CDOF);
  const double* cdof_dot = e.D(B2MJ_F_CDOF_DOT);
  const double* cvel = e.D(B2MJ_F_CVEL);
  const double* cinert = e.D(B2MJ_F_CINERT);
  const double* qvel = e.D(B2MJ_F_QVEL);
  double* A = e.XG(XF_IMPL_LU);
  double* dF = e.XG(XF_IMPL_D);  // [nv columns][nbody][6]: d cfrc_body / d qvel_c
  FORL(k, nv * nv) A[k] = 0;
  // (1) own-body force derivative of every (column, body) pair
  FORL(item, nv * nb) {
    const int c = item / nb, b = item - c * nb;
    double f[6] = {0, 0, 0, 0, 0, 0};
    const unsigned* chain = m.body_dofmask + b * m.nmaskword;
    // only bodies below (or carrying) dof c move with it
    if (b && maskBit(chain, c)) {
      const double* cc = cdof + 6 * c;
      double a[6];
      for (int t = 0; t < 6; t++) a[t] = cdof_dot[6 * c + t];
      FOR_MASK_BITS(k, chain, m.nmaskword, {
        if (maskBit(m.dof_premask + k * m.nmaskword, c)) {
          double x[6];
          crossMotion(x, cc, cdof + 6 * k);
          const double v = qvel[k];
          for (int t = 0; t < 6; t++) a[t] += x[t] * v;
        }
      })
      double Iv[6], Idv[6], f1[6], f2[6];
      mulInertVec(f, cinert + 10 * b, a);
      mulInertVec(Iv, cinert + 10 * b, cvel + 6 * b);
      mulInertVec(Idv, cinert + 10 * b, cc);
      crossForce(f1, cc, Iv);
      crossForce(f2, cvel + 6 * b, Idv);
      for (int t = 0; t < 6; t++) f[t] += f1[t] + f2[t];
    }
    for (int t = 0; t < 6; t++) dF[(size_t)item * 6 + t] = f[t];
  }
  WSYNC();
  // (2) A = M - h (smooth derivative - d bias / d qvel) on the chain pattern
  FORL(item, nv * nv) {
    const int i = item / nv, j = item - i * nv;
    if (!sameChain(i, j)) continue;
    // M(i, j): sparse row of the deeper dof, offset = number of chain steps from it up to the other dof
    const int hi = i > j ? i : j, lo = i > j ? j : i;
    int adr = m.dof_Madr[hi];
    for (int k = hi; k != lo; k = m.dof_parentid[k]) adr++;
    double s[6] = {0, 0, 0, 0, 0, 0};
    const unsigned* sub = m.body_submask + m.dof_bodyid[i] * m.nbodyword;
    FOR_MASK_BITS(b, sub, m.nbodyword, { for (int t = 0; t < 6; t++) s[t] += dF[((size_t)j * nb + b) * 6 + t]; })
    A[item] = qM[adr] - h * (smoothVelEntry(e, i, j) - dot6(cdof + 6 * i, s));
  }
  WSYNC();
  // (3) reverse-order LU without pivoting (mju_factorLUSparse): row j < i owned by a lane
  B2K_NOUNROLL for (int i = nv - 1; i > 0; i--) {
    const double piv = A[i * nv + i];
    FORL(j, i) {
      const double aji = A[j * nv + i];
      if (aji == 0) continue;
      const double f = aji / piv;
      A[j * nv + i] = f;
      B2K_NOUNROLL for (int k = 0; k < i; k++) A[j * nv + k] -= A[i * nv + k] * f;
    }
    WSYNC();
  }
  // (4) (U + I) y = rhs, then L acc = y
  FORL(i, nv) acc[i] = rhs[i];
  WSYNC();
  B2K_NOUNROLL for (int i = nv - 1; i > 0; i--) {
    const double xi = acc[i];
    FORL(j, i) acc[j] -= A[j * nv + i] * xi;
    WSYNC();
  }
  B2K_NOUNROLL for (int i = 0; i < nv; i++) {
    double s = 0;
    FORL(k, i) s += A[i * nv + k] * acc[k];
    s = warpSum(e.mask, s);
    if (e.lane == 0) acc[i] = (acc[i] - s) / A[i * nv + i];
    WSYNC();
  }
}

// mj_implicit: solve for the implicit acceleration, then mj_advance
__device__ __noinline__ void stage_implicit(const Env e) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  const double* act_dot = m.na ? e.D(B2MJ_F_ACT_DOT) : nullptr;
  double* rhs = e.X(XF_VEC1);
  double* acc = e.X(XF_VEC2);
  const double* qs = e.D(B2MJ_F_QFRC_SMOOTH);
  const double* qc = e.D(B2MJ_F_QFRC_CONSTRAINT);
  FORL(i, nv) rhs[i] = qs[i] + qc[i];
  WSYNC();
  if (m.opt.integrator == B2MJ_INT_IMPLICIT) implicitFullSolve(e, acc, rhs);
  else implicitFastSolve(e, acc, rhs);
  advance_warp(e, act_dot, acc, nullptr);
}

}  // namespace b2k
