// stages_solver.cuh — constraint solve (mj_fwdConstraint), one env per warp.
//
// Replaces mj_projectConstraint + mj_solPGS / mj_solNewton inside the reference's mj_step call
// (mujoco_env.cpp:498; rows M6/M9 of SURVEY 8a).
//
// PGS is run matrix-free: instead of the nefc x nefc matrix AR = J inv(M) J' + R (MuJoCo's efc_AR)
// the warp keeps the nefc x nv factor B = (inv(M) J')' and the running vector a = inv(M) J' f, so a row
// residual is b_i + J_i.a + R_i f_i (one nv-long dot product done with a warp shuffle reduction) and a
// force change costs one nv-long axpy.  Same fixed point and same sweep order as the dense form;
// memory per env drops from nefc^2 to nefc*nv doubles, which is what lets the arena stay in shared
// memory.  Projections: equality free, friction loss box, limits / frictionless / pyramidal rows f>=0,
// elliptic contacts as a block (ray update, then QCQP over the friction dims with the normal fixed).
#pragma once
#include "env_ctx.cuh"
#include "stages_constraint.cuh"
#include "stages_smooth.cuh"

namespace b2k {

// rows of B = inv(M) J' (row i = inv(M) J_i') and the diagonal of AR
__device__ void stage_projectConstraint(const Env& e, int nefc) {
  if (!nefc) return;
  const DevModel& m = e.m;
  const int nv = m.nv;
  const double* J = e.D(B2MJ_F_EFC_J);
  const double* R = e.D(B2MJ_F_EFC_R);
  double* B = e.X(XF_EFC_MINVJT);
  double* ard = e.X(XF_EFC_ARDIAG);
  const double* qLD = e.D(B2MJ_F_QLD);
  const double* dinv = e.D(B2MJ_F_QLDIAGINV);
  FORL(i, nefc) {
    double* b = B + i * nv;
    for (int k = 0; k < nv; k++) b[k] = J[i * nv + k];
    solveLD_lane(m, b, qLD, dinv);
    double s = 0;
    for (int k = 0; k < nv; k++) s += J[i * nv + k] * b[k];
    ard[i] = s + R[i];
  }
  WSYNC();
}

// dot of a constraint row with an nv-vector, result identical on all lanes
__device__ __forceinline__ double rowDot(const Env& e, const double* row, const double* vec, int nv) {
  double s = 0;
  FORL(k, nv) s += row[k] * vec[k];
  return warpSum(s);
}

// small SPD solve helpers for the elliptic QCQP (n <= 5), executed redundantly by every lane
__device__ __forceinline__ int cholFactorSmall(double* A, int n, double mindiag) {
  int bad = 0;
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    if (s < mindiag) { s = mindiag; bad++; }
    const double ljj = sqrt(s);
    A[j * n + j] = ljj;
    const double inv = 1 / ljj;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t * inv;
    }
  }
  return bad;
}
__device__ __forceinline__ void cholSolveSmall(double* x, const double* L, const double* b, int n) {
  for (int i = 0; i < n; i++) {
    double t = b[i];
    for (int k = 0; k < i; k++) t -= L[i * n + k] * x[k];
    x[i] = t / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double t = x[i];
    for (int k = i + 1; k < n; k++) t -= L[k * n + i] * x[k];
    x[i] = t / L[i * n + i];
  }
}
__device__ int QCQP(double* res, const double* Ain, const double* bin, const double* d, double r, int n) {
  double A[25], b[5], P[25], y[5], z[5], nb[5];
  for (int i = 0; i < n; i++) {
    b[i] = bin[i] * d[i];
    y[i] = 0;
    for (int j = 0; j < n; j++) A[i * n + j] = Ain[i * n + j] * d[i] * d[j];
  }
  double la = 0;
  const double r2 = r * r;
  for (int iter = 0; iter < 20; iter++) {
    for (int i = 0; i < n * n; i++) P[i] = A[i];
    for (int i = 0; i < n; i++) P[i * n + i] += la;
    if (cholFactorSmall(P, n, 1e-10)) { la = 0; for (int i = 0; i < n; i++) y[i] = 0; break; }
    for (int i = 0; i < n; i++) nb[i] = -b[i];
    cholSolveSmall(y, P, nb, n);
    double val = -r2;
    for (int i = 0; i < n; i++) val += y[i] * y[i];
    if (val < 1e-10) break;
    cholSolveSmall(z, P, y, n);
    double deriv = 0;
    for (int i = 0; i < n; i++) deriv += -2 * y[i] * z[i];
    const double delta = -val / deriv;
    if (delta < 1e-10) break;
    la += delta;
  }
  for (int i = 0; i < n; i++) res[i] = y[i] * d[i];
  return la != 0;
}

// mj_solPGS (matrix-free).  force holds the warm start on entry.  Returns iterations used.
__device__ int solvePGS(const Env& e, int nefc, double* avec) {
  const DevModel& m = e.m;
  const int nv = m.nv;
  EfcPtrs P = efcPtrs(e);
  const double* B = e.X(XF_EFC_MINVJT);
  const double* ard = e.X(XF_EFC_ARDIAG);
  const int* c_dim = e.I(B2MJ_F_CONTACT_DIM);
  const double* c_fri = e.D(B2MJ_F_CONTACT_FRICTION);
  const double scale = 1 / (m.meaninertia * max(1, nv));
  // a = inv(M) J' f
  FORL(k, nv) {
    double s = 0;
    for (int i = 0; i < nefc; i++) s += B[i * nv + k] * P.force[i];
    avec[k] = s;
  }
  WSYNC();
  int iter = 0;
  while (iter < m.opt.iterations) {
    double improvement = 0;
    for (int i = 0; i < nefc;) {
      const int type = P.type[i];
      if (type != B2MJ_CNSTR_CONTACT_ELLIPTIC) {
        const double fold = P.force[i];
        const double Aii = ard[i];
        const double res = P.b[i] + rowDot(e, P.J + i * nv, avec, nv) + P.R[i] * fold;
        double f = fold - res / Aii;
        if (type == B2MJ_CNSTR_FRICTION_DOF || type == B2MJ_CNSTR_FRICTION_TENDON) {
          const double fl = P.floss[i];
          f = clampd(f, -fl, fl);
        } else if (type != B2MJ_CNSTR_EQUALITY) {
          if (f < 0) f = 0;
        }
        double delta = f - fold;
        double change = 0.5 * delta * delta * Aii + delta * res;
        if (change > 1e-10) { delta = 0; change = 0; f = fold; }
        improvement -= change;
        if (delta != 0) {
          FORL(k, nv) avec[k] += delta * B[i * nv + k];
          if (e.lane == 0) P.force[i] = f;
          WSYNC();
        }
        i += 1;
      } else {
        const int c = P.id[i], dim = c_dim[c];
        const double* fri = c_fri + 5 * c;
        // dim x dim block of AR: J_(i+j) . B_(i+k) (+R on the diagonal)
        double Athis[36], res[6], oldf[6], f[6];
        for (int j = 0; j < dim; j++) {
          for (int k = 0; k < dim; k++) {
            double v = rowDot(e, P.J + (i + j) * nv, B + (i + k) * nv, nv);
            if (j == k) v += P.R[i + j];
            Athis[j * dim + k] = v;
          }
          oldf[j] = P.force[i + j];
          f[j] = oldf[j];
          res[j] = P.b[i + j] + rowDot(e, P.J + (i + j) * nv, avec, nv) + P.R[i + j] * oldf[j];
        }
        if (f[0] < B2K_MINVAL) {
          f[0] -= res[0] / Athis[0];
          if (f[0] < 0) f[0] = 0;
          for (int j = 1; j < dim; j++) f[j] = 0;
        } else {
          double v[6], v1[6];
          for (int j = 0; j < dim; j++) v[j] = f[j];
          double denom = 0, num = 0;
          for (int j = 0; j < dim; j++) {
            double s = 0;
            for (int k = 0; k < dim; k++) s += Athis[j * dim + k] * v[k];
            v1[j] = s;
          }
          for (int j = 0; j < dim; j++) { denom += v[j] * v1[j]; num += v[j] * res[j]; }
          if (denom >= B2K_MINVAL) {
            double x = -num / denom;
            if (f[0] + x * v[0] < 0) x = -v[0] / f[0];
            for (int j = 0; j < dim; j++) f[j] += x * v[j];
          }
        }
        if (f[0] < B2K_MINVAL) {
          for (int j = 1; j < dim; j++) f[j] = 0;
        } else {
          double Ac[25], bc[5], v[5];
          for (int j = 0; j < dim - 1; j++) {
            for (int k = 0; k < dim - 1; k++) Ac[j * (dim - 1) + k] = Athis[(j + 1) * dim + k + 1];
            double t = res[j + 1];
            for (int k = 0; k < dim; k++) t -= Athis[(j + 1) * dim + k] * oldf[k];
            t += Athis[(j + 1) * dim] * f[0];
            bc[j] = t;
          }
          const int active = QCQP(v, Ac, bc, fri, f[0], dim - 1);
          if (active) {
            double s = 0;
            for (int j = 0; j < dim - 1; j++) s += (v[j] / fri[j]) * (v[j] / fri[j]);
            s = sqrt(f[0] * f[0] / fmax(B2K_MINVAL, s));
            for (int j = 0; j < dim - 1; j++) v[j] *= s;
          }
          for (int j = 0; j < dim - 1; j++) f[1 + j] = v[j];
        }
        double change = 0, delta[6];
        for (int j = 0; j < dim; j++) delta[j] = f[j] - oldf[j];
        for (int j = 0; j < dim; j++) {
          double s = 0;
          for (int k = 0; k < dim; k++) s += Athis[j * dim + k] * delta[k];
          change += 0.5 * delta[j] * s + delta[j] * res[j];
        }
        if (change > 1e-10) {
          change = 0;
          for (int j = 0; j < dim; j++) { delta[j] = 0; f[j] = oldf[j]; }
        }
        improvement -= change;
        for (int j = 0; j < dim; j++) {
          if (delta[j] != 0) FORL(k, nv) avec[k] += delta[j] * B[(i + j) * nv + k];
        }
        if (e.lane < dim) P.force[i + e.lane] = f[e.lane];
        WSYNC();
        i += dim;
      }
    }
    improvement *= scale;
    iter++;
    if (improvement < m.opt.tolerance) break;
  }
  return iter;
}

// mj_fwdConstraint.  Returns solver iterations.
__device__ int stage_fwdConstraint(const Env& e, int nefc, int ncon) {
  const DevModel& m = e.m;
  const int nv = m.nv;
  double* qacc = e.D(B2MJ_F_QACC);
  double* warm = e.D(B2MJ_F_QACC_WARMSTART);
  double* qfc = e.D(B2MJ_F_QFRC_CONSTRAINT);
  const double* qas = e.D(B2MJ_F_QACC_SMOOTH);
  if (nefc == 0) {
    FORL(i, nv) { const double a = qas[i]; qacc[i] = a; warm[i] = a; qfc[i] = 0; }
    WSYNC();
    return 0;
  }
  EfcPtrs P = efcPtrs(e);
  // efc_b = J qacc_smooth - aref
  FORL(i, nefc) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += P.J[i * nv + k] * qas[k];
    P.b[i] = s - P.aref[i];
  }
  WSYNC();
  const bool warmstart = !(m.opt.disableflags & B2MJ_DSBL_WARMSTART);
  int iters = 0;
  if (m.opt.solver == B2MJ_SOL_PGS) {
    double* jar = e.X(XF_EFC_JAREF);
    double* avec = e.X(XF_VEC1);
    const double* B = e.X(XF_EFC_MINVJT);
    if (warmstart) {
      FORL(i, nefc) {
        double s = 0;
        for (int k = 0; k < nv; k++) s += P.J[i * nv + k] * warm[k];
        jar[i] = s - P.aref[i];
      }
      WSYNC();
      constraintUpdate_warp(e, nefc, ncon, jar, false);
      // dual cost 0.5 f'ARf + f'b with AR f = J (inv(M) J' f) + R f
      FORL(k, nv) {
        double s = 0;
        for (int i = 0; i < nefc; i++) s += B[i * nv + k] * P.force[i];
        avec[k] = s;
      }
      WSYNC();
      double cost = 0;
      FORL(i, nefc) {
        double s = 0;
        for (int k = 0; k < nv; k++) s += P.J[i * nv + k] * avec[k];
        cost += P.force[i] * (0.5 * (s + P.R[i] * P.force[i]) + P.b[i]);
      }
      cost = warpSum(cost);
      if (cost > 0) { FORL(i, nefc) P.force[i] = 0; }
      WSYNC();
    } else {
      FORL(i, nefc) P.force[i] = 0;
      WSYNC();
    }
    iters = solvePGS(e, nefc, avec);
    // dual finish: qfrc_constraint = J' f ; qacc = qacc_smooth + inv(M) qfrc_constraint
    mulJacTVec_warp(e, nefc, qfc, P.force);
    double* tmp = e.X(XF_VEC2);
    FORL(i, nv) tmp[i] = qfc[i];
    WSYNC();
    if (e.lane == 0) solveLD_lane(m, tmp, e.D(B2MJ_F_QLD), e.D(B2MJ_F_QLDIAGINV));
    WSYNC();
    FORL(i, nv) { const double a = qas[i] + tmp[i]; qacc[i] = a; warm[i] = a; }
    WSYNC();
  }
  return iters;
}

}  // namespace b2k
