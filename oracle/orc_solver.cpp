// orc_solver.cpp — constraint solvers of the CPU oracle (PGS dual; CG / Newton primal).
// TEST INFRASTRUCTURE ONLY (see orc_math.h).  Restates MuJoCo 2.3.7 engine_solver.c and
// mj_fwdConstraint (engine_forward.c): warm start, PGS sweep with pyramidal/limit/friction
// projections and the elliptic ray + QCQP block update, primal Newton/CG with exact line search on
// the piecewise-quadratic cost.  This is the M9 row of SURVEY 8(a) — inside the `mj_step` call at
// reference mujoco_env.cpp:498.  Termination as SURVEY A2: scale = 1/(meaninertia*max(1,nv)).
#include <algorithm>
#include <cmath>
#include <vector>

#include "orc_math.h"
#include "orc_types.h"

namespace orc {

// ----------------------------------------------------------------------------------------------
// small dense helpers
// ----------------------------------------------------------------------------------------------

// in-place Cholesky (lower) of an n x n SPD matrix; returns rank deficiency count
static int cholFactor(double* A, int n, double mindiag) {
  int bad = 0;
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    if (s < mindiag) { s = mindiag; bad++; }
    const double ljj = std::sqrt(s);
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

static void cholSolve(double* x, const double* L, const double* b, int n) {
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

// minimise 0.5 x'Ax + x'b  s.t.  sum (x_i/d_i)^2 <= r^2   (mju_QCQP family); returns 1 if the
// constraint is active.  n <= 5.
static int QCQP(double* res, const double* Ain, const double* bin, const double* d, double r, int n) {
  double A[25], b[5], P[25], y[5], z[5];
  for (int i = 0; i < n; i++) {
    b[i] = bin[i] * d[i];
    for (int j = 0; j < n; j++) A[i * n + j] = Ain[i * n + j] * d[i] * d[j];
  }
  double la = 0;
  const double r2 = r * r;
  for (int iter = 0; iter < 20; iter++) {
    for (int i = 0; i < n * n; i++) P[i] = A[i];
    for (int i = 0; i < n; i++) P[i * n + i] += la;
    if (cholFactor(P, n, 1e-10)) { la = 0; for (int i = 0; i < n; i++) y[i] = 0; break; }
    double nb[5];
    for (int i = 0; i < n; i++) nb[i] = -b[i];
    cholSolve(y, P, nb, n);
    double val = -r2;
    for (int i = 0; i < n; i++) val += y[i] * y[i];
    if (val < 1e-10) break;
    cholSolve(z, P, y, n);
    double deriv = 0;
    for (int i = 0; i < n; i++) deriv += -2 * y[i] * z[i];
    const double delta = -val / deriv;
    if (delta < 1e-10) break;
    la += delta;
  }
  for (int i = 0; i < n; i++) res[i] = y[i] * d[i];
  return la != 0;
}

// ----------------------------------------------------------------------------------------------
// PGS (dual)
// ----------------------------------------------------------------------------------------------

static double costChange(const double* A, double* force, const double* oldforce, const double* res, int dim, int lda) {
  double change;
  if (dim == 1) {
    const double delta = force[0] - oldforce[0];
    change = 0.5 * delta * delta * A[0] + delta * res[0];
  } else {
    double delta[6];
    for (int i = 0; i < dim; i++) delta[i] = force[i] - oldforce[i];
    change = 0;
    for (int i = 0; i < dim; i++) {
      double s = 0;
      for (int j = 0; j < dim; j++) s += A[i * lda + j] * delta[j];
      change += 0.5 * delta[i] * s + delta[i] * res[i];
    }
  }
  if (change > 1e-10) {
    for (int i = 0; i < dim; i++) force[i] = oldforce[i];
    change = 0;
  }
  return change;
}

static void solPGS(const b2mjModel* m, OrcData* d, int maxiter) {
  const int nefc = d->nefc();
  const double* AR = d->efc_AR;
  double* force = d->efc_force;
  const double* b = d->efc_b;
  const double scale = 1 / (m->stat.meaninertia * std::max(1, m->nv));
  std::vector<double> ARinv(nefc);
  for (int i = 0; i < nefc; i++) ARinv[i] = 1 / AR[i * nefc + i];
  int iter = 0;
  while (iter < maxiter) {
    double improvement = 0;
    for (int i = 0; i < nefc;) {
      int dim = 1;
      if (d->efc_type[i] == B2MJ_CNSTR_CONTACT_ELLIPTIC) dim = d->contact_dim[d->efc_id[i]];
      double res[6], oldforce[6];
      for (int j = 0; j < dim; j++) {
        res[j] = b[i + j] + dot(AR + (i + j) * nefc, force, nefc);
        oldforce[j] = force[i + j];
      }
      if (dim == 1) {
        force[i] -= res[0] * ARinv[i];
        const int t = d->efc_type[i];
        if (t == B2MJ_CNSTR_FRICTION_DOF || t == B2MJ_CNSTR_FRICTION_TENDON) {
          const double f = d->efc_frictionloss[i];
          force[i] = clampd(force[i], -f, f);
        } else if (t != B2MJ_CNSTR_EQUALITY) {
          if (force[i] < 0) force[i] = 0;
        }
      } else {
        const int c = d->efc_id[i];
        const double* fri = d->contact_friction + 5 * c;
        double Athis[36];
        for (int j = 0; j < dim; j++)
          for (int k = 0; k < dim; k++) Athis[j * dim + k] = AR[(i + j) * nefc + i + k];
        // normal or ray update
        if (force[i] < MINVAL) {
          force[i] -= res[0] * ARinv[i];
          if (force[i] < 0) force[i] = 0;
          for (int j = 1; j < dim; j++) force[i + j] = 0;
        } else {
          double v[6], v1[6];
          for (int j = 0; j < dim; j++) v[j] = force[i + j];
          for (int j = 0; j < dim; j++) v1[j] = dot(Athis + j * dim, v, dim);
          const double denom = dot(v, v1, dim);
          if (denom >= MINVAL) {
            double x = -dot(v, res, dim) / denom;
            if (force[i] + x * v[0] < 0) x = -v[0] / force[i];
            for (int j = 0; j < dim; j++) force[i + j] += x * v[j];
          }
        }
        // friction update with the normal fixed
        if (force[i] < MINVAL) {
          for (int j = 1; j < dim; j++) force[i + j] = 0;
        } else {
          double Ac[25], bc[5], v[5];
          for (int j = 0; j < dim - 1; j++) {
            for (int k = 0; k < dim - 1; k++) Ac[j * (dim - 1) + k] = Athis[(j + 1) * dim + k + 1];
            bc[j] = res[j + 1];
            for (int k = 0; k < dim; k++) bc[j] -= Athis[(j + 1) * dim + k] * oldforce[k];
            bc[j] += Athis[(j + 1) * dim] * force[i];
          }
          const int active = QCQP(v, Ac, bc, fri, force[i], dim - 1);
          if (active) {
            double s = 0;
            for (int j = 0; j < dim - 1; j++) s += (v[j] / fri[j]) * (v[j] / fri[j]);
            s = std::sqrt(force[i] * force[i] / std::fmax(MINVAL, s));
            for (int j = 0; j < dim - 1; j++) v[j] *= s;
          }
          for (int j = 0; j < dim - 1; j++) force[i + 1 + j] = v[j];
        }
      }
      // dim x dim block of AR starting at (i,i), leading dimension nefc
      improvement -= costChange(AR + i * nefc + i, force + i, oldforce, res, dim, nefc);
      i += dim;
    }
    improvement *= scale;
    iter++;
    if (improvement < m->opt.tolerance) break;
  }
  d->solver_iter_[0] = iter;
  // dual finish: map forces to joint space and accelerations
  mulJacTVec(m, d, d->qfrc_constraint, force);
  solveM(m, d, d->qacc, d->qfrc_constraint);
  for (int i = 0; i < m->nv; i++) d->qacc[i] += d->qacc_smooth[i];
}

// mj_solNoSlip: modified PGS over the friction-loss rows and the friction dimensions of contacts, on the dual problem
// WITHOUT the regulariser R (so converged friction forces leave no residual slip); normal forces stay as the main
// solver left them.  Runs after the main solver when opt.noslip_iterations > 0 -- the reference exposes both knobs in
// its option panel (mujoco_ros/src/viewer.cpp:590-591 "Noslip Iter" / "Noslip Tol").
static void solNoSlip(const b2mjModel* m, OrcData* d, int maxiter) {
  const int nefc = d->nefc();
  const double* AR = d->efc_AR;
  const double* R = d->efc_R;
  const double* b = d->efc_b;
  double* force = d->efc_force;
  const double scale = 1 / (m->stat.meaninertia * std::max(1, m->nv));
  auto residual = [&](double* res, int i, int dim) {  // residual of the unregularised system
    for (int j = 0; j < dim; j++) res[j] = b[i + j] + dot(AR + (i + j) * nefc, force, nefc) - R[i + j] * force[i + j];
  };
  auto block = [&](double* Ac, int i, int dim) {
    for (int j = 0; j < dim; j++)
      for (int k = 0; k < dim; k++) Ac[j * dim + k] = AR[(i + j) * nefc + i + k] - (j == k ? R[i + j] : 0.0);
  };
  int iter = 0;
  while (iter < maxiter) {
    double improvement = 0;
    if (iter == 0)  // the cost drops by the regulariser's share when R is removed
      for (int i = 0; i < nefc; i++) improvement += 0.5 * force[i] * force[i] * R[i];
    for (int i = 0; i < nefc; i++) {
      const int t = d->efc_type[i];
      double res[6], oldforce[6], Ac[36];
      if (t == B2MJ_CNSTR_FRICTION_DOF || t == B2MJ_CNSTR_FRICTION_TENDON) {
        residual(res, i, 1);
        oldforce[0] = force[i];
        block(Ac, i, 1);
        force[i] -= res[0] / Ac[0];
        const double fl = d->efc_frictionloss[i];
        force[i] = clampd(force[i], -fl, fl);
        improvement -= costChange(Ac, force + i, oldforce, res, 1, 1);
      } else if (t == B2MJ_CNSTR_CONTACT_PYRAMIDAL) {
        const int dim = d->contact_dim[d->efc_id[i]];
        // pairs of opposing pyramid edges: their sum (the normal load they carry) is kept, the split is re-optimised
        for (int j = i; j < i + 2 * (dim - 1); j += 2) {
          residual(res, j, 2);
          oldforce[0] = force[j]; oldforce[1] = force[j + 1];
          block(Ac, j, 2);
          const double bc0 = res[0] - Ac[0] * oldforce[0] - Ac[1] * oldforce[1];
          const double bc1 = res[1] - Ac[2] * oldforce[0] - Ac[3] * oldforce[1];
          const double mid = 0.5 * (force[j] + force[j + 1]);
          const double K1 = Ac[0] + Ac[3] - Ac[1] - Ac[2], K0 = mid * (Ac[0] - Ac[3]) + bc0 - bc1;
          if (K1 < MINVAL) {
            force[j] = force[j + 1] = mid;
          } else {
            const double y = -K0 / K1;
            if (y < -mid) { force[j] = 0; force[j + 1] = 2 * mid; }
            else if (y > mid) { force[j] = 2 * mid; force[j + 1] = 0; }
            else { force[j] = mid + y; force[j + 1] = mid - y; }
          }
          improvement -= costChange(Ac, force + j, oldforce, res, 2, 2);
        }
        i += 2 * (dim - 1) - 1;
      } else if (t == B2MJ_CNSTR_CONTACT_ELLIPTIC) {
        const int c = d->efc_id[i], dim = d->contact_dim[c];
        const double* mu = d->contact_friction + 5 * c;
        const int fd = dim - 1;
        if (fd > 0) {
          residual(res, i + 1, fd);
          for (int j = 0; j < fd; j++) oldforce[j] = force[i + 1 + j];
          block(Ac, i + 1, fd);
          double bc[5], v[5];
          for (int j = 0; j < fd; j++) {
            bc[j] = res[j];
            for (int k = 0; k < fd; k++) bc[j] -= Ac[j * fd + k] * oldforce[k];
          }
          if (force[i] < MINVAL) {
            for (int j = 0; j < fd; j++) force[i + 1 + j] = 0;
          } else {
            const int active = QCQP(v, Ac, bc, mu, force[i], fd);
            if (active) {
              double s = 0;
              for (int j = 0; j < fd; j++) s += (v[j] / mu[j]) * (v[j] / mu[j]);
              s = std::sqrt(force[i] * force[i] / std::fmax(MINVAL, s));
              for (int j = 0; j < fd; j++) v[j] *= s;
            }
            for (int j = 0; j < fd; j++) force[i + 1 + j] = v[j];
          }
          improvement -= costChange(Ac, force + i + 1, oldforce, res, fd, fd);
        }
        i += dim - 1;
      }
    }
    improvement *= scale;
    iter++;
    if (improvement < m->opt.noslip_tolerance) break;
  }
  d->solver_iter_[0] += iter;
  mulJacTVec(m, d, d->qfrc_constraint, force);
  solveM(m, d, d->qacc, d->qfrc_constraint);
  for (int i = 0; i < m->nv; i++) d->qacc[i] += d->qacc_smooth[i];
}

// ----------------------------------------------------------------------------------------------
// primal solvers (CG, Newton)
// ----------------------------------------------------------------------------------------------

struct Primal {
  const b2mjModel* m;
  OrcData* d;
  int nv, nefc;
  bool newton, cone;
  std::vector<double> Jaref, Jv, Ma, Mv, grad, Mgrad, search, quad, H, tmp;
  double quadGauss[3];
  double cost, gauss, scale;
};

struct LSPoint {
  double alpha, cost, deriv[2];
};

static void primalUpdateConstraint(Primal& c) {
  constraintUpdate(c.m, c.d, c.Jaref.data(), &c.cost, c.newton && c.cone);
  double g = 0;
  for (int i = 0; i < c.nv; i++) g += (c.Ma[i] - c.d->qfrc_smooth[i]) * (c.d->qacc[i] - c.d->qacc_smooth[i]);
  c.gauss = 0.5 * g;
  c.cost += c.gauss;
}

static void primalHessian(Primal& c) {
  const b2mjModel* m = c.m;
  OrcData* d = c.d;
  const int nv = c.nv, nefc = c.nefc;
  std::fill(c.H.begin(), c.H.end(), 0.0);
  // dense M
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      c.H[i * nv + j] = d->qM[adr];
      c.H[j * nv + i] = d->qM[adr];
      adr++;
    }
  }
  for (int r = 0; r < nefc; r++) {
    if (d->efc_state[r] == B2MJ_CSTATE_QUADRATIC) {
      const double* J = d->efc_J + r * nv;
      const double D = d->efc_D[r];
      for (int i = 0; i < nv; i++) {
        if (J[i] == 0) continue;
        const double s = D * J[i];
        for (int j = 0; j <= i; j++) c.H[i * nv + j] += s * J[j];
      }
    } else if (d->efc_state[r] == B2MJ_CSTATE_CONE) {
      const int con = d->efc_id[r], dim = d->contact_dim[con];
      const double* Hc = d->contact_H + 36 * con;
      for (int a = 0; a < dim; a++)
        for (int b = 0; b < dim; b++) {
          const double h = Hc[a * dim + b];
          if (h == 0) continue;
          const double *Ja = d->efc_J + (r + a) * nv, *Jb = d->efc_J + (r + b) * nv;
          for (int i = 0; i < nv; i++) {
            if (Ja[i] == 0) continue;
            const double s = h * Ja[i];
            for (int j = 0; j <= i; j++) c.H[i * nv + j] += s * Jb[j];
          }
        }
      r += dim - 1;
    }
  }
  // symmetrise lower -> full is not needed: cholFactor reads the lower triangle only
  cholFactor(c.H.data(), nv, MINVAL);
}

static void primalGradient(Primal& c) {
  for (int i = 0; i < c.nv; i++) c.grad[i] = c.Ma[i] - c.d->qfrc_smooth[i] - c.d->qfrc_constraint[i];
  if (c.newton) cholSolve(c.Mgrad.data(), c.H.data(), c.grad.data(), c.nv);
  else solveM(c.m, c.d, c.Mgrad.data(), c.grad.data());
}

static void primalPrepare(Primal& c) {
  const OrcData* d = c.d;
  mulM(c.m, d, c.Mv.data(), c.search.data());
  mulJacVec(c.m, d, c.Jv.data(), c.search.data());
  c.quadGauss[0] = c.gauss;
  c.quadGauss[1] = dot(c.search.data(), c.Ma.data(), c.nv) - dot(d->qfrc_smooth, c.search.data(), c.nv);
  c.quadGauss[2] = 0.5 * dot(c.search.data(), c.Mv.data(), c.nv);
  for (int i = 0; i < c.nefc; i++) {
    const double D = d->efc_D[i];
    c.quad[3 * i] = 0.5 * D * c.Jaref[i] * c.Jaref[i];
    c.quad[3 * i + 1] = D * c.Jaref[i] * c.Jv[i];
    c.quad[3 * i + 2] = 0.5 * D * c.Jv[i] * c.Jv[i];
  }
}

static void primalEval(const Primal& c, LSPoint* p) {
  const OrcData* d = c.d;
  const double alpha = p->alpha;
  double q[3] = {c.quadGauss[0], c.quadGauss[1], c.quadGauss[2]};
  double cost = 0, deriv0 = 0, deriv1 = 0;
  for (int i = 0; i < c.nefc; i++) {
    const double x = c.Jaref[i] + alpha * c.Jv[i];
    const double* qi = &c.quad[3 * i];
    switch (d->efc_type[i]) {
      case B2MJ_CNSTR_EQUALITY:
        q[0] += qi[0]; q[1] += qi[1]; q[2] += qi[2];
        break;
      case B2MJ_CNSTR_FRICTION_DOF:
      case B2MJ_CNSTR_FRICTION_TENDON: {
        const double f = d->efc_frictionloss[i], Rf = d->efc_R[i] * f;
        if (x <= -Rf) { q[0] += f * (-0.5 * Rf - c.Jaref[i]); q[1] += -f * c.Jv[i]; }
        else if (x >= Rf) { q[0] += f * (-0.5 * Rf + c.Jaref[i]); q[1] += f * c.Jv[i]; }
        else { q[0] += qi[0]; q[1] += qi[1]; q[2] += qi[2]; }
        break;
      }
      case B2MJ_CNSTR_CONTACT_ELLIPTIC: {
        const int con = d->efc_id[i], dim = d->contact_dim[con];
        const double mu = d->contact_mu[con];
        const double* fri = d->contact_friction + 5 * con;
        const double U0 = c.Jaref[i] * mu, V0 = c.Jv[i] * mu;
        double UU = 0, UV = 0, VV = 0;
        for (int j = 1; j < dim; j++) {
          const double U = c.Jaref[i + j] * fri[j - 1], V = c.Jv[i + j] * fri[j - 1];
          UU += U * U; UV += U * V; VV += V * V;
        }
        const double N = U0 + alpha * V0, Tsqr = UU + alpha * (2 * UV + alpha * VV);
        bool bottom = false;
        if (Tsqr <= 0) {
          if (N < 0) bottom = true;
        } else {
          const double T = std::sqrt(Tsqr);
          if (N >= mu * T) {
          } else if (mu * N + T <= 0) {
            bottom = true;
          } else {
            const double Dm = d->efc_D[i] / (mu * mu * (1 + mu * mu));
            const double N1 = V0, T1 = (UV + alpha * VV) / T, T2 = VV / T - (UV + alpha * VV) * T1 / (T * T);
            const double NmT = N - mu * T;
            cost += 0.5 * Dm * NmT * NmT;
            deriv0 += Dm * NmT * (N1 - mu * T1);
            deriv1 += Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NmT * (-mu * T2));
          }
        }
        if (bottom)
          for (int j = 0; j < dim; j++) { q[0] += qi[3 * j]; q[1] += qi[3 * j + 1]; q[2] += qi[3 * j + 2]; }
        i += dim - 1;
        break;
      }
      default:  // limits, frictionless and pyramidal contacts
        if (x < 0) { q[0] += qi[0]; q[1] += qi[1]; q[2] += qi[2]; }
    }
  }
  p->cost = cost + alpha * alpha * q[2] + alpha * q[1] + q[0];
  p->deriv[0] = deriv0 + 2 * alpha * q[2] + q[1];
  p->deriv[1] = deriv1 + 2 * q[2];
  if (p->deriv[1] < MINVAL) p->deriv[1] = MINVAL;
}

// exact line search on the convex piecewise-quadratic restriction; returns the step (0 = no progress)
static double primalSearch(Primal& c) {
  const b2mjModel* m = c.m;
  const double snorm = std::sqrt(dot(c.search.data(), c.search.data(), c.nv));
  if (snorm < MINVAL) return 0;
  const double gtol = m->opt.tolerance * m->opt.ls_tolerance * snorm / c.scale;
  primalPrepare(c);
  LSPoint p0, p1, p2, pn;
  p0.alpha = 0;
  primalEval(c, &p0);
  p1.alpha = p0.alpha - p0.deriv[0] / p0.deriv[1];
  primalEval(c, &p1);
  if (p0.cost < p1.cost) p1 = p0;
  if (std::fabs(p1.deriv[0]) < gtol) return p1.alpha;
  const double dir = p1.deriv[0] < 0 ? 1.0 : -1.0;
  int iter = 0;
  p2 = p1;
  // phase 1: Newton steps in one direction until the derivative changes sign
  while (p1.deriv[0] * dir <= -gtol && iter < m->opt.ls_iterations) {
    p2 = p1;
    pn.alpha = p1.alpha - p1.deriv[0] / p1.deriv[1];
    primalEval(c, &pn);
    p1 = pn;
    iter++;
    if (std::fabs(p1.deriv[0]) < gtol) return p1.alpha;
  }
  if (iter >= m->opt.ls_iterations || p1.deriv[0] * dir <= -gtol) return p1.cost < p0.cost ? p1.alpha : 0;
  // phase 2: root of the derivative bracketed by p2 (same sign as at start) and p1 (opposite sign)
  while (iter < m->opt.ls_iterations) {
    LSPoint cand[3];
    int nc = 0;
    const double lo = std::fmin(p1.alpha, p2.alpha), hi = std::fmax(p1.alpha, p2.alpha);
    double a1 = p1.alpha - p1.deriv[0] / p1.deriv[1], a2 = p2.alpha - p2.deriv[0] / p2.deriv[1];
    if (a1 > lo && a1 < hi) cand[nc++].alpha = a1;
    if (a2 > lo && a2 < hi) cand[nc++].alpha = a2;
    cand[nc++].alpha = 0.5 * (lo + hi);
    bool moved = false;
    for (int k = 0; k < nc; k++) {
      primalEval(c, &cand[k]);
      if (std::fabs(cand[k].deriv[0]) < gtol) return cand[k].alpha;
      // same side as p2 (derivative sign matches the initial direction) tightens p2, otherwise p1
      if (cand[k].deriv[0] * dir < 0) {
        if (std::fabs(cand[k].alpha - p1.alpha) < std::fabs(p2.alpha - p1.alpha)) { p2 = cand[k]; moved = true; }
      } else {
        if (std::fabs(cand[k].alpha - p2.alpha) < std::fabs(p1.alpha - p2.alpha)) { p1 = cand[k]; moved = true; }
      }
    }
    iter++;
    if (!moved || std::fabs(p1.alpha - p2.alpha) < MINVAL) break;
  }
  const LSPoint& best = p1.cost < p2.cost ? p1 : p2;
  return best.cost < p0.cost ? best.alpha : 0;
}

static void solPrimal(const b2mjModel* m, OrcData* d, int maxiter, bool newton) {
  Primal c;
  c.m = m; c.d = d;
  c.nv = m->nv; c.nefc = d->nefc();
  c.newton = newton;
  c.cone = m->opt.cone == B2MJ_CONE_ELLIPTIC;
  const int nv = c.nv, nefc = c.nefc;
  c.Jaref.resize(nefc); c.Jv.resize(nefc); c.Ma.resize(nv); c.Mv.resize(nv); c.grad.resize(nv);
  c.Mgrad.resize(nv); c.search.resize(nv); c.quad.resize(3 * nefc);
  if (newton) c.H.resize((size_t)nv * nv);
  c.scale = 1 / (m->stat.meaninertia * std::max(1, nv));
  std::vector<double> gradold(nv), Mgradold(nv);
  std::vector<int> oldstate(nefc);

  mulM(m, d, c.Ma.data(), d->qacc);
  mulJacVec(m, d, c.Jaref.data(), d->qacc);
  for (int i = 0; i < nefc; i++) c.Jaref[i] -= d->efc_aref[i];
  primalUpdateConstraint(c);
  if (newton) primalHessian(c);
  primalGradient(c);
  for (int i = 0; i < nv; i++) c.search[i] = -c.Mgrad[i];

  int iter = 0;
  while (iter < maxiter) {
    const double alpha = primalSearch(c);
    if (alpha == 0) break;
    for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * c.search[i]; c.Ma[i] += alpha * c.Mv[i]; }
    for (int i = 0; i < nefc; i++) c.Jaref[i] += alpha * c.Jv[i];
    const double oldcost = c.cost;
    gradold = c.grad;
    Mgradold = c.Mgrad;
    for (int i = 0; i < nefc; i++) oldstate[i] = d->efc_state[i];
    primalUpdateConstraint(c);
    if (newton) {
      bool changed = c.cone;
      for (int i = 0; i < nefc && !changed; i++) changed = oldstate[i] != d->efc_state[i];
      if (changed) primalHessian(c);
    }
    primalGradient(c);
    if (newton) {
      for (int i = 0; i < nv; i++) c.search[i] = -c.Mgrad[i];
    } else {
      double num = 0, den = 0;
      for (int i = 0; i < nv; i++) { num += c.grad[i] * (c.Mgrad[i] - Mgradold[i]); den += gradold[i] * Mgradold[i]; }
      double beta = num / std::fmax(MINVAL, den);
      if (beta < 0) beta = 0;
      for (int i = 0; i < nv; i++) c.search[i] = -c.Mgrad[i] + beta * c.search[i];
    }
    const double improvement = c.scale * (oldcost - c.cost);
    const double gradient = c.scale * std::sqrt(dot(c.grad.data(), c.grad.data(), nv));
    iter++;
    if (improvement < m->opt.tolerance || gradient < m->opt.tolerance) break;
  }
  d->solver_iter_[0] = iter;
}

// mj_fwdConstraint
void fwdConstraint(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv, nefc = d->nefc();
  d->solver_iter_[0] = 0;
  if (nefc == 0) {
    copy(d->qacc, d->qacc_smooth, nv);
    copy(d->qacc_warmstart, d->qacc_smooth, nv);
    zero(d->qfrc_constraint, nv);
    return;
  }
  // efc_b = J*qacc_smooth - aref
  mulJacVec(m, d, d->efc_b, d->qacc_smooth);
  for (int i = 0; i < nefc; i++) d->efc_b[i] -= d->efc_aref[i];
  const bool warm = !(m->opt.disableflags & B2MJ_DSBL_WARMSTART);
  std::vector<double> jar(nefc);
  if (m->opt.solver == B2MJ_SOL_PGS) {
    if (warm) {
      mulJacVec(m, d, jar.data(), d->qacc_warmstart);
      for (int i = 0; i < nefc; i++) jar[i] -= d->efc_aref[i];
      constraintUpdate(m, d, jar.data(), nullptr, 0);
      double cost = 0;
      for (int i = 0; i < nefc; i++)
        cost += d->efc_force[i] * (0.5 * dot(d->efc_AR + i * nefc, d->efc_force, nefc) + d->efc_b[i]);
      if (cost > 0) zero(d->efc_force, nefc);
    } else {
      zero(d->efc_force, nefc);
    }
    solPGS(m, d, m->opt.iterations);
  } else {
    if (warm) {
      // cost at the warm start vs at the unconstrained acceleration
      std::vector<double> Ma(nv);
      double cost_warm, cost_smooth;
      mulJacVec(m, d, jar.data(), d->qacc_warmstart);
      for (int i = 0; i < nefc; i++) jar[i] -= d->efc_aref[i];
      constraintUpdate(m, d, jar.data(), &cost_warm, 0);
      mulM(m, d, Ma.data(), d->qacc_warmstart);
      double g = 0;
      for (int i = 0; i < nv; i++) g += (Ma[i] - d->qfrc_smooth[i]) * (d->qacc_warmstart[i] - d->qacc_smooth[i]);
      cost_warm += 0.5 * g;
      constraintUpdate(m, d, d->efc_b, &cost_smooth, 0);
      if (cost_warm < cost_smooth) copy(d->qacc, d->qacc_warmstart, nv);
      else copy(d->qacc, d->qacc_smooth, nv);
    } else {
      copy(d->qacc, d->qacc_smooth, nv);
    }
    solPrimal(m, d, m->opt.iterations, m->opt.solver == B2MJ_SOL_NEWTON);
  }
  // the warm start of the next step is the main solver's result; the noslip pass runs after it is saved
  copy(d->qacc_warmstart, d->qacc, nv);
  if (m->opt.noslip_iterations > 0) solNoSlip(m, d, m->opt.noslip_iterations);
}

}  // namespace orc
