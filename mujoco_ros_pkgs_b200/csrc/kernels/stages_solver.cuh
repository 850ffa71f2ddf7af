// stages_solver.cuh — constraint solve (mj_fwdConstraint), one env per warp.
//
// Replaces mj_projectConstraint + mj_solPGS / mj_solNewton inside the reference's mj_step call
// (mujoco_env.cpp:498; rows M6/M9 of SURVEY 8a).
//
// PGS works on the explicit dual matrix AR = J inv(M) J' + R (MuJoCo's efc_AR), built wide-parallel from
// G = J W with W = inv(L) of the L'DL factor: AR = G diag(1/D) G' + R.  The sweep itself keeps the row
// residuals r = b + AR f distributed over the lanes' registers (lane j owns rows j, j+32): a row update
// is one shuffle to fetch r_i, a handful of scalar flops executed redundantly by every lane, and one
// FMA per lane to push the force change into the residuals -- no per-row reductions.  Same sweep order,
// projections and cost guard as mj_solPGS.  Rows beyond 64 fall back to the matrix-free form.
#pragma once
#include <math_constants.h>

#include "env_ctx.cuh"
#include "stages_constraint.cuh"
#include "stages_smooth.cuh"
#include "step_launch.h"
#include "team.cuh"

namespace b2k {

#define B2K_PGS_REGROWS 64
#define B2K_SPARSE_H_MIN_NV 48  // models wider than this build the Newton Hessian row-sparse

// Where AR (+ the 4 row constants per row that stage_projectConstraint appends after the matrix) lives:
//  1. the small shared-memory window XF_EFC_AR_S (nefc <= 17 for the C2 model: the common case);
//  2. else an OVERLAY on the kinematic frames [xpos .. crb] of the env's own arena: by the time the constraint solve
//     runs (AR is built at the head of stage_fwdConstraint, not in the position stage) nothing reads those fields any
//     more in a plain step (host-side rule in make_layout: no acc-stage sensor that needs frames; never when the arena
//     will be dumped).  Round 1 fetched AR from the L2 arena for 18..25 rows -- exactly the envs whose 100-iteration
//     Gauss-Seidel chains end every per-step launch -- at ~270 cycles per row update against ~140 from shared memory;
//  3. else the env's HBM / L2 arena.
__device__ __forceinline__ double* arPtr(const Env e, int nefc) {
  const int need = nefc * (nefc + 4);
  if (need <= c_dm.xsize[XF_EFC_AR_S]) return e.XG(XF_EFC_AR_S);
  if (!e.dump && need <= c_dm.ar_ovl_doubles) return e.sd() + c_dm.ar_ovl_off;
  return e.XG(XF_EFC_AR);
}
// shared-memory byte offset of AR if it is on chip (cases 1, 2), else 0xffffffff
__device__ __forceinline__ unsigned arSmemOffset(const Env e, int nefc) {
  const int need = nefc * (nefc + 4);
  if (need <= c_dm.xsize[XF_EFC_AR_S]) return e.sbd + 8u * (unsigned)c_dm.xoff_s[XF_EFC_AR_S];
  if (!e.dump && need <= c_dm.ar_ovl_doubles) return e.sbd + 8u * (unsigned)c_dm.ar_ovl_off;
  return 0xffffffffu;
}

// G = J W (rows of J pushed through inv(L)) and AR = G diag(1/D) G' + R
// want_ar = false (noslip pass after a primal solver): only G and the diagonal, which the matrix-free sweeps need
__device__ __noinline__ void stage_projectConstraint(const Env e, int nefc, bool want_ar = true) {
  if (!nefc) return;
  const DevModel& m = c_dm;
  const int nv = m.nv;
  const double* J = e.DG(B2MJ_F_EFC_J);
  const double* R = e.DG(B2MJ_F_EFC_R);
  const double* W = e.XG(XF_QW);
  const double* dinv = e.DG(B2MJ_F_QLDIAGINV);
  double* G = e.XG(XF_EFC_MINVJT);
  const bool dense = m.dense_small;
  if (dense) {
    // G = J inv(M) (dense), AR = G J' + R
    const double* Minv = e.X(XF_MINV);
    FORL(item, nefc * nv) {
      const int i = item / nv, k = item - i * nv;
      const double* Ji = J + i * nv;
      double s = 0;
      B2K_NOUNROLL for (int j = 0; j < nv; j++) s += Ji[j] * Minv[j * nv + k];
      G[item] = s;
    }
  } else {
    FORL(item, nefc * nv) {
      const int i = item / nv, k = item - i * nv;
      const double* Ji = J + i * nv;
      double s = Ji[k];
      B2K_NOUNROLL for (int p = m.dof_descadr[k]; p < m.dof_descadr[k + 1]; p++) s += W[m.dof_desc_adr[p]] * Ji[m.dof_desc_dof[p]];
      G[item] = s;
    }
  }
  WSYNC();
  double* ard = e.XG(XF_EFC_ARDIAG);
  if (want_ar && nefc <= B2K_PGS_REGROWS) {
    double* AR = arPtr(e, nefc);
    // lower triangle incl. diagonal, mirrored
    B2K_NOUNROLL for (int item = e.lane; item < nefc * nefc; item += B2K_G) {
      const int i = item / nefc, j = item - i * nefc;
      if (j > i) continue;
      double s = 0;
      if (dense) { B2K_NOUNROLL for (int k = 0; k < nv; k++) s += G[i * nv + k] * J[j * nv + k]; }
      else { B2K_NOUNROLL for (int k = 0; k < nv; k++) s += G[i * nv + k] * G[j * nv + k] * dinv[k]; }
      if (i == j) { s += R[i]; ard[i] = s; }
      AR[i * nefc + j] = s;
      AR[j * nefc + i] = s;
    }
    WSYNC();
    // row constants of the PGS sweep: {1/A_ii (negated on elliptic-cone rows), A_ii, lo, hi}
    const int* type = e.IG(B2MJ_F_EFC_TYPE);
    const double* floss = e.DG(B2MJ_F_EFC_FRICTIONLOSS);
    double* rowc = AR + nefc * nefc;
    FORL(i, nefc) {
      const int t = type[i];
      const double Aii = AR[i * nefc + i];
      double lo = 0, up = CUDART_INF;
      if (t == B2MJ_CNSTR_EQUALITY) lo = -CUDART_INF;
      else if (t == B2MJ_CNSTR_FRICTION_DOF || t == B2MJ_CNSTR_FRICTION_TENDON) { lo = -floss[i]; up = floss[i]; }
      rowc[4 * i] = t == B2MJ_CNSTR_CONTACT_ELLIPTIC ? -1.0 / Aii : 1.0 / Aii;
      rowc[4 * i + 1] = Aii;
      rowc[4 * i + 2] = lo;
      rowc[4 * i + 3] = up;
    }
  } else {
    FORL(i, nefc) {
      double s = 0;
      if (dense) { B2K_NOUNROLL for (int k = 0; k < nv; k++) s += G[i * nv + k] * J[i * nv + k]; }
      else { B2K_NOUNROLL for (int k = 0; k < nv; k++) s += G[i * nv + k] * G[i * nv + k] * dinv[k]; }
      ard[i] = s + R[i];
    }
  }
  WSYNC();
}

// dot of a constraint row with an nv-vector, result identical on all lanes
__device__ __forceinline__ double rowDot(const Env e, const double* row, const double* vec, int nv) {
  double s = 0;
  FORL(k, nv) s += row[k] * vec[k];
  return warpSum(e.mask, s);
}

// small SPD solve helpers for the elliptic QCQP (n <= 5), executed redundantly by every lane
__device__ __forceinline__ int cholFactorSmall(double* A, int n, double mindiag) {
  int bad = 0;
  B2K_NOUNROLL for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    B2K_NOUNROLL for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    if (s < mindiag) { s = mindiag; bad++; }
    const double ljj = sqrt(s);
    A[j * n + j] = ljj;
    const double inv = 1 / ljj;
    B2K_NOUNROLL for (int i = j + 1; i < n; i++) {
      double t = A[i * n + j];
      B2K_NOUNROLL for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t * inv;
    }
  }
  return bad;
}
__device__ __forceinline__ void cholSolveSmall(double* x, const double* L, const double* b, int n) {
  B2K_NOUNROLL for (int i = 0; i < n; i++) {
    double t = b[i];
    B2K_NOUNROLL for (int k = 0; k < i; k++) t -= L[i * n + k] * x[k];
    x[i] = t / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double t = x[i];
    B2K_NOUNROLL for (int k = i + 1; k < n; k++) t -= L[k * n + i] * x[k];
    x[i] = t / L[i * n + i];
  }
}
__device__ __noinline__ int QCQP(double* res, const double* Ain, const double* bin, const double* d, double r, int n) {
  double A[25], b[5], P[25], y[5], z[5], nb[5];
  B2K_NOUNROLL for (int i = 0; i < n; i++) {
    b[i] = bin[i] * d[i];
    y[i] = 0;
    B2K_NOUNROLL for (int j = 0; j < n; j++) A[i * n + j] = Ain[i * n + j] * d[i] * d[j];
  }
  double la = 0;
  const double r2 = r * r;
  for (int iter = 0; iter < 20; iter++) {
    B2K_NOUNROLL for (int i = 0; i < n * n; i++) P[i] = A[i];
    B2K_NOUNROLL for (int i = 0; i < n; i++) P[i * n + i] += la;
    if (cholFactorSmall(P, n, 1e-10)) { la = 0; B2K_NOUNROLL for (int i = 0; i < n; i++) y[i] = 0; break; }
    B2K_NOUNROLL for (int i = 0; i < n; i++) nb[i] = -b[i];
    cholSolveSmall(y, P, nb, n);
    double val = -r2;
    B2K_NOUNROLL for (int i = 0; i < n; i++) val += y[i] * y[i];
    if (val < 1e-10) break;
    cholSolveSmall(z, P, y, n);
    double deriv = 0;
    B2K_NOUNROLL for (int i = 0; i < n; i++) deriv += -2 * y[i] * z[i];
    const double delta = -val / deriv;
    if (delta < 1e-10) break;
    la += delta;
  }
  B2K_NOUNROLL for (int i = 0; i < n; i++) res[i] = y[i] * d[i];
  return la != 0;
}

// weighted row product sum_k a[k] b[k] w[k], identical on all lanes
__device__ __forceinline__ double rowDotW(const Env e, const double* a, const double* b, const double* w, int nv) {
  double s = 0;
  if (w) { FORL(k, nv) s += a[k] * b[k] * w[k]; }
  else { FORL(k, nv) s += a[k] * b[k]; }
  return warpSum(e.mask, s);
}

// elliptic-cone block update shared by both PGS forms: given the dim x dim block Athis of AR, the block
// residual res and the old forces, produce the new forces f (ray update, then QCQP on the friction dims)
__device__ __noinline__ void pgsConeBlock(int dim, const double* Athis, const double* res, const double* oldf, const double* fri, double* f) {
  B2K_NOUNROLL for (int j = 0; j < dim; j++) f[j] = oldf[j];
  if (f[0] < B2K_MINVAL) {
    f[0] -= res[0] / Athis[0];
    if (f[0] < 0) f[0] = 0;
    B2K_NOUNROLL for (int j = 1; j < dim; j++) f[j] = 0;
  } else {
    double v[6], v1[6];
    B2K_NOUNROLL for (int j = 0; j < dim; j++) v[j] = f[j];
    double denom = 0, num = 0;
    B2K_NOUNROLL for (int j = 0; j < dim; j++) {
      double s = 0;
      B2K_NOUNROLL for (int k = 0; k < dim; k++) s += Athis[j * dim + k] * v[k];
      v1[j] = s;
    }
    B2K_NOUNROLL for (int j = 0; j < dim; j++) { denom += v[j] * v1[j]; num += v[j] * res[j]; }
    if (denom >= B2K_MINVAL) {
      double x = -num / denom;
      if (f[0] + x * v[0] < 0) x = -v[0] / f[0];
      B2K_NOUNROLL for (int j = 0; j < dim; j++) f[j] += x * v[j];
    }
  }
  if (f[0] < B2K_MINVAL) {
    B2K_NOUNROLL for (int j = 1; j < dim; j++) f[j] = 0;
  } else {
    double Ac[25], bc[5], v[5];
    B2K_NOUNROLL for (int j = 0; j < dim - 1; j++) {
      B2K_NOUNROLL for (int k = 0; k < dim - 1; k++) Ac[j * (dim - 1) + k] = Athis[(j + 1) * dim + k + 1];
      double t = res[j + 1];
      B2K_NOUNROLL for (int k = 0; k < dim; k++) t -= Athis[(j + 1) * dim + k] * oldf[k];
      t += Athis[(j + 1) * dim] * f[0];
      bc[j] = t;
    }
    const int active = QCQP(v, Ac, bc, fri, f[0], dim - 1);
    if (active) {
      double s = 0;
      B2K_NOUNROLL for (int j = 0; j < dim - 1; j++) s += (v[j] / fri[j]) * (v[j] / fri[j]);
      s = sqrt(f[0] * f[0] / fmax(B2K_MINVAL, s));
      B2K_NOUNROLL for (int j = 0; j < dim - 1; j++) v[j] *= s;
    }
    B2K_NOUNROLL for (int j = 0; j < dim - 1; j++) f[1 + j] = v[j];
  }
}

// mj_solPGS on the explicit AR with register-resident residuals (nefc <= 32, or <= 64 with TWO).
// force holds the (already accepted) warm start on entry.  Per row: the row constants {1/A_ii, A_ii, lo,
// hi} come from one broadcast read, the residual and the old force from two shuffles, the projection is a
// clamp, and every lane folds the force change into its own residual with one FMA.  SM = AR and the row
// constants live in the shared-memory window (addresses formed from the shared symbol -> LDS), else in the
// env's HBM/L2 arena.  Returns iterations used.
template <int NS, bool SM>
__device__ __noinline__ int solvePGS_regT(const Env e, int nefc, const double* ARg, unsigned ARs) {
  const DevModel& m = c_dm;
  const double* AR = SM ? reinterpret_cast<const double*>(b2k_smem + ARs) : ARg;
  const double* rowc = AR + nefc * nefc;
  EfcPtrs P = efcPtrs(e);
  const double scale = 1 / (m.env_scalars[0] * max(1, m.nv));
  const double tol = m.opt.tolerance;
  const int maxiter = m.opt.iterations;
  const int lane = e.lane;
  // lane j owns rows j, j+G, ... (NS slots): residual r = b + AR f and force f live in registers
  double r[NS], f[NS];
  int col[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) {
    const int j = lane + B2K_G * s;
    col[s] = min(j, nefc - 1);  // clamped column index for AR row reads
    r[s] = 0; f[s] = 0;
    if (j < nefc) {
      f[s] = P.force[j];
      double acc = P.b[j];
      B2K_NOUNROLL for (int k = 0; k < nefc; k++) acc += AR[j * nefc + k] * P.force[k];
      r[s] = acc;
    }
  }
  int iter = 0;
  while (iter < maxiter) {
    double improvement = 0;
    const double* rc = rowc;
    const double* arow = AR;
    for (int i = 0; i < nefc;) {
      const double iA = rc[0];
      if (!(iA < 0)) {  // scalar row (1/A_ii is stored negated for the rows of an elliptic cone; NaN stays here)
        const double Aii = rc[1], lo = rc[2], up = rc[3];
        double a[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) a[s] = arow[col[s]];
        const int src = i % B2K_G, slot = i / B2K_G;
        double res = 0, fold = 0;
#pragma unroll
        for (int s = 0; s < NS; s++)
          if (NS == 1 || slot == s) {
            res = __shfl_sync(e.mask, r[s], src, B2K_G);
            fold = __shfl_sync(e.mask, f[s], src, B2K_G);
          }
        double fn = fold - res * iA;
        fn = fn < lo ? lo : fn;
        fn = fn > up ? up : fn;
        double delta = fn - fold;
        double change = delta * (0.5 * delta * Aii + res);
        if (change > 1e-10) { delta = 0; change = 0; fn = fold; }  // cost guard of mj_solPGS (uniform branch)
        improvement -= change;
#pragma unroll
        for (int s = 0; s < NS; s++) {
          r[s] += a[s] * delta;
          if ((NS == 1 || slot == s) && lane == src) f[s] = fn;
        }
        i += 1;
        rc += 4;
        arow += nefc;
      } else {
        const int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
        const double* c_fri = e.DG(B2MJ_F_CONTACT_FRICTION);
        const int c = P.id[i], dim = min(max(c_dim[c], 1), 6);
        const double* fri = c_fri + 5 * c;
        double Athis[36], res[6], oldf[6], fb[6];
        for (int j = 0; j < dim; j++) {
          const int row = i + j, rs = row % B2K_G, slot = row / B2K_G;
          for (int k = 0; k < dim; k++) Athis[j * dim + k] = AR[row * nefc + i + k];
          double rr = 0, ff = 0;
#pragma unroll
          for (int s = 0; s < NS; s++)
            if (NS == 1 || slot == s) {
              ff = __shfl_sync(e.mask, f[s], rs, B2K_G);
              rr = __shfl_sync(e.mask, r[s], rs, B2K_G);
            }
          oldf[j] = ff;
          res[j] = rr;
        }
        pgsConeBlock(dim, Athis, res, oldf, fri, fb);
        double change = 0, delta[6];
        for (int j = 0; j < dim; j++) delta[j] = fb[j] - oldf[j];
        for (int j = 0; j < dim; j++) {
          double acc = 0;
          for (int k = 0; k < dim; k++) acc += Athis[j * dim + k] * delta[k];
          change += 0.5 * delta[j] * acc + delta[j] * res[j];
        }
        if (change > 1e-10) {
          change = 0;
          for (int j = 0; j < dim; j++) { delta[j] = 0; fb[j] = oldf[j]; }
        }
        improvement -= change;
        for (int j = 0; j < dim; j++) {
          const int row = i + j, rs = row % B2K_G, slot = row / B2K_G;
#pragma unroll
          for (int s = 0; s < NS; s++) {
            r[s] += AR[row * nefc + col[s]] * delta[j];
            if ((NS == 1 || slot == s) && lane == rs) f[s] = fb[j];
          }
        }
        i += dim;
        rc += 4 * dim;
        arow += nefc * dim;
      }
    }
    iter++;
    if (improvement * scale < tol) break;
  }
#pragma unroll
  for (int s = 0; s < NS; s++) {
    const int j = lane + B2K_G * s;
    if (j < nefc) P.force[j] = f[s];
  }
  WSYNC();
  return iter;
}

// mj_solPGS for scalar rows only (no elliptic-cone blocks), nefc <= lanes per env: OWNER COMPUTES.
// Lane i owns row i: its residual, force and row constants {1/A_ii, A_ii, lo, hi} stay in registers.  A row update
// is: the owner clamps its own force (no operand shuffles, no row-constant loads), one shuffle broadcasts the
// force change, every lane folds it into its residual with one FMA.  The only memory operand, AR[i][lane], does
// not depend on the chain: it is fetched two rows ahead into a register ring.  The row loop stays rolled: a fully
// unrolled sweep measured 30% SLOWER in the desynchronised rollout (instruction fetch).
//
// Round 2: the Gauss-Seidel sweep of the heaviest env ends every per-step launch (21-25 rows x 70-100 iterations =
// 2/3 of the launch).  Measured in situ, a row costs (instructions per row) x ~5.2 cycles -- the warp shares its
// scheduler and the instruction caches with warps running other stages, so the row is ISSUE bound, not bound by its
// dependent chain (a variant that shortened the chain by speculating the cost guard but added six instructions per
// row ran 13 % slower: 261 vs 231 cycles per row).  Hence the row is written for instruction count (44 -> 24):
//  * the guard value is not broadcast: the owner zeroes its own force change before the one shuffle, and each lane
//    accumulates the improvement of its own row; one butterfly sum per SWEEP replaces two shuffles + an add per ROW
//    (the sum is taken in a different order than mj_solPGS takes it: it only feeds the convergence test);
//  * the next AR operand is fetched through a running pointer (the two rows read past the matrix fall into the row
//    constants stored behind it), no index clamp / multiply per row;
//  * 0.5 * A_ii is folded into one constant (exact), so the guard is one FMA and one multiply;
//  * when every row is an equality (no clamp) or a unilateral row (lo = 0, hi = inf) the clamp is a sign test on the
//    high word plus a select (SIMPLE = true; friction-loss rows take the generic two-sided form).
// Forces and residuals are bitwise those of the round-1 form.  Returns iterations used.
template <bool SIMPLE>
__device__ __forceinline__ int solvePGS_own_T(const Env e, int nefc, const double* AR, double iA, double Aii, double lo,
                                              double up, double f, double r, double* f_out) {
  const DevModel& m = c_dm;
  const double scale = 1 / (m.env_scalars[0] * max(1, m.nv));
  const double tol = m.opt.tolerance;
  const int maxiter = m.opt.iterations;
  const int lane = e.lane;
  const int me = lane < nefc ? lane : nefc - 1;
  const bool pos = !(lo < 0);  // unilateral row (lo == 0); equality rows have lo = -inf
  const double hA = 0.5 * Aii;
  // register ring: slot k holds AR[row][me] of the next row with row % 2 == k; refilled right after its row consumed
  // it, two rows ahead.  Rows 0,1 open every sweep and stay in q0,q1.
  const double* col = AR + me;
  const double q0 = col[0], q1 = col[nefc];
  int iter = 0;
#define B2K_PGS_OWN_ROW(SLOT)                                                  \
  {                                                                            \
    const double ai = SLOT;                                                    \
    const double x = f - r * iA;                                               \
    double fn;                                                                 \
    if (SIMPLE) {                                                              \
      fn = (pos && __double2hiint(x) < 0) ? 0.0 : x;                           \
    } else {                                                                   \
      fn = x < lo ? lo : x;                                                    \
      fn = fn > up ? up : fn;                                                  \
    }                                                                          \
    double delta = fn - f;                                                     \
    const double change = delta * (hA * delta + r);                            \
    const bool ok = !(change > 1e-10); /* cost guard of mj_solPGS */           \
    delta = ok ? delta : 0.0;                                                  \
    const double d = __shfl_sync(e.mask, delta, i, B2K_G);                     \
    r += ai * d;                                                               \
    if (ok && lane == i) { f = fn; improvement -= change; }                    \
    SLOT = *pcol;                                                              \
    pcol += nefc;                                                              \
    i++;                                                                       \
  }
  // The same row with the cost guard taken OFF the serial chain: the candidate change is broadcast at once, the guard
  // (change > 1e-10: only ever true through round-off, the clamped 1-D minimiser cannot raise the cost) is evaluated
  // beside the shuffle and only recorded.  A sweep in which some guard fired is thrown away and redone from the saved
  // (f, r) with the guarded row above, so the result is bitwise the guarded algorithm's; the serial chain per row drops
  // from FMA, clamp, ADD, FMA, MUL, SETP, SEL, SHFL, FMA to FMA, clamp, ADD, SHFL, FMA.
#define B2K_PGS_OWN_ROW_FAST(SLOT)                                             \
  {                                                                            \
    const double ai = SLOT;                                                    \
    const double x = f - r * iA;                                               \
    double fn;                                                                 \
    if (SIMPLE) {                                                              \
      fn = (pos && __double2hiint(x) < 0) ? 0.0 : x;                           \
    } else {                                                                   \
      fn = x < lo ? lo : x;                                                    \
      fn = fn > up ? up : fn;                                                  \
    }                                                                          \
    const double delta = fn - f;                                               \
    const double d = __shfl_sync(e.mask, delta, i, B2K_G);                     \
    const double change = delta * (hA * delta + r);                            \
    r += ai * d;                                                               \
    if (lane == i) { f = fn; improvement -= change; worst = max(worst, __double2hiint(change)); } \
    SLOT = *pcol;                                                              \
    pcol += nefc;                                                              \
    i++;                                                                       \
  }
  while (iter < maxiter) {
    double improvement = 0;  // this lane's own row only; summed over the warp once per sweep
    const double f_save = f, r_save = r;
    // largest high word of any row's cost change in this sweep: positive doubles order like their high words, so
    // worst >= hiword(1e-10) flags every change > 1e-10 (and, conservatively, a few just below: the redo is exact)
    int worst = (int)0x80000000;
    {
      double p0 = q0, p1 = q1;
      const double* pcol = col + 2 * nefc;
      int i = 0;
      B2K_NOUNROLL while (i + 2 <= nefc) {
        B2K_PGS_OWN_ROW_FAST(p0)
        B2K_PGS_OWN_ROW_FAST(p1)
      }
      if (i < nefc) B2K_PGS_OWN_ROW_FAST(p0)
    }
    if (__any_sync(e.mask, worst >= __double2hiint(1e-10))) {  // redo the sweep with the guard on the chain
      f = f_save; r = r_save; improvement = 0;
      double p0 = q0, p1 = q1;
      const double* pcol = col + 2 * nefc;
      int i = 0;
      B2K_NOUNROLL while (i + 2 <= nefc) {
        B2K_PGS_OWN_ROW(p0)
        B2K_PGS_OWN_ROW(p1)
      }
      if (i < nefc) B2K_PGS_OWN_ROW(p0)
    }
    iter++;
    if (warpSum(e.mask, improvement) * scale < tol) break;
  }
#undef B2K_PGS_OWN_ROW_FAST
#undef B2K_PGS_OWN_ROW
  *f_out = f;
  return iter;
}

__device__ __noinline__ int solvePGS_own(const Env e, int nefc, const double* AR) {
  const double* rowc = AR + nefc * nefc;
  EfcPtrs P = efcPtrs(e);
  const int lane = e.lane;
  const bool own = lane < nefc;
  const int me = own ? lane : nefc - 1;
  const double iA = rowc[4 * me];
  if (__any_sync(e.mask, iA < 0)) return -1;
  const double Aii = rowc[4 * me + 1], lo = rowc[4 * me + 2], up = rowc[4 * me + 3];
  double f = 0, r = 0;
  if (own) {
    f = P.force[lane];
    double acc = P.b[lane];
    B2K_NOUNROLL for (int k = 0; k < nefc; k++) acc += AR[lane * nefc + k] * P.force[k];
    r = acc;
  }
  // simple rows: (-inf, +inf) or (0, +inf)
  const bool simple_row = (up == CUDART_INF) && (lo == 0 || lo == -CUDART_INF);
  int iters;
  if (__all_sync(e.mask, simple_row)) iters = solvePGS_own_T<true>(e, nefc, AR, iA, Aii, lo, up, f, r, &f);
  else iters = solvePGS_own_T<false>(e, nefc, AR, iA, Aii, lo, up, f, r, &f);
  if (own) P.force[lane] = f;
  WSYNC();
  return iters;
}

// mj_solPGS, matrix-free form for large nefc: rows of G = J inv(L) and the running vector
// a = diag(1/D) G' f, so (AR f)_i = G_i . a + R_i f_i.  Returns iterations used.
__device__ __noinline__ int solvePGS_free(const Env e, int nefc, double* avec) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  EfcPtrs P = efcPtrs(e);
  const double* G = e.XG(XF_EFC_MINVJT);
  const double* ard = e.XG(XF_EFC_ARDIAG);
  // sparse models: AR = G diag(1/D) G' (G = J inv(L));  dense small models: AR = G J' (G = J inv(M))
  const bool dense = m.dense_small;
  const double* U = dense ? P.J : G;
  const double* dinv = dense ? nullptr : e.DG(B2MJ_F_QLDIAGINV);
  const int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
  const double* c_fri = e.DG(B2MJ_F_CONTACT_FRICTION);
  const double scale = 1 / (m.env_scalars[0] * max(1, nv));
  FORL(k, nv) {
    double s = 0;
    B2K_NOUNROLL for (int i = 0; i < nefc; i++) s += U[i * nv + k] * P.force[i];
    avec[k] = dinv ? s * dinv[k] : s;
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
        const double res = P.b[i] + rowDot(e, G + i * nv, avec, nv) + P.R[i] * fold;
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
          FORL(k, nv) avec[k] += delta * U[i * nv + k] * (dinv ? dinv[k] : 1.0);
          if (e.lane == 0) P.force[i] = f;
          WSYNC();
        }
        i += 1;
      } else {
        const int c = P.id[i], dim = c_dim[c];
        const double* fri = c_fri + 5 * c;
        double Athis[36], res[6], oldf[6], f[6];
        B2K_NOUNROLL for (int j = 0; j < dim; j++) {
          B2K_NOUNROLL for (int k = 0; k < dim; k++) {
            double v = rowDotW(e, G + (i + j) * nv, U + (i + k) * nv, dinv, nv);
            if (j == k) v += P.R[i + j];
            Athis[j * dim + k] = v;
          }
          oldf[j] = P.force[i + j];
          res[j] = P.b[i + j] + rowDot(e, G + (i + j) * nv, avec, nv) + P.R[i + j] * oldf[j];
        }
        pgsConeBlock(dim, Athis, res, oldf, fri, f);
        double change = 0, delta[6];
        B2K_NOUNROLL for (int j = 0; j < dim; j++) delta[j] = f[j] - oldf[j];
        B2K_NOUNROLL for (int j = 0; j < dim; j++) {
          double s = 0;
          B2K_NOUNROLL for (int k = 0; k < dim; k++) s += Athis[j * dim + k] * delta[k];
          change += 0.5 * delta[j] * s + delta[j] * res[j];
        }
        if (change > 1e-10) {
          change = 0;
          B2K_NOUNROLL for (int j = 0; j < dim; j++) { delta[j] = 0; f[j] = oldf[j]; }
        }
        improvement -= change;
        B2K_NOUNROLL for (int j = 0; j < dim; j++) {
          if (delta[j] != 0) FORL(k, nv) avec[k] += delta[j] * U[(i + j) * nv + k] * (dinv ? dinv[k] : 1.0);
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

// mj_solNoSlip: modified PGS over the friction-loss rows and the friction dimensions of contacts on the dual problem
// WITHOUT the regulariser R; normal forces stay as the main solver left them (opt.noslip_iterations > 0; the reference
// exposes "Noslip Iter" / "Noslip Tol" at mujoco_ros/src/viewer.cpp:590-591).  Matrix-free like solvePGS_free: the
// unregularised A = J inv(M) J' is never formed, blocks and residuals come from row products with G = J inv(L) and the
// running vector avec = inv(M) J' f.  Every lane computes the same scalars; lane 0 (or lanes < dim) commits the forces.
__device__ __noinline__ int solveNoSlip(const Env e, int nefc, double* avec) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  EfcPtrs P = efcPtrs(e);
  const double* G = e.XG(XF_EFC_MINVJT);
  const bool dense = m.dense_small;
  const double* U = dense ? P.J : G;
  const double* dinv = dense ? nullptr : e.DG(B2MJ_F_QLDIAGINV);
  const int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
  const double* c_fri = e.DG(B2MJ_F_CONTACT_FRICTION);
  const double scale = 1 / (m.env_scalars[0] * max(1, nv));
  FORL(k, nv) {
    double s = 0;
    B2K_NOUNROLL for (int i = 0; i < nefc; i++) s += U[i * nv + k] * P.force[i];
    avec[k] = dinv ? s * dinv[k] : s;
  }
  WSYNC();
  int iter = 0;
  while (iter < m.opt.noslip_iterations) {
    double improvement = 0;
    if (iter == 0) {
      double s = 0;
      FORL(i, nefc) s += 0.5 * P.force[i] * P.force[i] * P.R[i];
      improvement = warpSum(e.mask, s);
    }
    B2K_NOUNROLL for (int i = 0; i < nefc; i++) {
      const int type = P.type[i];
      int base, bd;  // block of rows [base, base + bd) updated together
      if (type == B2MJ_CNSTR_FRICTION_DOF || type == B2MJ_CNSTR_FRICTION_TENDON) { base = i; bd = 1; }
      else if (type == B2MJ_CNSTR_CONTACT_PYRAMIDAL) { base = i; bd = 2; }
      else if (type == B2MJ_CNSTR_CONTACT_ELLIPTIC) { base = i + 1; bd = c_dim[P.id[i]] - 1; }
      else continue;
      const int con = P.id[i];
      const int npair = type == B2MJ_CNSTR_CONTACT_PYRAMIDAL ? c_dim[con] - 1 : 1;
      B2K_NOUNROLL for (int pr = 0; pr < npair; pr++, base += (type == B2MJ_CNSTR_CONTACT_PYRAMIDAL ? 2 : 0)) {
        if (bd <= 0) break;
        double Ac[25], res[5], oldf[5], f[5];
        B2K_NOUNROLL for (int j = 0; j < bd; j++) {
          B2K_NOUNROLL for (int k = 0; k < bd; k++) Ac[j * bd + k] = rowDotW(e, G + (base + j) * nv, U + (base + k) * nv, dinv, nv);
          oldf[j] = P.force[base + j];
          res[j] = P.b[base + j] + rowDot(e, G + (base + j) * nv, avec, nv);
        }
        if (bd == 1 && type != B2MJ_CNSTR_CONTACT_ELLIPTIC) {
          const double fl = P.floss[base];
          f[0] = clampd(oldf[0] - res[0] / Ac[0], -fl, fl);
        } else if (type == B2MJ_CNSTR_CONTACT_PYRAMIDAL) {
          const double bc0 = res[0] - Ac[0] * oldf[0] - Ac[1] * oldf[1];
          const double bc1 = res[1] - Ac[2] * oldf[0] - Ac[3] * oldf[1];
          const double mid = 0.5 * (oldf[0] + oldf[1]);
          const double K1 = Ac[0] + Ac[3] - Ac[1] - Ac[2], K0 = mid * (Ac[0] - Ac[3]) + bc0 - bc1;
          if (K1 < B2K_MINVAL) { f[0] = mid; f[1] = mid; }
          else {
            const double y = -K0 / K1;
            if (y < -mid) { f[0] = 0; f[1] = 2 * mid; }
            else if (y > mid) { f[0] = 2 * mid; f[1] = 0; }
            else { f[0] = mid + y; f[1] = mid - y; }
          }
        } else {
          const double fn = P.force[i];
          const double* mu = c_fri + 5 * con;
          double bc[5];
          B2K_NOUNROLL for (int j = 0; j < bd; j++) {
            double t = res[j];
            B2K_NOUNROLL for (int k = 0; k < bd; k++) t -= Ac[j * bd + k] * oldf[k];
            bc[j] = t;
          }
          if (fn < B2K_MINVAL) {
            B2K_NOUNROLL for (int j = 0; j < bd; j++) f[j] = 0;
          } else {
            const int active = QCQP(f, Ac, bc, mu, fn, bd);
            if (active) {
              double t = 0;
              B2K_NOUNROLL for (int j = 0; j < bd; j++) t += (f[j] / mu[j]) * (f[j] / mu[j]);
              t = sqrt(fn * fn / fmax(B2K_MINVAL, t));
              B2K_NOUNROLL for (int j = 0; j < bd; j++) f[j] *= t;
            }
          }
        }
        double change = 0, delta[5];
        B2K_NOUNROLL for (int j = 0; j < bd; j++) delta[j] = f[j] - oldf[j];
        B2K_NOUNROLL for (int j = 0; j < bd; j++) {
          double t = 0;
          B2K_NOUNROLL for (int k = 0; k < bd; k++) t += Ac[j * bd + k] * delta[k];
          change += 0.5 * delta[j] * t + delta[j] * res[j];
        }
        if (change > 1e-10) {
          change = 0;
          B2K_NOUNROLL for (int j = 0; j < bd; j++) { delta[j] = 0; f[j] = oldf[j]; }
        }
        improvement -= change;
        B2K_NOUNROLL for (int j = 0; j < bd; j++) {
          if (delta[j] != 0) FORL(k, nv) avec[k] += delta[j] * U[(base + j) * nv + k] * (dinv ? dinv[k] : 1.0);
        }
        if (e.lane < bd) P.force[base + e.lane] = f[e.lane];
        WSYNC();
      }
      if (type == B2MJ_CNSTR_CONTACT_PYRAMIDAL) i += 2 * (c_dim[con] - 1) - 1;
      else if (type == B2MJ_CNSTR_CONTACT_ELLIPTIC) i += c_dim[con] - 1;
    }
    improvement *= scale;
    iter++;
    if (improvement < m.opt.noslip_tolerance) break;
  }
  return iter;
}

// ------------------------------------------------------------------------------------------------
// primal solvers (mj_solNewton / mj_solCG), warp-cooperative
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void mulJacVec_warp(const Env e, int nefc, double* res, const double* vec) {
  if (c_dm.team_warps > 1) { mulJacVec_sparse(e, nefc, res, vec); return; }  // column lists of this pass (TEAM_JCOLS)
  const int nv = c_dm.nv;
  const double* J = solveJ(e);
  FORL(i, nefc) {
    double s = 0;
    B2K_UNROLL4 for (int k = 0; k < nv; k++) s += J[i * nv + k] * vec[k];
    res[i] = s;
  }
  WSYNC();
}

__device__ __forceinline__ double dot_warp(const Env e, const double* a, const double* b, int n) {
  double s = 0;
  FORL(k, n) s += a[k] * b[k];
  return warpSum(e.mask, s);
}

// in-place dense Cholesky (lower, left-looking) of the nv x nv Hessian; invd[j] = 1 / L[j][j].
// Per column: the diagonal's dot product is split over the lanes (one butterfly sum instead of a j-long dependent
// chain executed redundantly by every lane), and each lane's row update runs four independent accumulators so the
// FMA chain does not wait on itself (the factorisation was 98% of a C5 step: nv 120, refactored every iteration).
// SM: the Hessian is in the shared arena -- its address is then formed from the shared symbol (LDS / STS, ~29 cycles);
// through the generic pointer of a possibly demoted field every access is a generic load (~100 cycles in situ), and
// the dependent chains of the factorisation / substitution multiply that.
template <bool SM>
__device__ __forceinline__ double* newtonH(const Env e) { return SM ? e.X(XF_NEWTON_H) : e.XG(XF_NEWTON_H); }

// Pair table of the trailing blocks (env_ctx.cuh::triTable): one table serves every pivot, because a row-major lower
// triangle of order m is the prefix p < m (m + 1) / 2 of the enumeration.
__device__ __forceinline__ const unsigned short* cholPairTable(const Env e, int n) {
  (void)n;
  return triTable(e);
}

// Right-looking Cholesky.  Per pivot the column is scaled, then the trailing block A(i, k) -= L(i, j) L(k, j),
// j < k <= i, is updated with ALL lanes busy: entry p of the block goes to lane p mod 32 through the pair table
// (row-major, so a warp's entries sit in one or two rows: consecutive addresses, L(i, j) a broadcast, L(k, j) a walk
// down column j with the odd leading dimension) -- n^3 / 192 warp passes instead of the n^2 / 2 of one row per lane.
// The only serial chain per pivot is pivot -> sqrt -> reciprocal -> column.  Every entry sees the same subtractions
// in the same order as in the row-per-lane form (tri == nullptr, kept for matrices whose table does not fit).
template <bool SM>
__device__ __noinline__ void cholFactor_warp(const Env e, int n, int ld, double mindiag, const unsigned short* tri) {
  double* A = newtonH<SM>(e);
  double* invd = e.X(XF_PRIMAL) + 7 * n;
  B2K_NOUNROLL for (int j = 0; j < n; j++) {
    double s = A[j * ld + j];
    if (s < mindiag) s = mindiag;
    const double ljj = sqrt(s), inv = 1 / ljj;
    B2K_NOUNROLL for (int i = j + 1 + e.lane; i < n; i += B2K_G) A[i * ld + j] *= inv;
    WSYNC();
    if (e.lane == 0) { A[j * ld + j] = ljj; invd[j] = inv; }
    if (tri) {
      const int mrem = n - j - 1, cnt = mrem * (mrem + 1) / 2;
      double* base = A + (j + 1) * ld;  // row j + 1
      B2K_NOUNROLL for (int p = e.lane; p < cnt; p += 2 * B2K_G) {
        const unsigned t0 = tri[p];
        double* r0 = base + (t0 & 255u) * ld;
        const double v0 = r0[j + 1 + (t0 >> 8)] - r0[j] * base[(t0 >> 8) * ld + j];
        if (p + B2K_G < cnt) {
          const unsigned t1 = tri[p + B2K_G];
          double* r1 = base + (t1 & 255u) * ld;
          r1[j + 1 + (t1 >> 8)] -= r1[j] * base[(t1 >> 8) * ld + j];
        }
        r0[j + 1 + (t0 >> 8)] = v0;
      }
    } else {
      const double* Lj = A + j;  // column j: L(k, j) = Lj[k * ld]
      B2K_NOUNROLL for (int i = j + 1 + e.lane; i < n; i += B2K_G) {
        double* Ai = A + i * ld;
        const double lij = Ai[j];
        int k = j + 1;
        B2K_NOUNROLL for (; k + 2 <= i + 1; k += 2) {
          const double a0 = Ai[k] - lij * Lj[k * ld];
          const double a1 = Ai[k + 1] - lij * Lj[(k + 1) * ld];
          Ai[k] = a0;
          Ai[k + 1] = a1;
        }
        if (k <= i) Ai[k] -= lij * Lj[k * ld];
      }
    }
    WSYNC();
  }
}

// x = inv(L L') b with the triangular sweeps held in registers: lane k owns entries k, k+32, k+64, k+96 (n <= 128).
// Written block-wise -- the outer loop over the 32-entry blocks is unrolled so that every register slot is named
// statically; the round-1 form indexed the slots through run-time predicates, which put the vector in LOCAL memory
// with eight branches per step (800 cycles per substitution step, 194 k cycles per solve at nv = 120, ~38 k at nv = 24).
#define B2K_CHOL_SLOTS_MAX (B2K_NEWTON_MAX_NV / 32)
// B2K_CHOL_SLOTS = register slots per lane: 1 serves n <= 32 without the predicated-off updates of the wider form
// PK: the factor is the packed lower triangle of team mode (team.cuh), row r at r (r + 1) / 2
template <bool PK>
__device__ __forceinline__ int hix(int r, int c, int ld) { return PK ? ((r * (r + 1)) >> 1) + c : r * ld + c; }
template <bool SM, int B2K_CHOL_SLOTS, bool PK = false>
__device__ __noinline__ void cholSolve_warpT(const Env e, int n, int ld) {
  // Mgrad = inv(H) grad on the solver's work vectors (XF_PRIMAL: Ma, Mv, grad, Mgrad, search, gradold, Mgradold, invdiag)
  const double* L = newtonH<SM>(e);
  double* wv = e.X(XF_PRIMAL);
  const double* b = wv + 2 * n;
  double* x = wv + 3 * n;
  const double* invd = wv + 7 * n;
  double t[B2K_CHOL_SLOTS];
  const int lane = e.lane;
#pragma unroll
  for (int s = 0; s < B2K_CHOL_SLOTS; s++) { const int k = lane + 32 * s; t[s] = k < n ? b[k] : 0.0; }
  // L y = b
#pragma unroll
  for (int bi = 0; bi < B2K_CHOL_SLOTS; bi++) {
    const int i0 = 32 * bi;
    if (i0 < n) {
      const int cnt = min(32, n - i0);
      B2K_NOUNROLL for (int ii = 0; ii < cnt; ii++) {
        const int i = i0 + ii;
        const double yi = __shfl_sync(e.mask, t[bi], ii) * invd[i];
        if (lane == ii) t[bi] = yi;
        else if (lane > ii && lane + i0 < n) t[bi] -= L[hix<PK>(lane + i0, i, ld)] * yi;
#pragma unroll
        for (int s = bi + 1; s < B2K_CHOL_SLOTS; s++) {
          const int k = lane + 32 * s;
          if (k < n) t[s] -= L[hix<PK>(k, i, ld)] * yi;
        }
      }
    }
  }
  // L' x = y
#pragma unroll
  for (int bi = B2K_CHOL_SLOTS - 1; bi >= 0; bi--) {
    const int i0 = 32 * bi;
    if (i0 < n) {
      const int cnt = min(32, n - i0);
      B2K_NOUNROLL for (int ii = cnt - 1; ii >= 0; ii--) {
        const int i = i0 + ii;
        const double xi = __shfl_sync(e.mask, t[bi], ii) * invd[i];
        const double* Li = L + hix<PK>(i, 0, ld);
        if (lane == ii) t[bi] = xi;
        else if (lane < ii) t[bi] -= Li[lane + i0] * xi;
#pragma unroll
        for (int s = 0; s < bi; s++) t[s] -= Li[lane + 32 * s] * xi;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < B2K_CHOL_SLOTS; s++) { const int k = lane + 32 * s; if (k < n) x[k] = t[s]; }
  WSYNC();
}

// developer build (-DB2K_SOLVE_PROF): cycle split of the primal solver, read with b2k_sprof_read (tools/solve_profile.py)
#ifdef B2K_SOLVE_PROF  /* g_sprof lives in team.cuh */
#define SPROF_DECL long long _sp_t = clock64();
#define SPROF(id) { const long long _n = clock64(); if (e.lane == 0) atomicAdd(&g_sprof[id], (unsigned long long)(_n - _sp_t)); _sp_t = _n; }
#else
#define SPROF_DECL
#define SPROF(id)
#endif

struct PrimalCtx {
  int nv, nefc, ncon;
  bool newton, cone;
  double *Jaref, *Jv, *quad, *Ma, *Mv, *grad, *Mgrad, *search, *gradold, *Mgradold, *invd, *H;
  double quadGauss[3];
  double cost, gauss, scale;
  const unsigned short* tri;  // pair table of the Cholesky (cholPairTable), or null
};

struct LSPoint {
  double alpha, cost, d0, d1;
};

// constraint update at the current qacc + Gauss term (primalUpdateConstraint)
__device__ __noinline__ void primalUpdate(const Env e, PrimalCtx& c, int* changed) {
  const double* qacc = e.D(B2MJ_F_QACC);
  const double* qas = e.D(B2MJ_F_QACC_SMOOTH);
  const double* qs = e.D(B2MJ_F_QFRC_SMOOTH);
  c.cost = constraintUpdate_warp(e, c.nefc, c.ncon, c.Jaref, c.newton && c.cone, changed);
  EfcPtrs P = efcPtrs(e);
  if (c_dm.team_warps > 1) mulJacTVec_sparse(e, c.nefc, e.D(B2MJ_F_QFRC_CONSTRAINT), P.force);
  else mulJacTVec_warp(e, c.nefc, e.D(B2MJ_F_QFRC_CONSTRAINT), P.force);
  double g = 0;
  FORL(i, c.nv) g += (c.Ma[i] - qs[i]) * (qacc[i] - qas[i]);
  c.gauss = 0.5 * warpSum(e.mask, g);
  c.cost += c.gauss;
}

// J' D J for models of up to 32 dofs: every lane keeps its share of the lower triangle (NI entries, lane-strided) in
// registers and the constraint rows are visited ONCE, in order, by the whole warp.  Per row the lane issues the loads of
// all its entries back to back, so the L2 / shared-memory latency of efc_J, efc_D and efc_state is paid once per row
// instead of once per (entry, row) as in the entry-outer loop this replaces (profiles/r2b_ncu_c4_lines.txt: that loop was
// 17.7 % of the C4 step, almost all of it long-scoreboard stalls).  Each entry still accumulates its rows in increasing
// order with the same expression, so the result is bitwise unchanged.
template <int NI>
__device__ __forceinline__ void hessianJTDJ_reg(const Env e, const PrimalCtx& c, const EfcPtrs& P, double* H, int ld,
                                                const int* c_dim, const double* cH) {
  const int nv = c.nv, nefc = c.nefc;
  const int ntri = nv * (nv + 1) / 2;
  // passes of 32 * NI entries (one pass up to nv = 24, two up to nv = 34): bounded register use, no spills
  B2K_NOUNROLL for (int base = 0; base < ntri; base += 32 * NI) {
    double s[NI];
    unsigned ij[NI];  // entry (i, j <= i) packed as i | j << 8
#pragma unroll
    for (int k = 0; k < NI; k++) {
      const int idx = base + e.lane + 32 * k;
      const bool ok = idx < ntri;
      // (i, j) of entry idx from the pair table the Cholesky uses (a square root and two search loops per entry
      // here were 3.9 % of a C4 step, profiles/r2b_ncu_lines_c4.txt)
      ij[k] = ok ? (unsigned)c.tri[idx] : 0u;
      s[k] = ok ? H[(ij[k] & 255u) * ld + (ij[k] >> 8)] : 0.0;
    }
    B2K_NOUNROLL for (int r = 0; r < nefc; r++) {
      const int st = P.state[r];
      if (st == B2MJ_CSTATE_QUADRATIC) {
        const double Dr = P.D[r];
        const double* Jr = P.J + r * nv;
#pragma unroll
        for (int k = 0; k < NI; k++) s[k] += Dr * Jr[ij[k] & 255u] * Jr[ij[k] >> 8];
      } else if (st == B2MJ_CSTATE_CONE) {
        const int con = P.id[r], dim = c_dim[con];
        const double* Hc = cH + c_dm.conh_stride * con;
        // B = Hc J_block (dim x nv) once per cone, by the whole warp; an entry then needs dim products instead of
        // dim^2.  B(a, j) is formed with the loop order the entry-wise form used, so the sums are bitwise the same.
        double* Bm = e.X(XF_SCRATCH);  // dim * nv <= 6 nv doubles; free during the dense Newton solve
        WSYNC();
        B2K_NOUNROLL for (int t = e.lane; t < dim * nv; t += B2K_G) {
          const int a = t / nv, j = t - a * nv;
          double u = 0;
          B2K_NOUNROLL for (int b = 0; b < dim; b++) u += Hc[a * dim + b] * P.J[(r + b) * nv + j];
          Bm[t] = u;
        }
        WSYNC();
#pragma unroll
        for (int k = 0; k < NI; k++) {
          double sk = s[k];
          const int oi = (int)(ij[k] & 255u), oj = (int)(ij[k] >> 8);
          B2K_NOUNROLL for (int a = 0; a < dim; a++) {
            const double Ja = P.J[(r + a) * nv + oi];
            if (Ja == 0) continue;
            sk += Ja * Bm[a * nv + oj];
          }
          s[k] = sk;
        }
        r += dim - 1;
      }
    }
#pragma unroll
    for (int k = 0; k < NI; k++)
      if (base + e.lane + 32 * k < ntri) H[(ij[k] & 255u) * ld + (ij[k] >> 8)] = s[k];
  }
}

// H = M + J' diag(D_active) J (+ cone blocks), then Cholesky
template <bool SM>
__device__ __noinline__ void primalHessianT(const Env e, PrimalCtx& c) {
  const DevModel& m = c_dm;
  if (m.team_warps > 1) {  // wide models: the whole team builds and factorises H in shared memory (team.cuh)
    team_call(e, TEAM_HESSIAN_CHOL, c.nefc, c.cone ? 1 : 0);
    return;
  }
  const int nv = c.nv, nefc = c.nefc, ld = m.ldh;
  EfcPtrs P = efcPtrs(e);
  P.J = const_cast<double*>(solveJ(e));  // active rows from the shared-memory window when staged
  const double* qM = e.D(B2MJ_F_QM);
  double* H = newtonH<SM>(e);
  FORL(k, nv * ld) H[k] = 0;
  WSYNC();
  FORL(t, m.nM) {
    const int i = m.M_row[t], j = m.M_col[t];
    H[i * ld + j] = qM[t];
    H[j * ld + i] = qM[t];
  }
  WSYNC();
  const int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
  const double* cH = c.cone ? e.XG(XF_CONTACT_H) : nullptr;
#if B2K_G == 32
  if (nv > B2K_SPARSE_H_MIN_NV) {
    // Row-sparse J' D J for large models (C5: nv 120, a contact row touches the <= 12 dofs of two free bodies).
    // Rows are visited in order; per row (or elliptic-cone block) the non-zero columns are compacted with a ballot
    // and every lane adds one (column, column) pair of the lower triangle.  Each H entry still accumulates its rows
    // in increasing order with the same expression as the dense loop below, which only adds exact zeros elsewhere:
    // the result is bitwise the same, at nnz^2 / 2 instead of nv^2 / 2 multiply-adds per row.
    int* cols = reinterpret_cast<int*>(e.X(XF_SCRATCH));
    B2K_NOUNROLL for (int r = 0; r < nefc; r++) {
      const int st = P.state[r];
      int dim = 1;
      if (st == B2MJ_CSTATE_CONE) dim = c_dim[P.id[r]];
      else if (st != B2MJ_CSTATE_QUADRATIC) continue;
      int nnz = 0;
      B2K_NOUNROLL for (int k0 = 0; k0 < nv; k0 += 32) {
        const int k = k0 + e.lane;
        bool nz = false;
        if (k < nv) { B2K_NOUNROLL for (int a = 0; a < dim; a++) nz |= P.J[(r + a) * nv + k] != 0; }
        const unsigned b = __ballot_sync(e.mask, nz);
        if (nz) cols[nnz + __popc(b & ((1u << e.lane) - 1u))] = k;
        nnz += __popc(b);
      }
      WSYNC();
      const int npair = nnz * (nnz + 1) / 2;
      B2K_NOUNROLL for (int idx = e.lane; idx < npair; idx += 32) {
        int a = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while (a * (a + 1) / 2 > idx) a--;
        while ((a + 1) * (a + 2) / 2 <= idx) a++;
        const int i = cols[a], j = cols[idx - a * (a + 1) / 2];  // cols ascending: i >= j
        double s = H[i * ld + j];
        if (st == B2MJ_CSTATE_QUADRATIC) {
          s += P.D[r] * P.J[r * nv + i] * P.J[r * nv + j];
        } else {
          const double* Hc = cH + c_dm.conh_stride * P.id[r];
          B2K_NOUNROLL for (int p = 0; p < dim; p++) {
            const double Ja = P.J[(r + p) * nv + i];
            if (Ja == 0) continue;
            double u = 0;
            B2K_NOUNROLL for (int q = 0; q < dim; q++) u += Hc[p * dim + q] * P.J[(r + q) * nv + j];
            s += Ja * u;
          }
        }
        H[i * ld + j] = s;
      }
      WSYNC();
      r += dim - 1;
    }
    cholFactor_warp<SM>(e, nv, ld, B2K_MINVAL, nullptr);  // XF_SCRATCH holds the column lists here
    return;
  }
#endif
  hessianJTDJ_reg<10>(e, c, P, H, ld, c_dim, cH);
  WSYNC();
  cholFactor_warp<SM>(e, nv, ld, B2K_MINVAL, c.tri);
}
__device__ __forceinline__ void primalHessian(const Env e, PrimalCtx& c) {
  if (c_dm.xoff_s[XF_NEWTON_H] >= 0) primalHessianT<true>(e, c);
  else primalHessianT<false>(e, c);
}

__device__ void primalGradient(const Env e, PrimalCtx& c) {
  const double* qs = e.D(B2MJ_F_QFRC_SMOOTH);
  const double* qc = e.D(B2MJ_F_QFRC_CONSTRAINT);
  FORL(i, c.nv) c.grad[i] = c.Ma[i] - qs[i] - qc[i];
  WSYNC();
  if (c.newton) {
    if (c_dm.team_warps > 1) {  // team mode: packed factor in shared memory
      cholSolve_warpT<true, B2K_CHOL_SLOTS_MAX, true>(e, c.nv, c_dm.ldh);
    } else if (c_dm.xoff_s[XF_NEWTON_H] >= 0) {
      if (c.nv <= 32) cholSolve_warpT<true, 1>(e, c.nv, c_dm.ldh);
      else cholSolve_warpT<true, B2K_CHOL_SLOTS_MAX>(e, c.nv, c_dm.ldh);
    } else {
      if (c.nv <= 32) cholSolve_warpT<false, 1>(e, c.nv, c_dm.ldh);
      else cholSolve_warpT<false, B2K_CHOL_SLOTS_MAX>(e, c.nv, c_dm.ldh);
    }
  } else {
    solveM_warp(e, c.Mgrad, c.grad);
  }
}

__device__ void primalPrepare(const Env e, PrimalCtx& c) {
  EfcPtrs P = efcPtrs(e);
  const double* qs = e.D(B2MJ_F_QFRC_SMOOTH);
  mulM_warp(e, c.Mv, c.search);
  mulJacVec_warp(e, c.nefc, c.Jv, c.search);
  c.quadGauss[0] = c.gauss;
  c.quadGauss[1] = dot_warp(e, c.search, c.Ma, c.nv) - dot_warp(e, qs, c.search, c.nv);
  c.quadGauss[2] = 0.5 * dot_warp(e, c.search, c.Mv, c.nv);
  FORL(i, c.nefc) {
    const double D = P.D[i];
    c.quad[3 * i] = 0.5 * D * c.Jaref[i] * c.Jaref[i];
    c.quad[3 * i + 1] = D * c.Jaref[i] * c.Jv[i];
    c.quad[3 * i + 2] = 0.5 * D * c.Jv[i] * c.Jv[i];
  }
  WSYNC();
}

__device__ __noinline__ LSPoint primalEval(const Env e, const PrimalCtx& c, double alpha) {
  EfcPtrs P = efcPtrs(e);
  const int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
  const int* c_adr = e.IG(B2MJ_F_CONTACT_EFC_ADDRESS);
  const double* c_mu = e.DG(B2MJ_F_CONTACT_MU);
  const double* c_fri = e.DG(B2MJ_F_CONTACT_FRICTION);
  double q0 = 0, q1 = 0, q2 = 0, cost = 0, deriv0 = 0, deriv1 = 0;
  FORL(i, c.nefc) {
    const double x = c.Jaref[i] + alpha * c.Jv[i];
    const double* qi = c.quad + 3 * i;
    const int type = P.type[i];
    if (type == B2MJ_CNSTR_EQUALITY) {
      q0 += qi[0]; q1 += qi[1]; q2 += qi[2];
    } else if (type == B2MJ_CNSTR_FRICTION_DOF || type == B2MJ_CNSTR_FRICTION_TENDON) {
      const double f = P.floss[i], Rf = P.R[i] * f;
      if (x <= -Rf) { q0 += f * (-0.5 * Rf - c.Jaref[i]); q1 += -f * c.Jv[i]; }
      else if (x >= Rf) { q0 += f * (-0.5 * Rf + c.Jaref[i]); q1 += f * c.Jv[i]; }
      else { q0 += qi[0]; q1 += qi[1]; q2 += qi[2]; }
    } else if (type == B2MJ_CNSTR_CONTACT_ELLIPTIC) {
      const int con = P.id[i];
      if (c_adr[con] != i) continue;  // the contact's first row handles the whole cone
      const int dim = c_dim[con];
      const double mu = c_mu[con];
      const double* fri = c_fri + 5 * con;
      const double U0 = c.Jaref[i] * mu, V0 = c.Jv[i] * mu;
      double UU = 0, UV = 0, VV = 0;
      B2K_NOUNROLL for (int j = 1; j < dim; j++) {
        const double U = c.Jaref[i + j] * fri[j - 1], V = c.Jv[i + j] * fri[j - 1];
        UU += U * U; UV += U * V; VV += V * V;
      }
      const double N = U0 + alpha * V0, Tsqr = UU + alpha * (2 * UV + alpha * VV);
      bool bottom = false;
      if (Tsqr <= 0) {
        if (N < 0) bottom = true;
      } else {
        const double T = sqrt(Tsqr);
        if (N >= mu * T) {
        } else if (mu * N + T <= 0) {
          bottom = true;
        } else {
          const double Dm = P.D[i] / (mu * mu * (1 + mu * mu));
          const double N1 = V0, T1 = (UV + alpha * VV) / T, T2 = VV / T - (UV + alpha * VV) * T1 / (T * T);
          const double NmT = N - mu * T;
          cost += 0.5 * Dm * NmT * NmT;
          deriv0 += Dm * NmT * (N1 - mu * T1);
          deriv1 += Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NmT * (-mu * T2));
        }
      }
      if (bottom)
        B2K_NOUNROLL for (int j = 0; j < dim; j++) { q0 += qi[3 * j]; q1 += qi[3 * j + 1]; q2 += qi[3 * j + 2]; }
    } else {
      if (x < 0) { q0 += qi[0]; q1 += qi[1]; q2 += qi[2]; }
    }
  }
  q0 = warpSum(e.mask, q0) + c.quadGauss[0];
  q1 = warpSum(e.mask, q1) + c.quadGauss[1];
  q2 = warpSum(e.mask, q2) + c.quadGauss[2];
  cost = warpSum(e.mask, cost);
  deriv0 = warpSum(e.mask, deriv0);
  deriv1 = warpSum(e.mask, deriv1);
  LSPoint p;
  p.alpha = alpha;
  p.cost = cost + alpha * alpha * q2 + alpha * q1 + q0;
  p.d0 = deriv0 + 2 * alpha * q2 + q1;
  p.d1 = deriv1 + 2 * q2;
  if (p.d1 < B2K_MINVAL) p.d1 = B2K_MINVAL;
  return p;
}

// exact line search on the convex piecewise-quadratic restriction; returns the step (0 = no progress)
__device__ __noinline__ double primalSearch(const Env e, PrimalCtx& c) {
  const DevModel& m = c_dm;
  const double snorm = sqrt(dot_warp(e, c.search, c.search, c.nv));
  if (snorm < B2K_MINVAL) return 0;
  const double gtol = m.opt.tolerance * m.opt.ls_tolerance * snorm / c.scale;
  primalPrepare(e, c);
  LSPoint p0 = primalEval(e, c, 0.0);
  LSPoint p1 = primalEval(e, c, p0.alpha - p0.d0 / p0.d1);
  if (p0.cost < p1.cost) p1 = p0;
  if (fabs(p1.d0) < gtol) return p1.alpha;
  const double dir = p1.d0 < 0 ? 1.0 : -1.0;
  int iter = 0;
  LSPoint p2 = p1;
  while (p1.d0 * dir <= -gtol && iter < m.opt.ls_iterations) {
    p2 = p1;
    p1 = primalEval(e, c, p1.alpha - p1.d0 / p1.d1);
    iter++;
    if (fabs(p1.d0) < gtol) return p1.alpha;
  }
  if (iter >= m.opt.ls_iterations || p1.d0 * dir <= -gtol) return p1.cost < p0.cost ? p1.alpha : 0;
  while (iter < m.opt.ls_iterations) {
    const double lo = fmin(p1.alpha, p2.alpha), hi = fmax(p1.alpha, p2.alpha);
    const double a1 = p1.alpha - p1.d0 / p1.d1, a2 = p2.alpha - p2.d0 / p2.d1;
    double cand[3];
    int nc = 0;
    if (a1 > lo && a1 < hi) cand[nc++] = a1;
    if (a2 > lo && a2 < hi) cand[nc++] = a2;
    cand[nc++] = 0.5 * (lo + hi);
    bool moved = false;
    B2K_NOUNROLL for (int k = 0; k < nc; k++) {
      const LSPoint pc = primalEval(e, c, cand[k]);
      if (fabs(pc.d0) < gtol) return pc.alpha;
      if (pc.d0 * dir < 0) {
        if (fabs(pc.alpha - p1.alpha) < fabs(p2.alpha - p1.alpha)) { p2 = pc; moved = true; }
      } else {
        if (fabs(pc.alpha - p2.alpha) < fabs(p1.alpha - p2.alpha)) { p1 = pc; moved = true; }
      }
    }
    iter++;
    if (!moved || fabs(p1.alpha - p2.alpha) < B2K_MINVAL) break;
  }
  const LSPoint best = p1.cost < p2.cost ? p1 : p2;
  return best.cost < p0.cost ? best.alpha : 0;
}

// mj_solNewton / mj_solCG.  qacc holds the starting point.  Returns iterations used.
__device__ __noinline__ int solvePrimal(const Env e, int nefc, int ncon, bool newton) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  PrimalCtx c;
  c.nv = nv; c.nefc = nefc; c.ncon = ncon;
  c.newton = newton;
  c.cone = m.opt.cone == B2MJ_CONE_ELLIPTIC;
  EfcPtrs P = efcPtrs(e);
  double* w = e.X(XF_PRIMAL);
  c.Ma = w; c.Mv = w + nv; c.grad = w + 2 * nv; c.Mgrad = w + 3 * nv; c.search = w + 4 * nv;
  c.gradold = w + 5 * nv; c.Mgradold = w + 6 * nv; c.invd = w + 7 * nv;
  c.Jaref = e.XG(XF_EFC_JAREF); c.Jv = e.XG(XF_EFC_JV); c.quad = e.XG(XF_EFC_QUAD);
  c.H = newton ? e.XG(XF_NEWTON_H) : nullptr;
  c.scale = 1 / (m.env_scalars[0] * max(1, nv));
  c.tri = (newton && m.team_warps == 1 && nv <= B2K_SPARSE_H_MIN_NV) ? cholPairTable(e, nv) : nullptr;
  double* qacc = e.D(B2MJ_F_QACC);

  SPROF_DECL
  mulM_warp(e, c.Ma, qacc);
  mulJacVec_warp(e, nefc, c.Jaref, qacc);
  FORL(i, nefc) c.Jaref[i] -= P.aref[i];
  WSYNC();
  SPROF(0)
  primalUpdate(e, c, nullptr);
  SPROF(1)
  if (newton) primalHessian(e, c);
  SPROF(2)
  primalGradient(e, c);
  SPROF(3)
  FORL(i, nv) c.search[i] = -c.Mgrad[i];
  WSYNC();
  int iter = 0;
  while (iter < m.opt.iterations) {
    const double alpha = primalSearch(e, c);
    SPROF(4)
    if (alpha == 0) break;
    FORL(i, nv) { qacc[i] += alpha * c.search[i]; c.Ma[i] += alpha * c.Mv[i]; }
    FORL(i, nefc) c.Jaref[i] += alpha * c.Jv[i];
    const double oldcost = c.cost;
    FORL(i, nv) { c.gradold[i] = c.grad[i]; c.Mgradold[i] = c.Mgrad[i]; }
    WSYNC();
    int changed = 0;
    primalUpdate(e, c, &changed);
    SPROF(1)
    if (newton && (c.cone || changed)) primalHessian(e, c);
    SPROF(2)
    primalGradient(e, c);
    SPROF(3)
    if (newton) {
      FORL(i, nv) c.search[i] = -c.Mgrad[i];
    } else {
      double num = 0, den = 0;
      FORL(i, nv) { num += c.grad[i] * (c.Mgrad[i] - c.Mgradold[i]); den += c.gradold[i] * c.Mgradold[i]; }
      num = warpSum(e.mask, num);
      den = warpSum(e.mask, den);
      double beta = num / fmax(B2K_MINVAL, den);
      if (beta < 0) beta = 0;
      FORL(i, nv) c.search[i] = -c.Mgrad[i] + beta * c.search[i];
    }
    WSYNC();
    const double improvement = c.scale * (oldcost - c.cost);
    const double gradient = c.scale * sqrt(dot_warp(e, c.grad, c.grad, nv));
    iter++;
    if (improvement < m.opt.tolerance || gradient < m.opt.tolerance) break;
  }
  return iter;
}

// mj_fwdConstraint.  Returns solver iterations.
__device__ __noinline__ int stage_fwdConstraint(const Env e, int nefc, int ncon) {
  const DevModel& m = c_dm;
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
    B2K_UNROLL4 for (int k = 0; k < nv; k++) s += P.J[i * nv + k] * qas[k];
    P.b[i] = s - P.aref[i];
  }
  WSYNC();
  const bool warmstart = !(m.opt.disableflags & B2MJ_DSBL_WARMSTART);
  int iters = 0;
  if (m.opt.solver == B2MJ_SOL_PGS) {
    // mj_projectConstraint, deferred from the position stage so that AR may overlay fields that are dead by now
    stage_projectConstraint(e, nefc);
    double* jar = e.XG(XF_EFC_JAREF);
    double* avec = e.X(XF_VEC1);
    const double* G = e.XG(XF_EFC_MINVJT);
    const double* dinv = e.DG(B2MJ_F_QLDIAGINV);
    const bool reg = nefc <= B2K_PGS_REGROWS;
    if (warmstart) {
      FORL(i, nefc) {
        double s = 0;
        B2K_UNROLL4 for (int k = 0; k < nv; k++) s += P.J[i * nv + k] * warm[k];
        jar[i] = s - P.aref[i];
      }
      WSYNC();
      constraintUpdate_warp(e, nefc, ncon, jar, false);
      // dual cost 0.5 f'ARf + f'b
      double cost = 0;
      if (reg) {
        const double* AR = arPtr(e, nefc);
        FORL(i, nefc) {
          double s = 0;
          B2K_NOUNROLL for (int k = 0; k < nefc; k++) s += AR[i * nefc + k] * P.force[k];
          cost += P.force[i] * (0.5 * s + P.b[i]);
        }
      } else {
        const double* U = m.dense_small ? P.J : G;
        FORL(k, nv) {
          double s = 0;
          B2K_NOUNROLL for (int i = 0; i < nefc; i++) s += U[i * nv + k] * P.force[i];
          avec[k] = m.dense_small ? s : s * dinv[k];
        }
        WSYNC();
        FORL(i, nefc) {
          double s = 0;
          B2K_NOUNROLL for (int k = 0; k < nv; k++) s += G[i * nv + k] * avec[k];
          cost += P.force[i] * (0.5 * (s + P.R[i] * P.force[i]) + P.b[i]);
        }
      }
      cost = warpSum(e.mask, cost);
      WSYNC();
      if (cost > 0) { FORL(i, nefc) P.force[i] = 0; }
      WSYNC();
    } else {
      FORL(i, nefc) P.force[i] = 0;
      WSYNC();
    }
    bool done = false;
    if (reg && nefc <= B2K_G) {
      const int it = solvePGS_own(e, nefc, arPtr(e, nefc));
      if (it >= 0) { iters = it; done = true; }
    }
    if (done) {
    } else if (reg) {
      const unsigned soff = arSmemOffset(e, nefc);
      const bool sm = soff != 0xffffffffu;
      const double* ARg = e.XG(XF_EFC_AR);
      if (sm && nefc <= B2K_G) iters = solvePGS_regT<1, true>(e, nefc, nullptr, soff);
      else if (sm) iters = solvePGS_regT<2, true>(e, nefc, nullptr, soff);   // window holds <= 17 rows
      else if (nefc <= B2K_G) iters = solvePGS_regT<1, false>(e, nefc, ARg, 0);
      else if (nefc <= 2 * B2K_G) iters = solvePGS_regT<2, false>(e, nefc, ARg, 0);
      else iters = solvePGS_regT<(64 / B2K_G > 2 ? 64 / B2K_G : 2), false>(e, nefc, ARg, 0);
    } else {
      iters = solvePGS_free(e, nefc, avec);
    }
    // dual finish: qfrc_constraint = J' f ; qacc = qacc_smooth + inv(M) qfrc_constraint
    mulJacTVec_warp(e, nefc, qfc, P.force);
    double* tmp = e.X(XF_VEC2);
    solveM_warp(e, tmp, qfc);
    FORL(i, nv) { const double a = qas[i] + tmp[i]; qacc[i] = a; warm[i] = a; }
    WSYNC();
  } else {
    const bool newton = m.opt.solver == B2MJ_SOL_NEWTON;
    if (m.team_warps > 1) team_call(e, TEAM_JCOLS, nefc, 0);  // column lists of J for this forward pass
    stageJWindow(e, nefc);                                     // active rows of J into shared memory when they fit
    if (warmstart) {
      // cost at the warm start vs at the unconstrained acceleration
      double* jar = e.XG(XF_EFC_JAREF);
      double* Ma = e.X(XF_PRIMAL);
      const double* qs = e.D(B2MJ_F_QFRC_SMOOTH);
      mulJacVec_warp(e, nefc, jar, warm);
      FORL(i, nefc) jar[i] -= P.aref[i];
      WSYNC();
      double cost_warm = constraintUpdate_warp(e, nefc, ncon, jar, false);
      mulM_warp(e, Ma, warm);
      double g = 0;
      FORL(i, nv) g += (Ma[i] - qs[i]) * (warm[i] - qas[i]);
      cost_warm += 0.5 * warpSum(e.mask, g);
      const double cost_smooth = constraintUpdate_warp(e, nefc, ncon, P.b, false);
      if (cost_warm < cost_smooth) { FORL(i, nv) qacc[i] = warm[i]; }
      else { FORL(i, nv) qacc[i] = qas[i]; }
    } else {
      FORL(i, nv) qacc[i] = qas[i];
    }
    WSYNC();
    iters = solvePrimal(e, nefc, ncon, newton);
    FORL(i, nv) warm[i] = qacc[i];
    WSYNC();
  }
  // noslip pass: after the warm start of the next step has been saved
  if (m.opt.noslip_iterations > 0) {
    if (m.opt.solver != B2MJ_SOL_PGS) stage_projectConstraint(e, nefc, false);
    iters += solveNoSlip(e, nefc, e.X(XF_VEC1));
    mulJacTVec_warp(e, nefc, qfc, P.force);
    double* tmp = e.X(XF_VEC2);
    solveM_warp(e, tmp, qfc);
    FORL(i, nv) qacc[i] = qas[i] + tmp[i];
    WSYNC();
  }
  return iters;
}

}  // namespace b2k
