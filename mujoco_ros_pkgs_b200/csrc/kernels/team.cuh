// team.cuh — several warps working on ONE env for the stages that dominate wide models.
//
// The step kernel gives every env one warp.  That is the right grain for the Panda (nv = 9), but the collision-heavy
// BASELINE config C5 (20 free boxes, nv = 120, Newton, elliptic cones, ~230 constraint rows) spent 94 % of its step in
// the Newton solve: a dense 120 x 120 Hessian built and factorised three to four times per step by 32 lanes out of an L2
// resident matrix (2.9 M cycles per factorisation, profiles/r2_*).  With 512 envs per GPU there are only 3.5 envs per SM,
// so the SM's other schedulers idle.  For such models (DevModel::team_warps > 1) a CTA carries ONE env and team_warps
// warps: warp 0 runs the step as before, the others park on a named barrier and are called in -- a command word in
// shared memory, two barrier crossings -- for the team stages:
//   TEAM_JCOLS         compact non-zero column lists of the constraint Jacobian rows (a contact row of C5 touches the
//                      12 dofs of two free bodies, not 120 columns)
//   TEAM_HESSIAN_CHOL  H = M + J' D J (+ elliptic cone blocks) accumulated row-sparse with every H row owned by one warp
//                      (deterministic: each entry adds its constraint rows in increasing order), then a right-looking
//                      Cholesky with the trailing update spread over all threads; H lives in shared memory with an odd
//                      leading dimension so column walks are bank-conflict free
// Reference semantics: mj_solNewton's Hessian / factor (the test oracle restates it); same matrix, different
// summation order in the factorisation (right-looking), within the solver's own tolerance.
#pragma once
#include "env_ctx.cuh"
#include "stages_constraint.cuh"

namespace b2k {

#ifdef B2K_SOLVE_PROF
__device__ unsigned long long g_sprof[16];  // cycle split of the primal solver (developer build only)
#define TPROF_DECL long long _tp = clock64();
#define TPROF(id) { const long long _n = clock64(); if (team_tid() == 0) atomicAdd(&g_sprof[id], (unsigned long long)(_n - _tp)); _tp = _n; }
#else
#define TPROF_DECL
#define TPROF(id)
#endif

#define B2K_JCOLS_K 16                       /* stored non-zero columns per Jacobian row; 255 in nnz = "dense row" */
#define B2K_JCOLS_STRIDE (B2K_JCOLS_K + 1)   /* bytes per row: nnz, then the columns */
#define B2K_TEAM_HDR 64                      /* bytes of per-env-slot header in team mode: mbarrier + control block */

enum { TEAM_EXIT = 0, TEAM_JCOLS = 1, TEAM_HESSIAN_CHOL = 2 };

struct TeamCtl {
  int cmd, nefc, ncon, newton_cone;
};

__device__ __forceinline__ int team_T() { return c_dm.team_warps * 32; }
__device__ __forceinline__ int team_tid() { return (int)threadIdx.x % (c_dm.team_warps * 32); }
// team mode runs ONE env slot per CTA (make_layout enforces it), so the slot is 0 and the team barrier is a fixed id:
// a run-time id would make ptxas reserve all 16 named barriers for every CTA of every model
__device__ __forceinline__ int team_slot() { return 0; }
__device__ __forceinline__ void team_bar() { asm volatile("bar.sync 2, %0;" ::"r"(team_T()) : "memory"); }
__device__ __forceinline__ TeamCtl* team_ctl() {
  return reinterpret_cast<TeamCtl*>(b2k_smem + (size_t)team_slot() * B2K_TEAM_HDR + 16);
}
__device__ __forceinline__ unsigned char* team_jcols(const Env e) { return reinterpret_cast<unsigned char*>(e.X(XF_JCOLS)); }
// XF_JVALS: J[r][cols[a]] for the first jvals_rows rows, [row][B2K_JCOLS_K], aligned with the row's column list
__device__ __forceinline__ double* team_jvals(const Env e) { return e.X(XF_JVALS); }

// ---- TEAM_JCOLS: per constraint row the sorted list of columns where the row (or, for a row of an elliptic contact,
// any row of that contact) is non-zero.  One warp per row, ballot compaction over 32-column chunks.
__device__ void team_build_jcols(const Env e, int nefc) {
  const DevModel& m = c_dm;
  const int nv = m.nv, w = team_tid() >> 5, lane = team_tid() & 31, TW = m.team_warps;
  const double* J = e.DG(B2MJ_F_EFC_J);
  const int* type = e.IG(B2MJ_F_EFC_TYPE);
  const int* id = e.IG(B2MJ_F_EFC_ID);
  const int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
  const int* c_adr = e.IG(B2MJ_F_CONTACT_EFC_ADDRESS);
  unsigned char* jc = team_jcols(e);
  for (int r = w; r < nefc; r += TW) {
    int r0 = r, dim = 1;
    if (type[r] == B2MJ_CNSTR_CONTACT_ELLIPTIC) { r0 = c_adr[id[r]]; dim = c_dim[id[r]]; }
    unsigned char* row = jc + (size_t)r * B2K_JCOLS_STRIDE;
    int nnz = 0;
    for (int k0 = 0; k0 < nv; k0 += 32) {
      const int k = k0 + lane;
      bool nz = false;
      if (k < nv)
        for (int a = 0; a < dim; a++) nz |= J[(size_t)(r0 + a) * nv + k] != 0;
      const unsigned b = __ballot_sync(0xffffffffu, nz);
      const int pos = nnz + __popc(b & ((1u << lane) - 1u));
      if (nz && pos < B2K_JCOLS_K) row[1 + pos] = (unsigned char)k;
      nnz += __popc(b);
    }
    if (lane == 0) row[0] = (unsigned char)(nnz > B2K_JCOLS_K ? 255 : nnz);
    if (r < m.jvals_rows && nnz <= B2K_JCOLS_K) {  // mirror the row's entries at the listed columns (one L2 trip, all rows in flight)
      __syncwarp();
      if (lane < nnz) team_jvals(e)[(size_t)r * B2K_JCOLS_K + lane] = J[(size_t)r * nv + row[1 + lane]];
    }
  }
}

// One sparse constraint row (DIM = 1, D_r J_r' J_r) or elliptic-cone block (DIM = contact dimension, J_b' Hc J_b) added to the
// H rows this warp owns.  Lane b holds column cols[b] of the block; the J entries of an owned row come from the lane that
// holds that column, by shuffle.  DIM is a template constant so that the 6 x 6 worst case costs nothing at dim 3 (the
// run-time-bounded form predicated 36 multiply-adds and 6 shuffles per owned row off and on: 3.1 k cycles per block);
// the sums are formed in the order of that form.
template <int DIM, bool CONE>
__device__ __forceinline__ void team_jtdj_sparse(double* H, const double* jvsrc, int jvstride, const double* Hc, double Dr, int myc,
                                                 bool own) {
  double jv[DIM], wv[DIM];
#pragma unroll
  for (int p = 0; p < DIM; p++) jv[p] = myc >= 0 ? jvsrc[p * jvstride] : 0.0;
  if (CONE) {
#pragma unroll
    for (int p = 0; p < DIM; p++) {
      double u = 0;
#pragma unroll
      for (int q = 0; q < DIM; q++) u += Hc[p * DIM + q] * jv[q];
      wv[p] = u;
    }
  }
  unsigned mine = __ballot_sync(0xffffffffu, own);
  while (mine) {
    const int a = __ffs(mine) - 1;
    mine &= mine - 1;
    const int i = __shfl_sync(0xffffffffu, myc, a);
    double acc = 0;
    if (!CONE) {
      acc = (Dr * __shfl_sync(0xffffffffu, jv[0], a)) * jv[0];
    } else {
#pragma unroll
      for (int p = 0; p < DIM; p++) acc += __shfl_sync(0xffffffffu, jv[p], a) * wv[p];
    }
    if (myc >= 0 && myc <= i) H[(((i) * ((i) + 1)) >> 1) + myc] += acc;
  }
}

// ---- TEAM_HESSIAN_CHOL
// The team's Hessian is the PACKED lower triangle, row i at i (i + 1) / 2: 58 KB instead of 116 KB at nv = 120, which is
// what lets the rest of the solver's working set (cone blocks, efc_type / id, friction) share the SM's shared memory
// with it.  Every access of the build, the factorisation and the substitution is to the lower triangle.
#define TRI(i) (((i) * ((i) + 1)) >> 1)
__device__ void team_hessian_chol(const Env e, int nefc, bool cone) {
  const DevModel& m = c_dm;
  const int nv = m.nv, T = team_T(), tid = team_tid(), w = tid >> 5, lane = tid & 31, TW = m.team_warps;
  double* H = e.X(XF_NEWTON_H);  // team mode: always in the shared arena (make_layout), addressed as shared memory
  const bool pow2 = (TW & (TW - 1)) == 0;
  double* invd = e.X(XF_PRIMAL) + 7 * nv;
  const double* qM = e.D(B2MJ_F_QM);
  EfcPtrs P = efcPtrs(e);
  const int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
  const double* cH = cone ? e.XG(XF_CONTACT_H) : nullptr;
  const unsigned char* jc = team_jcols(e);
  TPROF_DECL
  // H = M (lower triangle; everything else zero)
  for (int k = tid; k < TRI(nv); k += T) H[k] = 0;
  team_bar();
  for (int t = tid; t < m.nM; t += T) {
    const int a = max(m.M_row[t], m.M_col[t]), b = min(m.M_row[t], m.M_col[t]);
    H[TRI(a) + b] = qM[t];
  }
  team_bar();
  TPROF(5)
  // H += J' D J: warp w owns the H rows i with i % TW == w and visits the constraint rows in order
  for (int r = 0; r < nefc; r++) {
    const int st = P.state[r];
    int dim = 1;
    if (st == B2MJ_CSTATE_CONE) dim = c_dim[P.id[r]];
    else if (st != B2MJ_CSTATE_QUADRATIC) continue;
    const unsigned char* row = jc + (size_t)r * B2K_JCOLS_STRIDE;
    const int nnz = row[0];
    const double* Jr = P.J + (size_t)r * nv;
    if (nnz != 255) {
      // sparse row / block: lane b holds column cols[b].  All operands of the block are fetched up front in ONE L2
      // round trip (this lane's column of J for every row of the block, the cone block of the contact); the row
      // values J[.][i] of an owned H row i then come from the lane that holds column i, by shuffle.  Fetching them
      // inside the accumulation loop made the build a chain of dependent L2 gathers (465 k cycles per build).
      const int myc = lane < nnz ? row[1 + lane] : -1;
      const bool own = myc >= 0 && (pow2 ? (myc & (TW - 1)) : (myc % TW)) == w;
      const bool mirrored = r + dim <= m.jvals_rows;  // the block's entries are in shared memory (XF_JVALS)
      const double* jvsrc = mirrored ? team_jvals(e) + (size_t)r * B2K_JCOLS_K + lane : Jr + max(myc, 0);
      const int jvstride = mirrored ? B2K_JCOLS_K : nv;
      if (st == B2MJ_CSTATE_QUADRATIC) {
        team_jtdj_sparse<1, false>(H, jvsrc, jvstride, nullptr, P.D[r], myc, own);
      } else {
        const double* Hc = cH + c_dm.conh_stride * P.id[r];
        if (dim == 3) team_jtdj_sparse<3, true>(H, jvsrc, jvstride, Hc, 0.0, myc, own);
        else if (dim == 4) team_jtdj_sparse<4, true>(H, jvsrc, jvstride, Hc, 0.0, myc, own);
        else if (dim == 6) team_jtdj_sparse<6, true>(H, jvsrc, jvstride, Hc, 0.0, myc, own);
      }
    } else {
      // dense row / block: every owned H row i, lanes over the columns j <= i
      for (int i = w; i < nv; i += TW) {
        bool any = false;
        for (int p = 0; p < dim; p++) any |= Jr[p * nv + i] != 0;
        if (!any) continue;
        for (int j = lane; j <= i; j += 32) {
          double s = H[TRI(i) + j];
          if (st == B2MJ_CSTATE_QUADRATIC) {
            s += P.D[r] * Jr[i] * Jr[j];
          } else {
            const double* Hc = cH + c_dm.conh_stride * P.id[r];
            for (int p = 0; p < dim; p++) {
              const double Ja = Jr[p * nv + i];
              if (Ja == 0) continue;
              double u = 0;
              for (int q = 0; q < dim; q++) u += Hc[p * dim + q] * Jr[q * nv + j];
              s += Ja * u;
            }
          }
          H[TRI(i) + j] = s;
        }
      }
    }
    r += dim - 1;
  }
  team_bar();
  TPROF(6)
  // right-looking Cholesky, lower triangle in place; invd[j] = 1 / L[j][j]
  for (int j = 0; j < nv; j++) {
    double s = H[TRI(j) + j];
    if (s < B2K_MINVAL) s = B2K_MINVAL;
    const double ljj = sqrt(s), inv = 1 / ljj;
    for (int i = j + 1 + tid; i < nv; i += T) H[TRI(i) + j] *= inv;
    team_bar();
    if (tid == 0) { H[TRI(j) + j] = ljj; invd[j] = inv; }
    // trailing update: a warp takes rows i and i + TW together (two independent chains in flight and one load of
    // L(k, j) serving both), lanes over the columns j < k <= i.  Every entry still sees exactly one subtraction per pivot.
    for (int i = j + 1 + w; i < nv; i += 2 * TW) {
      const int i2 = i + TW;
      double* Hi = H + TRI(i);
      double* Hi2 = H + TRI(i2 < nv ? i2 : i);
      const double lij = Hi[j], lij2 = i2 < nv ? Hi2[j] : 0.0;
      // block structure: most of a contact-sparse factor is exact zeros
      const int kend = lij2 != 0 ? i2 : (lij != 0 ? i : -1);
      for (int k = j + 1 + lane; k <= kend; k += 32) {
        const double lkj = H[TRI(k) + j];
        if (lij != 0 && k <= i) Hi[k] -= lij * lkj;
        if (lij2 != 0) Hi2[k] -= lij2 * lkj;
      }
    }
    team_bar();
  }
  TPROF(7)
}

__device__ __noinline__ void team_exec(const Env e, int cmd, int nefc, int cone) {
  if (cmd == TEAM_JCOLS) team_build_jcols(e, nefc);
  else if (cmd == TEAM_HESSIAN_CHOL) team_hessian_chol(e, nefc, cone != 0);
}

// called by the env's main warp (warp 0 of the team)
__device__ __noinline__ void team_call(const Env e, int cmd, int nefc, int cone) {
  TeamCtl* c = team_ctl();
  if (e.lane == 0) { c->cmd = cmd; c->nefc = nefc; c->newton_cone = cone; }
  __syncwarp();
  team_bar();  // the helpers wait here
  team_exec(e, cmd, nefc, cone);
  team_bar();
}

// helper warps: serve team calls until the main warp retires the env slot
__device__ void team_worker(const Env e) {
  TeamCtl* c = team_ctl();
  for (;;) {
    team_bar();
    const int cmd = c->cmd;
    if (cmd == TEAM_EXIT) return;
    team_exec(e, cmd, c->nefc, c->newton_cone);
    team_bar();
  }
}
__device__ __forceinline__ void team_release(const Env e) {
  TeamCtl* c = team_ctl();
  if (e.lane == 0) c->cmd = TEAM_EXIT;
  __syncwarp();
  team_bar();
}

// ---- sparse Jacobian products for the main warp, on the column lists (res identical to the dense loops up to the
// order in which exact zeros are skipped)
__device__ __forceinline__ void mulJacVec_sparse(const Env e, int nefc, double* res, const double* vec) {
  const int nv = c_dm.nv;
  const double* J = e.DG(B2MJ_F_EFC_J);
  const unsigned char* jc = team_jcols(e);
  FORL(i, nefc) {
    const unsigned char* row = jc + (size_t)i * B2K_JCOLS_STRIDE;
    const double* Ji = J + (size_t)i * nv;
    double s = 0;
    if (row[0] != 255 && i < c_dm.jvals_rows) {
      const double* jvs = team_jvals(e) + (size_t)i * B2K_JCOLS_K;
      for (int a = 0; a < row[0]; a++) s += jvs[a] * vec[row[1 + a]];
    } else if (row[0] != 255) {
      for (int a = 0; a < row[0]; a++) { const int k = row[1 + a]; s += Ji[k] * vec[k]; }
    } else {
      for (int k = 0; k < nv; k++) s += Ji[k] * vec[k];
    }
    res[i] = s;
  }
  WSYNC();
}
__device__ __forceinline__ void mulJacTVec_sparse(const Env e, int nefc, double* res, const double* f) {
  const int nv = c_dm.nv;
  const double* J = e.DG(B2MJ_F_EFC_J);
  const unsigned char* jc = team_jcols(e);
  FORL(k, nv) res[k] = 0;
  WSYNC();
  // rows in order; within a row the columns are distinct, so the lanes never collide
  for (int i = 0; i < nefc; i++) {
    const double fi = f[i];
    if (fi == 0) continue;
    const unsigned char* row = jc + (size_t)i * B2K_JCOLS_STRIDE;
    const double* Ji = J + (size_t)i * nv;
    if (row[0] != 255 && i < c_dm.jvals_rows) {
      if (e.lane < row[0]) res[row[1 + e.lane]] += team_jvals(e)[(size_t)i * B2K_JCOLS_K + e.lane] * fi;
    } else if (row[0] != 255) {
      if (e.lane < row[0]) { const int k = row[1 + e.lane]; res[k] += Ji[k] * fi; }
    } else {
      FORL(k, nv) res[k] += Ji[k] * fi;
    }
    WSYNC();
  }
}

}  // namespace b2k
