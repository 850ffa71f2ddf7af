// stages_constraint.cuh — constraint rows, impedance, reference acceleration (one env per warp).
//
// Replaces mj_makeConstraint / mj_makeImpedance / mj_referenceConstraint / mj_constraintUpdate inside
// the reference's mj_step call (mujoco_env.cpp:498; rows M6, M7 of SURVEY 8a).  Row order is MuJoCo's:
// equality, friction loss, limits, contacts.  Data-dependent row counts (limits, contacts) get their
// addresses from ordered warp prefix sums, so the row order is a pure function of the state.
// Jacobian entries are produced one (constraint, dof) column per lane from cdof and the per-body
// dof-chain bitmask — no per-body Jacobian matrices are materialised.
#pragma once
#include "env_ctx.cuh"

namespace b2k {

// column k of the translational / rotational Jacobian of a world point on `body` (zero if off-chain)
__device__ __forceinline__ bool jacColumn(const double* cdof, const double* com, int body, int k,
                                          const double* point, double* jp, double* jr) {
  const DevModel& m = c_dm;
  const unsigned* mask = m.body_dofmask + body * m.nmaskword;
  if (!((mask[k >> 5] >> (k & 31)) & 1u)) {
    zero3(jp);
    if (jr) zero3(jr);
    return false;
  }
  const double* cd = cdof + 6 * k;
  double off[3], cr[3];
  sub3(off, point, com + 3 * m.body_rootid[body]);
  cross(cr, cd, off);
  jp[0] = cd[3] + cr[0]; jp[1] = cd[4] + cr[1]; jp[2] = cd[5] + cr[2];
  if (jr) copy3(jr, cd);
  return true;
}

// efc_J as the SOLVE stage reads it.  The Jacobian is sized for njmax rows and lives in the env's L2 arena, but a step
// uses a handful of rows (C3: ~9 of 96, C4: ~8 of 128): stage_fwdConstraint copies the active rows into a shared-memory
// window when they fit (XF_JWIN, sized by make_layout from the shared memory left over at the chosen residency) and
// every solver routine picks its base pointer here.  Round 2 profile: with J in L2 the Newton Hessian build was a
// chain of dependent ~600-cycle gathers (93 k cycles per build at nv = 24).
__device__ __forceinline__ const double* solveJ(const Env e) {
  if (c_dm.jwin_rows > 0) {
    const double* w = e.X(XF_JWIN);
    if (w[0] != 0.0) return w + 2;
  }
  return e.DG(B2MJ_F_EFC_J);
}
__device__ __forceinline__ void stageJWindow(const Env e, int nefc) {
  if (c_dm.jwin_rows <= 0) return;
  double* w = e.X(XF_JWIN);
  const bool fits = nefc <= c_dm.jwin_rows;
  if (fits) {
    const double* J = e.DG(B2MJ_F_EFC_J);
    const int n = nefc * c_dm.nv;
    FORL(k, n) w[2 + k] = J[k];
  }
  if (e.lane == 0) w[0] = fits ? 1.0 : 0.0;
  WSYNC();
}

struct EfcPtrs {
  double *J, *pos, *margin, *floss, *diag, *KBIP, *D, *R, *vel, *aref, *b, *force;
  int *type, *id, *state;
};

__device__ __forceinline__ EfcPtrs efcPtrs(const Env e) {
  EfcPtrs p;
  p.J = e.DG(B2MJ_F_EFC_J); p.pos = e.DG(B2MJ_F_EFC_POS); p.margin = e.DG(B2MJ_F_EFC_MARGIN);
  p.floss = e.DG(B2MJ_F_EFC_FRICTIONLOSS); p.diag = e.DG(B2MJ_F_EFC_DIAGAPPROX); p.KBIP = e.DG(B2MJ_F_EFC_KBIP);
  p.D = e.DG(B2MJ_F_EFC_D); p.R = e.DG(B2MJ_F_EFC_R); p.vel = e.DG(B2MJ_F_EFC_VEL); p.aref = e.DG(B2MJ_F_EFC_AREF);
  p.b = e.DG(B2MJ_F_EFC_B); p.force = e.DG(B2MJ_F_EFC_FORCE);
  p.type = e.IG(B2MJ_F_EFC_TYPE); p.id = e.IG(B2MJ_F_EFC_ID); p.state = e.IG(B2MJ_F_EFC_STATE);
  return p;
}

__device__ __forceinline__ double getImpedance(const double* solimp, double pos, double margin) {
  const double dmin = clampd(solimp[0], B2MJ_MINIMP, B2MJ_MAXIMP), dmax = clampd(solimp[1], B2MJ_MINIMP, B2MJ_MAXIMP);
  const double width = fmax(B2K_MINVAL, solimp[2]), mid = clampd(solimp[3], B2MJ_MINIMP, B2MJ_MAXIMP);
  const double power = fmax(1.0, solimp[4]);
  if (dmin == dmax) return 0.5 * (dmin + dmax);
  const double x = fabs(pos - margin) / width;
  double y;
  if (x >= 1) y = 1;
  else if (x <= 0) y = 0;
  else if (power == 1) y = x;
  else if (x <= mid) y = pow(x / mid, power) * mid;
  else y = 1 - pow((1 - x) / (1 - mid), power) * (1 - mid);
  return dmin + y * (dmax - dmin);
}

// mj_makeConstraint; returns nefc (also stored)
__device__ __noinline__ int stage_makeConstraint(const Env e, int ncon, int* warning) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  int* nefc_p = e.I(B2MJ_F_NEFC);
  int* c_adr = e.IG(B2MJ_F_CONTACT_EFC_ADDRESS);
  FORL(c, ncon) c_adr[c] = -1;
  if ((m.opt.disableflags & B2MJ_DSBL_CONSTRAINT) || m.njmax == 0 || nv == 0) {
    if (e.lane == 0) nefc_p[0] = 0;
    WSYNC();
    return 0;
  }
  EfcPtrs P = efcPtrs(e);
  const double* cdof = e.D(B2MJ_F_CDOF);
  const double* com = e.D(B2MJ_F_SUBTREE_COM);
  const double* qpos = e.D(B2MJ_F_QPOS);
  int row = 0, full = 0;

  // ---------------- equality ----------------
  if (m.neq && !(m.opt.disableflags & B2MJ_DSBL_EQUALITY)) {
    const double* xpos = e.D(B2MJ_F_XPOS);
    const double* xmat = e.D(B2MJ_F_XMAT);
    const double* xquat = e.D(B2MJ_F_XQUAT);
    B2K_NOUNROLL for (int q = 0; q < m.neq; q++) {
      if (!m.eq_active[q]) continue;
      const int et = m.eq_type[q], id0 = m.eq_obj1id[q], id1 = m.eq_obj2id[q];
      const double* data = m.eq_data + B2MJ_NEQDATA * q;
      const int size = et == B2MJ_EQ_CONNECT ? 3 : et == B2MJ_EQ_WELD ? 6 : 1;
      if (row + size > m.njmax) { full = 1; continue; }
      double cpos[6] = {0, 0, 0, 0, 0, 0};
      if (et == B2MJ_EQ_CONNECT || et == B2MJ_EQ_WELD) {
        double pos0[3], pos1[3];
        const double* a0 = et == B2MJ_EQ_CONNECT ? data : data + 3;
        const double* a1 = et == B2MJ_EQ_CONNECT ? data + 3 : data;
        rotVecMat(pos0, a0, xmat + 9 * id0); addTo3(pos0, xpos + 3 * id0);
        rotVecMat(pos1, a1, xmat + 9 * id1); addTo3(pos1, xpos + 3 * id1);
        sub3(cpos, pos0, pos1);
        double quat[4], quat1[4];
        const double ts = data[10];
        if (et == B2MJ_EQ_WELD) {
          double quat2[4];
          mulQuat(quat, xquat + 4 * id0, data + 6);
          negQuat(quat1, xquat + 4 * id1);
          mulQuat(quat2, quat1, quat);
          scl3(cpos + 3, quat2 + 1, ts);
        }
        FORL(k, nv) {
          double jp0[3], jp1[3], jr0[3], jr1[3];
          jacColumn(cdof, com, id0, k, pos0, jp0, jr0);
          jacColumn(cdof, com, id1, k, pos1, jp1, jr1);
          for (int r = 0; r < 3; r++) P.J[(row + r) * nv + k] = jp0[r] - jp1[r];
          if (et == B2MJ_EQ_WELD) {
            double axis[3], q2[4], q3[4];
            sub3(axis, jr0, jr1);
            mulQuatAxis(q2, quat1, axis);
            mulQuat(q3, q2, quat);
            for (int r = 0; r < 3; r++) P.J[(row + 3 + r) * nv + k] = 0.5 * q3[1 + r] * ts;
          }
        }
        if (e.lane == 0) {
          const double tran = m.body_invweight0[2 * id0] + m.body_invweight0[2 * id1];
          const double rot = m.body_invweight0[2 * id0 + 1] + m.body_invweight0[2 * id1 + 1];
          B2K_NOUNROLL for (int r = 0; r < size; r++) P.diag[row + r] = r < 3 ? tran : rot;
        }
      } else {
        const bool isj = et == B2MJ_EQ_JOINT;
        double pos0, pos1 = 0, ref0, ref1 = 0, deriv = 0;
        const double* tl = m.ntendon ? e.D(B2MJ_F_TEN_LENGTH) : nullptr;
        const double* tJ = m.ntendon ? e.D(B2MJ_F_TEN_J) : nullptr;
        if (isj) {
          pos0 = qpos[m.jnt_qposadr[id0]]; ref0 = m.qpos0[m.jnt_qposadr[id0]];
          if (id1 >= 0) { pos1 = qpos[m.jnt_qposadr[id1]]; ref1 = m.qpos0[m.jnt_qposadr[id1]]; }
        } else {
          pos0 = tl[id0]; ref0 = m.tendon_length0[id0];
          if (id1 >= 0) { pos1 = tl[id1]; ref1 = m.tendon_length0[id1]; }
        }
        if (id1 >= 0) {
          const double dif = pos1 - ref1;
          const double dif2 = dif * dif, dif3 = dif2 * dif, dif4 = dif3 * dif;
          cpos[0] = pos0 - ref0 - data[0] - (data[1] * dif + data[2] * dif2 + data[3] * dif3 + data[4] * dif4);
          deriv = data[1] + 2 * data[2] * dif + 3 * data[3] * dif2 + 4 * data[4] * dif3;
        } else {
          cpos[0] = pos0 - ref0 - data[0];
        }
        FORL(k, nv) {
          double v = 0;
          if (isj) {
            if (id1 >= 0 && k == m.jnt_dofadr[id1]) v = -deriv;
            if (k == m.jnt_dofadr[id0]) v = 1;
          } else {
            v = tJ[id0 * nv + k];
            if (id1 >= 0) v -= deriv * tJ[id1 * nv + k];
          }
          P.J[row * nv + k] = v;
        }
        if (e.lane == 0) {
          if (isj) P.diag[row] = m.dof_invweight0[m.jnt_dofadr[id0]] + (id1 >= 0 ? m.dof_invweight0[m.jnt_dofadr[id1]] : 0.0);
          else P.diag[row] = m.tendon_invweight0[id0] + (id1 >= 0 ? m.tendon_invweight0[id1] : 0.0);
        }
      }
      if (e.lane < size) {
        const int r = row + e.lane;
        P.pos[r] = cpos[e.lane]; P.margin[r] = 0; P.floss[r] = 0;
        P.type[r] = B2MJ_CNSTR_EQUALITY; P.id[r] = q;
      }
      row += size;
    }
  }
  // ---------------- friction loss ----------------
  if (!(m.opt.disableflags & B2MJ_DSBL_FRICTIONLOSS)) {
    B2K_NOUNROLL for (int i = 0; i < nv; i++) {
      const double fl = m.dof_frictionloss[i];
      if (fl <= 0) continue;
      if (row + 1 > m.njmax) { full = 1; continue; }
      FORL(k, nv) P.J[row * nv + k] = (k == i) ? 1.0 : 0.0;
      if (e.lane == 0) {
        P.pos[row] = 0; P.margin[row] = 0; P.floss[row] = fl; P.type[row] = B2MJ_CNSTR_FRICTION_DOF; P.id[row] = i;
        P.diag[row] = m.dof_invweight0[i];
      }
      row++;
    }
    B2K_NOUNROLL for (int i = 0; i < m.ntendon; i++) {
      const double fl = m.tendon_frictionloss[i];
      if (fl <= 0) continue;
      if (row + 1 > m.njmax) { full = 1; continue; }
      const double* tJ = e.D(B2MJ_F_TEN_J);
      FORL(k, nv) P.J[row * nv + k] = tJ[i * nv + k];
      if (e.lane == 0) {
        P.pos[row] = 0; P.margin[row] = 0; P.floss[row] = fl; P.type[row] = B2MJ_CNSTR_FRICTION_TENDON; P.id[row] = i;
        P.diag[row] = m.tendon_invweight0[i];
      }
      row++;
    }
  }
  // ---------------- limits ----------------
  if (!(m.opt.disableflags & B2MJ_DSBL_LIMIT)) {
    B2K_NOUNROLL for (int base = 0; base < m.njnt; base += B2K_G) {
      const int j = base + e.lane;
      int cnt = 0;
      double dist[2], sgn[2], axis[3];
      bool ball = false;
      if (j < m.njnt && m.jnt_limited[j]) {
        const double margin = m.jnt_margin[j];
        const int jt = m.jnt_type[j];
        if (jt == B2MJ_JNT_SLIDE || jt == B2MJ_JNT_HINGE) {
          const double value = qpos[m.jnt_qposadr[j]];
          for (int side = -1; side <= 1; side += 2) {
            const double d = side * (m.jnt_range[2 * j + (side + 1) / 2] - value);
            if (d < margin) { dist[cnt] = d; sgn[cnt] = -(double)side; cnt++; }
          }
        } else if (jt == B2MJ_JNT_BALL) {
          double quat[4];
          copy4(quat, qpos + m.jnt_qposadr[j]);
          normalize4(quat);
          quat2Vel(axis, quat, 1);
          const double value = normalize3(axis);
          const double d = fmax(m.jnt_range[2 * j], m.jnt_range[2 * j + 1]) - value;
          if (d < margin) { dist[0] = d; cnt = 1; ball = true; }
        }
      }
      const int incl = warpInclusiveScan(e.mask, cnt, e.lane);
      const int total = __shfl_sync(e.mask, incl, B2K_G - 1, B2K_G);
      int r0 = row + incl - cnt;
      B2K_NOUNROLL for (int s = 0; s < cnt; s++) {
        const int r = r0 + s;
        if (r >= m.njmax) { full = 1; break; }
        const int da = m.jnt_dofadr[j];
        B2K_NOUNROLL for (int k = 0; k < nv; k++) P.J[r * nv + k] = 0;
        if (ball) for (int k = 0; k < 3; k++) P.J[r * nv + da + k] = -axis[k];
        else P.J[r * nv + da] = sgn[s];
        P.pos[r] = dist[s]; P.margin[r] = m.jnt_margin[j]; P.floss[r] = 0;
        P.type[r] = B2MJ_CNSTR_LIMIT_JOINT; P.id[r] = j;
        P.diag[r] = m.dof_invweight0[da];
      }
      row = min(row + total, m.njmax);
    }
    if (m.ntendon) {
      const double* tl = e.D(B2MJ_F_TEN_LENGTH);
      const double* tJ = e.D(B2MJ_F_TEN_J);
      B2K_NOUNROLL for (int i = 0; i < m.ntendon; i++) {
        if (!m.tendon_limited[i]) continue;
        const double value = tl[i], margin = m.tendon_margin[i];
        for (int side = -1; side <= 1; side += 2) {
          const double d = side * (m.tendon_range[2 * i + (side + 1) / 2] - value);
          if (d < margin) {
            if (row + 1 > m.njmax) { full = 1; continue; }
            FORL(k, nv) P.J[row * nv + k] = -side * tJ[i * nv + k];
            if (e.lane == 0) {
              P.pos[row] = d; P.margin[row] = margin; P.floss[row] = 0; P.type[row] = B2MJ_CNSTR_LIMIT_TENDON; P.id[row] = i;
              P.diag[row] = m.tendon_invweight0[i];
            }
            row++;
          }
        }
      }
    }
  }
  // ---------------- contacts ----------------
  if (ncon > 0 && !(m.opt.disableflags & B2MJ_DSBL_CONTACT)) {
    const bool pyramid = m.opt.cone == B2MJ_CONE_PYRAMIDAL;
    const int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
    const int* c_excl = e.IG(B2MJ_F_CONTACT_EXCLUDE);
    const int* c_g1 = e.IG(B2MJ_F_CONTACT_GEOM1);
    const int* c_g2 = e.IG(B2MJ_F_CONTACT_GEOM2);
    const double* c_dist = e.DG(B2MJ_F_CONTACT_DIST);
    const double* c_inc = e.DG(B2MJ_F_CONTACT_INCLUDEMARGIN);
    const double* c_pos = e.DG(B2MJ_F_CONTACT_POS);
    const double* c_frame = e.DG(B2MJ_F_CONTACT_FRAME);
    const double* c_fri = e.DG(B2MJ_F_CONTACT_FRICTION);
    bool contacts_full = false;  // uniform: once a contact does not fit, all later ones are dropped
    B2K_NOUNROLL for (int base = 0; base < ncon && !contacts_full; base += B2K_G) {
      const int c = base + e.lane;
      int cnt = 0, dim = 0;
      if (c < ncon && !c_excl[c]) {
        dim = c_dim[c];
        cnt = dim == 1 ? 1 : (pyramid ? 2 * (dim - 1) : dim);
      }
      const int incl = warpInclusiveScan(e.mask, cnt, e.lane);
      int endrow = row;
      if (cnt > 0) {
        const int adr = row + incl - cnt;
        if (adr + cnt > m.njmax) {
          full = 1;  // this and (by monotonicity of addresses) all later contacts are dropped
        } else {
          endrow = adr + cnt;
          c_adr[c] = adr;
          const int type = dim == 1 ? B2MJ_CNSTR_CONTACT_FRICTIONLESS
                                    : (pyramid ? B2MJ_CNSTR_CONTACT_PYRAMIDAL : B2MJ_CNSTR_CONTACT_ELLIPTIC);
          const int b1 = m.geom_bodyid[c_g1[c]], b2 = m.geom_bodyid[c_g2[c]];
          const double tran = m.body_invweight0[2 * b1] + m.body_invweight0[2 * b2];
          const double rot = m.body_invweight0[2 * b1 + 1] + m.body_invweight0[2 * b2 + 1];
          const double* fri = c_fri + 5 * c;
          B2K_NOUNROLL for (int r = 0; r < cnt; r++) {
            const bool first_or_pyr = (r == 0) || type != B2MJ_CNSTR_CONTACT_ELLIPTIC;
            P.pos[adr + r] = first_or_pyr ? c_dist[c] : 0.0;
            P.margin[adr + r] = first_or_pyr ? c_inc[c] : 0.0;
            P.floss[adr + r] = 0;
            P.type[adr + r] = type;
            P.id[adr + r] = c;
            double dg;
            if (type == B2MJ_CNSTR_CONTACT_PYRAMIDAL) { const int kk = r / 2; dg = tran + fri[kk] * fri[kk] * (kk < 2 ? tran : rot); }
            else dg = r < 3 ? tran : rot;
            P.diag[adr + r] = dg;
          }
        }
      }
      contacts_full = __any_sync(e.mask, full);
      row = warpMaxInt(e.mask, endrow);
    }
    WSYNC();
    // Jacobian entries: one (contact, dof) column per lane
    FORL(item, ncon * nv) {
      const int c = item / nv, k = item - c * nv;
      const int adr = c_adr[c];
      if (adr < 0) continue;
      const int dim = c_dim[c];
      const int b1 = m.geom_bodyid[c_g1[c]], b2 = m.geom_bodyid[c_g2[c]];
      double jp1[3], jp2[3], jr1[3], jr2[3], dp[3], dr[3], v[6];
      jacColumn(cdof, com, b1, k, c_pos + 3 * c, jp1, jr1);
      jacColumn(cdof, com, b2, k, c_pos + 3 * c, jp2, jr2);
      sub3(dp, jp2, jp1);
      sub3(dr, jr2, jr1);
      const double* fr = c_frame + 9 * c;
      for (int r = 0; r < 3; r++) { v[r] = dot3(fr + 3 * r, dp); v[3 + r] = dot3(fr + 3 * r, dr); }
      if (dim == 1) {
        P.J[adr * nv + k] = v[0];
      } else if (pyramid) {
        const double* fri = c_fri + 5 * c;
        B2K_NOUNROLL for (int kk = 1; kk < dim; kk++) {
          P.J[(adr + 2 * (kk - 1)) * nv + k] = v[0] + fri[kk - 1] * v[kk];
          P.J[(adr + 2 * (kk - 1) + 1) * nv + k] = v[0] - fri[kk - 1] * v[kk];
        }
      } else {
        B2K_NOUNROLL for (int r = 0; r < dim; r++) P.J[(adr + r) * nv + k] = v[r];
      }
    }
  }
  if (__any_sync(e.mask, full)) {
    if (e.lane == 0) warning[B2MJ_WARN_CNSTRFULL]++;
  }
  const int nefc = row;
  if (e.lane == 0) nefc_p[0] = nefc;
  WSYNC();

  // ---------------- impedance: KBIP, R, D (mj_makeImpedance) ----------------
  const bool refsafe = !(m.opt.disableflags & B2MJ_DSBL_REFSAFE);
  const double* c_solref = e.DG(B2MJ_F_CONTACT_SOLREF);
  const double* c_solimp = e.DG(B2MJ_F_CONTACT_SOLIMP);
  FORL(i, nefc) {
    const int id = P.id[i], type = P.type[i];
    const double *solref, *solimp;
    bool fr_row = false;
    switch (type) {
      case B2MJ_CNSTR_EQUALITY: solref = m.eq_solref + 2 * id; solimp = m.eq_solimp + 5 * id; break;
      case B2MJ_CNSTR_FRICTION_DOF: solref = m.dof_solref + 2 * id; solimp = m.dof_solimp + 5 * id; fr_row = true; break;
      case B2MJ_CNSTR_FRICTION_TENDON: solref = m.tendon_solref_fri + 2 * id; solimp = m.tendon_solimp_fri + 5 * id; fr_row = true; break;
      case B2MJ_CNSTR_LIMIT_JOINT: solref = m.jnt_solref + 2 * id; solimp = m.jnt_solimp + 5 * id; break;
      case B2MJ_CNSTR_LIMIT_TENDON: solref = m.tendon_solref_lim + 2 * id; solimp = m.tendon_solimp_lim + 5 * id; break;
      default:
        solref = c_solref + 2 * id; solimp = c_solimp + 5 * id;
        if (type == B2MJ_CNSTR_CONTACT_ELLIPTIC && i > c_adr[id]) fr_row = true;
    }
    const double dmax = clampd(solimp[1], B2MJ_MINIMP, B2MJ_MAXIMP);
    const double imp = getImpedance(solimp, P.pos[i], P.margin[i]);
    double K, B;
    if (solref[0] > 0) {
      double tc = solref[0];
      if (refsafe) tc = fmax(tc, 2 * m.opt.timestep);
      const double dr = solref[1];
      K = 1 / fmax(B2K_MINVAL, dmax * dmax * tc * tc * dr * dr);
      B = 2 / fmax(B2K_MINVAL, dmax * tc);
    } else {
      K = -solref[0] / fmax(B2K_MINVAL, dmax * dmax);
      B = -solref[1] / fmax(B2K_MINVAL, dmax);
    }
    if (fr_row) K = 0;
    P.KBIP[4 * i] = K; P.KBIP[4 * i + 1] = B; P.KBIP[4 * i + 2] = imp; P.KBIP[4 * i + 3] = 0;
    P.R[i] = fmax(B2K_MINVAL, (1 - imp) * P.diag[i] / imp);
  }
  WSYNC();
  if (ncon > 0) {
    const int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
    const double* c_fri = e.DG(B2MJ_F_CONTACT_FRICTION);
    double* c_mu = e.DG(B2MJ_F_CONTACT_MU);
    FORL(c, ncon) {
      const int adr = c_adr[c], dim = c_dim[c];
      if (adr < 0 || dim == 1) continue;
      const double* fri = c_fri + 5 * c;
      if (m.opt.cone == B2MJ_CONE_ELLIPTIC) {
        P.R[adr + 1] = P.R[adr] / fmax(B2K_MINVAL, m.opt.impratio);
        c_mu[c] = fri[0] * sqrt(P.R[adr + 1] / P.R[adr]);
        B2K_NOUNROLL for (int j = 2; j < dim; j++) P.R[adr + j] = P.R[adr + 1] * fri[0] * fri[0] / (fri[j - 1] * fri[j - 1]);
      } else {
        const double mu = fri[0] * sqrt(1 / fmax(B2K_MINVAL, m.opt.impratio));
        c_mu[c] = mu;
        const double Rpy = 2 * mu * mu * P.R[adr];
        B2K_NOUNROLL for (int j = 0; j < 2 * (dim - 1); j++) P.R[adr + j] = Rpy;
      }
    }
    WSYNC();
  }
  FORL(i, nefc) P.D[i] = 1 / P.R[i];
  WSYNC();
  return nefc;
}

// mj_referenceConstraint: efc_vel = J qvel, aref = -B vel - K imp (pos - margin)
__device__ __noinline__ void stage_referenceConstraint(const Env e, int nefc) {
  if (!nefc) return;
  const DevModel& m = c_dm;
  const int nv = m.nv;
  EfcPtrs P = efcPtrs(e);
  const double* qvel = e.D(B2MJ_F_QVEL);
  FORL(i, nefc) {
    double s = 0;
    B2K_UNROLL4 for (int k = 0; k < nv; k++) s += P.J[i * nv + k] * qvel[k];
    P.vel[i] = s;
    P.aref[i] = -P.KBIP[4 * i + 1] * s - P.KBIP[4 * i] * P.KBIP[4 * i + 2] * (P.pos[i] - P.margin[i]);
  }
  WSYNC();
}

// res[k] = sum_i J[i][k] * f[i]   (J' f), one lane per dof
__device__ __noinline__ void mulJacTVec_warp(const Env e, int nefc, double* res, const double* f) {
  const int nv = c_dm.nv;
  const double* J = solveJ(e);  // only called from the solve stage (window staged) or with the window invalid
  FORL(k, nv) {
    double s = 0;
    B2K_NOUNROLL for (int i = 0; i < nefc; i++) {
      const double fi = f[i];
      if (fi != 0) s += J[i * nv + k] * fi;
    }
    res[k] = s;
  }
  WSYNC();
}

// mj_constraintUpdate: force / state / cost for jar = J qacc - aref; returns the constraint cost
// (identical on all lanes).  Does NOT compute qfrc_constraint (callers do, when they need it).
__device__ __noinline__ double constraintUpdate_warp(const Env e, int nefc, int ncon, const double* jar, bool coneHessian,
                                        int* changed = nullptr) {
  const DevModel& m = c_dm;
  EfcPtrs P = efcPtrs(e);
  double s = 0;
  int ch = 0;
  FORL(i, nefc) {
    const double D = P.D[i], R = P.R[i], x = jar[i];
    const int oldstate = P.state[i];
    switch (P.type[i]) {
      case B2MJ_CNSTR_EQUALITY:
        P.force[i] = -D * x; P.state[i] = B2MJ_CSTATE_QUADRATIC; s += 0.5 * D * x * x;
        break;
      case B2MJ_CNSTR_FRICTION_DOF:
      case B2MJ_CNSTR_FRICTION_TENDON: {
        const double f = P.floss[i];
        if (x <= -R * f) { P.force[i] = f; P.state[i] = B2MJ_CSTATE_LINEARNEG; s += -0.5 * R * f * f - f * x; }
        else if (x >= R * f) { P.force[i] = -f; P.state[i] = B2MJ_CSTATE_LINEARPOS; s += -0.5 * R * f * f + f * x; }
        else { P.force[i] = -D * x; P.state[i] = B2MJ_CSTATE_QUADRATIC; s += 0.5 * D * x * x; }
        break;
      }
      case B2MJ_CNSTR_CONTACT_ELLIPTIC: break;  // handled per contact below
      default:
        if (x >= 0) { P.force[i] = 0; P.state[i] = B2MJ_CSTATE_SATISFIED; }
        else { P.force[i] = -D * x; P.state[i] = B2MJ_CSTATE_QUADRATIC; s += 0.5 * D * x * x; }
    }
    if (P.type[i] != B2MJ_CNSTR_CONTACT_ELLIPTIC) ch |= (P.state[i] != oldstate);
  }
  if (m.opt.cone == B2MJ_CONE_ELLIPTIC && ncon > 0) {
    const int* c_adr = e.IG(B2MJ_F_CONTACT_EFC_ADDRESS);
    const int* c_dim = e.IG(B2MJ_F_CONTACT_DIM);
    const double* c_mu = e.DG(B2MJ_F_CONTACT_MU);
    const double* c_fri = e.DG(B2MJ_F_CONTACT_FRICTION);
    double* cH = e.XG(XF_CONTACT_H);
    FORL(c, ncon) {
      const int i = c_adr[c], dim = c_dim[c];
      if (i < 0 || dim == 1) continue;
      const double mu = c_mu[c];
      const double* fri = c_fri + 5 * c;
      double U[6];
      U[0] = jar[i] * mu;
      double TT = 0;
      B2K_NOUNROLL for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * fri[j - 1]; TT += U[j] * U[j]; }
      const double N = U[0], T = sqrt(TT);
      if ((N >= mu * T) || (T <= 0 && N >= 0)) {
        B2K_NOUNROLL for (int j = 0; j < dim; j++) { P.force[i + j] = 0; P.state[i + j] = B2MJ_CSTATE_SATISFIED; }
      } else if ((mu * N + T <= 0) || (T <= 0 && N < 0)) {
        B2K_NOUNROLL for (int j = 0; j < dim; j++) {
          const double x = jar[i + j];
          P.force[i + j] = -P.D[i + j] * x;
          P.state[i + j] = B2MJ_CSTATE_QUADRATIC;
          s += 0.5 * P.D[i + j] * x * x;
        }
      } else {
        const double Dm = P.D[i] / (mu * mu * (1 + mu * mu));
        const double NmT = N - mu * T;
        s += 0.5 * Dm * NmT * NmT;
        const double f0 = -Dm * NmT * mu;
        P.force[i] = f0;
        B2K_NOUNROLL for (int j = 1; j < dim; j++) P.force[i + j] = -f0 / T * U[j] * fri[j - 1];
        B2K_NOUNROLL for (int j = 0; j < dim; j++) P.state[i + j] = B2MJ_CSTATE_CONE;
        if (coneHessian) {
          double* H = cH + c_dm.conh_stride * c;
          double g[6];
          g[0] = 0;
          B2K_NOUNROLL for (int j = 1; j < dim; j++) g[j] = U[j] * fri[j - 1] / T;
          B2K_NOUNROLL for (int a = 0; a < dim; a++)
            B2K_NOUNROLL for (int b = 0; b < dim; b++) {
              const double da = (a == 0 ? mu : -mu * g[a]), db = (b == 0 ? mu : -mu * g[b]);
              double h = Dm * da * db;
              if (a > 0 && b > 0) {
                const double d2T = ((a == b ? fri[a - 1] * fri[a - 1] : 0.0) - g[a] * g[b]) / T;
                h += Dm * NmT * (-mu) * d2T;
              }
              H[a * dim + b] = h;
            }
        }
      }
    }
  }
  WSYNC();
  if (changed) *changed = __any_sync(e.mask, ch);
  return warpSum(e.mask, s);
}

}  // namespace b2k
