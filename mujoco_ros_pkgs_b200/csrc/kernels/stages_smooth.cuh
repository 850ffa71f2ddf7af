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

// v' = q v q* for a unit quaternion, without forming the matrix; exact for identity q / zero v
B2K_DI void rotQ(double* r, const double* v, const double* q) {
  const double tx = 2 * (q[2] * v[2] - q[3] * v[1]);
  const double ty = 2 * (q[3] * v[0] - q[1] * v[2]);
  const double tz = 2 * (q[1] * v[1] - q[2] * v[0]);
  const double x = v[0] + q[0] * tx + (q[2] * tz - q[3] * ty);
  const double y = v[1] + q[0] * ty + (q[3] * tx - q[1] * tz);
  const double z = v[2] + q[0] * tz + (q[1] * ty - q[2] * tx);
  r[0] = x; r[1] = y; r[2] = z;
}

// iterate the set bits of a multi-word mask: BODY is executed with `IDX` = bit index
#define FOR_MASK_BITS(IDX, MASKPTR, NWORD, BODY)                \
  B2K_NOUNROLL for (int _w = 0; _w < (NWORD); _w++) {                        \
    unsigned _bits = (MASKPTR)[_w];                             \
    while (_bits) {                                             \
      const int IDX = (_w << 5) + __ffs(_bits) - 1;             \
      _bits &= _bits - 1;                                       \
      BODY                                                      \
    }                                                           \
  }

// mj_kinematics.  The tree recursion is replaced by (A) body-local transforms for all bodies at once,
// (B) a pointer-jumping scan that composes them along the ancestor chains in ceil(log2(depth)) rounds,
// (C) one parallel pass for matrices, inertial / joint / geom / site frames.
__device__ __noinline__ void stage_kinematics(const Env e) {
  const DevModel& m = c_dm;
  double* qpos = e.D(B2MJ_F_QPOS);
  double* xpos = e.D(B2MJ_F_XPOS);
  double* xquat = e.D(B2MJ_F_XQUAT);
  double* xmat = e.D(B2MJ_F_XMAT);
  double* xipos = e.D(B2MJ_F_XIPOS);
  double* ximat = e.DG(B2MJ_F_XIMAT);
  double* xanchor = e.D(B2MJ_F_XANCHOR);
  double* xaxis = e.D(B2MJ_F_XAXIS);
  double* T0 = e.X(XF_SCRATCH);
  double* T1 = T0 + 7 * m.nbody;
  double* mocap_pos = m.nmocap ? e.D(B2MJ_F_MOCAP_POS) : nullptr;
  double* mocap_quat = m.nmocap ? e.D(B2MJ_F_MOCAP_QUAT) : nullptr;

  // (A) transform of every body relative to its parent; joint anchors / axes in the parent frame
  FORL(i, m.nbody) {
    double p[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0};
    if (i) {
      const int jntadr = m.body_jntadr[i], jntnum = m.body_jntnum[i];
      if (jntnum == 1 && m.jnt_type[jntadr] == B2MJ_JNT_FREE) {
        const int qa = m.jnt_qposadr[jntadr];
        normalize4(qpos + qa + 3);
        copy3(p, qpos + qa);
        copy4(q, qpos + qa + 3);
        copy3(xanchor + 3 * jntadr, p);
        copy3(xaxis + 3 * jntadr, m.jnt_axis + 3 * jntadr);
      } else {
        const int mid = m.body_mocapid[i];
        if (mid >= 0) {
          normalize4(mocap_quat + 4 * mid);
          copy3(p, mocap_pos + 3 * mid);
          copy4(q, mocap_quat + 4 * mid);
        } else {
          copy3(p, m.body_pos + 3 * i);
          copy4(q, m.body_quat + 4 * i);
        }
        B2K_NOUNROLL for (int j = 0; j < jntnum; j++) {
          const int jid = jntadr + j, jt = m.jnt_type[jid], qa = m.jnt_qposadr[jid];
          double anchor[3], axis[3];
          rotQ(axis, m.jnt_axis + 3 * jid, q);
          rotQ(anchor, m.jnt_pos + 3 * jid, q);
          addTo3(anchor, p);
          if (jt == B2MJ_JNT_SLIDE) {
            addToScl3(p, axis, qpos[qa] - m.qpos0[qa]);
          } else {
            double ql[4], qn[4], v[3];
            if (jt == B2MJ_JNT_BALL) { normalize4(qpos + qa); copy4(ql, qpos + qa); }
            else axisAngle2Quat(ql, m.jnt_axis + 3 * jid, qpos[qa] - m.qpos0[qa]);
            mulQuat(qn, q, ql);
            copy4(q, qn);
            rotQ(v, m.jnt_pos + 3 * jid, q);
            sub3(p, anchor, v);
          }
          copy3(xanchor + 3 * jid, anchor);
          copy3(xaxis + 3 * jid, axis);
        }
      }
    }
    copy3(T0 + 7 * i, p);
    copy4(T0 + 7 * i + 3, q);
  }
  WSYNC();
  // (B) pointer jumping: after round r, T[i] maps body i to the frame of its ancestor 2^(r+1) levels up
  double* src = T0;
  double* dst = T1;
  B2K_NOUNROLL for (int r = 0; r < m.njump; r++) {
    const int* jump = m.body_jump + r * m.nbody;
    FORL(i, m.nbody) {
      const int a = jump[i];
      const double* ti = src + 7 * i;
      if (a == 0) {
        for (int k = 0; k < 7; k++) dst[7 * i + k] = ti[k];
      } else {
        const double* ta = src + 7 * a;
        double v[3], q[4];
        rotQ(v, ti, ta + 3);
        mulQuat(q, ta + 3, ti + 3);
        dst[7 * i] = ta[0] + v[0]; dst[7 * i + 1] = ta[1] + v[1]; dst[7 * i + 2] = ta[2] + v[2];
        copy4(dst + 7 * i + 3, q);
      }
    }
    WSYNC();
    double* t = src; src = dst; dst = t;
  }
  // (C) world frames
  FORL(i, m.nbody) {
    double q[4], p[3], mat[9];
    copy3(p, src + 7 * i);
    copy4(q, src + 7 * i + 3);
    normalize4(q);
    quat2Mat(mat, q);
    copy3(xpos + 3 * i, p);
    copy4(xquat + 4 * i, q);
    for (int k = 0; k < 9; k++) xmat[9 * i + k] = mat[k];
    if (i) {
      double v[3], qi[4];
      rotVecMat(v, m.body_ipos + 3 * i, mat);
      add3(xipos + 3 * i, v, p);
      if (e.dump) {
        mulQuat(qi, q, m.body_iquat + 4 * i);
        quat2Mat(ximat + 9 * i, qi);
      }
    } else {
      zero3(xipos);
      if (e.dump) for (int k = 0; k < 9; k++) ximat[k] = (k % 4 == 0) ? 1.0 : 0.0;
    }
  }
  WSYNC();
  FORL(j, m.njnt) {
    if (m.jnt_type[j] == B2MJ_JNT_FREE) continue;  // already in the world frame
    const int pid = m.body_parentid[m.jnt_bodyid[j]];
    if (pid) {
      double a[3], x[3];
      rotVecMat(a, xanchor + 3 * j, xmat + 9 * pid);
      add3(xanchor + 3 * j, a, xpos + 3 * pid);
      rotVecMat(x, xaxis + 3 * j, xmat + 9 * pid);
      copy3(xaxis + 3 * j, x);
    }
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

// mj_comPos: subtree sums are gathers over the subtree bit masks (one lane per body)
__device__ __noinline__ void stage_comPos(const Env e) {
  const DevModel& m = c_dm;
  const double* xipos = e.D(B2MJ_F_XIPOS);
  const double* xquat = e.D(B2MJ_F_XQUAT);
  const double* xmat = e.D(B2MJ_F_XMAT);
  const double* xanchor = e.D(B2MJ_F_XANCHOR);
  const double* xaxis = e.D(B2MJ_F_XAXIS);
  double* com = e.D(B2MJ_F_SUBTREE_COM);
  double* cinert = e.D(B2MJ_F_CINERT);
  double* cdof = e.D(B2MJ_F_CDOF);
  FORL(b, m.nbody) {
    const double sm = m.body_subtreemass[b];
    if (sm < B2K_MINVAL) {
      copy3(com + 3 * b, xipos + 3 * b);
    } else {
      double s0 = 0, s1 = 0, s2 = 0;
      const unsigned* mask = m.body_submask + b * m.nbodyword;
      FOR_MASK_BITS(i, mask, m.nbodyword, {
        const double mi = m.body_mass[i];
        s0 += xipos[3 * i] * mi; s1 += xipos[3 * i + 1] * mi; s2 += xipos[3 * i + 2] * mi;
      })
      const double inv = 1.0 / fmax(B2K_MINVAL, sm);
      com[3 * b] = s0 * inv; com[3 * b + 1] = s1 * inv; com[3 * b + 2] = s2 * inv;
    }
  }
  WSYNC();
  FORL(i, m.nbody) {
    if (i == 0) { for (int k = 0; k < 10; k++) cinert[k] = 0; continue; }
    double off[3], qi[4], imat[9];
    sub3(off, xipos + 3 * i, com + 3 * m.body_rootid[i]);
    mulQuat(qi, xquat + 4 * i, m.body_iquat + 4 * i);   // inertial frame recomputed: ximat is not kept on chip
    quat2Mat(imat, qi);
    inertCom(cinert + 10 * i, m.body_inertia + 3 * i, imat, off, m.body_mass[i]);
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
__device__ __noinline__ void stage_tendon_transmission(const Env e) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  const double* qpos = e.D(B2MJ_F_QPOS);
  if (m.ntendon) {
    double* tl = e.D(B2MJ_F_TEN_LENGTH);
    double* tJ = e.D(B2MJ_F_TEN_J);
    // one lane per (tendon, dof) entry of ten_J; the lane of dof 0 also accumulates the length.  Fixed tendons: joint
    // coefficients (mjWRAP_JOINT = 1).  Spatial tendons: site - site segments (mjWRAP_SITE = 3) scaled by the pulley
    // divisor in force (mjWRAP_PULLEY = 2); the Jacobian entry is the segment direction dotted with the difference of
    // the two sites' point-Jacobian columns (mj_jacDifPair), formed from cdof and the chain bit masks.
    const double* sxpos = m.nsite ? e.D(B2MJ_F_SITE_XPOS) : nullptr;
    const double* cdof = e.D(B2MJ_F_CDOF);
    const double* com = e.D(B2MJ_F_SUBTREE_COM);
    FORL(item, m.ntendon * nv) {
      const int i = item / nv, k = item - i * nv;
      double len = 0, jk = 0, divisor = 1;
      const int w0 = m.tendon_adr[i], w1 = w0 + m.tendon_num[i];
      B2K_NOUNROLL for (int w = w0; w < w1; w++) {
        const int type = m.wrap_type[w];
        if (type == 1) {
          const int jid = m.wrap_objid[w];
          len += m.wrap_prm[w] * qpos[m.jnt_qposadr[jid]];
          if (m.jnt_dofadr[jid] == k) jk = m.wrap_prm[w];
        } else if (type == 2) {
          divisor = m.wrap_prm[w];
        } else if (type == 3 && w + 1 < w1 && m.wrap_type[w + 1] == 3) {
          const int s0 = m.wrap_objid[w], s1 = m.wrap_objid[w + 1], b0 = m.site_bodyid[s0], b1 = m.site_bodyid[s1];
          const double *p0 = sxpos + 3 * s0, *p1 = sxpos + 3 * s1;
          double dif[3];
          sub3(dif, p1, p0);
          const double seg = normalize3(dif);
          len += seg / divisor;
          if (b0 != b1) {
            const double* cd = cdof + 6 * k;
            double s = 0;
            const unsigned* m1 = m.body_dofmask + b1 * m.nmaskword;
            const unsigned* m0 = m.body_dofmask + b0 * m.nmaskword;
            if ((m1[k >> 5] >> (k & 31)) & 1u) {
              double off[3], cr[3];
              sub3(off, p1, com + 3 * m.body_rootid[b1]);
              cross(cr, cd, off);
              s += dif[0] * (cd[3] + cr[0]) + dif[1] * (cd[4] + cr[1]) + dif[2] * (cd[5] + cr[2]);
            }
            if ((m0[k >> 5] >> (k & 31)) & 1u) {
              double off[3], cr[3];
              sub3(off, p0, com + 3 * m.body_rootid[b0]);
              cross(cr, cd, off);
              s -= dif[0] * (cd[3] + cr[0]) + dif[1] * (cd[4] + cr[1]) + dif[2] * (cd[5] + cr[2]);
            }
            jk += s / divisor;
          }
        }
      }
      tJ[item] = jk;
      if (k == 0) tl[i] = len;
    }
    WSYNC();
  }
  if (m.nu) {
    double* al = e.D(B2MJ_F_ACTUATOR_LENGTH);
    double* am = e.D(B2MJ_F_ACTUATOR_MOMENT);
    const double* tl = m.ntendon ? e.D(B2MJ_F_TEN_LENGTH) : nullptr;
    const double* tJ = m.ntendon ? e.D(B2MJ_F_TEN_J) : nullptr;
    // one lane per (actuator, dof) entry of the moment matrix
    FORL(item, m.nu * nv) {
      const int i = item / nv, k = item - i * nv;
      const int id = m.actuator_trnid[2 * i];
      const double gear = m.actuator_gear[6 * i];
      double v;
      if (m.actuator_trntype[i] == B2MJ_TRN_TENDON) v = tJ[id * nv + k] * gear;
      else v = (k == m.jnt_dofadr[id]) ? gear : 0.0;
      am[item] = v;
    }
    FORL(i, m.nu) {
      const int id = m.actuator_trnid[2 * i];
      const double gear = m.actuator_gear[6 * i];
      al[i] = (m.actuator_trntype[i] == B2MJ_TRN_TENDON ? tl[id] : qpos[m.jnt_qposadr[id]]) * gear;
    }
    WSYNC();
  }
}

// In-place sparse L'DL factorisation (mj_factorI) of up to two matrices with the same sparsity at once:
// lanes 0-15 factor A, lanes 16-31 factor B (B = null: all 32 lanes on A).  Per pivot k the rank-1
// update of the ancestor rows runs one lane per (ancestor row, column) pair.
__device__ __noinline__ void factorLD2(const Env e, double* A, double* Bm, double* dinvA, double* sqrtinvA, double* dinvB) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  const int half = Bm ? (e.lane / (B2K_G / 2)) : 0, sub = Bm ? (e.lane % (B2K_G / 2)) : e.lane, stride = Bm ? B2K_G / 2 : B2K_G;
  double* LD = half ? Bm : A;
  const unsigned short* tri = triTable(e);
  for (int k = nv - 1; k >= 0; k--) {
    const int d = m.dof_nanc[k];
    if (d == 0) continue;
    const int Mkk = m.dof_Madr[k];
    const double inv = 1.0 / LD[Mkk];
    if (tri) {
      // entries (a, c < d - a) of the ancestor block through the pair table: (x, y <= x) -> a = d - 1 - x, c = y
      B2K_NOUNROLL for (int p = sub; p < d * (d + 1) / 2; p += stride) {
        const unsigned t = tri[p];
        const int a = d - 1 - (int)(t & 255u), c = (int)(t >> 8);
        LD[m.M_ancadr[Mkk + 1 + a] + c] -= LD[Mkk + 1 + a] * LD[Mkk + 1 + a + c] * inv;
      }
    } else {
      B2K_NOUNROLL for (int idx = sub; idx < d * d; idx += stride) {
        const int a = idx / d, c = idx - a * d;
        if (c < d - a) LD[m.M_ancadr[Mkk + 1 + a] + c] -= LD[Mkk + 1 + a] * LD[Mkk + 1 + a + c] * inv;
      }
    }
    WSYNC();
    B2K_NOUNROLL for (int c = sub; c < d; c += stride) LD[Mkk + 1 + c] *= inv;
    WSYNC();
  }
  B2K_NOUNROLL for (int i = sub; i < nv; i += stride) {
    const double Dv = LD[m.dof_Madr[i]];
    if (half) dinvB[i] = 1.0 / Dv;
    else {
      dinvA[i] = 1.0 / Dv;
      if (sqrtinvA) sqrtinvA[i] = 1.0 / sqrt(Dv);
    }
  }
  WSYNC();
}

// W = inv(L) (unit lower triangular, same tree sparsity as L) for up to two factorisations at once.
// Off-diagonals only, stored in the sparse layout of qM.  Rows depend on the rows of their ancestors,
// so dofs are processed by depth level, one lane per (dof, ancestor) entry.
__device__ __noinline__ void invL2(const Env e, const double* LA, double* WA, const double* LB, double* WB) {
  const DevModel& m = c_dm;
  const int half = LB ? (e.lane / (B2K_G / 2)) : 0, sub = LB ? (e.lane % (B2K_G / 2)) : e.lane, stride = LB ? B2K_G / 2 : B2K_G;
  const double* LD = half ? LB : LA;
  double* W = half ? WB : WA;
  B2K_NOUNROLL for (int l = 1; l < m.ndoflevel; l++) {
    const int adr0 = m.doflevel_adr[l], cnt = m.doflevel_adr[l + 1] - adr0;
    B2K_NOUNROLL for (int item = sub; item < cnt * l; item += stride) {
      const int i = m.doflevel_dof[adr0 + item / l], a = item % l;
      const int row = m.dof_Madr[i] + 1;
      double s = -LD[row + a];
      B2K_NOUNROLL for (int b = 0; b < a; b++) s -= LD[row + b] * W[m.M_ancadr[row + b] + (a - b)];
      W[row + a] = s;
    }
    WSYNC();
  }
}

// y = W' x (gather over the descendants of each dof); in place is NOT allowed
__device__ __forceinline__ void mulWT(const Env e, double* y, const double* W, const double* x) {
  const DevModel& m = c_dm;
  FORL(k, m.nv) {
    double s = x[k];
    B2K_UNROLL4 for (int p = m.dof_descadr[k]; p < m.dof_descadr[k + 1]; p++) s += W[m.dof_desc_adr[p]] * x[m.dof_desc_dof[p]];
    y[k] = s;
  }
}

// x <- inv(L'DL) x = W diag(dinv) W' x for one right-hand side, whole warp; tmp: nv scratch
__device__ __noinline__ void solveW_warp(const Env e, double* x, const double* W, const double* dinv, double* tmp) {
  const DevModel& m = c_dm;
  FORL(k, m.nv) {
    double s = x[k];
    B2K_UNROLL4 for (int p = m.dof_descadr[k]; p < m.dof_descadr[k + 1]; p++) s += W[m.dof_desc_adr[p]] * x[m.dof_desc_dof[p]];
    tmp[k] = s * dinv[k];
  }
  WSYNC();
  FORL(i, m.nv) {
    const int row = m.dof_Madr[i] + 1, d = m.dof_nanc[i];
    double s = tmp[i];
    B2K_UNROLL4 for (int a = 0; a < d; a++) s += W[row + a] * tmp[m.M_col[row + a]];
    x[i] = s;
  }
  WSYNC();
}

// x <- inv(L'DL) x, executed by ONE lane (callers distribute independent right-hand sides over lanes)
__device__ __forceinline__ void solveLD_lane(double* x, const double* LD, const double* diaginv) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  for (int i = nv - 1; i >= 0; i--) {
    const double t = x[i];
    if (t == 0) continue;
    int adr = m.dof_Madr[i] + 1;
    for (int j = m.dof_parentid[i]; j >= 0; j = m.dof_parentid[j]) x[j] -= LD[adr++] * t;
  }
  B2K_NOUNROLL for (int i = 0; i < nv; i++) x[i] *= diaginv[i];
  B2K_NOUNROLL for (int i = 0; i < nv; i++) {
    int adr = m.dof_Madr[i] + 1;
    double xi = x[i];
    for (int j = m.dof_parentid[i]; j >= 0; j = m.dof_parentid[j]) xi -= LD[adr++] * x[j];
    x[i] = xi;
  }
}

// res = M * vec: one lane per output row; row part over ancestors, column part over descendants
__device__ __noinline__ void mulM_warp(const Env e, double* res, const double* vec) {
  const DevModel& m = c_dm;
  const double* qM = e.D(B2MJ_F_QM);
  FORL(i, m.nv) {
    const int row = m.dof_Madr[i], d = m.dof_nanc[i];
    double s = qM[row] * vec[i];
    B2K_UNROLL4 for (int a = 0; a < d; a++) s += qM[row + 1 + a] * vec[m.M_col[row + 1 + a]];
    B2K_UNROLL4 for (int p = m.dof_descadr[i]; p < m.dof_descadr[i + 1]; p++) s += qM[m.dof_desc_adr[p]] * vec[m.dof_desc_dof[p]];
    res[i] = s;
  }
  WSYNC();
}

// ---- dense small-model path (nv <= 16) ----------------------------------------------------------
// In-place Gauss-Jordan inversion of up to two SPD n x n matrices (n <= NV <= 16 <= B2K_G): no index tables,
// no factor / triangular-solve chain; every later solve is one dense mat-vec.  Lane i owns row i of both
// matrices IN REGISTERS; per pivot the pivot row travels by shuffles, so the dependent chain per pivot is
// shuffle -> reciprocal -> FMA instead of a shared-memory round trip per element (the shared-memory version
// was 20% of the step's instructions).  Same operation order as the row-by-row elimination it replaces.
template <int NV>
__device__ __forceinline__ void invertSPD_regT(const Env e, double* Mx, int n, int i) {
  // i = row owned by this lane (0..15 within its 16-lane segment); Mx = the matrix this segment inverts
  const bool own = Mx != nullptr && i < n;
  double a[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) a[j] = (own && j < n) ? Mx[i * n + j] : (i == j ? 1.0 : 0.0);
  // Rolled pivot loop (the megakernel is instruction-fetch bound: straight-line code measured slower in the
  // desynchronised rollout).  The row is rotated left by one column per pivot, so the pivot column is always
  // register 0 and the finished column re-enters at register NV-1: static register indices in a rolled loop.
  // Rows / columns n..NV-1 are identity padding; their pivots leave the n x n block untouched (f == 0 exactly)
  // and complete the NV rotations that bring the columns back in place.
  B2K_NOUNROLL for (int k = 0; k < NV; k++) {
    const bool piv = i == k;
    const double ia = 1.0 / __shfl_sync(e.mask, a[0], k, 16);
    const double fa = a[0] * ia;
#pragma unroll
    for (int j = 1; j < NV; j++) {
      const double pj = __shfl_sync(e.mask, a[j], k, 16);
      a[j - 1] = piv ? pj * ia : a[j] - fa * pj;
    }
    a[NV - 1] = piv ? ia : -fa;
  }
#pragma unroll
  for (int j = 0; j < NV; j++)
    if (own && j < n) Mx[i * n + j] = a[j];
}

// Inverts A and (optionally) Bm in place.  One env per warp (B2K_G == 32): the two 16-lane segments invert the
// two matrices side by side; two envs per warp (B2K_G == 16): one after the other.
__device__ __noinline__ void invertSPD2(const Env e, double* A, double* Bm, int n) {
  const int i = e.lane & 15;
#if B2K_G == 32
  double* Mx = (e.lane >> 4) ? Bm : A;
  if (n <= 8) invertSPD_regT<8>(e, Mx, n, i);
  else if (n <= 12) invertSPD_regT<12>(e, Mx, n, i);
  else invertSPD_regT<16>(e, Mx, n, i);
#else
  B2K_NOUNROLL for (int mat = 0; mat < 2; mat++) {
    double* Mx = mat ? Bm : A;
    if (!Mx) continue;
    if (n <= 8) invertSPD_regT<8>(e, Mx, n, i);
    else if (n <= 12) invertSPD_regT<12>(e, Mx, n, i);
    else invertSPD_regT<16>(e, Mx, n, i);
  }
#endif
  WSYNC();
}

// out = Ainv * in for a dense nv x nv matrix (out != in)
__device__ __forceinline__ void mulDense_warp(const Env e, double* out, const double* Ainv, const double* in, int n) {
  FORL(i, n) {
    double s = 0;
    B2K_NOUNROLL for (int j = 0; j < n; j++) s += Ainv[i * n + j] * in[j];
    out[i] = s;
  }
  WSYNC();
}

// mj_crb + mj_factorM, plus the Euler-damping matrix qH = qM + h diag(damping) factored alongside
__device__ __noinline__ void stage_crb_factor(const Env e, bool want_ld) {
  const DevModel& m = c_dm;
  const double* cinert = e.D(B2MJ_F_CINERT);
  const double* cdof = e.D(B2MJ_F_CDOF);
  double* crb = e.D(B2MJ_F_CRB);
  double* qM = e.D(B2MJ_F_QM);
  double* qLD = e.DG(B2MJ_F_QLD);
  double* buf = e.X(XF_SCRATCH) + 10 * 0;
  const bool damped = m.any_damping && !(m.opt.disableflags & B2MJ_DSBL_EULERDAMP) && m.opt.integrator == B2MJ_INT_EULER;
  const bool dense = m.dense_small;
  double* qH = (damped && (!dense || want_ld)) ? e.XG(XF_QH) : nullptr;
  double* Minv = dense ? e.X(XF_MINV) : nullptr;
  double* Hinv = (dense && damped) ? e.X(XF_HINV) : nullptr;
  const int nvv = m.nv;
  FORL(b, m.nbody) {
    double s[10];
    for (int k = 0; k < 10; k++) s[k] = 0;
    if (b) {
      const unsigned* mask = m.body_submask + b * m.nbodyword;
      FOR_MASK_BITS(i, mask, m.nbodyword, { for (int k = 0; k < 10; k++) s[k] += cinert[10 * i + k]; })
    }
    for (int k = 0; k < 10; k++) crb[10 * b + k] = s[k];
  }
  WSYNC();
  FORL(k, m.nv) mulInertVec(buf + 6 * k, crb + 10 * m.dof_bodyid[k], cdof + 6 * k);
  WSYNC();
  FORL(t, m.nM) {
    const int i = m.M_row[t], j = m.M_col[t];
    double v = dot6(cdof + 6 * j, buf + 6 * i);
    double hv = v;
    if (i == j) { v += m.dof_armature[i]; hv = v + m.opt.timestep * m.dof_damping[i]; }
    qM[t] = v;
    if (!dense || want_ld) qLD[t] = v;
    if (qH) qH[t] = hv;
    if (dense) {
      Minv[i * nvv + j] = v;
      Minv[j * nvv + i] = v;
      if (Hinv) { Hinv[i * nvv + j] = hv; Hinv[j * nvv + i] = hv; }
    }
  }
  WSYNC();
  if (dense) {
    // dense inverses; the sparse L'DL factor is only produced when someone will read it (arena dump)
    const int nv = m.nv;
    // entries between dofs of different branches are structural zeros of M
    FORL(item, nv * nv) {
      const int i = item / nv, j = item - i * nv;
      const int hi = i > j ? i : j, lo = i > j ? j : i;
      const unsigned* mask = m.body_dofmask + m.dof_bodyid[hi] * m.nmaskword;
      if (!((mask[lo >> 5] >> (lo & 31)) & 1u)) { Minv[item] = 0; if (Hinv) Hinv[item] = 0; }
    }
    WSYNC();
    invertSPD2(e, Minv, Hinv, nv);
    if (!want_ld) return;
  }
  factorLD2(e, qLD, qH, e.DG(B2MJ_F_QLDIAGINV), e.DG(B2MJ_F_QLDIAGSQRTINV), qH ? e.XG(XF_QHDIAGINV) : nullptr);
  invL2(e, qLD, e.XG(XF_QW), qH, qH ? e.XG(XF_QHW) : nullptr);
}

// out = inv(M) in  /  out = inv(M + h diag(damping)) in   (out != in)
__device__ __forceinline__ void solveM_warp(const Env e, double* out, const double* in) {
  const DevModel& m = c_dm;
  if (m.dense_small) { mulDense_warp(e, out, e.X(XF_MINV), in, m.nv); return; }
  FORL(i, m.nv) out[i] = in[i];
  WSYNC();
  solveW_warp(e, out, e.XG(XF_QW), e.DG(B2MJ_F_QLDIAGINV), e.X(XF_VEC0));
}
__device__ __forceinline__ void solveH_warp(const Env e, double* out, const double* in) {
  const DevModel& m = c_dm;
  if (m.dense_small) { mulDense_warp(e, out, e.X(XF_HINV), in, m.nv); return; }
  FORL(i, m.nv) out[i] = in[i];
  WSYNC();
  solveW_warp(e, out, e.XG(XF_QHW), e.XG(XF_QHDIAGINV), e.X(XF_VEC0));
}

__device__ __forceinline__ void mulDofVec(double* res, const double* dof, const double* vec, int n) {
  for (int k = 0; k < 6; k++) res[k] = 0;
  B2K_NOUNROLL for (int i = 0; i < n; i++)
    for (int k = 0; k < 6; k++) res[k] += dof[6 * i + k] * vec[i];
}

// mj_comVel: cvel of a body is the sum of cdof*qvel over the dofs of its chain (all cdof share the
// subtree-com frame), so every body / dof is computed independently from the chain bit masks.
__device__ __noinline__ void stage_comVel(const Env e) {
  const DevModel& m = c_dm;
  const double* cdof = e.D(B2MJ_F_CDOF);
  const double* qvel = e.D(B2MJ_F_QVEL);
  double* cvelA = e.D(B2MJ_F_CVEL);
  double* cdof_dot = e.D(B2MJ_F_CDOF_DOT);
  FORL(b, m.nbody) {
    double s[6] = {0, 0, 0, 0, 0, 0};
    const unsigned* mask = m.body_dofmask + b * m.nmaskword;
    FOR_MASK_BITS(k, mask, m.nmaskword, { const double v = qvel[k]; for (int c = 0; c < 6; c++) s[c] += cdof[6 * k + c] * v; })
    for (int c = 0; c < 6; c++) cvelA[6 * b + c] = s[c];
  }
  FORL(k, m.nv) {
    const int jid = m.dof_jntid[k];
    if (m.jnt_type[jid] == B2MJ_JNT_FREE && k < m.jnt_dofadr[jid] + 3) {
      for (int c = 0; c < 6; c++) cdof_dot[6 * k + c] = 0;
      continue;
    }
    double s[6] = {0, 0, 0, 0, 0, 0};
    const unsigned* mask = m.dof_premask + k * m.nmaskword;
    FOR_MASK_BITS(j, mask, m.nmaskword, { const double v = qvel[j]; for (int c = 0; c < 6; c++) s[c] += cdof[6 * j + c] * v; })
    crossMotion(cdof_dot + 6 * k, s, cdof + 6 * k);
  }
  WSYNC();
}

// Jacobian-transpose application of a force/torque at a point of a body: one lane per dof.
// qfrc[k] += jacp[:,k].force + jacr[:,k].torque for dofs on the body's chain.
__device__ __noinline__ void applyFT_warp(const Env e, const double* force, const double* torque, const double* point, int body,
                             double* qfrc) {
  const DevModel& m = c_dm;
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
__device__ __noinline__ void stage_passive(const Env e) {
  const DevModel& m = c_dm;
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
    B2K_NOUNROLL for (int i = 0; i < m.ntendon; i++) {
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
    B2K_NOUNROLL for (int i = 1; i < m.nbody; i++) {
      const double gc = m.body_gravcomp[i];
      if (gc == 0) continue;
      double force[3], torque[3] = {0, 0, 0};
      scl3(force, m.env_gravity + 0, -(m.body_mass[i] * gc));
      applyFT_warp(e, force, torque, xipos + 3 * i, i, qp);
    }
  }
  // body-level viscosity, lift and drag (mj_inertiaBoxFluidModel; option density / viscosity / wind, which the reference
  // exposes at mujoco_ros/src/viewer.cpp:597-600): one lane per body forms the world-frame wrench at the body's centre of
  // mass, the warp then projects each wrench onto the body's dof chain
  if (m.opt.viscosity > 0 || m.opt.density > 0) {
    const double* xipos = e.D(B2MJ_F_XIPOS);
    const double* xquat = e.D(B2MJ_F_XQUAT);
    const double* cvel = e.D(B2MJ_F_CVEL);
    const double* com = e.D(B2MJ_F_SUBTREE_COM);
    double* buf = e.X(XF_SCRATCH);  // [nbody][6]: torque, force
    const double rho = m.opt.density, mu = m.opt.viscosity, kPi = 3.14159265358979323846;
    FORL(i, m.nbody) {
      double lfrc[6] = {0, 0, 0, 0, 0, 0}, bfrc[6] = {0, 0, 0, 0, 0, 0};
      const double mass = m.body_mass[i];
      if (i && mass >= B2K_MINVAL) {
        const double* inertia = m.body_inertia + 3 * i;
        double box[3], qi[4], ximat[9], lvel[6], wind[6] = {0, 0, 0, 0, 0, 0}, lwind[6];
        box[0] = sqrt(fmax(B2K_MINVAL, inertia[1] + inertia[2] - inertia[0]) / mass * 6.0);
        box[1] = sqrt(fmax(B2K_MINVAL, inertia[0] + inertia[2] - inertia[1]) / mass * 6.0);
        box[2] = sqrt(fmax(B2K_MINVAL, inertia[0] + inertia[1] - inertia[2]) / mass * 6.0);
        mulQuat(qi, xquat + 4 * i, m.body_iquat + 4 * i);
        quat2Mat(ximat, qi);
        const double* c0 = com + 3 * m.body_rootid[i];
        transformSpatial(lvel, cvel + 6 * i, 0, xipos + 3 * i, c0, ximat);
        wind[3] = m.opt.wind[0]; wind[4] = m.opt.wind[1]; wind[5] = m.opt.wind[2];
        transformSpatial(lwind, wind, 0, xipos + 3 * i, c0, ximat);
        for (int k = 0; k < 3; k++) lvel[3 + k] -= lwind[3 + k];
        if (mu > 0) {
          const double diam = (box[0] + box[1] + box[2]) / 3.0;
          scl3(lfrc, lvel, -kPi * diam * diam * diam * mu);
          scl3(lfrc + 3, lvel + 3, -3.0 * kPi * diam * mu);
        }
        if (rho > 0) {
          lfrc[3] -= 0.5 * rho * box[1] * box[2] * fabs(lvel[3]) * lvel[3];
          lfrc[4] -= 0.5 * rho * box[0] * box[2] * fabs(lvel[4]) * lvel[4];
          lfrc[5] -= 0.5 * rho * box[0] * box[1] * fabs(lvel[5]) * lvel[5];
          const double b0 = box[0] * box[0] * box[0] * box[0], b1 = box[1] * box[1] * box[1] * box[1],
                       b2 = box[2] * box[2] * box[2] * box[2];
          lfrc[0] -= rho * box[0] * (b1 + b2) * fabs(lvel[0]) * lvel[0] / 64.0;
          lfrc[1] -= rho * box[1] * (b0 + b2) * fabs(lvel[1]) * lvel[1] / 64.0;
          lfrc[2] -= rho * box[2] * (b0 + b1) * fabs(lvel[2]) * lvel[2] / 64.0;
        }
        rotVecMat(bfrc, lfrc, ximat);
        rotVecMat(bfrc + 3, lfrc + 3, ximat);
      }
      for (int k = 0; k < 6; k++) buf[6 * i + k] = bfrc[k];
    }
    WSYNC();
    B2K_NOUNROLL for (int i = 1; i < m.nbody; i++) {
      if (m.body_mass[i] < B2K_MINVAL) continue;
      applyFT_warp(e, buf + 6 * i + 3, buf + 6 * i, xipos + 3 * i, i, qp);
    }
  }
}

// mj_rne(flg_acc = 0): bias forces, without the tree recursions: cacc of a body is the sum of
// cdof_dot*qvel over its chain, the backward force accumulation is folded into the projection
// bias[k] = cdof[k] . sum_{i in subtree(body(k))} f_i.
__device__ __noinline__ void stage_rne_bias(const Env e) {
  const DevModel& m = c_dm;
  const double* cdof = e.D(B2MJ_F_CDOF);
  const double* cdof_dot = e.D(B2MJ_F_CDOF_DOT);
  const double* cvel = e.D(B2MJ_F_CVEL);
  const double* cinert = e.D(B2MJ_F_CINERT);
  const double* qvel = e.D(B2MJ_F_QVEL);
  double* fb = e.X(XF_SCRATCH);
  double* bias = e.D(B2MJ_F_QFRC_BIAS);
  const bool grav = !(m.opt.disableflags & B2MJ_DSBL_GRAVITY);
  FORL(b, m.nbody) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (grav) { acc[3] = -m.env_gravity[0]; acc[4] = -m.env_gravity[1]; acc[5] = -m.env_gravity[2]; }
    const unsigned* mask = m.body_dofmask + b * m.nmaskword;
    FOR_MASK_BITS(k, mask, m.nmaskword, { const double v = qvel[k]; for (int c = 0; c < 6; c++) acc[c] += cdof_dot[6 * k + c] * v; })
    double f[6], tmp[6], tmp1[6];
    if (b) {
      mulInertVec(f, cinert + 10 * b, acc);
      mulInertVec(tmp, cinert + 10 * b, cvel + 6 * b);
      crossForce(tmp1, cvel + 6 * b, tmp);
      for (int c = 0; c < 6; c++) fb[6 * b + c] = f[c] + tmp1[c];
    } else {
      for (int c = 0; c < 6; c++) fb[c] = 0;
    }
  }
  WSYNC();
  FORL(k, m.nv) {
    double s[6] = {0, 0, 0, 0, 0, 0};
    const unsigned* mask = m.body_submask + m.dof_bodyid[k] * m.nbodyword;
    FOR_MASK_BITS(i, mask, m.nbodyword, { for (int c = 0; c < 6; c++) s[c] += fb[6 * i + c]; })
    bias[k] = dot6(cdof + 6 * k, s);
  }
  WSYNC();
}

// tendon / actuator velocities (head of mj_fwdVelocity)
__device__ __noinline__ void stage_velocity_head(const Env e) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  const double* qvel = e.D(B2MJ_F_QVEL);
  if (m.ntendon) {
    const double* tJ = e.D(B2MJ_F_TEN_J);
    double* tv = e.D(B2MJ_F_TEN_VELOCITY);
    FORL(i, m.ntendon) {
      double s = 0;
      B2K_NOUNROLL for (int k = 0; k < nv; k++) s += tJ[i * nv + k] * qvel[k];
      tv[i] = s;
    }
  }
  if (m.nu) {
    const double* am = e.D(B2MJ_F_ACTUATOR_MOMENT);
    double* av = e.D(B2MJ_F_ACTUATOR_VELOCITY);
    FORL(i, m.nu) {
      double s = 0;
      B2K_NOUNROLL for (int k = 0; k < nv; k++) s += am[i * nv + k] * qvel[k];
      av[i] = s;
    }
  }
  WSYNC();
}

// mj_fwdActuation
__device__ __noinline__ void stage_actuation(const Env e, int* warning) {
  const DevModel& m = c_dm;
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
  if (__any_sync(e.mask, badc)) {
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
    B2K_NOUNROLL for (int i = 0; i < nu; i++) s += am[i * nv + k] * af[i];
    qa[k] = s;
  }
  WSYNC();
}

// mj_fwdAcceleration
__device__ __noinline__ void stage_acceleration(const Env e, const double* xfrc) {
  const DevModel& m = c_dm;
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
    B2K_NOUNROLL for (int i = 1; i < m.nbody; i++) {
      const double* x = xfrc + 6 * i;
      if (x[0] == 0 && x[1] == 0 && x[2] == 0 && x[3] == 0 && x[4] == 0 && x[5] == 0) continue;
      applyFT_warp(e, x, x + 3, xipos + 3 * i, i, qs);
    }
  }
  solveM_warp(e, qas, qs);
}

// mj_integratePos for the joints handled by this lane
__device__ void integratePos_warp(const Env e, double* qpos, const double* qvel, double dt) {
  const DevModel& m = c_dm;
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
__device__ void advance_warp(const Env e, const double* act_dot, const double* qacc, const double* qvel_for_pos) {
  const DevModel& m = c_dm;
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
__device__ __noinline__ void stage_euler(const Env e) {
  const DevModel& m = c_dm;
  const int nv = m.nv;
  const double* act_dot = m.na ? e.D(B2MJ_F_ACT_DOT) : nullptr;
  if (!m.any_damping || (m.opt.disableflags & B2MJ_DSBL_EULERDAMP)) {
    advance_warp(e, act_dot, e.D(B2MJ_F_QACC), nullptr);
    return;
  }
  double* rhs = e.X(XF_VEC1);
  double* acc = e.X(XF_VEC2);
  const double* qs = e.D(B2MJ_F_QFRC_SMOOTH);
  const double* qc = e.D(B2MJ_F_QFRC_CONSTRAINT);
  FORL(i, nv) rhs[i] = qs[i] + qc[i];
  WSYNC();
  // qH = qM + h diag(damping) was factored / inverted next to qM in stage_crb_factor
  solveH_warp(e, acc, rhs);
  advance_warp(e, act_dot, acc, nullptr);
}

}  // namespace b2k
