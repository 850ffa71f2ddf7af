// orc_constraint.cpp — constraint construction of the CPU oracle.
// TEST INFRASTRUCTURE ONLY (see orc_math.h).  Restates MuJoCo 2.3.7 engine_core_constraint.c:
// mj_makeConstraint (equality, friction loss, limits, contacts; dense Jacobian), mj_makeImpedance,
// mj_referenceConstraint, mj_projectConstraint, mj_constraintUpdate — stages inside the `mj_step`
// call at reference mujoco_env.cpp:498; formulas in SURVEY Appendix A / A2.
#include <algorithm>
#include <cmath>
#include <vector>

#include "orc_math.h"
#include "orc_types.h"

namespace orc {

static bool addConstraint(const b2mjModel* m, OrcData* d, const double* jac, const double* pos, const double* margin,
                          double frictionloss, int size, int type, int id) {
  const int nv = m->nv;
  if (d->nefc() + size > m->njmax) {
    d->warning[B2MJ_WARN_CNSTRFULL]++;
    return false;
  }
  const int base = d->nefc();
  for (int i = 0; i < size; i++) {
    copy(d->efc_J + (base + i) * nv, jac + i * nv, nv);
    d->efc_pos[base + i] = pos ? pos[i] : 0;
    d->efc_margin[base + i] = margin ? margin[i] : 0;
    d->efc_frictionloss[base + i] = frictionloss;
    d->efc_type[base + i] = type;
    d->efc_id[base + i] = id;
  }
  d->nefc() += size;
  return true;
}

static void instantiateEquality(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv;
  if ((m->opt.disableflags & B2MJ_DSBL_EQUALITY) || !m->neq) return;
  std::vector<double> jac(6 * nv), jp0(3 * nv), jp1(3 * nv), jr0(3 * nv), jr1(3 * nv);
  for (int i = 0; i < m->neq; i++) {
    if (!m->eq_active[i]) continue;
    const double* data = m->eq_data + B2MJ_NEQDATA * i;
    const int id0 = m->eq_obj1id[i], id1 = m->eq_obj2id[i];
    double cpos[6] = {0, 0, 0, 0, 0, 0};
    switch (m->eq_type[i]) {
      case B2MJ_EQ_CONNECT: {
        double pos0[3], pos1[3];
        rotVecMat(pos0, data, d->xmat + 9 * id0); addTo3(pos0, d->xpos + 3 * id0);
        rotVecMat(pos1, data + 3, d->xmat + 9 * id1); addTo3(pos1, d->xpos + 3 * id1);
        sub3(cpos, pos0, pos1);
        orc::jac(m, d, jp0.data(), nullptr, pos0, id0);
        orc::jac(m, d, jp1.data(), nullptr, pos1, id1);
        for (int k = 0; k < 3 * nv; k++) jac[k] = jp0[k] - jp1[k];
        addConstraint(m, d, jac.data(), cpos, nullptr, 0, 3, B2MJ_CNSTR_EQUALITY, i);
        break;
      }
      case B2MJ_EQ_WELD: {
        double pos0[3], pos1[3];
        // body1 carries the relative-pose offset, body2 the anchor
        rotVecMat(pos0, data + 3, d->xmat + 9 * id0); addTo3(pos0, d->xpos + 3 * id0);
        rotVecMat(pos1, data, d->xmat + 9 * id1); addTo3(pos1, d->xpos + 3 * id1);
        sub3(cpos, pos0, pos1);
        orc::jac(m, d, jp0.data(), jr0.data(), pos0, id0);
        orc::jac(m, d, jp1.data(), jr1.data(), pos1, id1);
        for (int k = 0; k < 3 * nv; k++) { jac[k] = jp0[k] - jp1[k]; jr0[k] -= jr1[k]; }
        const double torquescale = data[10];
        double quat[4], quat1[4], quat2[4], quat3[4];
        mulQuat(quat, d->xquat + 4 * id0, data + 6);  // q0 * relpose
        negQuat(quat1, d->xquat + 4 * id1);           // neg(q1)
        mulQuat(quat2, quat1, quat);                  // neg(q1) * q0 * relpose
        scl3(cpos + 3, quat2 + 1, torquescale);
        // rotational rows: 0.5 * neg(q1) * (jacr0 - jacr1) * q0 * relpose, axis components
        for (int j = 0; j < nv; j++) {
          double axis[3] = {jr0[j], jr0[nv + j], jr0[2 * nv + j]};
          mulQuatAxis(quat2, quat1, axis);
          mulQuat(quat3, quat2, quat);
          jac[3 * nv + j] = 0.5 * quat3[1] * torquescale;
          jac[4 * nv + j] = 0.5 * quat3[2] * torquescale;
          jac[5 * nv + j] = 0.5 * quat3[3] * torquescale;
        }
        addConstraint(m, d, jac.data(), cpos, nullptr, 0, 6, B2MJ_CNSTR_EQUALITY, i);
        break;
      }
      case B2MJ_EQ_JOINT:
      case B2MJ_EQ_TENDON: {
        const bool isj = m->eq_type[i] == B2MJ_EQ_JOINT;
        double pos0, pos1 = 0, ref0, ref1 = 0;
        zero(jac.data(), nv);
        if (isj) {
          pos0 = d->qpos[m->jnt_qposadr[id0]]; ref0 = m->qpos0[m->jnt_qposadr[id0]];
          if (id1 >= 0) { pos1 = d->qpos[m->jnt_qposadr[id1]]; ref1 = m->qpos0[m->jnt_qposadr[id1]]; }
        } else {
          pos0 = d->ten_length[id0]; ref0 = m->tendon_length0[id0];
          if (id1 >= 0) { pos1 = d->ten_length[id1]; ref1 = m->tendon_length0[id1]; }
        }
        if (id1 >= 0) {
          const double dif = pos1 - ref1;
          const double dif2 = dif * dif, dif3 = dif2 * dif, dif4 = dif3 * dif;
          cpos[0] = pos0 - ref0 - data[0] - (data[1] * dif + data[2] * dif2 + data[3] * dif3 + data[4] * dif4);
          const double deriv = data[1] + 2 * data[2] * dif + 3 * data[3] * dif2 + 4 * data[4] * dif3;
          if (isj) {
            jac[m->jnt_dofadr[id1]] = -deriv;   // written first: joint1 wins if both map to one dof
            jac[m->jnt_dofadr[id0]] = 1;
          } else {
            for (int k = 0; k < nv; k++) jac[k] = d->ten_J[id0 * nv + k] - deriv * d->ten_J[id1 * nv + k];
          }
        } else {
          cpos[0] = pos0 - ref0 - data[0];
          if (isj) jac[m->jnt_dofadr[id0]] = 1;
          else copy(jac.data(), d->ten_J + id0 * nv, nv);
        }
        addConstraint(m, d, jac.data(), cpos, nullptr, 0, 1, B2MJ_CNSTR_EQUALITY, i);
        break;
      }
    }
  }
}

static void instantiateFriction(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv;
  if (m->opt.disableflags & B2MJ_DSBL_FRICTIONLOSS) return;
  std::vector<double> jac(nv);
  for (int i = 0; i < nv; i++) {
    if (m->dof_frictionloss[i] <= 0) continue;
    zero(jac.data(), nv);
    jac[i] = 1;
    addConstraint(m, d, jac.data(), nullptr, nullptr, m->dof_frictionloss[i], 1, B2MJ_CNSTR_FRICTION_DOF, i);
  }
  for (int i = 0; i < m->ntendon; i++) {
    if (m->tendon_frictionloss[i] <= 0) continue;
    addConstraint(m, d, d->ten_J + i * nv, nullptr, nullptr, m->tendon_frictionloss[i], 1, B2MJ_CNSTR_FRICTION_TENDON, i);
  }
}

static void instantiateLimit(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv;
  if (m->opt.disableflags & B2MJ_DSBL_LIMIT) return;
  std::vector<double> jac(nv);
  for (int i = 0; i < m->njnt; i++) {
    if (!m->jnt_limited[i]) continue;
    const double margin = m->jnt_margin[i];
    const int type = m->jnt_type[i];
    if (type == B2MJ_JNT_SLIDE || type == B2MJ_JNT_HINGE) {
      const double value = d->qpos[m->jnt_qposadr[i]];
      for (int side = -1; side <= 1; side += 2) {
        const double dist = side * (m->jnt_range[2 * i + (side + 1) / 2] - value);
        if (dist < margin) {
          zero(jac.data(), nv);
          jac[m->jnt_dofadr[i]] = -(double)side;
          addConstraint(m, d, jac.data(), &dist, &margin, 0, 1, B2MJ_CNSTR_LIMIT_JOINT, i);
        }
      }
    } else if (type == B2MJ_JNT_BALL) {
      double quat[4], angleAxis[3];
      copy4(quat, d->qpos + m->jnt_qposadr[i]);
      normalize4(quat);
      quat2Vel(angleAxis, quat, 1);
      const double value = normalize3(angleAxis);
      const double dist = std::fmax(m->jnt_range[2 * i], m->jnt_range[2 * i + 1]) - value;
      if (dist < margin) {
        zero(jac.data(), nv);
        for (int k = 0; k < 3; k++) jac[m->jnt_dofadr[i] + k] = -angleAxis[k];
        addConstraint(m, d, jac.data(), &dist, &margin, 0, 1, B2MJ_CNSTR_LIMIT_JOINT, i);
      }
    }
  }
  for (int i = 0; i < m->ntendon; i++) {
    if (!m->tendon_limited[i]) continue;
    const double value = d->ten_length[i], margin = m->tendon_margin[i];
    for (int side = -1; side <= 1; side += 2) {
      const double dist = side * (m->tendon_range[2 * i + (side + 1) / 2] - value);
      if (dist < margin) {
        for (int k = 0; k < nv; k++) jac[k] = -side * d->ten_J[i * nv + k];
        addConstraint(m, d, jac.data(), &dist, &margin, 0, 1, B2MJ_CNSTR_LIMIT_TENDON, i);
      }
    }
  }
}

static void instantiateContact(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv;
  if ((m->opt.disableflags & B2MJ_DSBL_CONTACT) || d->ncon() == 0 || nv == 0) return;
  const bool pyramid = m->opt.cone == B2MJ_CONE_PYRAMIDAL;
  std::vector<double> jp1(3 * nv), jp2(3 * nv), jr1(3 * nv), jr2(3 * nv), jd(6 * nv), jac(6 * nv), jrow(2 * nv);
  for (int c = 0; c < d->ncon(); c++) {
    if (d->contact_exclude[c]) continue;
    const int dim = d->contact_dim[c];
    const int b1 = m->geom_bodyid[d->contact_geom1[c]], b2 = m->geom_bodyid[d->contact_geom2[c]];
    const double* pos = d->contact_pos + 3 * c;
    const double* frame = d->contact_frame + 9 * c;
    orc::jac(m, d, jp1.data(), jr1.data(), pos, b1);
    orc::jac(m, d, jp2.data(), jr2.data(), pos, b2);
    for (int k = 0; k < 3 * nv; k++) { jd[k] = jp2[k] - jp1[k]; jd[3 * nv + k] = jr2[k] - jr1[k]; }
    // rotate into the contact frame: rows 0..2 translational, 3..5 rotational
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < nv; k++) {
        jac[r * nv + k] = frame[3 * r] * jd[k] + frame[3 * r + 1] * jd[nv + k] + frame[3 * r + 2] * jd[2 * nv + k];
        if (dim > 3)
          jac[(3 + r) * nv + k] = frame[3 * r] * jd[3 * nv + k] + frame[3 * r + 1] * jd[4 * nv + k] + frame[3 * r + 2] * jd[5 * nv + k];
      }
    const double dist = d->contact_dist[c], inc = d->contact_includemargin[c];
    const int adr = d->nefc();
    bool ok = true;
    if (dim == 1) {
      ok = addConstraint(m, d, jac.data(), &dist, &inc, 0, 1, B2MJ_CNSTR_CONTACT_FRICTIONLESS, c);
    } else if (pyramid) {
      if (d->nefc() + 2 * (dim - 1) > m->njmax) { d->warning[B2MJ_WARN_CNSTRFULL]++; ok = false; }
      const double cpos[2] = {dist, dist}, cmargin[2] = {inc, inc};
      const double* fri = d->contact_friction + 5 * c;
      for (int k = 1; k < dim && ok; k++) {
        for (int j = 0; j < nv; j++) {
          jrow[j] = jac[j] + fri[k - 1] * jac[k * nv + j];
          jrow[nv + j] = jac[j] - fri[k - 1] * jac[k * nv + j];
        }
        addConstraint(m, d, jrow.data(), cpos, cmargin, 0, 2, B2MJ_CNSTR_CONTACT_PYRAMIDAL, c);
      }
    } else {
      double cpos[6] = {dist, 0, 0, 0, 0, 0}, cmargin[6] = {inc, 0, 0, 0, 0, 0};
      ok = addConstraint(m, d, jac.data(), cpos, cmargin, 0, dim, B2MJ_CNSTR_CONTACT_ELLIPTIC, c);
    }
    d->contact_efc_address[c] = ok ? adr : -1;
    if (!ok) break;  // row capacity reached: this and all later contacts are dropped (CNSTRFULL)
  }
}

static void diagApprox(const b2mjModel* m, OrcData* d) {
  for (int i = 0; i < d->nefc(); i++) {
    const int id = d->efc_id[i];
    switch (d->efc_type[i]) {
      case B2MJ_CNSTR_EQUALITY: {
        const int b1 = m->eq_obj1id[id], b2 = m->eq_obj2id[id];
        switch (m->eq_type[id]) {
          case B2MJ_EQ_CONNECT:
            for (int k = 0; k < 3; k++) d->efc_diagApprox[i + k] = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
            i += 2;
            break;
          case B2MJ_EQ_WELD:
            for (int k = 0; k < 3; k++) {
              d->efc_diagApprox[i + k] = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
              d->efc_diagApprox[i + 3 + k] = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
            }
            i += 5;
            break;
          case B2MJ_EQ_JOINT:
            d->efc_diagApprox[i] = m->dof_invweight0[m->jnt_dofadr[b1]] + (b2 >= 0 ? m->dof_invweight0[m->jnt_dofadr[b2]] : 0);
            break;
          case B2MJ_EQ_TENDON:
            d->efc_diagApprox[i] = m->tendon_invweight0[b1] + (b2 >= 0 ? m->tendon_invweight0[b2] : 0);
            break;
        }
        break;
      }
      case B2MJ_CNSTR_FRICTION_DOF: d->efc_diagApprox[i] = m->dof_invweight0[id]; break;
      case B2MJ_CNSTR_LIMIT_JOINT: d->efc_diagApprox[i] = m->dof_invweight0[m->jnt_dofadr[id]]; break;
      case B2MJ_CNSTR_FRICTION_TENDON:
      case B2MJ_CNSTR_LIMIT_TENDON: d->efc_diagApprox[i] = m->tendon_invweight0[id]; break;
      case B2MJ_CNSTR_CONTACT_FRICTIONLESS:
      case B2MJ_CNSTR_CONTACT_PYRAMIDAL:
      case B2MJ_CNSTR_CONTACT_ELLIPTIC: {
        const int b1 = m->geom_bodyid[d->contact_geom1[id]], b2 = m->geom_bodyid[d->contact_geom2[id]];
        const int dim = d->contact_dim[id];
        const double tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
        const double rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
        if (d->efc_type[i] == B2MJ_CNSTR_CONTACT_FRICTIONLESS) {
          d->efc_diagApprox[i] = tran;
        } else if (d->efc_type[i] == B2MJ_CNSTR_CONTACT_PYRAMIDAL) {
          const double* fri = d->contact_friction + 5 * id;
          for (int j = 0; j < 2 * (dim - 1); j++) {
            const int k = j / 2;
            d->efc_diagApprox[i + j] = tran + fri[k] * fri[k] * (k < 2 ? tran : rot);
          }
          i += 2 * (dim - 1) - 1;
        } else {
          for (int j = 0; j < dim; j++) d->efc_diagApprox[i + j] = j < 3 ? tran : rot;
          i += dim - 1;
        }
        break;
      }
    }
  }
}

static double getImpedance(const double* solimp, double pos, double margin) {
  double dmin = clampd(solimp[0], B2MJ_MINIMP, B2MJ_MAXIMP), dmax = clampd(solimp[1], B2MJ_MINIMP, B2MJ_MAXIMP);
  double width = std::fmax(MINVAL, solimp[2]), mid = clampd(solimp[3], B2MJ_MINIMP, B2MJ_MAXIMP);
  double power = std::fmax(1.0, solimp[4]);
  if (dmin == dmax) return 0.5 * (dmin + dmax);
  double x = std::fabs(pos - margin) / width, y;
  if (x >= 1) y = 1;
  else if (x <= 0) y = 0;
  else if (power == 1) y = x;
  else if (x <= mid) y = std::pow(x / mid, power) * mid;        // = x^p / mid^(p-1)
  else y = 1 - std::pow((1 - x) / (1 - mid), power) * (1 - mid);
  return dmin + y * (dmax - dmin);
}

static void makeImpedance(const b2mjModel* m, OrcData* d) {
  const bool refsafe = !(m->opt.disableflags & B2MJ_DSBL_REFSAFE);
  for (int i = 0; i < d->nefc(); i++) {
    const int id = d->efc_id[i], type = d->efc_type[i];
    const double *solref, *solimp;
    bool friction_row = false;
    switch (type) {
      case B2MJ_CNSTR_EQUALITY: solref = m->eq_solref + 2 * id; solimp = m->eq_solimp + 5 * id; break;
      case B2MJ_CNSTR_FRICTION_DOF: solref = m->dof_solref + 2 * id; solimp = m->dof_solimp + 5 * id; friction_row = true; break;
      case B2MJ_CNSTR_FRICTION_TENDON: solref = m->tendon_solref_fri + 2 * id; solimp = m->tendon_solimp_fri + 5 * id; friction_row = true; break;
      case B2MJ_CNSTR_LIMIT_JOINT: solref = m->jnt_solref + 2 * id; solimp = m->jnt_solimp + 5 * id; break;
      case B2MJ_CNSTR_LIMIT_TENDON: solref = m->tendon_solref_lim + 2 * id; solimp = m->tendon_solimp_lim + 5 * id; break;
      default:
        solref = d->contact_solref + 2 * id; solimp = d->contact_solimp + 5 * id;
        if (type == B2MJ_CNSTR_CONTACT_ELLIPTIC && i > d->contact_efc_address[id]) friction_row = true;
    }
    const double dmax = clampd(solimp[1], B2MJ_MINIMP, B2MJ_MAXIMP);
    const double imp = getImpedance(solimp, d->efc_pos[i], d->efc_margin[i]);
    double K, B;
    if (solref[0] > 0) {
      double tc = solref[0];
      if (refsafe) tc = std::fmax(tc, 2 * m->opt.timestep);
      const double dr = solref[1];
      K = 1 / std::fmax(MINVAL, dmax * dmax * tc * tc * dr * dr);
      B = 2 / std::fmax(MINVAL, dmax * tc);
    } else {
      K = -solref[0] / std::fmax(MINVAL, dmax * dmax);
      B = -solref[1] / std::fmax(MINVAL, dmax);
    }
    if (friction_row) K = 0;
    d->efc_KBIP[4 * i] = K; d->efc_KBIP[4 * i + 1] = B; d->efc_KBIP[4 * i + 2] = imp; d->efc_KBIP[4 * i + 3] = 0;
    d->efc_R[i] = std::fmax(MINVAL, (1 - imp) * d->efc_diagApprox[i] / imp);
  }
  // friction regularisation of contacts
  for (int c = 0; c < d->ncon(); c++) {
    const int adr = d->contact_efc_address[c], dim = d->contact_dim[c];
    if (adr < 0 || dim == 1) continue;
    const double* fri = d->contact_friction + 5 * c;
    if (m->opt.cone == B2MJ_CONE_ELLIPTIC) {
      d->efc_R[adr + 1] = d->efc_R[adr] / std::fmax(MINVAL, m->opt.impratio);
      d->contact_mu[c] = fri[0] * std::sqrt(d->efc_R[adr + 1] / d->efc_R[adr]);
      for (int j = 2; j < dim; j++) d->efc_R[adr + j] = d->efc_R[adr + 1] * fri[0] * fri[0] / (fri[j - 1] * fri[j - 1]);
    } else {
      d->contact_mu[c] = fri[0] * std::sqrt(1 / std::fmax(MINVAL, m->opt.impratio));
      const double Rpy = 2 * d->contact_mu[c] * d->contact_mu[c] * d->efc_R[adr];
      for (int j = 0; j < 2 * (dim - 1); j++) d->efc_R[adr + j] = Rpy;
    }
  }
  for (int i = 0; i < d->nefc(); i++) d->efc_D[i] = 1 / d->efc_R[i];
}

// mj_makeConstraint
void makeConstraint(const b2mjModel* m, OrcData* d) {
  d->nefc() = 0;
  for (int c = 0; c < d->ncon(); c++) d->contact_efc_address[c] = -1;
  if (m->opt.disableflags & B2MJ_DSBL_CONSTRAINT) return;
  instantiateEquality(m, d);
  instantiateFriction(m, d);
  instantiateLimit(m, d);
  instantiateContact(m, d);
  diagApprox(m, d);
  makeImpedance(m, d);
}

void mulJacVec(const b2mjModel* m, const OrcData* d, double* res, const double* vec) {
  for (int i = 0; i < d->nefc_[0]; i++) res[i] = dot(d->efc_J + i * m->nv, vec, m->nv);
}

void mulJacTVec(const b2mjModel* m, const OrcData* d, double* res, const double* vec) {
  const int nv = m->nv;
  zero(res, nv);
  for (int i = 0; i < d->nefc_[0]; i++) {
    const double f = vec[i];
    if (f == 0) continue;
    for (int k = 0; k < nv; k++) res[k] += d->efc_J[i * nv + k] * f;
  }
}

// mj_projectConstraint: AR = J inv(M) J' + diag(R) via JM2 = J inv(L) sqrt(inv(D))  (dual solvers)
void projectConstraint(const b2mjModel* m, OrcData* d) {
  const int nv = m->nv, nefc = d->nefc();
  if (nefc == 0 || !(m->opt.solver == B2MJ_SOL_PGS || m->opt.noslip_iterations > 0)) return;  // mj_isDual
  std::vector<double> JM2((size_t)nefc * nv);
  for (int i = 0; i < nefc; i++) solveM2(m, d, &JM2[(size_t)i * nv], d->efc_J + i * nv);
  for (int i = 0; i < nefc; i++)
    for (int j = 0; j <= i; j++) {
      double s = dot(&JM2[(size_t)i * nv], &JM2[(size_t)j * nv], nv);
      d->efc_AR[i * nefc + j] = s;
      d->efc_AR[j * nefc + i] = s;
    }
  for (int i = 0; i < nefc; i++) d->efc_AR[i * nefc + i] += d->efc_R[i];
}

// mj_referenceConstraint: efc_vel = J qvel ; aref = -B vel - K imp (pos - margin)
void referenceConstraint(const b2mjModel* m, OrcData* d) {
  mulJacVec(m, d, d->efc_vel, d->qvel);
  for (int i = 0; i < d->nefc(); i++)
    d->efc_aref[i] = -d->efc_KBIP[4 * i + 1] * d->efc_vel[i] -
                     d->efc_KBIP[4 * i] * d->efc_KBIP[4 * i + 2] * (d->efc_pos[i] - d->efc_margin[i]);
}

// mj_constraintUpdate: forces, states and cost of the primal problem for jar = J qacc - aref
void constraintUpdate(const b2mjModel* m, OrcData* d, const double* jar, double* cost, int flg_coneHessian) {
  const int nefc = d->nefc();
  double s = 0;
  for (int i = 0; i < nefc; i++) {
    const double D = d->efc_D[i], R = d->efc_R[i];
    switch (d->efc_type[i]) {
      case B2MJ_CNSTR_EQUALITY:
        d->efc_force[i] = -D * jar[i];
        d->efc_state[i] = B2MJ_CSTATE_QUADRATIC;
        s += 0.5 * D * jar[i] * jar[i];
        break;
      case B2MJ_CNSTR_FRICTION_DOF:
      case B2MJ_CNSTR_FRICTION_TENDON: {
        const double f = d->efc_frictionloss[i];
        if (jar[i] <= -R * f) {
          d->efc_force[i] = f; d->efc_state[i] = B2MJ_CSTATE_LINEARNEG;
          s += -0.5 * R * f * f - f * jar[i];
        } else if (jar[i] >= R * f) {
          d->efc_force[i] = -f; d->efc_state[i] = B2MJ_CSTATE_LINEARPOS;
          s += -0.5 * R * f * f + f * jar[i];
        } else {
          d->efc_force[i] = -D * jar[i]; d->efc_state[i] = B2MJ_CSTATE_QUADRATIC;
          s += 0.5 * D * jar[i] * jar[i];
        }
        break;
      }
      case B2MJ_CNSTR_LIMIT_JOINT:
      case B2MJ_CNSTR_LIMIT_TENDON:
      case B2MJ_CNSTR_CONTACT_FRICTIONLESS:
      case B2MJ_CNSTR_CONTACT_PYRAMIDAL:
        if (jar[i] >= 0) {
          d->efc_force[i] = 0; d->efc_state[i] = B2MJ_CSTATE_SATISFIED;
        } else {
          d->efc_force[i] = -D * jar[i]; d->efc_state[i] = B2MJ_CSTATE_QUADRATIC;
          s += 0.5 * D * jar[i] * jar[i];
        }
        break;
      case B2MJ_CNSTR_CONTACT_ELLIPTIC: {
        const int c = d->efc_id[i], dim = d->contact_dim[c];
        const double mu = d->contact_mu[c];
        const double* fri = d->contact_friction + 5 * c;
        double U[6];
        U[0] = jar[i] * mu;
        double TT = 0;
        for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * fri[j - 1]; TT += U[j] * U[j]; }
        const double N = U[0], T = std::sqrt(TT);
        if ((N >= mu * T) || (T <= 0 && N >= 0)) {
          for (int j = 0; j < dim; j++) { d->efc_force[i + j] = 0; d->efc_state[i + j] = B2MJ_CSTATE_SATISFIED; }
        } else if ((mu * N + T <= 0) || (T <= 0 && N < 0)) {
          for (int j = 0; j < dim; j++) {
            d->efc_force[i + j] = -d->efc_D[i + j] * jar[i + j];
            d->efc_state[i + j] = B2MJ_CSTATE_QUADRATIC;
            s += 0.5 * d->efc_D[i + j] * jar[i + j] * jar[i + j];
          }
        } else {
          const double Dm = d->efc_D[i] / (mu * mu * (1 + mu * mu));
          const double NmT = N - mu * T;
          s += 0.5 * Dm * NmT * NmT;
          d->efc_force[i] = -Dm * NmT * mu;
          for (int j = 1; j < dim; j++) d->efc_force[i + j] = -d->efc_force[i] / T * U[j] * fri[j - 1];
          for (int j = 0; j < dim; j++) d->efc_state[i + j] = B2MJ_CSTATE_CONE;
          if (flg_coneHessian) {
            // H = d2/djar2 of 0.5*Dm*(N - mu*T)^2, N = mu*jar0, T = |fri .* jar_t|
            double* H = d->contact_H + 36 * c;
            double g[6];  // dT/djar
            g[0] = 0;
            for (int j = 1; j < dim; j++) g[j] = U[j] * fri[j - 1] / T;
            for (int a = 0; a < dim; a++)
              for (int b = 0; b < dim; b++) {
                // d(NmT)/djar_a
                const double da = (a == 0 ? mu : -mu * g[a]), db = (b == 0 ? mu : -mu * g[b]);
                double h = Dm * da * db;
                if (a > 0 && b > 0) {
                  // second derivative of T
                  const double d2T = ((a == b ? fri[a - 1] * fri[a - 1] : 0) - g[a] * g[b]) / T;
                  h += Dm * NmT * (-mu) * d2T;
                }
                H[a * dim + b] = h;
              }
          }
        }
        i += dim - 1;
        break;
      }
    }
  }
  if (cost) *cost = s;
  mulJacTVec(m, d, d->qfrc_constraint, d->efc_force);
}

}  // namespace orc
