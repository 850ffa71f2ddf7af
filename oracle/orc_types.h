// orc_types.h — data container of the CPU oracle (one env; mjData field names).
// TEST INFRASTRUCTURE ONLY (see orc_math.h header).
#pragma once
#include <vector>

#include "b2mj.h"
#include "oracle.h"

// X(field enum, member, ctype, per-env count expression using `m`)
#define ORC_FIELDS(X)                                                                              \
  X(B2MJ_F_QPOS, qpos, double, m->nq) X(B2MJ_F_QVEL, qvel, double, m->nv)                          \
  X(B2MJ_F_ACT, act, double, m->na) X(B2MJ_F_CTRL, ctrl, double, m->nu)                            \
  X(B2MJ_F_QFRC_APPLIED, qfrc_applied, double, m->nv)                                              \
  X(B2MJ_F_XFRC_APPLIED, xfrc_applied, double, 6 * m->nbody)                                       \
  X(B2MJ_F_MOCAP_POS, mocap_pos, double, 3 * m->nmocap)                                            \
  X(B2MJ_F_MOCAP_QUAT, mocap_quat, double, 4 * m->nmocap)                                          \
  X(B2MJ_F_QACC_WARMSTART, qacc_warmstart, double, m->nv) X(B2MJ_F_TIME, time, double, 1)          \
  X(B2MJ_F_QACC, qacc, double, m->nv) X(B2MJ_F_SENSORDATA, sensordata, double, m->nsensordata)     \
  X(B2MJ_F_ACT_DOT, act_dot, double, m->na)                                                        \
  X(B2MJ_F_XPOS, xpos, double, 3 * m->nbody) X(B2MJ_F_XQUAT, xquat, double, 4 * m->nbody)          \
  X(B2MJ_F_XMAT, xmat, double, 9 * m->nbody) X(B2MJ_F_XIPOS, xipos, double, 3 * m->nbody)          \
  X(B2MJ_F_XIMAT, ximat, double, 9 * m->nbody) X(B2MJ_F_XANCHOR, xanchor, double, 3 * m->njnt)     \
  X(B2MJ_F_XAXIS, xaxis, double, 3 * m->njnt) X(B2MJ_F_GEOM_XPOS, geom_xpos, double, 3 * m->ngeom) \
  X(B2MJ_F_GEOM_XMAT, geom_xmat, double, 9 * m->ngeom)                                             \
  X(B2MJ_F_SITE_XPOS, site_xpos, double, 3 * m->nsite)                                             \
  X(B2MJ_F_SITE_XMAT, site_xmat, double, 9 * m->nsite)                                             \
  X(B2MJ_F_SUBTREE_COM, subtree_com, double, 3 * m->nbody)                                         \
  X(B2MJ_F_CINERT, cinert, double, 10 * m->nbody) X(B2MJ_F_CDOF, cdof, double, 6 * m->nv)          \
  X(B2MJ_F_CRB, crb, double, 10 * m->nbody) X(B2MJ_F_TEN_LENGTH, ten_length, double, m->ntendon)   \
  X(B2MJ_F_TEN_J, ten_J, double, m->ntendon * m->nv)                                               \
  X(B2MJ_F_ACTUATOR_LENGTH, actuator_length, double, m->nu)                                        \
  X(B2MJ_F_ACTUATOR_MOMENT, actuator_moment, double, m->nu * m->nv)                                \
  X(B2MJ_F_QM, qM, double, m->nM) X(B2MJ_F_QLD, qLD, double, m->nM)                                \
  X(B2MJ_F_QLDIAGINV, qLDiagInv, double, m->nv) X(B2MJ_F_QLDIAGSQRTINV, qLDiagSqrtInv, double, m->nv) \
  X(B2MJ_F_TEN_VELOCITY, ten_velocity, double, m->ntendon)                                         \
  X(B2MJ_F_ACTUATOR_VELOCITY, actuator_velocity, double, m->nu)                                    \
  X(B2MJ_F_CVEL, cvel, double, 6 * m->nbody) X(B2MJ_F_CDOF_DOT, cdof_dot, double, 6 * m->nv)       \
  X(B2MJ_F_QFRC_BIAS, qfrc_bias, double, m->nv) X(B2MJ_F_QFRC_PASSIVE, qfrc_passive, double, m->nv)\
  X(B2MJ_F_ACTUATOR_FORCE, actuator_force, double, m->nu)                                          \
  X(B2MJ_F_QFRC_ACTUATOR, qfrc_actuator, double, m->nv)                                            \
  X(B2MJ_F_QFRC_SMOOTH, qfrc_smooth, double, m->nv) X(B2MJ_F_QACC_SMOOTH, qacc_smooth, double, m->nv) \
  X(B2MJ_F_QFRC_CONSTRAINT, qfrc_constraint, double, m->nv)                                        \
  X(B2MJ_F_CACC, cacc, double, 6 * m->nbody) X(B2MJ_F_CFRC_INT, cfrc_int, double, 6 * m->nbody)    \
  X(B2MJ_F_CFRC_EXT, cfrc_ext, double, 6 * m->nbody)                                               \
  X(B2MJ_F_CONTACT_DIST, contact_dist, double, m->nconmax)                                         \
  X(B2MJ_F_CONTACT_POS, contact_pos, double, 3 * m->nconmax)                                       \
  X(B2MJ_F_CONTACT_FRAME, contact_frame, double, 9 * m->nconmax)                                   \
  X(B2MJ_F_CONTACT_INCLUDEMARGIN, contact_includemargin, double, m->nconmax)                       \
  X(B2MJ_F_CONTACT_FRICTION, contact_friction, double, 5 * m->nconmax)                             \
  X(B2MJ_F_CONTACT_SOLREF, contact_solref, double, 2 * m->nconmax)                                 \
  X(B2MJ_F_CONTACT_SOLIMP, contact_solimp, double, 5 * m->nconmax)                                 \
  X(B2MJ_F_CONTACT_MU, contact_mu, double, m->nconmax)                                             \
  X(B2MJ_F_CONTACT_DIM, contact_dim, int, m->nconmax)                                              \
  X(B2MJ_F_CONTACT_GEOM1, contact_geom1, int, m->nconmax)                                          \
  X(B2MJ_F_CONTACT_GEOM2, contact_geom2, int, m->nconmax)                                          \
  X(B2MJ_F_CONTACT_EXCLUDE, contact_exclude, int, m->nconmax)                                      \
  X(B2MJ_F_CONTACT_EFC_ADDRESS, contact_efc_address, int, m->nconmax)                              \
  X(B2MJ_F_EFC_TYPE, efc_type, int, m->njmax) X(B2MJ_F_EFC_ID, efc_id, int, m->njmax)              \
  X(B2MJ_F_EFC_J, efc_J, double, m->njmax * m->nv) X(B2MJ_F_EFC_POS, efc_pos, double, m->njmax)    \
  X(B2MJ_F_EFC_MARGIN, efc_margin, double, m->njmax)                                               \
  X(B2MJ_F_EFC_FRICTIONLOSS, efc_frictionloss, double, m->njmax)                                   \
  X(B2MJ_F_EFC_DIAGAPPROX, efc_diagApprox, double, m->njmax)                                       \
  X(B2MJ_F_EFC_KBIP, efc_KBIP, double, 4 * m->njmax) X(B2MJ_F_EFC_D, efc_D, double, m->njmax)      \
  X(B2MJ_F_EFC_R, efc_R, double, m->njmax) X(B2MJ_F_EFC_VEL, efc_vel, double, m->njmax)            \
  X(B2MJ_F_EFC_AREF, efc_aref, double, m->njmax) X(B2MJ_F_EFC_B, efc_b, double, m->njmax)          \
  X(B2MJ_F_EFC_FORCE, efc_force, double, m->njmax) X(B2MJ_F_EFC_STATE, efc_state, int, m->njmax)   \
  X(B2MJ_F_EFC_AR, efc_AR, double, ((m->opt.solver == B2MJ_SOL_PGS || m->opt.noslip_iterations > 0) ? m->njmax * m->njmax : 0))      \
  X(B2MJ_F_NCON, ncon_, int, 1) X(B2MJ_F_NEFC, nefc_, int, 1) X(B2MJ_F_SOLVER_ITER, solver_iter_, int, 1) \
  X(B2MJ_F_WARNING, warning, int, B2MJ_NWARNING)

struct OrcData {
#define X(e, n, t, c) t* n;
  ORC_FIELDS(X)
#undef X
  // scratch outside the field table
  double* qH;         // nM: M + h*diag(damping), factorised (Euler implicit damping)
  double* qHDiagInv;  // nv
  double* contact_H;  // 36*nconmax elliptic cone Hessians (Newton)
  std::vector<std::vector<double>> bufs;
  std::vector<std::vector<int>> ibufs;
  orc_callback cb_control;
  orc_callback cb_passive;
  void* cb_user;
  int n_control_calls, n_passive_calls;
  // convenience accessors
  int& ncon() { return ncon_[0]; }
  int& nefc() { return nefc_[0]; }
};

namespace orc {
// orc_smooth.cpp
void kinematics(const b2mjModel* m, OrcData* d);
void comPos(const b2mjModel* m, OrcData* d);
void tendon(const b2mjModel* m, OrcData* d);
void transmission(const b2mjModel* m, OrcData* d);
void crb(const b2mjModel* m, OrcData* d);
void factorM(const b2mjModel* m, OrcData* d);
void factorI(const b2mjModel* m, const double* M, double* LD, double* diaginv, double* sqrtdiaginv);
void solveLD(const b2mjModel* m, double* x, const double* LD, const double* diaginv);
void solveM(const b2mjModel* m, const OrcData* d, double* x, const double* y);
void solveM2(const b2mjModel* m, const OrcData* d, double* x, const double* y);
void mulM(const b2mjModel* m, const OrcData* d, double* res, const double* vec);
void comVel(const b2mjModel* m, OrcData* d);
void passive(const b2mjModel* m, OrcData* d);
void rne(const b2mjModel* m, OrcData* d, int flg_acc, double* result);
void rnePostConstraint(const b2mjModel* m, OrcData* d);
void jac(const b2mjModel* m, const OrcData* d, double* jacp, double* jacr, const double* point, int body);
void applyFT(const b2mjModel* m, const OrcData* d, const double* force, const double* torque, const double* point,
             int body, double* qfrc);
void fwdActuation(const b2mjModel* m, OrcData* d);
void fwdAcceleration(const b2mjModel* m, OrcData* d);
void objectVelocity(const b2mjModel* m, const OrcData* d, int objtype, int objid, double* res, int flg_local);
void objectAcceleration(const b2mjModel* m, const OrcData* d, int objtype, int objid, double* res, int flg_local);
// orc_collision.cpp
void collision(const b2mjModel* m, OrcData* d);
// orc_constraint.cpp
void makeConstraint(const b2mjModel* m, OrcData* d);
void projectConstraint(const b2mjModel* m, OrcData* d);
void referenceConstraint(const b2mjModel* m, OrcData* d);
void constraintUpdate(const b2mjModel* m, OrcData* d, const double* jar, double* cost, int flg_coneHessian);
void mulJacVec(const b2mjModel* m, const OrcData* d, double* res, const double* vec);
void mulJacTVec(const b2mjModel* m, const OrcData* d, double* res, const double* vec);
// orc_solver.cpp
void fwdConstraint(const b2mjModel* m, OrcData* d);
// orc_forward.cpp
void sensorPos(const b2mjModel* m, OrcData* d);
void sensorVel(const b2mjModel* m, OrcData* d);
void sensorAcc(const b2mjModel* m, OrcData* d);
void integratePos(const b2mjModel* m, double* qpos, const double* qvel, double dt);
// orc_ray.cpp
double ray(const b2mjModel* m, const OrcData* d, const double* pnt, const double* vec, int bodyexclude, int* geomid);
// orc_implicit.cpp
void implicitQacc(const b2mjModel* m, OrcData* d, double* qacc_out);
}  // namespace orc
