/* b2mj.h — C-ABI of the B200-native batched physics step.
 *
 * This is the drop-in boundary for the one hot path of ubi-agni/mujoco_ros_pkgs: the
 * `mj_step(model, data)` call made by MujocoEnv::physicsLoop (reference
 * mujoco_ros/src/mujoco_env.cpp:498,552,593) plus the data paths either side of it
 * (mj_resetData :252, mj_forward :329/:621, the plugin-visible mjModel/mjData fields of
 * SURVEY.md Appendix B).  Plain C: pointers, sizes, ints.  No torch / C++ types cross it.
 *
 * Conventions: every function returns 0 on success and a negative B2MJ_E* code on error (never
 * aborts, never throws); b2mj_last_error() returns a thread-local description.  One stepping
 * thread per handle (the reference serialises on physics_thread_mutex_, mujoco_env.cpp:456).
 *
 * Field names follow mjModel / mjData (MuJoCo 2.3.7, the version the reference pins:
 * mujoco_ros/CMakeLists.txt:61) so plugin code ports by search-and-replace.
 */
#ifndef B2MJ_H_
#define B2MJ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2MJ_VERSION 100

/* ---- constants (values pinned by reference mujoco_env_fixture.h:145-156, GeomType.msg:2-9) ---- */
#define B2MJ_MINVAL 1E-15
#define B2MJ_MAXVAL 1E+10
#define B2MJ_MINMU 1E-5
#define B2MJ_MINIMP 0.0001
#define B2MJ_MAXIMP 0.9999
#define B2MJ_NEQDATA 11
#define B2MJ_NIMP 5
#define B2MJ_NREF 2
#define B2MJ_NGAIN 10
#define B2MJ_NBIAS 10
#define B2MJ_NDYN 10
#define B2MJ_NWARNING 8

enum { B2MJ_JNT_FREE = 0, B2MJ_JNT_BALL = 1, B2MJ_JNT_SLIDE = 2, B2MJ_JNT_HINGE = 3 };
enum {
  B2MJ_GEOM_PLANE = 0, B2MJ_GEOM_HFIELD = 1, B2MJ_GEOM_SPHERE = 2, B2MJ_GEOM_CAPSULE = 3,
  B2MJ_GEOM_ELLIPSOID = 4, B2MJ_GEOM_CYLINDER = 5, B2MJ_GEOM_BOX = 6, B2MJ_GEOM_MESH = 7
};
enum { B2MJ_INT_EULER = 0, B2MJ_INT_RK4 = 1, B2MJ_INT_IMPLICIT = 2, B2MJ_INT_IMPLICITFAST = 3 };
enum { B2MJ_CONE_PYRAMIDAL = 0, B2MJ_CONE_ELLIPTIC = 1 };
enum { B2MJ_SOL_PGS = 0, B2MJ_SOL_CG = 1, B2MJ_SOL_NEWTON = 2 };
enum { B2MJ_EQ_CONNECT = 0, B2MJ_EQ_WELD = 1, B2MJ_EQ_JOINT = 2, B2MJ_EQ_TENDON = 3 };
enum {
  B2MJ_CNSTR_EQUALITY = 0, B2MJ_CNSTR_FRICTION_DOF = 1, B2MJ_CNSTR_FRICTION_TENDON = 2,
  B2MJ_CNSTR_LIMIT_JOINT = 3, B2MJ_CNSTR_LIMIT_TENDON = 4, B2MJ_CNSTR_CONTACT_FRICTIONLESS = 5,
  B2MJ_CNSTR_CONTACT_PYRAMIDAL = 6, B2MJ_CNSTR_CONTACT_ELLIPTIC = 7
};
enum {
  B2MJ_CSTATE_SATISFIED = 0, B2MJ_CSTATE_QUADRATIC = 1, B2MJ_CSTATE_LINEARNEG = 2,
  B2MJ_CSTATE_LINEARPOS = 3, B2MJ_CSTATE_CONE = 4
};
enum { B2MJ_TRN_JOINT = 0, B2MJ_TRN_JOINTINPARENT = 1, B2MJ_TRN_TENDON = 3 };
enum { B2MJ_DYN_NONE = 0, B2MJ_DYN_INTEGRATOR = 1, B2MJ_DYN_FILTER = 2 };
enum { B2MJ_GAIN_FIXED = 0, B2MJ_GAIN_AFFINE = 1 };
enum { B2MJ_BIAS_NONE = 0, B2MJ_BIAS_AFFINE = 1 };
enum { B2MJ_OBJ_UNKNOWN = 0, B2MJ_OBJ_BODY = 1, B2MJ_OBJ_XBODY = 2, B2MJ_OBJ_JOINT = 3, B2MJ_OBJ_DOF = 4,
       B2MJ_OBJ_GEOM = 5, B2MJ_OBJ_SITE = 6, B2MJ_OBJ_CAMERA = 7, B2MJ_OBJ_TENDON = 16,
       B2MJ_OBJ_ACTUATOR = 17, B2MJ_OBJ_SENSOR = 18, B2MJ_OBJ_PAIR = 13, B2MJ_OBJ_EQUALITY = 15, B2MJ_OBJ_KEY = 21 };
enum { B2MJ_STAGE_NONE = 0, B2MJ_STAGE_POS = 1, B2MJ_STAGE_VEL = 2, B2MJ_STAGE_ACC = 3 };
enum { B2MJ_DATATYPE_REAL = 0, B2MJ_DATATYPE_POSITIVE = 1, B2MJ_DATATYPE_AXIS = 2, B2MJ_DATATYPE_QUATERNION = 3 };
/* sensor enum: the 36 types the reference's sensor plugin names (mujoco_sensor_handler_plugin.cpp:70-105) */
enum {
  B2MJ_SENS_TOUCH = 0, B2MJ_SENS_ACCELEROMETER, B2MJ_SENS_VELOCIMETER, B2MJ_SENS_GYRO, B2MJ_SENS_FORCE,
  B2MJ_SENS_TORQUE, B2MJ_SENS_MAGNETOMETER, B2MJ_SENS_RANGEFINDER, B2MJ_SENS_JOINTPOS, B2MJ_SENS_JOINTVEL,
  B2MJ_SENS_TENDONPOS, B2MJ_SENS_TENDONVEL, B2MJ_SENS_ACTUATORPOS, B2MJ_SENS_ACTUATORVEL,
  B2MJ_SENS_ACTUATORFRC, B2MJ_SENS_JOINTACTFRC, B2MJ_SENS_BALLQUAT, B2MJ_SENS_BALLANGVEL,
  B2MJ_SENS_JOINTLIMITPOS, B2MJ_SENS_JOINTLIMITVEL, B2MJ_SENS_JOINTLIMITFRC, B2MJ_SENS_TENDONLIMITPOS,
  B2MJ_SENS_TENDONLIMITVEL, B2MJ_SENS_TENDONLIMITFRC, B2MJ_SENS_FRAMEPOS, B2MJ_SENS_FRAMEQUAT,
  B2MJ_SENS_FRAMEXAXIS, B2MJ_SENS_FRAMEYAXIS, B2MJ_SENS_FRAMEZAXIS, B2MJ_SENS_FRAMELINVEL,
  B2MJ_SENS_FRAMEANGVEL, B2MJ_SENS_FRAMELINACC, B2MJ_SENS_FRAMEANGACC, B2MJ_SENS_SUBTREECOM,
  B2MJ_SENS_SUBTREELINVEL, B2MJ_SENS_SUBTREEANGMOM, B2MJ_SENS_CLOCK, B2MJ_NSENSORTYPE
};
/* disable flags (subset of mjtDisableBit, same bit positions as MuJoCo 2.3.7) */
enum {
  B2MJ_DSBL_CONSTRAINT = 1 << 0, B2MJ_DSBL_EQUALITY = 1 << 1, B2MJ_DSBL_FRICTIONLOSS = 1 << 2,
  B2MJ_DSBL_LIMIT = 1 << 3, B2MJ_DSBL_CONTACT = 1 << 4, B2MJ_DSBL_PASSIVE = 1 << 5,
  B2MJ_DSBL_GRAVITY = 1 << 6, B2MJ_DSBL_CLAMPCTRL = 1 << 7, B2MJ_DSBL_WARMSTART = 1 << 8,
  B2MJ_DSBL_FILTERPARENT = 1 << 9, B2MJ_DSBL_ACTUATION = 1 << 10, B2MJ_DSBL_REFSAFE = 1 << 11,
  B2MJ_DSBL_SENSOR = 1 << 12, B2MJ_DSBL_MIDPHASE = 1 << 13, B2MJ_DSBL_EULERDAMP = 1 << 14
};
/* mjtEnableBit.  Only OVERRIDE has an effect (applied by the model compiler: every geom / pair takes the o_* contact
 * parameters); ENERGY / FWDINV add outputs this library does not expose; the others are refused by the compiler. */
enum { B2MJ_ENBL_OVERRIDE = 1 << 0, B2MJ_ENBL_ENERGY = 1 << 1, B2MJ_ENBL_FWDINV = 1 << 2, B2MJ_ENBL_SENSORNOISE = 1 << 3,
       B2MJ_ENBL_MULTICCD = 1 << 4 };
/* warnings (mjtWarning order) */
enum {
  B2MJ_WARN_INERTIA = 0, B2MJ_WARN_CONTACTFULL = 1, B2MJ_WARN_CNSTRFULL = 2, B2MJ_WARN_VGEOMFULL = 3,
  B2MJ_WARN_BADQPOS = 4, B2MJ_WARN_BADQVEL = 5, B2MJ_WARN_BADQACC = 6, B2MJ_WARN_BADCTRL = 7
};

/* error codes */
enum {
  B2MJ_OK = 0, B2MJ_AGAIN = 1 /* b2mj_step_end: RK4 sub-step done, run the hooks and call it again */, B2MJ_EINVAL = -1, B2MJ_ENOMEM = -2, B2MJ_ECUDA = -3, B2MJ_EPARSE = -4,
  B2MJ_EUNSUPPORTED = -5, B2MJ_ESTATE = -6, B2MJ_ENODEVICE = -7
};

/* ---- mjOption mirror (fields the reference exposes: viewer.cpp:1624-1647) ---- */
typedef struct b2mjOption {
  double timestep;
  double impratio;
  double tolerance;
  double ls_tolerance;
  double noslip_tolerance;
  double mpr_tolerance;
  double gravity[3];
  double wind[3];
  double magnetic[3];
  double density;
  double viscosity;
  double o_margin;
  double o_solref[2];
  double o_solimp[5];
  int integrator;
  int collision;
  int cone;
  int jacobian;
  int solver;
  int iterations;
  int ls_iterations;
  int noslip_iterations;
  int mpr_iterations;
  int disableflags;
  int enableflags;
  int _pad;
} b2mjOption;

typedef struct b2mjStatistic {
  double meaninertia;
  double meanmass;
  double meansize;
  double extent;
  double center[3];
} b2mjStatistic;

/* ---- model sizes and arrays, X-macro lists (one source for struct, alloc, device upload, reflection) ---- */
#define B2MJ_MODEL_SIZES(X)                                                                       \
  X(nq) X(nv) X(nu) X(na) X(nbody) X(njnt) X(ngeom) X(nsite) X(ntendon) X(nwrap) X(neq)          \
  X(nsensor) X(nsensordata) X(nM) X(nmocap) X(nexclude) X(ncollpair) X(nconmax) X(njmax)          \
  X(nnames) X(nlevel) X(ntree)                                                                    \
  X(npair) X(nmesh) X(nmeshvert) X(nhfield) X(nhfielddata)                                                 \
  X(nkey) X(nkeyq) X(nkeyv) X(nkeya) X(nkeyu) X(nkeymp) X(nkeymq) /* keyframes; nkeyq = nkey*nq, ... (flat arrays) */

/* X(ctype, name, rows(size field), cols) */
#define B2MJ_MODEL_ARRAYS(X)                                                                      \
  X(double, qpos0, nq, 1) X(double, qpos_spring, nq, 1)                                           \
  X(int, body_parentid, nbody, 1) X(int, body_rootid, nbody, 1) X(int, body_weldid, nbody, 1)     \
  X(int, body_mocapid, nbody, 1) X(int, body_jntnum, nbody, 1) X(int, body_jntadr, nbody, 1)      \
  X(int, body_dofnum, nbody, 1) X(int, body_dofadr, nbody, 1) X(int, body_geomnum, nbody, 1)      \
  X(int, body_geomadr, nbody, 1) X(int, body_level, nbody, 1)                                     \
  X(double, body_pos, nbody, 3) X(double, body_quat, nbody, 4)                                    \
  X(double, body_ipos, nbody, 3) X(double, body_iquat, nbody, 4) X(double, body_mass, nbody, 1)   \
  X(double, body_subtreemass, nbody, 1) X(double, body_inertia, nbody, 3)                         \
  X(double, body_invweight0, nbody, 2) X(double, body_gravcomp, nbody, 1)                         \
  X(int, level_bodyadr, nlevel, 1) X(int, level_bodynum, nlevel, 1) X(int, level_body, nbody, 1) \
  X(int, jnt_type, njnt, 1) X(int, jnt_qposadr, njnt, 1) X(int, jnt_dofadr, njnt, 1)              \
  X(int, jnt_bodyid, njnt, 1) X(int, jnt_limited, njnt, 1)                                        \
  X(double, jnt_solref, njnt, 2) X(double, jnt_solimp, njnt, 5) X(double, jnt_pos, njnt, 3)       \
  X(double, jnt_axis, njnt, 3) X(double, jnt_stiffness, njnt, 1) X(double, jnt_range, njnt, 2)    \
  X(double, jnt_margin, njnt, 1)                                                                  \
  X(int, dof_bodyid, nv, 1) X(int, dof_jntid, nv, 1) X(int, dof_parentid, nv, 1)                  \
  X(int, dof_Madr, nv, 1) X(int, dof_treeid, nv, 1)                                               \
  X(double, dof_solref, nv, 2) X(double, dof_solimp, nv, 5)                                       \
  X(double, dof_frictionloss, nv, 1) X(double, dof_armature, nv, 1) X(double, dof_damping, nv, 1) \
  X(double, dof_invweight0, nv, 1) X(double, dof_M0, nv, 1)                                       \
  X(int, geom_type, ngeom, 1) X(int, geom_contype, ngeom, 1) X(int, geom_conaffinity, ngeom, 1)   \
  X(int, geom_condim, ngeom, 1) X(int, geom_bodyid, ngeom, 1) X(int, geom_priority, ngeom, 1)     \
  X(double, geom_solmix, ngeom, 1) X(double, geom_solref, ngeom, 2)                               \
  X(double, geom_solimp, ngeom, 5) X(double, geom_size, ngeom, 3) X(double, geom_rbound, ngeom, 1)\
  X(double, geom_pos, ngeom, 3) X(double, geom_quat, ngeom, 4) X(double, geom_friction, ngeom, 3) \
  X(double, geom_margin, ngeom, 1) X(double, geom_gap, ngeom, 1) X(double, geom_rgba, ngeom, 4)   \
  X(int, geom_dataid, ngeom, 1) /* mesh id of a mesh geom, else -1 */                             \
  X(int, mesh_vertadr, nmesh, 1) X(int, mesh_vertnum, nmesh, 1)                                   \
  X(double, mesh_vert, nmeshvert, 3) /* convex-hull vertices, mesh frame = centre of mass + principal axes */ \
  /* height fields: geom_dataid of an hfield geom is the hfield id; size = (x radius, y radius, elevation, base); */ \
  /* data row-major [nrow][ncol], normalised to [0, 1], (0, 0) at the (-x, -y) corner (mjModel.hfield_*) */         \
  X(int, hfield_nrow, nhfield, 1) X(int, hfield_ncol, nhfield, 1) X(int, hfield_adr, nhfield, 1)  \
  X(double, hfield_size, nhfield, 4) X(double, hfield_data, nhfielddata, 1)                       \
  X(int, site_bodyid, nsite, 1) X(int, site_type, nsite, 1) X(double, site_size, nsite, 3)        \
  X(double, site_pos, nsite, 3) X(double, site_quat, nsite, 4)                                    \
  X(int, tendon_adr, ntendon, 1) X(int, tendon_num, ntendon, 1) X(int, tendon_limited, ntendon, 1)\
  X(double, tendon_solref_lim, ntendon, 2) X(double, tendon_solimp_lim, ntendon, 5)               \
  X(double, tendon_solref_fri, ntendon, 2) X(double, tendon_solimp_fri, ntendon, 5)               \
  X(double, tendon_range, ntendon, 2) X(double, tendon_margin, ntendon, 1)                        \
  X(double, tendon_stiffness, ntendon, 1) X(double, tendon_damping, ntendon, 1)                   \
  X(double, tendon_frictionloss, ntendon, 1) X(double, tendon_lengthspring, ntendon, 2)           \
  X(double, tendon_length0, ntendon, 1) X(double, tendon_invweight0, ntendon, 1)                  \
  X(int, wrap_type, nwrap, 1) X(int, wrap_objid, nwrap, 1) X(double, wrap_prm, nwrap, 1)          \
  X(int, actuator_trntype, nu, 1) X(int, actuator_dyntype, nu, 1) X(int, actuator_gaintype, nu, 1)\
  X(int, actuator_biastype, nu, 1) X(int, actuator_trnid, nu, 2) X(int, actuator_actadr, nu, 1)   \
  X(int, actuator_ctrllimited, nu, 1) X(int, actuator_forcelimited, nu, 1)                        \
  X(int, actuator_actlimited, nu, 1)                                                              \
  X(double, actuator_dynprm, nu, B2MJ_NDYN) X(double, actuator_gainprm, nu, B2MJ_NGAIN)           \
  X(double, actuator_biasprm, nu, B2MJ_NBIAS) X(double, actuator_ctrlrange, nu, 2)                \
  X(double, actuator_forcerange, nu, 2) X(double, actuator_actrange, nu, 2)                       \
  X(double, actuator_gear, nu, 6) X(double, actuator_acc0, nu, 1) X(double, actuator_length0, nu, 1) \
  X(int, eq_type, neq, 1) X(int, eq_obj1id, neq, 1) X(int, eq_obj2id, neq, 1)                     \
  X(int, eq_active, neq, 1) X(double, eq_solref, neq, 2) X(double, eq_solimp, neq, 5)             \
  X(double, eq_data, neq, B2MJ_NEQDATA)                                                           \
  X(int, exclude_signature, nexclude, 1)                                                          \
  X(int, collpair_geom1, ncollpair, 1) X(int, collpair_geom2, ncollpair, 1)                       \
  X(int, collpair_slotadr, ncollpair, 1) X(int, collpair_maxcon, ncollpair, 1)                    \
  X(int, collpair_pairid, ncollpair, 1) /* explicit <contact><pair> behind the candidate, or -1 */ \
  X(int, pair_geom1, npair, 1) X(int, pair_geom2, npair, 1) X(int, pair_dim, npair, 1)            \
  X(double, pair_solref, npair, 2) X(double, pair_solimp, npair, 5) X(double, pair_margin, npair, 1) \
  X(double, pair_gap, npair, 1) X(double, pair_friction, npair, 5) X(int, name_pairadr, npair, 1) \
  X(int, sensor_type, nsensor, 1) X(int, sensor_datatype, nsensor, 1)                             \
  X(int, sensor_needstage, nsensor, 1) X(int, sensor_objtype, nsensor, 1)                         \
  X(int, sensor_objid, nsensor, 1) X(int, sensor_reftype, nsensor, 1)                             \
  X(int, sensor_refid, nsensor, 1) X(int, sensor_dim, nsensor, 1) X(int, sensor_adr, nsensor, 1)  \
  X(double, sensor_cutoff, nsensor, 1) X(double, sensor_noise, nsensor, 1)                        \
  X(int, name_bodyadr, nbody, 1) X(int, name_jntadr, njnt, 1) X(int, name_geomadr, ngeom, 1)      \
  X(int, name_siteadr, nsite, 1) X(int, name_tendonadr, ntendon, 1)                               \
  X(int, name_actuatoradr, nu, 1) X(int, name_sensoradr, nsensor, 1) X(int, name_eqadr, neq, 1)   \
  X(double, key_time, nkey, 1) X(double, key_qpos, nkeyq, 1) X(double, key_qvel, nkeyv, 1)        \
  X(double, key_act, nkeya, 1) X(double, key_ctrl, nkeyu, 1) X(double, key_mpos, nkeymp, 1)       \
  X(double, key_mquat, nkeymq, 1) X(int, name_keyadr, nkey, 1)                                    \
  X(char, names, nnames, 1)

typedef struct b2mjModel {
#define B2MJ_X_SIZE(n) int n;
  B2MJ_MODEL_SIZES(B2MJ_X_SIZE)
#undef B2MJ_X_SIZE
  int _pad0;
  b2mjOption opt;
  b2mjStatistic stat;
#define B2MJ_X_ARR(t, n, r, c) t* n;
  B2MJ_MODEL_ARRAYS(B2MJ_X_ARR)
#undef B2MJ_X_ARR
  void* _owner; /* library-private allocation record */
} b2mjModel;

/* ---- per-env data fields (mjData names).  One list drives: the device arena layout, b2mj_get/set,
 *      the CPU oracle's data struct, and the Python reflection used by the tests. ---- */
typedef enum b2mj_field {
  /* state + inputs (resident in HBM between steps) */
  B2MJ_F_QPOS = 0, B2MJ_F_QVEL, B2MJ_F_ACT, B2MJ_F_CTRL, B2MJ_F_QFRC_APPLIED, B2MJ_F_XFRC_APPLIED,
  B2MJ_F_MOCAP_POS, B2MJ_F_MOCAP_QUAT, B2MJ_F_QACC_WARMSTART, B2MJ_F_TIME,
  /* outputs written every step */
  B2MJ_F_QACC, B2MJ_F_SENSORDATA, B2MJ_F_ACT_DOT,
  /* position stage */
  B2MJ_F_XPOS, B2MJ_F_XQUAT, B2MJ_F_XMAT, B2MJ_F_XIPOS, B2MJ_F_XIMAT, B2MJ_F_XANCHOR, B2MJ_F_XAXIS,
  B2MJ_F_GEOM_XPOS, B2MJ_F_GEOM_XMAT, B2MJ_F_SITE_XPOS, B2MJ_F_SITE_XMAT, B2MJ_F_SUBTREE_COM,
  B2MJ_F_CINERT, B2MJ_F_CDOF, B2MJ_F_CRB, B2MJ_F_TEN_LENGTH, B2MJ_F_TEN_J, B2MJ_F_ACTUATOR_LENGTH,
  B2MJ_F_ACTUATOR_MOMENT, B2MJ_F_QM, B2MJ_F_QLD, B2MJ_F_QLDIAGINV, B2MJ_F_QLDIAGSQRTINV,
  /* velocity stage */
  B2MJ_F_TEN_VELOCITY, B2MJ_F_ACTUATOR_VELOCITY, B2MJ_F_CVEL, B2MJ_F_CDOF_DOT, B2MJ_F_QFRC_BIAS,
  B2MJ_F_QFRC_PASSIVE,
  /* acceleration stage */
  B2MJ_F_ACTUATOR_FORCE, B2MJ_F_QFRC_ACTUATOR, B2MJ_F_QFRC_SMOOTH, B2MJ_F_QACC_SMOOTH,
  B2MJ_F_QFRC_CONSTRAINT, B2MJ_F_CACC, B2MJ_F_CFRC_INT, B2MJ_F_CFRC_EXT,
  /* contacts: fixed-stride records, nconmax per env */
  B2MJ_F_CONTACT_DIST, B2MJ_F_CONTACT_POS, B2MJ_F_CONTACT_FRAME, B2MJ_F_CONTACT_INCLUDEMARGIN,
  B2MJ_F_CONTACT_FRICTION, B2MJ_F_CONTACT_SOLREF, B2MJ_F_CONTACT_SOLIMP, B2MJ_F_CONTACT_MU,
  B2MJ_F_CONTACT_DIM, B2MJ_F_CONTACT_GEOM1, B2MJ_F_CONTACT_GEOM2, B2MJ_F_CONTACT_EXCLUDE,
  B2MJ_F_CONTACT_EFC_ADDRESS,
  /* constraints: njmax rows per env */
  B2MJ_F_EFC_TYPE, B2MJ_F_EFC_ID, B2MJ_F_EFC_J, B2MJ_F_EFC_POS, B2MJ_F_EFC_MARGIN,
  B2MJ_F_EFC_FRICTIONLOSS, B2MJ_F_EFC_DIAGAPPROX, B2MJ_F_EFC_KBIP, B2MJ_F_EFC_D, B2MJ_F_EFC_R,
  B2MJ_F_EFC_VEL, B2MJ_F_EFC_AREF, B2MJ_F_EFC_B, B2MJ_F_EFC_FORCE, B2MJ_F_EFC_STATE, B2MJ_F_EFC_AR,
  /* counters */
  B2MJ_F_NCON, B2MJ_F_NEFC, B2MJ_F_SOLVER_ITER, B2MJ_F_WARNING,
  B2MJ_NFIELD
} b2mj_field;

/* element type and per-env element count of a field for a given model; returns count (<0 on error);
 * *is_int = 1 for int32 fields, 0 for float64 */
int b2mj_field_size(const b2mjModel* m, b2mj_field f, int* is_int);
const char* b2mj_field_name(b2mj_field f);
int b2mj_field_by_name(const char* name); /* -1 if unknown */

/* ---- model compile / ownership (replaces mj_loadXML / mj_deleteModel at mujoco_env.cpp:840-843,747) ---- */
int b2mj_model_from_xml_file(const char* path, b2mjModel** out);
int b2mj_model_from_xml_string(const char* xml, b2mjModel** out);
void b2mj_model_free(b2mjModel* m);
/* Binary model files: the counterpart of the reference's .mjb load path (mj_loadModel, mujoco_env.cpp:771-911 picks
 * it by file extension).  The format is this library's own (header "B2MJB", version, size fields, option, statistic,
 * then every array of B2MJ_MODEL_ARRAYS) -- NOT MuJoCo's .mjb layout, which only libmujoco can produce. */
int b2mj_model_save_binary(const b2mjModel* m, const char* path);
int b2mj_model_load_binary(const char* path, b2mjModel** out);
/* file loader that dispatches on the extension like the reference: ".b2mjb" -> binary, anything else -> MJCF XML */
int b2mj_model_from_file(const char* path, b2mjModel** out);
/* re-derive qpos0-dependent constants after mass / geometry edits (replaces mj_setConst,
 * reference callbacks.cpp:254,582) */
int b2mj_model_set_const(b2mjModel* m);
/* name lookup (replaces mj_name2id / mj_id2name; reference uses them 45x) */
int b2mj_name2id(const b2mjModel* m, int objtype, const char* name);
const char* b2mj_id2name(const b2mjModel* m, int objtype, int id);
/* reflection over B2MJ_MODEL_ARRAYS / B2MJ_MODEL_SIZES (used by bindings; index 0..n-1) */
int b2mj_model_narrays(void);
int b2mj_model_array_info(const b2mjModel* m, int idx, const char** name, int* elem_kind /*0 f64,1 i32,2 char*/,
                          int* rows, int* cols, void** ptr);
int b2mj_model_nsizes(void);
int b2mj_model_size_info(const b2mjModel* m, int idx, const char** name, int* value);

/* ---- batched environment handle ---- */
typedef struct b2mj_handle b2mj_handle;

/* create nenv independent envs of model m on CUDA device `device` (one process per GPU).
 * The model is copied; later host-side edits need b2mj_model_update.  Replaces mj_makeData
 * (mujoco_env.cpp:872).  State starts as after mj_resetData. */
int b2mj_create(const b2mjModel* m, int nenv, int device, b2mj_handle** out);
void b2mj_destroy(b2mj_handle* h);
int b2mj_nenv(const b2mj_handle* h);
const b2mjModel* b2mj_model(const b2mj_handle* h);
/* run launches on this CUDA stream (cudaStream_t as void*); default: the legacy default stream */
int b2mj_set_stream(b2mj_handle* h, void* cuda_stream);

/* mj_resetData semantics per env (mujoco_env.cpp:252); env_mask NULL = all, else nenv bytes (host) */
int b2mj_reset(b2mj_handle* h, const uint8_t* env_mask);
/* mj_resetDataKeyframe per env: mj_resetData, then time / qpos / qvel / act / ctrl / mocap pose from keyframe `key`
 * (<keyframe><key .../>) of the model.  The reference reaches keyframes through its viewer (viewer.cpp:1735-1751:
 * "load key" copies key_qpos / key_qvel / key_act / key_mpos / key_mquat into mjData).  env_mask as in b2mj_reset. */
int b2mj_reset_keyframe(b2mj_handle* h, int key, const uint8_t* env_mask);
/* mj_forward on all envs (mujoco_env.cpp:329,621): all stages, no integration */
int b2mj_forward(b2mj_handle* h);
/* nsteps x mj_step on all envs, asynchronous on the handle's stream; no host callbacks inside
 * (mujoco_env.cpp:498,552,593).  ctrl/qfrc_applied/xfrc_applied are read as currently set. */
int b2mj_step(b2mj_handle* h, int nsteps);
/* fused open-loop rollout: nsteps x mj_step in ONE step-kernel launch, every env's state resident on chip for the
 * whole rollout (static schedule: one env per warp, stages lock-stepped within a CTA; a ticketed persistent grid is
 * used instead when the batch would leave the last wave mostly empty).  dev_ctrl (DEVICE, [nsteps][nenv][nu], may be
 * NULL = keep current ctrl) supplies fresh controls for every env at every step; the optional DEVICE outputs receive
 * the per-step trajectory ([nsteps][nenv][nq] / [nv] / [nsensordata]) -- what the reference's lastStageCallback
 * consumers would have read after each step (mujoco_env.cpp:500,554,595).  Same arithmetic, bit for bit, as nsteps
 * calls of b2mj_step (tests/test_gpu_parity.py::test_rollout_equals_stepping). */
int b2mj_rollout(b2mj_handle* h, int nsteps, const double* dev_ctrl, double* dev_qpos_out, double* dev_qvel_out,
                 double* dev_sensor_out);
/* split step around the control hook (mjcb_control fires between the velocity stage and actuation:
 * mujoco_env.h:242-246).  step_begin: checks + position + velocity stages (+pos/vel sensors);
 * step_end: actuation, acceleration, constraint solve, acc sensors, check, integrate.
 * RK4 (mj_RungeKutta makes four forward passes and the callbacks fire in each, plugin_utils.h:89-105): b2mj_step_end
 * returns B2MJ_AGAIN after each of the first three sub-steps -- the caller runs its passive / control hooks again and
 * calls b2mj_step_end again; the fourth call integrates and returns 0:
 *     b2mj_step_begin(h); hooks(); while (b2mj_step_end(h) == B2MJ_AGAIN) hooks();   */
int b2mj_step_begin(b2mj_handle* h);
int b2mj_step_end(b2mj_handle* h);
/* one closed-loop exchange with HOST buffers in a single call: upload ctrl ([nenv][nu], may be NULL = keep), run
 * nsteps steps, download qpos / qvel / sensordata ([nenv][nq|nv|nsensordata], any may be NULL), synchronise once.
 * The reference does exactly this around every mj_step -- control plugins write d->ctrl, the step runs, lastStage
 * plugins read the state out (mujoco_env.cpp:593-595) -- as separate accesses to host memory; here the four
 * transfers and the launch are queued back to back on the handle's stream (use pinned buffers for full-speed DMA). */
int b2mj_step_host(b2mj_handle* h, int nsteps, const double* host_ctrl, double* host_qpos, double* host_qvel,
                   double* host_sensordata);
/* block until all queued work on the handle's stream has finished */
int b2mj_sync(b2mj_handle* h);

/* keep the full per-env mjData arena (every field of b2mj_field) readable after step/forward.
 * Off by default: the fused step then touches HBM only for the state record. */
int b2mj_set_keep_intermediates(b2mj_handle* h, int on);

/* host <-> device field copies.  Layout: [nenv][count] row-major, float64 or int32 (see b2mj_field_size).
 * bytes must equal nenv*count*elemsize.  Synchronous w.r.t. the handle's stream. */
int b2mj_get(b2mj_handle* h, b2mj_field f, void* host_dst, size_t bytes);
int b2mj_set(b2mj_handle* h, b2mj_field f, const void* host_src, size_t bytes);
/* device-side write of a settable field from a DEVICE buffer laid out [nenv][src_pitch_elems]
 * (src_pitch_elems 0 = tightly packed), asynchronous on the handle's stream: what a device-resident
 * controller / policy uses where a host plugin would write d->ctrl inside controlCallback
 * (plugin_utils.h:89-97). */
int b2mj_set_device(b2mj_handle* h, b2mj_field f, const void* dev_src, size_t src_pitch_elems);
/* device pointer + env pitch (in elements) of a resident state/input/output field, for device-side
 * plugins and policies (valid until destroy).  Element k of env e is ptr[e*pitch + k]. */
int b2mj_device_ptr(b2mj_handle* h, b2mj_field f, void** dev_ptr, size_t* pitch_elems);

/* re-upload model constants after host-side edits (mirrors the mutating services,
 * callbacks.cpp:462-738: gravity, body_mass, geom_*, eq_*) */
int b2mj_model_update(b2mj_handle* h, const b2mjModel* m);

/* narrowphase override per geom-type pair: the batched counterpart of MujocoEnv::registerCollisionFunction
 * (mujoco_env.h:211, mujoco_env.cpp:163-176), which lets a plugin overwrite mjCOLLISIONFUNC[type1][type2] and restores
 * the defaults on reload (:949-954).  The narrowphase is device code, so a host function pointer cannot be called from
 * it; the table instead selects among functions the library provides:
 *   B2MJ_COLLFN_DEFAULT           the built-in function of the pair
 *   B2MJ_COLLFN_NONE              no contacts for this pair type (what registering a function that returns 0 does)
 *   B2MJ_COLLFN_BOUNDING_SPHERES  both geoms collide as their bounding spheres (geom_rbound); planes stay planes
 * type1 <= type2 (the table is upper triangular, like mjCOLLISIONFUNC).  Takes effect from the next step. */
enum { B2MJ_COLLFN_DEFAULT = 0, B2MJ_COLLFN_NONE = 1, B2MJ_COLLFN_BOUNDING_SPHERES = 2 };
int b2mj_register_collision_function(b2mj_handle* h, int geom_type1, int geom_type2, int collfn);
int b2mj_reset_collision_functions(b2mj_handle* h); /* all pairs back to B2MJ_COLLFN_DEFAULT (prepareReload) */

/* per-env model variants: domain randomisation, and the reference's mutating services applied to ONE env rather than
 * to all (set_body_state mass / inertia, set_geom_properties friction / size / -> mj_setConst, set_gravity,
 * set_equality_constraint_parameters: callbacks.cpp:210-370, 462-592, 641-738).  models[0..nmodels) are edited copies of
 * the handle's model -- same sizes, topology, qpos0, solver / integrator / cone / timestep; parameters, gravity and the
 * constants b2mj_model_set_const derives may differ -- and env_model[e] (HOST, nenv ints) picks the variant of env e.
 * nmodels = 0 returns to the one shared model.  Stepping then uses the per-env-model build of the step kernel. */
int b2mj_set_env_models(b2mj_handle* h, const b2mjModel* const* models, int nmodels, const int* env_model);

/* batched plugin data paths (device kernels):
 *  robot_hw_write — DefaultRobotHWSim::writeSim (default_robot_hw_sim.cpp:248-326)
 *  sensor_readout — MujocoRosSensorsPlugin::lastStageCallback arithmetic (mujoco_sensor_handler_plugin.cpp:175-437) */
enum { B2MJ_CTRL_EFFORT = 0, B2MJ_CTRL_POSITION = 1, B2MJ_CTRL_POSITION_PID = 2, B2MJ_CTRL_VELOCITY = 3,
       B2MJ_CTRL_VELOCITY_PID = 4 };
/* joint_limits_interface::JointLimits + SoftJointLimits of one joint (what registerJointLimits reads from the URDF
 * and the parameter server, default_robot_hw_sim.cpp:340-446).  A joint with limits gets the saturation handle of its
 * hardware interface, or the soft-limits handle when has_soft_limits is set; enforced on the command before every
 * write (:262-267). */
typedef struct b2mjJointLimits {
  int has_position_limits, has_velocity_limits, has_acceleration_limits, has_effort_limits;
  int has_soft_limits, angle_wraparound;
  double min_position, max_position, max_velocity, max_acceleration, max_effort;
  double soft_min_position, soft_max_position, k_position, k_velocity;
} b2mjJointLimits;
typedef struct b2mjRobotHW {
  int njoint;                /* controlled joints (transmissions) */
  const int* joint_id;       /* [njoint] model joint ids (hinge/slide) */
  const int* control_mode;   /* [njoint] B2MJ_CTRL_* */
  const double* effort_limit;/* [njoint] or NULL (limits' max_effort, else none) */
  const double* pid_gains;   /* [njoint][5] p,i,d,i_max,i_min (control_toolbox::Pid) or NULL */
  const double* lower_limit; /* [njoint] joint limits for the limit-aware position error, or NULL (limits / jnt_range) */
  const double* upper_limit; /* [njoint] */
  const int* joint_kind;     /* [njoint] 0 revolute (limited), 1 continuous, 2 prismatic; NULL = from the model */
  const b2mjJointLimits* limits; /* [njoint] or NULL = no joint_limits_interface handles */
  const int* pid_antiwindup; /* [njoint] or NULL (0) */
} b2mjRobotHW;
/* Joint state starts as DefaultRobotHWSim::initSim leaves it (position 1.0, effort 1.0, :132-137). */
int b2mj_robot_hw_configure(b2mj_handle* h, const b2mjRobotHW* cfg);
/* DefaultRobotHWSim::writeSim (:248-326): e-stop hold, enforceLimits, mode switch.  Uses the joint state of the LAST
 * b2mj_robot_hw_read (the reference reads once per control period and writes every step).
 * cmd: DEVICE or HOST pointer [nenv][njoint] (is_device flag); e_stop: 0/1; period: time since the last write */
int b2mj_robot_hw_write(b2mj_handle* h, const double* cmd, int is_device, int e_stop, double period);
/* DefaultRobotHWSim::readSim (:230-246): refresh position (revolute: unwrapped) / velocity / effort from the state,
 * then copy them to HOST [nenv][njoint] arrays (any may be NULL; all NULL = device-side refresh only, asynchronous) */
int b2mj_robot_hw_read(b2mj_handle* h, double* pos, double* vel, double* eff);
/* DEVICE pointers ([nenv][njoint]) of the joint-state interface values, for device-resident controllers */
int b2mj_robot_hw_state_ptrs(b2mj_handle* h, double** dev_pos, double** dev_vel, double** dev_eff);

typedef struct b2mjSensorNoise {
  int sensor_id;
  double mean[3];
  double sigma[3];
  int set_flag; /* bit k set => dimension k noisy (SensorNoiseModel.msg semantics) */
} b2mjSensorNoise;
int b2mj_sensor_configure_noise(b2mj_handle* h, const b2mjSensorNoise* models, int nmodels, uint64_t seed);
/* values/gt: HOST float64 [nenv][nsensordata] holding float32-rounded numbers exactly as the reference
 * publishes them: value = float(sensordata/cutoff), noisy value = float(sensordata + noise/cutoff)
 * (noisy quaternions stay float64, as tf2 composes them); gt is the noise-free value.
 * gt may be NULL (eval mode publishes no ground truth: mujoco_sensor_handler_plugin.cpp:65-67). */
int b2mj_sensor_readout(b2mj_handle* h, double* values, double* gt);
/* device-resident form: same kernel, no copy, no synchronisation; returns DEVICE pointers ([nenv][nsensordata]) that
 * stay valid until destroy (dev_gt may be NULL) */
int b2mj_sensor_readout_device(b2mj_handle* h, double** dev_values, double** dev_gt);

/* multi-GPU publish: gather this rank's [nenv][count] slab of field f into dev_dst_all
 * ([world][nenv][count], device pointer) using the NCCL communicator passed as void* (ncclComm_t).
 * Only exchange step on the path. */
int b2mj_allgather_publish(b2mj_handle* h, b2mj_field f, void* nccl_comm, void* dev_dst_all);
/* same for several float64 fields packed per env ([nenv][sum of counts], fields in the order given) and gathered with
 * ONE collective: the aggregated state + sensor publish of a step (what the reference's lastStage consumers read after
 * mj_step, mujoco_env.cpp:593-595).  dev_dst_all: [world][nenv][sum of counts]. */
int b2mj_allgather_publish_multi(b2mj_handle* h, const b2mj_field* fields, int nfields, void* nccl_comm, void* dev_dst_all);
/* the packing step alone (no collective): this rank's slab, DEVICE pointer valid until the next publish call;
 * *count_per_env = sum of the field counts */
int b2mj_publish_pack(b2mj_handle* h, const b2mj_field* fields, int nfields, double** dev_slab, int* count_per_env);

/* ---- fused step + publish over NVLink peer memory (the all-gather folded into the step kernel) ----
 * The reference publishes every env's state after the step (mujoco_env.cpp:593-595 lastStage callbacks ->
 * mujoco_sensor_handler_plugin.cpp:175-437).  With one process per GPU the gathered slab [world][nenv][count] is what a
 * device-side consumer of the aggregated state reads; b2mj_allgather_publish builds it with a pack kernel + ncclAllGather
 * AFTER the step.  The fused form makes the step kernel itself the collective: when an env's step is done its warp
 * stores the env's row straight into EVERY rank's slab through peer (CUDA-IPC mapped, NVLink) pointers, so the exchange
 * rides under the launch instead of following it; the env whose row leaves last raises this rank's sequence flag in
 * every peer (system-scope fences), and b2mj_publish_fused_wait is a stream-ordered spin on the local flags.  Slabs are double buffered by sequence
 * parity: with every rank issuing step_publish -> wait -> (consumer) per step on one stream no slab is overwritten
 * while a peer still reads it.  Fields must live in the state record (qpos, qvel, act, qacc, sensordata, time, ...).
 *   create : allocates this rank's slabs + flags, returns their 64-byte CUDA IPC handle (exchange it out of band)
 *   connect: all_handles = [world][64] bytes in rank order; opens the peers' memory
 *   step_publish: one b2mj_step with the fused stores;  wait: returns the DEVICE pointer of the slab just completed */
int b2mj_publish_fused_create(b2mj_handle* h, int world, int rank, const b2mj_field* fields, int nfields,
                              unsigned char* ipc_handle_out /* [64] */);
int b2mj_publish_fused_connect(b2mj_handle* h, const unsigned char* all_handles /* [world][64] */);
int b2mj_step_publish(b2mj_handle* h);
int b2mj_publish_fused_wait(b2mj_handle* h, double** dev_gathered, int* count_per_env);

/* introspection for the benchmark / roofline */
typedef struct b2mjLaunchInfo {
  int warps_per_cta;
  int ctas;
  int smem_bytes_per_cta;
  int arena_doubles_per_env;
  int arena_in_smem;      /* 1 if the whole arena is shared-memory resident */
  int state_record_bytes; /* bytes per env of the HBM-resident state record (read + written per step) */
  int regs_per_thread;
  uint64_t launches;      /* kernels launched through this handle so far */
} b2mjLaunchInfo;
int b2mj_launch_info(b2mj_handle* h, b2mjLaunchInfo* out);

/* per-stage SM-cycle profile of the fused step (the kernel-internal counterpart of MuJoCo's d->timer[],
 * which the reference's viewer profiler pane reads: viewer.cpp:283-380).  enable=1 starts / clears the
 * counters, enable=0 stops; cycles (may be NULL) receives the totals accumulated so far, summed over
 * envs, one entry per stage (b2mj_stage_name).  Returns the number of stages. */
int b2mj_stage_profile(b2mj_handle* h, int enable, uint64_t* cycles, int ncycles);
const char* b2mj_stage_name(int stage);
/* diagnostic: SM residency of each env's last work item (last launch, or last rollout chunk), in units of
 * 1024 cycles; host_kcycles: HOST int32 [nenv].  Shows the load imbalance between envs. */
int b2mj_env_cycles(b2mj_handle* h, int* host_kcycles);

/* measured FP64 peak of a device (dependent-free DFMA chains on every SM): the denominator of the FP64 roofline
 * bench.py reports beside the HBM one */
int b2mj_ubench_dfma(int device, double* tflops, double* dfma_per_clk_per_sm);

const char* b2mj_last_error(void);
int b2mj_version(void);
int b2mj_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B2MJ_H_ */
