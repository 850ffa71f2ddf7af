/* oracle.h — C interface of the CPU oracle (liboracle.so).
 *
 * TEST INFRASTRUCTURE ONLY.  A plain, single-env, FP64 restatement of the physics step the reference
 * delegates to MuJoCo 2.3.7 (`mj_step`, reference mujoco_ros/src/mujoco_env.cpp:498,552,593), plus
 * the wrapper loop semantics of MujocoEnv's step path (mujoco_env.cpp:585-618).  Only tests/,
 * __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load this library; the
 * product (mujoco_ros_pkgs_b200/libb2mj.so) never does.
 *
 * PARITY UNPINNED for post-step dynamics: MuJoCo is an un-vendored binary dependency of the reference
 * (mujoco_ros/CMakeLists.txt:61, .docker/mujoco_installer.sh:9), absent from /root/reference and from
 * this image, and no reference test asserts post-step state.  What IS pinned by the
 * reference's tests is checked in tests/ (time arithmetic, qpos0, reset, callback order, sensor
 * readout arithmetic, bitwise-static pendulum), together with analytic known answers.
 */
#ifndef ORACLE_H_
#define ORACLE_H_

#include "b2mj.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcData OrcData;
typedef void (*orc_callback)(const b2mjModel* m, OrcData* d, void* user);

OrcData* orc_make_data(const b2mjModel* m);            /* mj_makeData + mj_resetData */
void orc_free_data(OrcData* d);
void orc_reset_data(const b2mjModel* m, OrcData* d);   /* mj_resetData */
void orc_reset_keyframe(const b2mjModel* m, OrcData* d, int key); /* mj_resetDataKeyframe */
void orc_forward(const b2mjModel* m, OrcData* d);      /* mj_forward */
void orc_step(const b2mjModel* m, OrcData* d);         /* mj_step */
void orc_step1(const b2mjModel* m, OrcData* d);        /* mj_step1: up to (excluding) the control callback */
void orc_step2(const b2mjModel* m, OrcData* d);        /* mj_step2: actuation .. integration (Euler) */
/* mjcb_control / mjcb_passive equivalents (reference mujoco_env.cpp:152-153) */
void orc_set_callbacks(OrcData* d, orc_callback control, orc_callback passive, void* user);
int orc_callback_counts(const OrcData* d, int* ncontrol, int* npassive);

/* implicit integrators (orc_implicit.cpp): dense [nv][nv] velocity derivatives after a forward pass (test hooks) */
void orc_rne_vel_derivative(const b2mjModel* m, const OrcData* d, double* dbias);
void orc_smooth_vel_derivative(const b2mjModel* m, const OrcData* d, int flg_bias, double* qderiv);

/* mj_ray (geomgroup NULL, flg_static 1): distance to the nearest geom along vec, -1 if none (orc_ray.cpp) */
double orc_ray(const b2mjModel* m, const OrcData* d, const double* pnt, const double* vec, int bodyexclude, int* geomid);

/* field access by b2mj_field id; count = number of elements copied; returns per-env element count */
int orc_get(const b2mjModel* m, const OrcData* d, int field, void* dst, int max_elems);
int orc_set(const b2mjModel* m, OrcData* d, int field, const void* src, int nelems);
void* orc_field_ptr(OrcData* d, int field);

/* CPU baseline driver: nenv envs x nsteps steps with the wrapper-loop semantics of the reference's
 * step path (step -> time publish stub -> last-stage sensor readout -> counter decrement).
 * qpos/qvel: [nenv][nq]/[nenv][nv] in/out; ctrl: [nsteps][nenv][nu] or NULL; nthreads >= 1.
 * Returns wall seconds spent stepping. */
double orc_rollout(const b2mjModel* m, int nenv, int nsteps, double* qpos, double* qvel, const double* ctrl,
                   int nthreads, float* sensor_out /* [nenv][nsensordata] or NULL */);

/* ---- plugin data paths restated from the reference's own sources (orc_plugins.cpp) ---- */
typedef struct OrcRobotHW OrcRobotHW;
OrcRobotHW* orc_hw_create(const b2mjModel* m, int nj, const int* joint_id, const int* method, const int* type,
                          const double* lower, const double* upper, const double* effort_limit, const double* pid6,
                          const b2mjJointLimits* limits, int literal_indexing);
void orc_hw_free(OrcRobotHW* hw);
void orc_hw_read(OrcRobotHW* hw, const b2mjModel* m, const OrcData* d);   /* DefaultRobotHWSim::readSim */
void orc_hw_write(OrcRobotHW* hw, const b2mjModel* m, OrcData* d, const double* cmd, int e_stop, double period);
void orc_hw_state(const OrcRobotHW* hw, double* pos, double* vel, double* eff);
void orc_sensor_readout(const b2mjModel* m, const OrcData* d, const int* flag, const double* mean, const double* sigma,
                        const double* normals, double* values, double* gt);   /* lastStageCallback arithmetic */

typedef struct OrcRolloutArgs {
  int nenv, nsteps, nthreads;
  double *qpos, *qvel, *act, *warm, *time; /* in/out, [nenv][nq|nv|na|nv|1]; act / warm / time may be NULL */
  const double* ctrl;                      /* [nsteps][nenv][nu] or NULL */
  float* sensor_out;                       /* [nenv][nsensordata] or NULL */
  int hw_njoint;                           /* 0 = no actuator-write path */
  const int *hw_joint_id, *hw_mode, *hw_kind;
  const double *hw_lower, *hw_upper, *hw_effort, *hw_pid6;
  const b2mjJointLimits* hw_limits;
  const double* hw_cmd;                    /* [nsteps][nenv][hw_njoint] */
  int hw_control_every;
} OrcRolloutArgs;
double orc_rollout_ex(const b2mjModel* m, const OrcRolloutArgs* a);

#ifdef __cplusplus
}
#endif
#endif
