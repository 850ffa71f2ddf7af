// b2mj_compat.hpp — a single env of the batch seen through an mjData-shaped struct.
//
// "existing ROS plugins see the same mjModel/mjData surface".  The reference's plugin callbacks receive
// (const mjModel*, mjData*) of its ONE env (plugin_utils.h:97-135) and their bodies are written against mjData member
// names: d->qpos[i], d->qvel[i], d->time, d->ctrl[i], d->qfrc_applied[dof], d->xfrc_applied[6*body+k], d->mocap_pos[...],
// d->sensordata[adr] (e.g. mujoco_ros_control default_robot_hw_sim.cpp:230-338, mujoco_ros_mocap mocap_plugin.cpp,
// mujoco_ros_sensors mujoco_sensor_handler_plugin.cpp:175-437).  EnvDataView carries exactly those member names for env
// `e` of a batch, and b2mjModel already carries the mjModel member names (include/b2mj.h), so such a body ports by
// changing the two type names.  SingleEnvPluginAdapter hosts ported bodies on the batched plugin interface: it
// materialises the view of one env (or of every env in turn), runs the body, and commits what the body may write --
// ctrl / qfrc_applied / xfrc_applied / mocap_* in the control hook, qfrc_passive in the passive hook -- the write
// contract of plugin_utils.h:89-105.  Host-side convenience for porting; device-resident plugins use b2mj_device_ptr.
#pragma once

#include <functional>
#include <utility>

#include "b2mj_env.hpp"

namespace b2mj_ros {

struct EnvDataView {
  int env = 0;
  // sizes the bodies commonly consult through the model; repeated here for convenience
  int nq = 0, nv = 0, na = 0, nu = 0, nbody = 0, nmocap = 0, nsensordata = 0;
  double time = 0;
  // state and inputs (mjData names); writable where the callback contract allows it
  double *qpos = nullptr, *qvel = nullptr, *act = nullptr, *qacc_warmstart = nullptr;
  double *ctrl = nullptr, *qfrc_applied = nullptr, *xfrc_applied = nullptr, *mocap_pos = nullptr, *mocap_quat = nullptr;
  // outputs of the last step
  const double *qacc = nullptr, *sensordata = nullptr, *act_dot = nullptr;
  // intermediates: filled inside the hooks of a split step (the arena image is readable there) or when the handle
  // keeps intermediates; nullptr otherwise
  const double *xpos = nullptr, *xquat = nullptr, *xmat = nullptr, *xipos = nullptr, *geom_xpos = nullptr,
               *geom_xmat = nullptr, *site_xpos = nullptr, *site_xmat = nullptr, *subtree_com = nullptr, *cvel = nullptr,
               *qfrc_bias = nullptr, *actuator_length = nullptr, *actuator_velocity = nullptr, *ten_length = nullptr;
  double* qfrc_passive = nullptr;  // passive hook: ADD to it
  int ncon = 0, nefc = 0;
};

// fill a view of env e from the batch mirrors (downloads each field once per hook invocation)
inline EnvDataView makeEnvDataView(const b2mjModel* m, BatchData* d, int e, bool with_intermediates) {
  EnvDataView v;
  v.env = e;
  v.nq = m->nq; v.nv = m->nv; v.na = m->na; v.nu = m->nu; v.nbody = m->nbody; v.nmocap = m->nmocap;
  v.nsensordata = m->nsensordata;
  v.time = d->time(e);
  v.qpos = d->row(B2MJ_F_QPOS, e);
  v.qvel = d->row(B2MJ_F_QVEL, e);
  v.act = m->na ? d->row(B2MJ_F_ACT, e) : nullptr;
  v.qacc_warmstart = d->row(B2MJ_F_QACC_WARMSTART, e);
  v.ctrl = m->nu ? d->row(B2MJ_F_CTRL, e) : nullptr;
  v.qfrc_applied = d->row(B2MJ_F_QFRC_APPLIED, e);
  v.xfrc_applied = d->row(B2MJ_F_XFRC_APPLIED, e);
  if (m->nmocap) { v.mocap_pos = d->row(B2MJ_F_MOCAP_POS, e); v.mocap_quat = d->row(B2MJ_F_MOCAP_QUAT, e); }
  v.qacc = d->row(B2MJ_F_QACC, e);
  v.sensordata = m->nsensordata ? d->row(B2MJ_F_SENSORDATA, e) : nullptr;
  v.act_dot = m->na ? d->row(B2MJ_F_ACT_DOT, e) : nullptr;
  if (with_intermediates) {
    v.xpos = d->row(B2MJ_F_XPOS, e); v.xquat = d->row(B2MJ_F_XQUAT, e); v.xmat = d->row(B2MJ_F_XMAT, e);
    v.xipos = d->row(B2MJ_F_XIPOS, e); v.geom_xpos = d->row(B2MJ_F_GEOM_XPOS, e); v.geom_xmat = d->row(B2MJ_F_GEOM_XMAT, e);
    v.site_xpos = d->row(B2MJ_F_SITE_XPOS, e); v.site_xmat = d->row(B2MJ_F_SITE_XMAT, e);
    v.subtree_com = d->row(B2MJ_F_SUBTREE_COM, e); v.cvel = d->row(B2MJ_F_CVEL, e);
    v.qfrc_bias = d->row(B2MJ_F_QFRC_BIAS, e); v.qfrc_passive = d->row(B2MJ_F_QFRC_PASSIVE, e);
    v.actuator_length = m->nu ? d->row(B2MJ_F_ACTUATOR_LENGTH, e) : nullptr;
    v.actuator_velocity = m->nu ? d->row(B2MJ_F_ACTUATOR_VELOCITY, e) : nullptr;
    v.ten_length = m->ntendon ? d->row(B2MJ_F_TEN_LENGTH, e) : nullptr;
    std::vector<int> nc = d->getInt(B2MJ_F_NCON), ne = d->getInt(B2MJ_F_NEFC);
    if (!nc.empty()) v.ncon = nc[e];
    if (!ne.empty()) v.nefc = ne[e];
  }
  return v;
}

// Hosts single-env callback bodies (reference signature, view instead of mjData) on the batched plugin interface.
class SingleEnvPluginAdapter : public BatchPlugin {
 public:
  using Body = std::function<void(const b2mjModel*, EnvDataView*)>;
  enum { ALL_ENVS = -1 };

  // env: which env the bodies see; ALL_ENVS runs them once per env (each sees its own view)
  explicit SingleEnvPluginAdapter(int env = 0) : env_(env) {}
  void setControlCallback(Body b) { control_ = std::move(b); }
  void setPassiveCallback(Body b) { passive_ = std::move(b); }
  void setLastStageCallback(Body b) { last_stage_ = std::move(b); }
  void setLoad(std::function<bool(const b2mjModel*, EnvDataView*)> f) { load_ = std::move(f); }
  void setReset(std::function<void()> f) { reset_ = std::move(f); }

  void controlCallback(const b2mjModel* m, BatchData* d) override {
    if (!control_) return;
    forEachEnv(m, d, true, control_);
    // what a control callback may write (plugin_utils.h:89-95)
    if (m->nu) d->commit(B2MJ_F_CTRL);
    d->commit(B2MJ_F_QFRC_APPLIED);
    d->commit(B2MJ_F_XFRC_APPLIED);
    if (m->nmocap) { d->commit(B2MJ_F_MOCAP_POS); d->commit(B2MJ_F_MOCAP_QUAT); }
  }
  void passiveCallback(const b2mjModel* m, BatchData* d) override {
    if (!passive_) return;
    forEachEnv(m, d, true, passive_);
    d->commit(B2MJ_F_QFRC_PASSIVE);  // plugin_utils.h:99-105: passive forces are ADDED to qfrc_passive
  }
  void lastStageCallback(const b2mjModel* m, BatchData* d) override {
    if (last_stage_) forEachEnv(m, d, false, last_stage_);  // read-only by convention (plugin_utils.h:118-124)
  }

 protected:
  bool load(const b2mjModel* m, BatchData* d) override {
    if (env_ != ALL_ENVS && (env_ < 0 || env_ >= d->nenv())) return false;
    if (!load_) return true;
    EnvDataView v = makeEnvDataView(m, d, env_ == ALL_ENVS ? 0 : env_, false);
    return load_(m, &v);
  }
  void reset() override { if (reset_) reset_(); }

 private:
  void forEachEnv(const b2mjModel* m, BatchData* d, bool intermediates, const Body& body) {
    const int lo = env_ == ALL_ENVS ? 0 : env_, hi = env_ == ALL_ENVS ? d->nenv() : env_ + 1;
    for (int e = lo; e < hi; e++) {
      EnvDataView v = makeEnvDataView(m, d, e, intermediates);
      body(m, &v);
    }
  }
  int env_;
  Body control_, passive_, last_stage_;
  std::function<bool(const b2mjModel*, EnvDataView*)> load_;
  std::function<void()> reset_;
};

}  // namespace b2mj_ros
