// b2mj_plugins.hpp — the reference's two shipped plugins, re-hosted on the batched plugin interface
// (b2mj_env.hpp) with their per-step arithmetic running as device kernels behind the C-ABI:
//
//   BatchRosControlPlugin  <- mujoco_ros_control::MujocoRosControlPlugin::controlCallback
//                             (mujoco_ros_control/src/mujoco_ros_control_plugin.cpp:153-194) driving
//                             DefaultRobotHWSim::readSim / writeSim (default_robot_hw_sim.cpp:230-326)
//                             -> b2mj_robot_hw_read / b2mj_robot_hw_write
//   BatchSensorsPlugin     <- mujoco_ros::sensors::MujocoRosSensorsPlugin
//                             (mujoco_ros_sensors/src/mujoco_sensor_handler_plugin.cpp: load :60-118,
//                             registerNoiseModelsCB :120-173, lastStageCallback :175-437)
//                             -> b2mj_sensor_configure_noise / b2mj_sensor_readout
//
// ROS transport (controller_manager, publishers, services) is out of scope: the
// controller_manager::update call becomes a user-supplied BatchController, the per-sensor publishers become a
// typed record table + one [nenv][nsensordata] slab a thin publisher can memcpy from.
#pragma once

#include <functional>
#include <string>
#include <vector>

#include "b2mj_env.hpp"

namespace b2mj_ros {

// ---- mujoco_ros_control ---------------------------------------------------------------------------------
// What ControllerManager::update(time, period, reset_ctrls) is to the reference: reads joint state, writes
// commands.  All arrays are [nenv][njoint] row-major; cmd is what writeSim consumes (effort, position or
// velocity set-points according to each joint's control mode).
using BatchController = std::function<void(double time, double period, bool reset_ctrls, int nenv, int njoint,
                                           const double* pos, const double* vel, const double* eff, double* cmd)>;

class BatchRosControlPlugin : public BatchPlugin {
 public:
  struct Joint {
    std::string name;            // MuJoCo joint name (transmission joint)
    int control_mode;            // B2MJ_CTRL_* (hardware interface of the transmission)
    double effort_limit = 1e30;  // URDF <limit effort=...>
    double pid[5] = {0, 0, 0, 0, 0};  // p, i, d, i_max, i_min (gazebo_ros_control/pid_gains)
    int joint_kind = 0;          // 0 revolute (limited), 1 continuous, 2 prismatic
    double lower = 0, upper = 0;
    // joint_limits_interface data of the joint (URDF <limit>/<safety_controller> + joint_limits rosparams,
    // default_robot_hw_sim.cpp:340-446).  With has_limits the joint gets the saturation handle of its hardware interface,
    // or the soft-limits handle when limits.has_soft_limits is set; enforced before every write (:262-267).
    bool has_limits = false;
    b2mjJointLimits limits{};
    bool pid_antiwindup = false;
  };

  BatchRosControlPlugin(std::vector<Joint> joints, BatchController controller, double control_period = 0.0)
      : joints_(std::move(joints)), controller_(std::move(controller)), control_period_(control_period) {}

  void eStopActive(bool active) { e_stop_active_ = active; }  // emergency-stop topic (:63-66)
  int njoint() const { return (int)joints_.size(); }
  const std::vector<double>& lastCommand() const { return cmd_; }

  // mujoco_ros_control_plugin.cpp:153-194, with sim time taken from the batch (env 0; all envs share dt)
  void controlCallback(const b2mjModel* /*model*/, BatchData* data) override {
    const double sim_time = data->time(0);
    if (sim_time < last_update_sim_time_) {  // time reset (:160-169)
      last_update_sim_time_ = sim_time;
      last_write_sim_time_ = sim_time;
    }
    const double sim_period = sim_time - last_update_sim_time_;
    bool reset_ctrls = last_update_sim_time_ == 0.0;  // ros::Time::isZero()
    if (sim_period >= control_period_ || (reset_ctrls && sim_period != 0.0)) {
      last_update_sim_time_ = sim_time;
      b2mj_robot_hw_read(data->handle(), pos_.data(), vel_.data(), eff_.data());  // readSim
      if (e_stop_active_) last_e_stop_active_ = true;
      else if (last_e_stop_active_) { reset_ctrls = true; last_e_stop_active_ = false; }
      if (controller_) controller_(sim_time, sim_period, reset_ctrls, data->nenv(), njoint(), pos_.data(), vel_.data(), eff_.data(), cmd_.data());
      n_updates_++;
    }
    if (last_update_sim_time_ != 0.0 && sim_time > last_write_sim_time_) {  // writeSim every step (:190-193)
      b2mj_robot_hw_write(data->handle(), cmd_.data(), /*is_device=*/0, e_stop_active_ ? 1 : 0, sim_time - last_write_sim_time_);
      last_write_sim_time_ = sim_time;
      n_writes_++;
    }
  }
  int updates() const { return n_updates_; }
  int writes() const { return n_writes_; }

 protected:
  bool load(const b2mjModel* m, BatchData* d) override {
    const int n = njoint();
    std::vector<int> ids(n), modes(n), kinds(n), aw(n);
    std::vector<double> elim(n), pid(5 * n), lo(n), hi(n);
    std::vector<b2mjJointLimits> lims(n);
    bool any_limits = false;
    for (int j = 0; j < n; j++) {
      ids[j] = b2mj_name2id(m, B2MJ_OBJ_JOINT, joints_[j].name.c_str());
      if (ids[j] < 0) return false;  // "This transmission has no associated joints" -> plugin quarantined
      modes[j] = joints_[j].control_mode;
      kinds[j] = joints_[j].joint_kind;
      elim[j] = joints_[j].effort_limit;
      for (int k = 0; k < 5; k++) pid[5 * j + k] = joints_[j].pid[k];
      lo[j] = joints_[j].lower;
      hi[j] = joints_[j].upper;
      aw[j] = joints_[j].pid_antiwindup ? 1 : 0;
      lims[j] = joints_[j].limits;
      any_limits |= joints_[j].has_limits;
    }
    if (any_limits) {
      // the C-ABI takes limits for all joints or none: a joint without limits gets a handle that cannot bind
      // (no position / acceleration limits, infinite velocity / effort bounds)
      for (int j = 0; j < n; j++)
        if (!joints_[j].has_limits) {
          lims[j] = b2mjJointLimits{};
          lims[j].has_velocity_limits = lims[j].has_effort_limits = 1;
          lims[j].max_velocity = lims[j].max_effort = 1.7976931348623157e308;
        }
    }
    b2mjRobotHW cfg{n, ids.data(), modes.data(), elim.data(), pid.data(), lo.data(), hi.data(), kinds.data(),
                    any_limits ? lims.data() : nullptr, aw.data()};
    if (b2mj_robot_hw_configure(d->handle(), &cfg) != B2MJ_OK) return false;
    const size_t sz = (size_t)d->nenv() * n;
    pos_.assign(sz, 0); vel_.assign(sz, 0); eff_.assign(sz, 0); cmd_.assign(sz, 0);
    reset();
    return true;
  }
  void reset() override {
    last_update_sim_time_ = 0; last_write_sim_time_ = 0; last_e_stop_active_ = false;
    n_updates_ = 0; n_writes_ = 0;
  }

 private:
  std::vector<Joint> joints_;
  BatchController controller_;
  double control_period_;
  bool e_stop_active_ = false, last_e_stop_active_ = false;
  double last_update_sim_time_ = 0, last_write_sim_time_ = 0;
  int n_updates_ = 0, n_writes_ = 0;
  std::vector<double> pos_, vel_, eff_, cmd_;
};

// ---- mujoco_ros_sensors ---------------------------------------------------------------------------------
class BatchSensorsPlugin : public BatchPlugin {
 public:
  // message type each MuJoCo sensor type is published as (mujoco_sensor_handler_plugin.cpp:439-620 initSensors)
  enum MsgType { SCALAR_STAMPED = 0, VECTOR3_STAMPED = 1, POINT_STAMPED = 2, QUATERNION_STAMPED = 3 };
  struct Record {
    std::string name;      // sensor name = topic suffix
    std::string frame_id;  // site / body / world frame the reference stamps the message with
    int type;              // B2MJ_SENS_*
    MsgType msg;
    int adr, dim;          // slice of the [nsensordata] row
  };
  struct NoiseModel {  // mujoco_ros_msgs/SensorNoiseModel.msg
    std::string sensor_name;
    double mean[3] = {0, 0, 0}, std[3] = {0, 0, 0};
    int set_flag = 0;
  };

  explicit BatchSensorsPlugin(uint64_t seed = 0) : seed_(seed) {}

  // RegisterSensorNoiseModels service (:120-173); unknown sensor names are skipped with a warning there
  bool registerNoiseModels(const std::vector<NoiseModel>& models) {
    if (!data_) return false;
    for (const NoiseModel& nm : models) {
      const int id = b2mj_name2id(data_->model(), B2MJ_OBJ_SENSOR, nm.sensor_name.c_str());
      if (id < 0) continue;
      b2mjSensorNoise* slot = nullptr;
      for (auto& s : noise_) if (s.sensor_id == id) slot = &s;
      if (!slot) { noise_.push_back(b2mjSensorNoise{id, {0, 0, 0}, {0, 0, 0}, 0}); slot = &noise_.back(); }
      int idx = 0;  // the message packs only the flagged dimensions, in order (:150-166)
      for (int bit = 0; bit < 3; bit++)
        if (nm.set_flag & (1 << bit)) { slot->mean[idx] = nm.mean[idx]; slot->sigma[idx] = nm.std[idx]; if (bit < 2) idx++; }
      slot->set_flag |= nm.set_flag;
    }
    return b2mj_sensor_configure_noise(data_->handle(), noise_.data(), (int)noise_.size(), seed_) == B2MJ_OK;
  }

  // :175-437 — one device kernel fills both slabs: values as published (float32-rounded, noise applied) and the
  // ground truth (train mode only; eval mode publishes none: :65-67)
  void lastStageCallback(const b2mjModel* /*model*/, BatchData* data) override {
    const bool want_gt = !eval_mode_;
    b2mj_sensor_readout(data->handle(), values_.data(), want_gt ? gt_.data() : nullptr);
    n_readouts_++;
  }

  const std::vector<Record>& records() const { return records_; }
  // [nenv][nsensordata]
  const std::vector<double>& values() const { return values_; }
  const std::vector<double>& groundTruth() const { return gt_; }
  const double* value(int env, int sensor) const { return values_.data() + (size_t)env * nsd_ + records_[sensor].adr; }
  const double* groundTruth(int env, int sensor) const { return gt_.data() + (size_t)env * nsd_ + records_[sensor].adr; }
  int readouts() const { return n_readouts_; }

 protected:
  bool load(const b2mjModel* m, BatchData* d) override {
    data_ = d;
    nsd_ = m->nsensordata;
    eval_mode_ = config_.count("eval_mode") && config_["eval_mode"] == "true";
    records_.clear();
    for (int i = 0; i < m->nsensor; i++) {
      Record r;
      const char* nm = b2mj_id2name(m, B2MJ_OBJ_SENSOR, i);
      r.name = nm ? nm : "";
      r.type = m->sensor_type[i];
      r.adr = m->sensor_adr[i];
      r.dim = m->sensor_dim[i];
      const char* fr = nullptr;
      switch (r.type) {  // :455-600: frame sensors report in the reference frame (world if none), site sensors in the site
        case B2MJ_SENS_FRAMEPOS: r.msg = POINT_STAMPED; break;
        case B2MJ_SENS_FRAMEQUAT: case B2MJ_SENS_BALLQUAT: r.msg = QUATERNION_STAMPED; break;
        default: r.msg = r.dim == 3 ? VECTOR3_STAMPED : SCALAR_STAMPED;
      }
      if (m->sensor_objtype[i] == B2MJ_OBJ_SITE) fr = b2mj_id2name(m, B2MJ_OBJ_SITE, m->sensor_objid[i]);
      if (r.type >= B2MJ_SENS_FRAMEPOS && r.type <= B2MJ_SENS_FRAMEANGACC)
        fr = m->sensor_refid[i] >= 0 ? b2mj_id2name(m, m->sensor_reftype[i], m->sensor_refid[i]) : "world";
      if (r.type == B2MJ_SENS_SUBTREECOM || r.type == B2MJ_SENS_SUBTREELINVEL || r.type == B2MJ_SENS_SUBTREEANGMOM) fr = "world";
      r.frame_id = fr ? fr : "world";
      records_.push_back(r);
    }
    values_.assign((size_t)d->nenv() * nsd_, 0.0);
    gt_.assign((size_t)d->nenv() * nsd_, 0.0);
    noise_.clear();
    return true;
  }
  void reset() override {}

 private:
  uint64_t seed_;
  BatchData* data_ = nullptr;
  int nsd_ = 0;
  bool eval_mode_ = false;
  int n_readouts_ = 0;
  std::vector<Record> records_;
  std::vector<b2mjSensorNoise> noise_;
  std::vector<double> values_, gt_;
};

}  // namespace b2mj_ros
