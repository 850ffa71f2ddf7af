// b2mj_env.hpp — C++ host side above the C-ABI: the batched counterpart of the reference's
// `mujoco_ros::MujocoEnv` stepping surface and `mujoco_ros::MujocoPlugin` callback contract.
//
// Header-only, C++17, depends on nothing but <b2mj.h> + libb2mj.so.  Names, argument meaning and error
// behaviour mirror the reference so its tests read the same here:
//   BatchEnv::step(num_steps, blocking)   <- MujocoEnv::step            mujoco_ros/src/mujoco_env.cpp:913-945
//   BatchEnv::physicsLoop()               <- MujocoEnv::physicsLoop     mujoco_env.cpp:436-639 (ROS time sync,
//                                            viewers and offscreen rendering dropped: SURVEY.md section 8 "out")
//   BatchEnv::resetSim()                  <- MujocoEnv::resetSim        mujoco_env.cpp:246-264 (no 100 ms ROS flush)
//   BatchEnv::loadInitialJointStates()    <- MujocoEnv::loadInitialJointStates  mujoco_env.cpp:266-402
//   BatchEnv::togglePaused / settings_    <- mujoco_env.h:168-206, mujoco_env.cpp (togglePaused)
//   BatchPlugin                           <- MujocoPlugin               mujoco_ros/include/mujoco_ros/plugin_utils.h:45-161
//   runControlCbs / runPassiveCbs / runLastStageCbs <- callbacks.cpp:131-157
//
// What differs, by design: one BatchEnv owns `nenv` independent environments on one GPU, so callbacks
// receive (const b2mjModel*, BatchData*) instead of (const mjModel*, mjData*).  BatchData gives the
// mjData field names as [nenv][n] host mirrors (download on first read, upload on commit) plus the
// device handle for device-resident plugins.  The control hook sits exactly where mjcb_control fires
// (after the velocity stage, before actuation: mujoco_env.h:242-246) by splitting the step with
// b2mj_step_begin / b2mj_step_end; passive callbacks add to qfrc_passive at the same point (nothing
// between mj_passive and the hook reads qfrc_passive); lastStage callbacks run once per full step.
// Unlike the reference (process-global mjcb_* routed through MujocoEnv::instance, mujoco_env.h:241-251)
// there are no globals: any number of BatchEnvs may live in one process.
#pragma once

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "b2mj.h"

namespace b2mj_ros {

class BatchEnv;

// ---- BatchData: the mjData surface of a whole batch --------------------------------------------------
// Field mirrors are row-major [nenv][count] (b2mj_field_size).  `get` downloads once per hook invocation
// (the env invalidates the cache after every device step); `commit` uploads a field the plugin changed.
class BatchData {
 public:
  BatchData(b2mj_handle* h, const b2mjModel* m) : h_(h), m_(m), nenv_(b2mj_nenv(h)) {}
  b2mj_handle* handle() const { return h_; }
  const b2mjModel* model() const { return m_; }
  int nenv() const { return nenv_; }

  // elements per env of a field (0 if absent for this model)
  int count(b2mj_field f) const { int is_int = 0; const int n = b2mj_field_size(m_, f, &is_int); return n < 0 ? 0 : n; }

  // host mirror of a float64 field, downloaded on first use since the last device step; nullptr on error
  double* get(b2mj_field f) {
    Mirror& mr = mirror_[(int)f];
    const int n = count(f);
    if (!mr.valid) {
      mr.buf.resize((size_t)nenv_ * n);
      if (n && b2mj_get(h_, f, mr.buf.data(), mr.buf.size() * sizeof(double)) != B2MJ_OK) return nullptr;
      mr.valid = true;
    }
    return mr.buf.data();
  }
  // row of env `e` (mjData-style access: d->row(B2MJ_F_QPOS, e)[i] is env e's qpos[i])
  double* row(b2mj_field f, int e) {
    double* p = get(f);
    return p ? p + (size_t)e * count(f) : nullptr;
  }
  // upload the (edited) mirror of a settable field: ctrl, qfrc_applied, xfrc_applied, mocap_*, qpos, qvel,
  // act, qacc_warmstart, time; qfrc_passive only inside the passive hook of a split step
  bool commit(b2mj_field f) {
    Mirror& mr = mirror_[(int)f];
    if (!mr.valid) return false;
    return mr.buf.empty() || b2mj_set(h_, f, mr.buf.data(), mr.buf.size() * sizeof(double)) == B2MJ_OK;
  }
  // int32 fields (ncon, nefc, solver_iter, warning, contact_geom1, ...)
  std::vector<int> getInt(b2mj_field f) {
    std::vector<int> out((size_t)nenv_ * count(f));
    if (!out.empty() && b2mj_get(h_, f, out.data(), out.size() * sizeof(int)) != B2MJ_OK) out.clear();
    return out;
  }
  double time(int e = 0) { double* t = get(B2MJ_F_TIME); return t ? t[e] : 0.0; }
  void invalidate() { for (auto& kv : mirror_) kv.second.valid = false; }

 private:
  struct Mirror { std::vector<double> buf; bool valid = false; };
  b2mj_handle* h_;
  const b2mjModel* m_;
  int nenv_;
  std::map<int, Mirror> mirror_;
};

// ---- BatchPlugin: mirror of mujoco_ros::MujocoPlugin (plugin_utils.h:45-161) ----------------------------
class BatchPlugin {
 public:
  using Config = std::map<std::string, std::string>;  // stands in for the XmlRpcValue plugin config
  virtual ~BatchPlugin() = default;

  // called directly after plugin creation (plugin_utils.h:51-57)
  void init(const Config& config, BatchEnv* env_ptr) {
    config_ = config;
    env_ptr_ = env_ptr;
    auto it = config_.find("type");
    type_ = it == config_.end() ? std::string() : it->second;
  }
  std::string type_;

  // plugin_utils.h:69-78: a plugin whose load() fails is ignored until the next load attempt
  bool safe_load(const b2mjModel* m, BatchData* d) {
    loading_successful_ = load(m, d);
    return loading_successful_;
  }
  // plugin_utils.h:83-87
  void safe_reset() {
    if (loading_successful_) reset();
  }
  bool loaded() const { return loading_successful_; }

  // plugin_utils.h:89-97: write ctrl / qfrc_applied / xfrc_applied (then d->commit(field))
  virtual void controlCallback(const b2mjModel* /*model*/, BatchData* /*data*/) {}
  // plugin_utils.h:99-107: ADD to qfrc_passive (then d->commit(B2MJ_F_QFRC_PASSIVE))
  virtual void passiveCallback(const b2mjModel* /*model*/, BatchData* /*data*/) {}
  // plugin_utils.h:118-126: end of a full step; not called for RK4 sub-steps
  virtual void lastStageCallback(const b2mjModel* /*model*/, BatchData* /*data*/) {}
  // plugin_utils.h:128-135
  virtual void onGeomChanged(const b2mjModel* /*model*/, BatchData* /*data*/, const int /*geom_id*/) {}

 protected:
  BatchPlugin() = default;
  virtual bool load(const b2mjModel* m, BatchData* d) = 0;  // plugin_utils.h:146
  virtual void reset() = 0;                                  // plugin_utils.h:151
  Config config_;
  BatchEnv* env_ptr_ = nullptr;

 private:
  bool loading_successful_ = false;
};
using BatchPluginPtr = std::unique_ptr<BatchPlugin>;

// ---- BatchEnv: mirror of mujoco_ros::MujocoEnv's stepping surface ---------------------------------------
class BatchEnv {
 public:
  using MutexLock = std::unique_lock<std::recursive_mutex>;

  explicit BatchEnv(int nenv = 1, int device = 0) : nenv_(nenv), device_(device) {}
  ~BatchEnv() {
    settings_.exit_request.store(1);
    waitForPhysicsJoin();
    unload();
  }
  BatchEnv(const BatchEnv&) = delete;
  BatchEnv& operator=(const BatchEnv&) = delete;

  std::recursive_mutex physics_thread_mutex_;  // mujoco_env.h:157

  struct {  // mujoco_env.h:168-199 (render / viewer / real-time fields dropped)
    bool eval_mode = false;
    std::atomic_int run = {0};
    std::atomic_int exit_request = {0};
    std::atomic_int reset_request = {0};
    std::atomic_int env_steps_request = {0};
  } settings_;

  // ---- model loading (mujoco_env.cpp:771-911: file path or XML string through the VFS) ----
  // viewer.cpp:1735-1745 ("load key": copy of key_qpos / key_qvel / key_act / key_mpos / key_mquat into mjData, here
  // with mj_resetDataKeyframe semantics) for the whole batch.  Returns false for a key the model does not have.
  bool loadKeyframe(int key) {
    MutexLock lock(physics_thread_mutex_);
    if (!handle_ || b2mj_reset_keyframe(handle_, key, nullptr) != B2MJ_OK) return false;
    data_->invalidate();
    return true;
  }

  bool load(const std::string& filename) { return loadImpl(filename, false); }
  bool loadFromString(const std::string& xml) { return loadImpl(xml, true); }
  const std::string& loadError() const { return load_error_; }

  // ---- plugins (mujoco_env.h:265-267; registration replaces pluginlib discovery) ----
  void registerPlugin(BatchPluginPtr plugin, const BatchPlugin::Config& config = {}) {
    MutexLock lock(physics_thread_mutex_);
    plugin->init(config, this);
    if (model_ && plugin->safe_load(model_, data_.get())) cb_ready_plugins_.push_back(plugin.get());
    plugins_.push_back(std::move(plugin));
  }
  const std::vector<BatchPluginPtr>& getPlugins() const { return plugins_; }

  // ---- threads (mujoco_env.cpp:641-660) ----
  void startPhysicsLoop() {
    if (physics_thread_handle_.joinable()) return;
    settings_.exit_request.store(0);
    is_physics_running_.store(1);  // before the thread exists: a blocking step() right after must wait for it
    physics_thread_handle_ = std::thread(&BatchEnv::physicsLoop, this);
  }
  void waitForPhysicsJoin() {
    if (physics_thread_handle_.joinable()) physics_thread_handle_.join();
  }
  bool isPhysicsRunning() const { return is_physics_running_.load() != 0; }

  // mujoco_env.cpp:913-945 — same guards, same order, same return values
  bool step(int num_steps = 1, bool blocking = true) {
    if (!model_) return false;                 // "No model loaded. Cannot step"
    if (settings_.run.load()) return false;    // "Simulation is already running. Ignoring request"
    if (num_steps <= 0) return false;          // "Number of steps must be positive. Ignoring request"
    if (blocking && std::this_thread::get_id() == physics_thread_handle_.get_id())
      return false;                            // "Simulation is running in the same thread. Cannot block!"
    if (!physics_thread_handle_.joinable()) {
      // no physics thread (library use without startPhysicsLoop): run the request inline
      MutexLock lock(physics_thread_mutex_);
      for (int k = 0; k < num_steps; k++)
        if (!simStep()) return false;
      return true;
    }
    settings_.env_steps_request.store(num_steps);
    if (blocking)
      while (settings_.env_steps_request.load() > 0 && is_physics_running_.load())
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    return true;
  }

  // mujoco_env.cpp togglePaused (admin hash / eval mode dropped)
  bool togglePaused(bool paused) {
    settings_.run.store(paused ? 0 : 1);
    return true;
  }

  // request handled by the physics thread between steps, or inline when no thread runs (mujoco_env.cpp:246-264)
  void reset() {
    if (physics_thread_handle_.joinable() && is_physics_running_.load()) {
      settings_.reset_request.store(1);
      while (settings_.reset_request.load() && is_physics_running_.load())
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    } else {
      MutexLock lock(physics_thread_mutex_);
      resetSim();
    }
  }

  // initial joint states: joint name -> space-separated values, exactly the rosparam maps
  // initial_joint_positions/joint_map and initial_joint_velocities/joint_map (mujoco_env.cpp:266-389)
  void setInitialJointPositions(const std::map<std::string, std::string>& m) { initial_joint_positions_ = m; }
  void setInitialJointVelocities(const std::map<std::string, std::string>& m) { initial_joint_velocities_ = m; }

  // mujoco_env.h:211 / mujoco_env.cpp:163-176: overwrite the narrowphase of a geom-type pair; the defaults come back
  // on the next load (prepareReload, :949-954).  collfn: B2MJ_COLLFN_* (device code cannot call a host function).
  bool registerCollisionFunction(int geom_type1, int geom_type2, int collfn) {
    if (!handle_) return false;  // like the reference: plugins register from load(), i.e. once a model is loaded
    custom_collisions_.push_back({geom_type1, geom_type2, collfn});
    return b2mj_register_collision_function(handle_, geom_type1, geom_type2, collfn) == B2MJ_OK;
  }

  size_t numCustomCollisions() const { return custom_collisions_.size(); }
  void setNumStepsUntilExit(int n) { num_steps_until_exit_ = n; }  // mujoco_env.h:271 (-1 = no limit)

  const b2mjModel* getModelPtr() const { return model_; }
  BatchData* getDataPtr() { return data_.get(); }
  b2mj_handle* getHandle() { return handle_; }
  int nenv() const { return nenv_; }
  // 0 = ready, 1 = loading (mujoco_env.h getOperationalStatus, without the visual-init state)
  int getOperationalStatus() const { return model_ ? 0 : 1; }
  uint64_t stepCount() const { return step_count_.load(); }

  // proxies to the step hooks (callbacks.cpp:131-157)
  void runControlCbs() { for (BatchPlugin* p : cb_ready_plugins_) p->controlCallback(model_, data_.get()); }
  void runPassiveCbs() { for (BatchPlugin* p : cb_ready_plugins_) p->passiveCallback(model_, data_.get()); }
  void runLastStageCbs() { for (BatchPlugin* p : cb_ready_plugins_) p->lastStageCallback(model_, data_.get()); }
  void notifyGeomChanged(int geom_id) { for (BatchPlugin* p : cb_ready_plugins_) p->onGeomChanged(model_, data_.get(), geom_id); }

  // re-derive constants and re-upload the model after host-side edits (callbacks.cpp:254,582: mj_setConst)
  bool updateModel(b2mjModel* edited) {
    MutexLock lock(physics_thread_mutex_);
    if (!handle_ || b2mj_model_set_const(edited) != B2MJ_OK) return false;
    return b2mj_model_update(handle_, edited) == B2MJ_OK;
  }

 protected:
  // one full batched step with the hooks in reference order:
  //   [check, position, velocity incl. mj_passive] -> passive cbs -> control cbs -> [actuation, acceleration,
  //   constraint, integrate] -> time publish stub -> lastStage cbs      (mujoco_env.cpp:593-595 + mjcb_* order)
  bool simStep() {
    bool ok;
    const bool hooks = !cb_ready_plugins_.empty();
    if (hooks) {
      // split step: the hooks run between the velocity stage and actuation of EVERY forward pass -- once for Euler,
      // four times for RK4 (b2mj_step_end returns B2MJ_AGAIN after each of the first three sub-steps), exactly where
      // mj_step fires mjcb_passive / mjcb_control (plugin_utils.h:89-105)
      ok = b2mj_step_begin(handle_) == B2MJ_OK;
      while (ok) {
        data_->invalidate();
        runPassiveCbs();
        runControlCbs();
        const int rc = b2mj_step_end(handle_);
        if (rc == B2MJ_AGAIN) continue;
        ok = rc == B2MJ_OK;
        break;
      }
    } else {
      ok = b2mj_step(handle_, 1) == B2MJ_OK;
    }
    if (!ok) { load_error_ = b2mj_last_error(); return false; }
    data_->invalidate();
    step_count_.fetch_add(1);
    runLastStageCbs();
    return true;
  }

  void resetSim() {
    if (!handle_) return;
    b2mj_reset(handle_, nullptr);
    data_->invalidate();
    loadInitialJointStates();
    for (auto& plugin : plugins_) plugin->safe_reset();
    settings_.reset_request.store(0);
  }

  // mujoco_env.cpp:266-402.  Values are broadcast to every env of the batch.
  void loadInitialJointStates() {
    auto parse = [](const std::string& s, std::vector<double>& out) {
      std::stringstream ss(s);
      std::string tok;
      while (std::getline(ss, tok, ' '))
        if (!tok.empty()) out.push_back(std::stod(tok));
    };
    for (const auto& kv : initial_joint_positions_) {
      const int id = b2mj_name2id(model_, B2MJ_OBJ_JOINT, kv.first.c_str());
      if (id == -1) continue;  // "Joint with name ... could not be found"
      int num_axes = 0;
      switch (model_->jnt_type[id]) {
        case B2MJ_JNT_FREE: num_axes = 7; break;
        case B2MJ_JNT_BALL: num_axes = 4; break;
        case B2MJ_JNT_SLIDE: case B2MJ_JNT_HINGE: num_axes = 1; break;
        default: continue;
      }
      std::vector<double> vals;
      parse(kv.second, vals);
      if ((int)vals.size() != num_axes) continue;  // "... don't match the degrees of freedom of the joint"
      for (int a = 0; a < num_axes; a++) setJointPosition(vals[a], id, a);
      data_->commit(B2MJ_F_QPOS); data_->commit(B2MJ_F_QVEL); data_->commit(B2MJ_F_QFRC_APPLIED);
      b2mj_forward(handle_);  // "Apply changes in forward dynamics" (:329)
      data_->invalidate();
    }
    for (const auto& kv : initial_joint_velocities_) {
      const int id = b2mj_name2id(model_, B2MJ_OBJ_JOINT, kv.first.c_str());
      if (id == -1) continue;
      int num_axes = 0;
      switch (model_->jnt_type[id]) {
        case B2MJ_JNT_FREE: num_axes = 6; break;
        case B2MJ_JNT_BALL: num_axes = 3; break;
        case B2MJ_JNT_SLIDE: case B2MJ_JNT_HINGE: num_axes = 1; break;
        default: continue;
      }
      std::vector<double> vals;
      parse(kv.second, vals);
      if ((int)vals.size() != num_axes) continue;
      for (int a = 0; a < num_axes; a++) setJointVelocity(vals[a], id, a);
      data_->commit(B2MJ_F_QVEL); data_->commit(B2MJ_F_QFRC_APPLIED);
      data_->invalidate();
    }
  }
  // mujoco_env.cpp:391-402: setting a position zeroes that dof's velocity and applied force
  void setJointPosition(double pos, int joint_id, int jnt_axis = 0) {
    const int qa = model_->jnt_qposadr[joint_id] + jnt_axis, da = model_->jnt_dofadr[joint_id] + jnt_axis;
    const bool has_dof = da < model_->nv;
    for (int e = 0; e < nenv_; e++) {
      data_->row(B2MJ_F_QPOS, e)[qa] = pos;
      // the reference indexes qvel / qfrc_applied with dofadr + axis for all 7 (free) / 4 (ball) position axes;
      // clamp to the joint's dofs so the last quaternion axis does not touch the next joint's dof
      const int ndof = model_->jnt_type[joint_id] == B2MJ_JNT_FREE ? 6 : model_->jnt_type[joint_id] == B2MJ_JNT_BALL ? 3 : 1;
      if (has_dof && jnt_axis < ndof) {
        data_->row(B2MJ_F_QVEL, e)[da] = 0;
        data_->row(B2MJ_F_QFRC_APPLIED, e)[da] = 0;
      }
    }
  }
  void setJointVelocity(double vel, int joint_id, int jnt_axis = 0) {
    const int da = model_->jnt_dofadr[joint_id] + jnt_axis;
    for (int e = 0; e < nenv_; e++) {
      data_->row(B2MJ_F_QVEL, e)[da] = vel;
      data_->row(B2MJ_F_QFRC_APPLIED, e)[da] = 0;
    }
  }

  // mujoco_env.cpp:436-639 without wall-clock sync: running -> step as fast as possible; paused -> serve
  // env_steps_request, else mj_forward (keeps derived quantities fresh, :619-623)
  void physicsLoop() {
    is_physics_running_.store(1);
    while (!settings_.exit_request.load() && num_steps_until_exit_ != 0) {
      if (settings_.run.load()) std::this_thread::yield();
      else std::this_thread::sleep_for(std::chrono::milliseconds(1));
      if (!model_) continue;
      if (!physics_thread_mutex_.try_lock()) continue;
      if (settings_.reset_request.load()) resetSim();
      if (settings_.run.load()) {
        if (!simStep()) settings_.run.store(0);
        if (num_steps_until_exit_ > 0) num_steps_until_exit_--;
      } else if (settings_.env_steps_request.load() > 0) {
        const double t_sync = data_->time(0);
        while (settings_.env_steps_request.load() > 0 && !settings_.exit_request.load()) {
          if (!simStep()) { settings_.env_steps_request.store(0); break; }
          settings_.env_steps_request.fetch_sub(1);
          if (data_->time(0) < t_sync) break;  // "Break if reset" (:614-616)
        }
      } else {
        b2mj_forward(handle_);
        b2mj_sync(handle_);
        data_->invalidate();
      }
      physics_thread_mutex_.unlock();
    }
    is_physics_running_.store(0);
  }

  bool loadImpl(const std::string& src, bool is_string) {
    MutexLock lock(physics_thread_mutex_);
    b2mjModel* m = nullptr;
    // files dispatch on their extension like the reference's loader (mujoco_env.cpp:771-911: .mjb -> mj_loadModel,
    // else mj_loadXML): ".b2mjb" is this library's binary model format, anything else is MJCF
    const int rc = is_string ? b2mj_model_from_xml_string(src.c_str(), &m) : b2mj_model_from_file(src.c_str(), &m);
    if (rc != B2MJ_OK) { load_error_ = b2mj_last_error(); return false; }
    b2mj_handle* h = nullptr;
    if (b2mj_create(m, nenv_, device_, &h) != B2MJ_OK) {
      load_error_ = b2mj_last_error();
      b2mj_model_free(m);
      return false;
    }
    unload();  // prepareReload (mujoco_env.cpp:947-960): drop callbacks of the old model first
    model_ = m;
    handle_ = h;
    data_.reset(new BatchData(handle_, model_));
    loadInitialJointStates();
    cb_ready_plugins_.clear();
    for (auto& plugin : plugins_)
      if (plugin->safe_load(model_, data_.get())) cb_ready_plugins_.push_back(plugin.get());
    return true;
  }
  void unload() {
    custom_collisions_.clear();  // prepareReload: "Resetting collision cbs to default" (mujoco_env.cpp:949-954)
    cb_ready_plugins_.clear();
    data_.reset();
    if (handle_) { b2mj_destroy(handle_); handle_ = nullptr; }
    if (model_) { b2mj_model_free(model_); model_ = nullptr; }
  }

  int nenv_, device_;
  b2mjModel* model_ = nullptr;        // common_types.h:74 mjModelPtr
  b2mj_handle* handle_ = nullptr;     // common_types.h:79 mjDataPtr, for the whole batch
  std::unique_ptr<BatchData> data_;
  std::vector<BatchPlugin*> cb_ready_plugins_;  // objects managed by plugins_ (mujoco_env.h:265)
  std::vector<BatchPluginPtr> plugins_;
  int num_steps_until_exit_ = -1;
  struct CustomCollision { int t1, t2, fn; };
  std::vector<CustomCollision> custom_collisions_;  // mujoco_env.h custom_collisions_ (cleared by every load)
  std::map<std::string, std::string> initial_joint_positions_, initial_joint_velocities_;
  std::thread physics_thread_handle_;
  std::atomic_int is_physics_running_ = {0};
  std::atomic<uint64_t> step_count_ = {0};
  std::string load_error_;
};

}  // namespace b2mj_ros
