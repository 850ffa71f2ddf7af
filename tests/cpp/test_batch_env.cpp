// C++ tests of the host-side mirror (include/b2mj_env.hpp) written to read like the reference's gtests:
//   mujoco_ros/test/mujoco_env_test.cpp        (step guards :150-183, StepSingle/MultiWhilePaused :185-227,
//                                               StepUnblocked :229-260, num_steps exit :395-424, reset :480-530)
//   mujoco_ros/test/mujoco_ros_plugin_test.cpp (Control/Passive/LastCallback :97-121, failing load :141-170, reset)
//   mujoco_ros/test/ros_interface_test.cpp     (initial joint states :300-351)
//   mujoco_ros_sensors/test/mujoco_sensors_test.cpp (GT == sensordata/cutoff :326-328, noise :335-391)
//   mujoco_ros_control: control_period gating + writeSim modes (mujoco_ros_control_plugin.cpp:153-194)
// No gtest in the image: a tiny EXPECT macro set; exit code = number of failed expectations.
// Usage: test_batch_env <models_dir> [nenv]
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "b2mj_compat.hpp"
#include "b2mj_env.hpp"
#include "b2mj_plugins.hpp"

using namespace b2mj_ros;

static int g_fail = 0, g_checks = 0;
#define EXPECT_TRUE(c) do { g_checks++; if (!(c)) { g_fail++; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)
#define EXPECT_FALSE(c) EXPECT_TRUE(!(c))
#define EXPECT_NEAR(a, b, tol) do { g_checks++; const double _a = (a), _b = (b); if (!(std::fabs(_a - _b) <= (tol))) { g_fail++; \
  std::printf("FAIL %s:%d  |%s - %s| = |%.17g - %.17g| > %g\n", __FILE__, __LINE__, #a, #b, _a, _b, (double)(tol)); } } while (0)
#define EXPECT_DOUBLE_EQ(a, b) do { g_checks++; const double _a = (a), _b = (b); if (!(_a == _b)) { g_fail++; \
  std::printf("FAIL %s:%d  %s == %s  (%.17g vs %.17g)\n", __FILE__, __LINE__, #a, #b, _a, _b); } } while (0)

// mirror of mujoco_ros/test/test_plugin (test_plugin.h:45-75): records which hooks ran; optional failing load
class TestPlugin : public BatchPlugin {
 public:
  std::atomic_bool ran_reset{false}, ran_control_cb{false}, ran_passive_cb{false}, ran_last_cb{false},
      ran_on_geom_changed_cb{false};
  std::atomic_int n_control{0}, n_passive{0}, n_last{0};
  std::vector<int> order;  // 0 passive, 1 control, 2 last
  bool should_fail = false;
  double ctrl_value = 0.0;      // written into ctrl[0] of every env by the control hook
  double passive_add = 0.0;     // added to qfrc_passive[dof] of every env by the passive hook
  int passive_dof = 0;
  double seen_qvel_in_control = 0.0;

  void controlCallback(const b2mjModel* m, BatchData* d) override {
    ran_control_cb = true; n_control++; order.push_back(1);
    if (m->nv) seen_qvel_in_control = d->row(B2MJ_F_QVEL, 0)[0];
    if (m->nu && ctrl_value != 0.0) {
      for (int e = 0; e < d->nenv(); e++) d->row(B2MJ_F_CTRL, e)[0] = ctrl_value;
      d->commit(B2MJ_F_CTRL);
    }
  }
  void passiveCallback(const b2mjModel* /*m*/, BatchData* d) override {
    ran_passive_cb = true; n_passive++; order.push_back(0);
    if (passive_add != 0.0) {
      for (int e = 0; e < d->nenv(); e++) d->row(B2MJ_F_QFRC_PASSIVE, e)[passive_dof] += passive_add;
      d->commit(B2MJ_F_QFRC_PASSIVE);
    }
  }
  void lastStageCallback(const b2mjModel* /*m*/, BatchData* /*d*/) override { ran_last_cb = true; n_last++; order.push_back(2); }
  void onGeomChanged(const b2mjModel*, BatchData*, const int) override { ran_on_geom_changed_cb = true; }

 protected:
  bool load(const b2mjModel*, BatchData*) override { return !should_fail; }
  void reset() override { ran_reset = true; }
};

static std::string g_models;
static int g_nenv = 4;

static void test_step_guards() {  // mujoco_env_test.cpp:150-183
  BatchEnv env(g_nenv);
  EXPECT_FALSE(env.step(1));  // no model loaded
  EXPECT_TRUE(env.load(g_models + "/pendulum_scene.xml"));
  env.startPhysicsLoop();
  EXPECT_FALSE(env.step(0));
  EXPECT_FALSE(env.step(-10));
  env.settings_.run.store(1);
  EXPECT_FALSE(env.step(1));  // already running
  env.settings_.run.store(0);
  env.settings_.exit_request.store(1);
  env.waitForPhysicsJoin();
}

static void test_step_while_paused() {  // :185-227
  BatchEnv env(g_nenv);
  EXPECT_TRUE(env.load(g_models + "/pendulum_scene.xml"));
  env.startPhysicsLoop();
  EXPECT_TRUE(env.getOperationalStatus() == 0);
  { BatchEnv::MutexLock lock(env.physics_thread_mutex_); EXPECT_DOUBLE_EQ(env.getDataPtr()->time(0), 0.0); }
  EXPECT_TRUE(env.step(1));
  const double dt = env.getModelPtr()->opt.timestep;
  {
    BatchEnv::MutexLock lock(env.physics_thread_mutex_);
    env.getDataPtr()->invalidate();
    for (int e = 0; e < env.nenv(); e++) EXPECT_DOUBLE_EQ(env.getDataPtr()->time(e), dt);
  }
  EXPECT_TRUE(env.step(99));
  {
    BatchEnv::MutexLock lock(env.physics_thread_mutex_);
    env.getDataPtr()->invalidate();
    EXPECT_NEAR(env.getDataPtr()->time(0), 100 * dt, 1e-6);
  }
  // StepUnblocked (:229-260)
  EXPECT_TRUE(env.step(100, false));
  double waited = 0;
  while (env.settings_.env_steps_request.load() > 0 && waited < 5) { std::this_thread::sleep_for(std::chrono::milliseconds(2)); waited += 0.002; }
  EXPECT_TRUE(waited < 5);
  {
    BatchEnv::MutexLock lock(env.physics_thread_mutex_);
    env.getDataPtr()->invalidate();
    EXPECT_NEAR(env.getDataPtr()->time(0), 200 * dt, 1e-6);
  }
}

static void test_num_steps_exit() {  // :395-424: num_steps=100 => the loop stops at time = 100 dt
  BatchEnv env(g_nenv);
  EXPECT_TRUE(env.load(g_models + "/pendulum_scene.xml"));
  env.setNumStepsUntilExit(100);
  env.settings_.run.store(1);
  env.startPhysicsLoop();
  env.waitForPhysicsJoin();
  const double dt = env.getModelPtr()->opt.timestep;
  env.getDataPtr()->invalidate();
  EXPECT_NEAR(env.getDataPtr()->time(0), dt * 100, dt * 0.1);
  EXPECT_TRUE(env.stepCount() == 100);
}

static void test_reset() {  // :480-530
  BatchEnv env(g_nenv);
  EXPECT_TRUE(env.load(g_models + "/pendulum_scene.xml"));
  env.startPhysicsLoop();
  EXPECT_TRUE(env.step(100));
  const b2mjModel* m = env.getModelPtr();
  env.reset();
  EXPECT_FALSE(env.settings_.run.load());  // "Model should stay paused after reset!"
  {
    BatchEnv::MutexLock lock(env.physics_thread_mutex_);
    env.getDataPtr()->invalidate();
    EXPECT_NEAR(env.getDataPtr()->time(0), 0, 1e-6);
  }
  const int id2 = b2mj_name2id(m, B2MJ_OBJ_JOINT, "joint2");
  EXPECT_TRUE(id2 != -1);
  {
    BatchEnv::MutexLock lock(env.physics_thread_mutex_);
    BatchData* d = env.getDataPtr();
    d->invalidate();
    for (int e = 0; e < env.nenv(); e++) { d->row(B2MJ_F_QPOS, e)[m->jnt_qposadr[id2]] = 0.5; d->row(B2MJ_F_QVEL, e)[m->jnt_dofadr[id2]] = 0.1; }
    EXPECT_TRUE(d->commit(B2MJ_F_QPOS));
    EXPECT_TRUE(d->commit(B2MJ_F_QVEL));
    d->invalidate();
    EXPECT_DOUBLE_EQ(d->row(B2MJ_F_QPOS, env.nenv() - 1)[m->jnt_qposadr[id2]], 0.5);
  }
  env.reset();
  {
    BatchEnv::MutexLock lock(env.physics_thread_mutex_);
    BatchData* d = env.getDataPtr();
    d->invalidate();
    for (int e = 0; e < env.nenv(); e++) {
      EXPECT_DOUBLE_EQ(d->row(B2MJ_F_QPOS, e)[m->jnt_qposadr[id2]], 0.0);
      EXPECT_DOUBLE_EQ(d->row(B2MJ_F_QVEL, e)[m->jnt_dofadr[id2]], 0.0);
    }
  }
}

static void test_keyframes_and_binary_models() {  // viewer.cpp:1735-1745 "load key"; mujoco_env.cpp:771-911 extension dispatch
  const char* xml =
      "<mujoco><option timestep=\"0.002\"/><worldbody><body pos=\"0 0 1\"><joint name=\"h\" type=\"hinge\" axis=\"0 1 0\"/>"
      "<geom type=\"capsule\" fromto=\"0 0 0 0.3 0 0\" size=\"0.02\"/></body></worldbody>"
      "<keyframe><key name=\"bent\" time=\"0.5\" qpos=\"0.7\" qvel=\"-1.5\"/></keyframe></mujoco>";
  BatchEnv env(g_nenv);
  EXPECT_TRUE(env.loadFromString(xml));
  const b2mjModel* m = env.getModelPtr();
  EXPECT_TRUE(m->nkey == 1);
  EXPECT_TRUE(b2mj_name2id(m, B2MJ_OBJ_KEY, "bent") == 0);
  EXPECT_FALSE(env.loadKeyframe(3));
  EXPECT_TRUE(env.loadKeyframe(0));
  {
    BatchEnv::MutexLock lock(env.physics_thread_mutex_);
    BatchData* d = env.getDataPtr();
    d->invalidate();
    for (int e = 0; e < env.nenv(); e++) {
      EXPECT_DOUBLE_EQ(d->row(B2MJ_F_QPOS, e)[0], 0.7);
      EXPECT_DOUBLE_EQ(d->row(B2MJ_F_QVEL, e)[0], -1.5);
      EXPECT_DOUBLE_EQ(d->time(e), 0.5);
    }
  }
  // save as a binary model, load it through the extension dispatch, step both: identical
  const std::string path = "/tmp/b2mj_test_model.b2mjb";
  EXPECT_TRUE(b2mj_model_save_binary(m, path.c_str()) == B2MJ_OK);
  BatchEnv env2(g_nenv);
  EXPECT_TRUE(env2.load(path));
  EXPECT_TRUE(env2.getModelPtr()->nkey == 1);
  EXPECT_TRUE(env2.loadKeyframe(0));
  env.startPhysicsLoop();
  env2.startPhysicsLoop();
  EXPECT_TRUE(env.step(50));
  EXPECT_TRUE(env2.step(50));
  {
    BatchEnv::MutexLock l1(env.physics_thread_mutex_);
    BatchEnv::MutexLock l2(env2.physics_thread_mutex_);
    env.getDataPtr()->invalidate();
    env2.getDataPtr()->invalidate();
    EXPECT_DOUBLE_EQ(env.getDataPtr()->row(B2MJ_F_QPOS, 0)[0], env2.getDataPtr()->row(B2MJ_F_QPOS, 0)[0]);
    EXPECT_NEAR(env.getDataPtr()->time(0), 0.5 + 50 * 0.002, 1e-12);
  }
  BatchEnv env3(g_nenv);
  EXPECT_FALSE(env3.load("/tmp/does_not_exist.b2mjb"));
  std::remove(path.c_str());
}

static void test_plugin_callbacks() {  // mujoco_ros_plugin_test.cpp:97-121 + order + counts
  BatchEnv env(g_nenv);
  auto* tp = new TestPlugin();
  env.registerPlugin(BatchPluginPtr(tp), {{"type", "mujoco_ros/TestPlugin"}});
  auto* failing = new TestPlugin();
  failing->should_fail = true;
  env.registerPlugin(BatchPluginPtr(failing), {{"type", "mujoco_ros/TestPlugin"}, {"should_fail", "true"}});
  EXPECT_TRUE(env.load(g_models + "/pendulum_scene.xml"));
  EXPECT_TRUE(tp->type_ == "mujoco_ros/TestPlugin");
  EXPECT_TRUE(tp->loaded());
  EXPECT_FALSE(failing->loaded());
  env.startPhysicsLoop();
  EXPECT_FALSE(tp->ran_control_cb.load());
  EXPECT_FALSE(tp->ran_last_cb.load());
  EXPECT_TRUE(env.step());
  EXPECT_TRUE(tp->ran_control_cb.load());
  EXPECT_TRUE(tp->ran_passive_cb.load());
  EXPECT_TRUE(tp->ran_last_cb.load());
  // hook order within one step: passive (velocity stage) -> control (before actuation) -> lastStage
  EXPECT_TRUE(tp->order.size() == 3 && tp->order[0] == 0 && tp->order[1] == 1 && tp->order[2] == 2);
  EXPECT_TRUE(env.step(10));
  EXPECT_TRUE(tp->n_control.load() == 11 && tp->n_passive.load() == 11 && tp->n_last.load() == 11);
  // quarantined plugin never runs (plugin_utils.h:69-78)
  EXPECT_FALSE(failing->ran_control_cb.load());
  EXPECT_FALSE(failing->ran_last_cb.load());
  EXPECT_FALSE(tp->ran_reset.load());
  env.reset();
  EXPECT_TRUE(tp->ran_reset.load());
  EXPECT_FALSE(failing->ran_reset.load());
  env.notifyGeomChanged(0);
  EXPECT_TRUE(tp->ran_on_geom_changed_cb.load());
}

// the hooks act on the physics exactly where mjcb_control / mjcb_passive sit: a control written in the hook of
// step k drives step k; a passive force added in the hook equals the same force applied through qfrc_applied
static void test_hook_semantics() {
  const double dt_tol = 1e-12;
  // (a) ctrl from the control hook == ctrl set before a plain step
  double q_hook, q_plain;
  {
    BatchEnv env(g_nenv);
    auto* tp = new TestPlugin();
    tp->ctrl_value = 0.3;
    env.registerPlugin(BatchPluginPtr(tp));
    EXPECT_TRUE(env.load(g_models + "/panda_like.xml"));
    EXPECT_TRUE(env.step(5));
    env.getDataPtr()->invalidate();
    q_hook = env.getDataPtr()->row(B2MJ_F_QPOS, g_nenv - 1)[0];
    // the control hook sees the velocity-stage state of the step it is in
    EXPECT_TRUE(std::isfinite(tp->seen_qvel_in_control));
  }
  {
    BatchEnv env(g_nenv);
    EXPECT_TRUE(env.load(g_models + "/panda_like.xml"));
    BatchData* d = env.getDataPtr();
    for (int e = 0; e < g_nenv; e++) d->row(B2MJ_F_CTRL, e)[0] = 0.3;
    EXPECT_TRUE(d->commit(B2MJ_F_CTRL));
    EXPECT_TRUE(env.step(5));
    d->invalidate();
    q_plain = d->row(B2MJ_F_QPOS, g_nenv - 1)[0];
  }
  EXPECT_NEAR(q_hook, q_plain, dt_tol);
  EXPECT_TRUE(q_hook != 0.0);
  // (b) passive hook adding tau to qfrc_passive[joint2 dof] == qfrc_applied[joint2 dof] = tau
  double v_passive, v_applied;
  {
    BatchEnv env(g_nenv);
    auto* tp = new TestPlugin();
    env.registerPlugin(BatchPluginPtr(tp));
    EXPECT_TRUE(env.load(g_models + "/pendulum_scene.xml"));
    const b2mjModel* m = env.getModelPtr();
    tp->passive_dof = m->jnt_dofadr[b2mj_name2id(m, B2MJ_OBJ_JOINT, "joint2")];
    tp->passive_add = 0.05;
    EXPECT_TRUE(env.step(20));
    env.getDataPtr()->invalidate();
    v_passive = env.getDataPtr()->row(B2MJ_F_QVEL, 0)[tp->passive_dof];
  }
  {
    BatchEnv env(g_nenv);
    EXPECT_TRUE(env.load(g_models + "/pendulum_scene.xml"));
    const b2mjModel* m = env.getModelPtr();
    const int dof = m->jnt_dofadr[b2mj_name2id(m, B2MJ_OBJ_JOINT, "joint2")];
    BatchData* d = env.getDataPtr();
    for (int e = 0; e < g_nenv; e++) d->row(B2MJ_F_QFRC_APPLIED, e)[dof] = 0.05;
    EXPECT_TRUE(d->commit(B2MJ_F_QFRC_APPLIED));
    EXPECT_TRUE(env.step(20));
    d->invalidate();
    v_applied = d->row(B2MJ_F_QVEL, 0)[dof];
  }
  EXPECT_TRUE(v_applied != 0.0);
  EXPECT_NEAR(v_passive, v_applied, 1e-12);
}

static void test_initial_joint_states() {  // ros_interface_test.cpp:300-351
  BatchEnv env(g_nenv);
  env.setInitialJointPositions({{"joint1", "-1.57"}, {"joint2", "-0.66"}, {"ball_freejoint", "2.0 1.0 1.06 0.0 0.707 0.0 0.707"},
                                {"no_such_joint", "1.0"}, {"balljoint", "1 2"}});
  env.setInitialJointVelocities({{"joint2", "1.05"}, {"ball_freejoint", "1.0 2.0 3.0 10 20 30"}});
  EXPECT_TRUE(env.load(g_models + "/pendulum_scene.xml"));
  const b2mjModel* m = env.getModelPtr();
  BatchData* d = env.getDataPtr();
  d->invalidate();
  const int j1 = b2mj_name2id(m, B2MJ_OBJ_JOINT, "joint1"), j2 = b2mj_name2id(m, B2MJ_OBJ_JOINT, "joint2"),
            fj = b2mj_name2id(m, B2MJ_OBJ_JOINT, "ball_freejoint"), bj = b2mj_name2id(m, B2MJ_OBJ_JOINT, "balljoint");
  for (int e = 0; e < g_nenv; e += g_nenv - 1 > 0 ? g_nenv - 1 : 1) {
    const double* q = d->row(B2MJ_F_QPOS, e);
    const double* v = d->row(B2MJ_F_QVEL, e);
    EXPECT_DOUBLE_EQ(q[m->jnt_qposadr[j1]], -1.57);
    EXPECT_DOUBLE_EQ(q[m->jnt_qposadr[j2]], -0.66);
    const int qa = m->jnt_qposadr[fj], da = m->jnt_dofadr[fj];
    EXPECT_DOUBLE_EQ(q[qa], 2.0); EXPECT_DOUBLE_EQ(q[qa + 1], 1.0); EXPECT_DOUBLE_EQ(q[qa + 2], 1.06);
    // the forward pass after injection normalises the quaternion (reference tolerance 9e-4)
    EXPECT_NEAR(q[qa + 3], 0.0, 9e-4); EXPECT_NEAR(q[qa + 4], 0.707, 9e-4); EXPECT_NEAR(q[qa + 5], 0.0, 9e-4); EXPECT_NEAR(q[qa + 6], 0.707, 9e-4);
    EXPECT_DOUBLE_EQ(v[m->jnt_dofadr[j2]], 1.05);
    EXPECT_DOUBLE_EQ(v[da], 1.0); EXPECT_DOUBLE_EQ(v[da + 1], 2.0); EXPECT_DOUBLE_EQ(v[da + 2], 3.0);
    EXPECT_DOUBLE_EQ(v[da + 3], 10.0); EXPECT_DOUBLE_EQ(v[da + 4], 20.0); EXPECT_DOUBLE_EQ(v[da + 5], 30.0);
    // wrong value count ("1 2" for a ball joint) is ignored
    EXPECT_DOUBLE_EQ(q[m->jnt_qposadr[bj]], 1.0);
    if (g_nenv == 1) break;
  }
  // reset re-applies the injected state (mujoco_env.cpp:253)
  EXPECT_TRUE(env.step(3));
  env.reset();
  d->invalidate();
  EXPECT_DOUBLE_EQ(d->row(B2MJ_F_QPOS, 0)[m->jnt_qposadr[j1]], -1.57);
  EXPECT_NEAR(d->time(0), 0.0, 1e-12);
}

static void test_load_errors() {
  BatchEnv env(g_nenv);
  EXPECT_FALSE(env.load(g_models + "/does_not_exist.xml"));
  EXPECT_FALSE(env.loadError().empty());
  EXPECT_FALSE(env.loadFromString("<mujoco><worldbody><body><geom type='mesh'/></body></worldbody></mujoco>"));
  EXPECT_TRUE(env.getOperationalStatus() == 1);
  EXPECT_TRUE(env.loadFromString("<mujoco><option timestep='0.002'/><worldbody><body pos='0 0 1'><freejoint/><geom type='sphere' size='0.1'/></body></worldbody></mujoco>"));
  EXPECT_TRUE(env.step(10));
  env.getDataPtr()->invalidate();
  EXPECT_NEAR(env.getDataPtr()->time(0), 0.02, 1e-12);
}


// mujoco_ros_sensors as a BatchPlugin: the lastStage readout publishes float(sensordata / cutoff); registered
// noise shifts only the flagged dimensions (mujoco_sensors_test.cpp:326-328, 335-391)
static void test_sensors_plugin() {
  BatchEnv env(g_nenv);
  auto* sp = new BatchSensorsPlugin(/*seed=*/7);
  env.registerPlugin(BatchPluginPtr(sp), {{"type", "mujoco_ros_sensors/MujocoRosSensorsPlugin"}});
  EXPECT_TRUE(env.load(g_models + "/pendulum_scene.xml"));
  const b2mjModel* m = env.getModelPtr();
  EXPECT_TRUE((int)sp->records().size() == m->nsensor);
  int s_pos = -1, s_quat = -1, s_vel = -1, s_jv = -1;
  for (int i = 0; i < m->nsensor; i++) {
    const auto& r = sp->records()[i];
    if (r.name == "immovable_pos") s_pos = i;
    if (r.name == "immovable_quat") s_quat = i;
    if (r.name == "vel_EE") s_vel = i;
    if (r.name == "vel_joint2") s_jv = i;
  }
  EXPECT_TRUE(s_pos >= 0 && s_quat >= 0 && s_vel >= 0 && s_jv >= 0);
  EXPECT_TRUE(sp->records()[s_pos].msg == BatchSensorsPlugin::POINT_STAMPED && sp->records()[s_pos].frame_id == "world");
  EXPECT_TRUE(sp->records()[s_quat].msg == BatchSensorsPlugin::QUATERNION_STAMPED);
  EXPECT_TRUE(sp->records()[s_vel].msg == BatchSensorsPlugin::VECTOR3_STAMPED && sp->records()[s_vel].frame_id == "joint2_site");
  EXPECT_TRUE(sp->records()[s_jv].msg == BatchSensorsPlugin::SCALAR_STAMPED);
  // give joint2 a velocity so the velocimeter reads something
  {
    BatchData* d = env.getDataPtr();
    const int dof = m->jnt_dofadr[b2mj_name2id(m, B2MJ_OBJ_JOINT, "joint2")];
    for (int e = 0; e < g_nenv; e++) d->row(B2MJ_F_QVEL, e)[dof] = 0.5 + 0.01 * e;
    EXPECT_TRUE(d->commit(B2MJ_F_QVEL));
  }
  EXPECT_TRUE(env.step(5));
  EXPECT_TRUE(sp->readouts() == 5);
  BatchData* d = env.getDataPtr();
  d->invalidate();
  for (int e = 0; e < g_nenv; e++) {
    const double* sd = d->row(B2MJ_F_SENSORDATA, e);
    for (int i = 0; i < m->nsensor; i++) {
      const double cutoff = m->sensor_cutoff[i] > 0 ? m->sensor_cutoff[i] : 1.0;
      for (int k = 0; k < m->sensor_dim[i]; k++) {
        const double expect = (double)(float)(sd[m->sensor_adr[i] + k] / cutoff);
        EXPECT_DOUBLE_EQ(sp->value(e, i)[k], expect);
        EXPECT_DOUBLE_EQ(sp->groundTruth(e, i)[k], expect);
      }
    }
  }
  EXPECT_TRUE(std::fabs(sp->value(0, s_jv)[0]) > 1e-3);
  // noise on dims x (sigma only) and y (mean only) of vel_EE; z untouched
  BatchSensorsPlugin::NoiseModel nm;
  nm.sensor_name = "vel_EE"; nm.set_flag = 0x03; nm.mean[0] = 0.0; nm.std[0] = 0.025; nm.mean[1] = 1.0; nm.std[1] = 0.0;
  BatchSensorsPlugin::NoiseModel unknown;
  unknown.sensor_name = "no_such_sensor"; unknown.set_flag = 1;
  EXPECT_TRUE(sp->registerNoiseModels({nm, unknown}));
  EXPECT_TRUE(env.step(1));
  double sx = 0, sxx = 0;
  for (int e = 0; e < g_nenv; e++) {
    const double* v = sp->value(e, s_vel);
    const double* g = sp->groundTruth(e, s_vel);
    EXPECT_NEAR(v[1] - g[1], 1.0, 1e-6);
    EXPECT_DOUBLE_EQ(v[2], g[2]);
    sx += v[0] - g[0]; sxx += (v[0] - g[0]) * (v[0] - g[0]);
    EXPECT_DOUBLE_EQ(sp->value(e, s_pos)[0], sp->groundTruth(e, s_pos)[0]);  // other sensors stay noise-free
  }
  if (g_nenv >= 32) { EXPECT_TRUE(std::fabs(sx / g_nenv) < 0.02); EXPECT_TRUE(sxx / g_nenv > 1e-5 && sxx / g_nenv < 0.003); }
}

// mujoco_ros_control as a BatchPlugin: update gated to control_period, write every step, e-stop zeroes efforts
static void test_control_plugin() {
  BatchEnv env(g_nenv);
  std::vector<BatchRosControlPlugin::Joint> joints(2);
  joints[0].name = "joint1"; joints[0].control_mode = B2MJ_CTRL_EFFORT; joints[0].effort_limit = 87; joints[0].lower = -2.8973; joints[0].upper = 2.8973;
  joints[1].name = "joint4"; joints[1].control_mode = B2MJ_CTRL_POSITION_PID; joints[1].effort_limit = 87;
  joints[1].pid[0] = 300; joints[1].pid[2] = 30; joints[1].lower = -3.0718; joints[1].upper = -0.0698;
  int n_calls = 0; bool saw_reset = false; double last_period = -1;
  auto controller = [&](double /*time*/, double period, bool reset_ctrls, int nenv, int nj, const double* pos, const double* /*vel*/,
                        const double* /*eff*/, double* cmd) {
    n_calls++; saw_reset |= reset_ctrls; last_period = period;
    for (int e = 0; e < nenv; e++) { cmd[e * nj + 0] = 2.5; cmd[e * nj + 1] = -1.2; }
    (void)pos;
  };
  auto* cp = new BatchRosControlPlugin(joints, controller, /*control_period=*/0.01);
  env.registerPlugin(BatchPluginPtr(cp), {{"type", "mujoco_ros_control/MujocoRosControlPlugin"}});
  auto* missing = new BatchRosControlPlugin({{"no_such_joint", B2MJ_CTRL_EFFORT}}, controller);
  env.registerPlugin(BatchPluginPtr(missing));
  EXPECT_TRUE(env.load(g_models + "/panda_like.xml"));
  EXPECT_TRUE(cp->loaded());
  EXPECT_FALSE(missing->loaded());
  const b2mjModel* m = env.getModelPtr();
  const double dt = m->opt.timestep;  // 0.002
  EXPECT_TRUE(env.step(1));           // t = 0 inside the hook: no update, no write (ros::Time zero)
  EXPECT_TRUE(cp->updates() == 0 && cp->writes() == 0);
  EXPECT_TRUE(env.step(1));           // t = dt: first update (reset_ctrls) + first write
  EXPECT_TRUE(cp->updates() == 1 && cp->writes() == 1 && saw_reset);
  EXPECT_NEAR(last_period, dt, 1e-12);
  EXPECT_TRUE(env.step(20));
  EXPECT_TRUE(cp->writes() == 21);                       // writeSim every step
  EXPECT_TRUE(cp->updates() >= 4 && cp->updates() <= 5);  // controller only every control_period (5 steps)
  EXPECT_TRUE(n_calls == cp->updates());
  BatchData* d = env.getDataPtr();
  d->invalidate();
  const int d1 = m->jnt_dofadr[b2mj_name2id(m, B2MJ_OBJ_JOINT, "joint1")], j4 = b2mj_name2id(m, B2MJ_OBJ_JOINT, "joint4");
  for (int e = 0; e < g_nenv; e++) EXPECT_DOUBLE_EQ(d->row(B2MJ_F_QFRC_APPLIED, e)[d1], 2.5);  // EFFORT: qfrc_applied = cmd
  // POSITION_PID pulls joint4 from qpos0 towards the target (the model's own position actuator holds against it)
  const double q4_0 = m->qpos0[m->jnt_qposadr[j4]];
  EXPECT_TRUE(env.step(600));
  d->invalidate();
  const double q4 = d->row(B2MJ_F_QPOS, g_nenv - 1)[m->jnt_qposadr[j4]];
  EXPECT_TRUE(std::fabs(q4 - (-1.2)) < 0.8 * std::fabs(q4_0 - (-1.2)));
  // e-stop: efforts go to zero on the next write (default_robot_hw_sim.cpp:271-275)
  cp->eStopActive(true);
  EXPECT_TRUE(env.step(6));  // spans a controller update, where the e-stop is latched (:177-179)
  d->invalidate();
  EXPECT_DOUBLE_EQ(d->row(B2MJ_F_QFRC_APPLIED, 0)[d1], 0.0);
  cp->eStopActive(false);
  saw_reset = false;
  EXPECT_TRUE(env.step(6));
  EXPECT_TRUE(saw_reset);  // controllers are reset after an e-stop release (:181-184)
  d->invalidate();
  EXPECT_DOUBLE_EQ(d->row(B2MJ_F_QFRC_APPLIED, 0)[d1], 2.5);
}

// A single-env callback body written against mjData member names (what the reference's plugins look like: they get
// (const mjModel*, mjData*) of the one env, plugin_utils.h:97-135) hosted on the batch through the compat view.
// The body below is a joint-space PD controller + a Cartesian push on a body + a passive damper, i.e. it touches
// d->time, d->qpos, d->qvel, d->ctrl, d->xfrc_applied, d->xpos and d->qfrc_passive by name.
static int g_body_calls = 0;
static void single_env_control_body(const b2mjModel* m, EnvDataView* d) {
  g_body_calls++;
  const int j = b2mj_name2id(m, B2MJ_OBJ_JOINT, "joint2");
  const int qadr = m->jnt_qposadr[j], dadr = m->jnt_dofadr[j];
  const double target = -0.3 + 0.02 * d->env;                      // each env its own target
  d->ctrl[1] = target + 0.2 * (target - d->qpos[qadr]) - 0.01 * d->qvel[dadr];
  const int b = b2mj_name2id(m, B2MJ_OBJ_BODY, "link7");
  if (b > 0 && d->xpos) d->xfrc_applied[6 * b + 2] = d->xpos[3 * b + 2] > 0.2 ? -1.0 : 0.0;  // push down while above 20 cm
}
static void single_env_passive_body(const b2mjModel* m, EnvDataView* d) {
  if (d->qfrc_passive) d->qfrc_passive[m->nv - 1] += -0.05 * d->qvel[m->nv - 1];
}
static void test_single_env_compat_view() {
  BatchEnv env(g_nenv);
  auto* ad = new SingleEnvPluginAdapter(SingleEnvPluginAdapter::ALL_ENVS);
  double seen_time = -1;
  int last_calls = 0;
  ad->setControlCallback(single_env_control_body);
  ad->setPassiveCallback(single_env_passive_body);
  ad->setLastStageCallback([&](const b2mjModel*, EnvDataView* d) { seen_time = d->time; last_calls++; });
  env.registerPlugin(BatchPluginPtr(ad), {{"type", "compat/SingleEnv"}});
  EXPECT_TRUE(env.load(g_models + "/panda_like.xml"));
  EXPECT_TRUE(ad->loaded());
  const b2mjModel* m = env.getModelPtr();
  g_body_calls = 0;
  EXPECT_TRUE(env.step(200));
  EXPECT_TRUE(g_body_calls == 200 * g_nenv);
  EXPECT_TRUE(last_calls == 200 * g_nenv);
  EXPECT_NEAR(seen_time, 200 * m->opt.timestep, 1e-9);
  BatchData* d = env.getDataPtr();
  d->invalidate();
  const int j = b2mj_name2id(m, B2MJ_OBJ_JOINT, "joint2");
  for (int e = 0; e < g_nenv; e++) {
    const double target = -0.3 + 0.02 * e;
    // the position servo on joint2 follows the per-env target the body wrote into ITS env's ctrl row
    EXPECT_NEAR(d->row(B2MJ_F_QPOS, e)[m->jnt_qposadr[j]], target, 0.08);
    EXPECT_TRUE(std::fabs(d->row(B2MJ_F_CTRL, e)[1] - target) < 0.1);
  }
  if (g_nenv > 1) EXPECT_TRUE(std::fabs(d->row(B2MJ_F_QPOS, 0)[m->jnt_qposadr[j]] - d->row(B2MJ_F_QPOS, g_nenv - 1)[m->jnt_qposadr[j]]) > 1e-3);
  // one env only: the others keep zero controls
  BatchEnv env1(g_nenv);
  auto* one = new SingleEnvPluginAdapter(0);
  one->setControlCallback([](const b2mjModel*, EnvDataView* v) { v->ctrl[0] = 0.7; });
  env1.registerPlugin(BatchPluginPtr(one));
  EXPECT_TRUE(env1.load(g_models + "/panda_like.xml"));
  EXPECT_TRUE(env1.step(3));
  BatchData* d1 = env1.getDataPtr();
  d1->invalidate();
  EXPECT_DOUBLE_EQ(d1->row(B2MJ_F_CTRL, 0)[0], 0.7);
  if (g_nenv > 1) EXPECT_DOUBLE_EQ(d1->row(B2MJ_F_CTRL, g_nenv - 1)[0], 0.0);
}

int main(int argc, char** argv) {
  if (argc < 2) { std::printf("usage: %s <models_dir> [nenv]\n", argv[0]); return 2; }
  g_models = argv[1];
  if (argc > 2) g_nenv = std::atoi(argv[2]);
  if (b2mj_device_count() <= 0) {
    // no GPU: the product path must fail loudly, never fall back (checked by the CPU test-suite)
    BatchEnv env(g_nenv);
    const bool ok = env.load(g_models + "/pendulum_scene.xml");
    std::printf("NO_DEVICE load=%d error=%s\n", (int)ok, env.loadError().c_str());
    return ok ? 1 : 0;
  }
  test_step_guards();
  test_step_while_paused();
  test_num_steps_exit();
  test_reset();
  test_keyframes_and_binary_models();
  test_plugin_callbacks();
  test_hook_semantics();
  test_initial_joint_states();
  test_load_errors();
  test_sensors_plugin();
  test_control_plugin();
  test_single_env_compat_view();
  std::printf("%s: %d checks, %d failed\n", g_fail ? "FAILED" : "OK", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
