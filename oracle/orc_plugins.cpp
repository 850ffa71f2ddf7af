// orc_plugins.cpp — CPU restatement of the two plugin data paths either side of the step.  TEST INFRASTRUCTURE ONLY.
//
// Unlike the physics (libmujoco is absent), the sources of these paths ARE in /root/reference, so this file follows
// them line by line and the GPU kernels (mujoco_ros_pkgs_b200/csrc/host/plugins.cu) are compared against it bitwise:
//
//   DefaultRobotHWSim::readSim        mujoco_ros_control/src/default_robot_hw_sim.cpp:230-246
//   DefaultRobotHWSim::writeSim       mujoco_ros_control/src/default_robot_hw_sim.cpp:248-326
//   DefaultRobotHWSim::getJointData   :333-338
//   registerJointLimits               :340-446 (which limit handle a joint gets)
//   initial joint state               :132-137 (joint_position_ = 1.0, joint_effort_ = 1.0, commands 0)
//   MujocoRosSensorsPlugin::lastStageCallback   mujoco_ros_sensors/src/mujoco_sensor_handler_plugin.cpp:175-437
//
// Three ROS packages the reference links are NOT in /root/reference; their published algorithms are restated here
// (ROS noetic versions, the distribution the reference's CI builds against: .github/workflows/ci.yaml):
//   angles 1.9.13                     normalize_angle, shortest_angular_distance, two_pi_complement,
//                                     find_min_max_delta, shortest_angular_distance_with_limits
//   control_toolbox 1.19.0            Pid::computeCommand(error, dt) incl. antiwindup
//   ros_control 0.19 joint_limits_interface   Position/Velocity/Effort JointSaturationHandle and
//                                     JointSoftLimitsHandle ::enforceLimits
//
// Indexing: the reference writes d->qfrc_applied[m->jnt_dofadr[j]] with j = TRANSMISSION index and d->qpos with the
// same dof address (:273-321); reads use the MuJoCo joint id (:237).  `literal_indexing` reproduces that; the default
// (0) uses the joint id for writes as well.  The two agree whenever transmissions are listed in MuJoCo joint order
// over a hinge/slide-only prefix (every model the reference ships; tests/test_gpu_plugins.py checks it on hand_like).
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "orc_types.h"

namespace {

// ---------------------------------------------------------------- angles (ros/angles, angles.h)
inline double normalize_angle(double angle) {
  const double result = std::fmod(angle + M_PI, 2.0 * M_PI);
  if (result <= 0.0) return result + M_PI;
  return result - M_PI;
}
inline double shortest_angular_distance(double from, double to) { return normalize_angle(to - from); }
inline double two_pi_complement(double angle) {
  if (angle > 2 * M_PI || angle < -2.0 * M_PI) angle = std::fmod(angle, 2.0 * M_PI);
  if (angle < 0) return (2 * M_PI + angle);
  else if (angle > 0) return (-2 * M_PI + angle);
  return (2 * M_PI);
}
bool find_min_max_delta(double from, double left_limit, double right_limit, double& result_min_delta,
                        double& result_max_delta) {
  double delta[4];
  delta[0] = shortest_angular_distance(from, left_limit);
  delta[1] = shortest_angular_distance(from, right_limit);
  delta[2] = two_pi_complement(delta[0]);
  delta[3] = two_pi_complement(delta[1]);
  if (delta[0] == 0) {
    result_min_delta = delta[0];
    result_max_delta = std::fmax(delta[1], delta[3]);
    return true;
  }
  if (delta[1] == 0) {
    result_max_delta = delta[1];
    result_min_delta = std::fmin(delta[0], delta[2]);
    return true;
  }
  double delta_min = delta[0], delta_min_2pi = delta[2];
  if (delta[2] < delta_min) { delta_min = delta[2]; delta_min_2pi = delta[0]; }
  double delta_max = delta[1], delta_max_2pi = delta[3];
  if (delta[3] > delta_max) { delta_max = delta[3]; delta_max_2pi = delta[1]; }
  if ((delta_min <= delta_max_2pi) || (delta_max >= delta_min_2pi)) {
    result_min_delta = delta_max_2pi;
    result_max_delta = delta_min_2pi;
    if (left_limit == -M_PI && right_limit == M_PI) return true;
    return false;
  }
  result_min_delta = delta_min;
  result_max_delta = delta_max;
  return true;
}
bool shortest_angular_distance_with_limits(double from, double to, double left_limit, double right_limit,
                                           double& shortest_angle) {
  double min_delta = -2 * M_PI, max_delta = 2 * M_PI, min_delta_to = -2 * M_PI, max_delta_to = 2 * M_PI;
  const bool flag = find_min_max_delta(from, left_limit, right_limit, min_delta, max_delta);
  const double delta = shortest_angular_distance(from, to);
  const double delta_mod_2pi = two_pi_complement(delta);
  if (flag) {  // from position is within the limits
    if (delta >= min_delta && delta <= max_delta) { shortest_angle = delta; return true; }
    if (delta_mod_2pi >= min_delta && delta_mod_2pi <= max_delta) { shortest_angle = delta_mod_2pi; return true; }
    find_min_max_delta(to, left_limit, right_limit, min_delta_to, max_delta_to);
    if (std::fabs(min_delta_to) < std::fabs(max_delta_to)) shortest_angle = std::fmax(delta, delta_mod_2pi);
    else if (std::fabs(min_delta_to) > std::fabs(max_delta_to)) shortest_angle = std::fmin(delta, delta_mod_2pi);
    else shortest_angle = std::fabs(delta) < std::fabs(delta_mod_2pi) ? delta : delta_mod_2pi;
    return false;
  }
  // from position is outside the limits
  find_min_max_delta(to, left_limit, right_limit, min_delta_to, max_delta_to);
  if (std::fabs(min_delta) < std::fabs(max_delta)) shortest_angle = std::fmin(delta, delta_mod_2pi);
  else if (std::fabs(min_delta) > std::fabs(max_delta)) shortest_angle = std::fmax(delta, delta_mod_2pi);
  else shortest_angle = std::fabs(delta) < std::fabs(delta_mod_2pi) ? delta : delta_mod_2pi;
  return false;
}

inline double saturate(double val, double lo, double hi) { return std::fmin(std::fmax(val, lo), hi); }

// ---------------------------------------------------------------- control_toolbox::Pid
struct Pid {
  double p = 0, i = 0, d = 0, i_max = 0, i_min = 0;
  int antiwindup = 0;
  double p_error_last = 0, p_error = 0, i_error = 0, d_error = 0;
  double computeCommand(double error, double dt) {
    if (dt == 0.0 || std::isnan(error) || std::isinf(error)) return 0.0;
    double error_dot = d_error;
    if (dt > 0.0) {
      error_dot = (error - p_error_last) / dt;
      p_error_last = error;
    }
    p_error = error;
    d_error = error_dot;
    if (std::isnan(error_dot) || std::isinf(error_dot)) return 0.0;
    const double p_term = p * p_error;
    i_error += dt * p_error;
    if (antiwindup && i != 0) {
      const double a = i_min / i, b = i_max / i;
      i_error = saturate(i_error, std::fmin(a, b), std::fmax(a, b));
    }
    double i_term = i * i_error;
    if (!antiwindup) i_term = saturate(i_term, i_min, i_max);
    const double d_term = d * d_error;
    return p_term + i_term + d_term;
  }
};

enum { EFFORT = B2MJ_CTRL_EFFORT, POSITION = B2MJ_CTRL_POSITION, POSITION_PID = B2MJ_CTRL_POSITION_PID,
       VELOCITY = B2MJ_CTRL_VELOCITY, VELOCITY_PID = B2MJ_CTRL_VELOCITY_PID };
enum { REVOLUTE = 0, CONTINUOUS = 1, PRISMATIC = 2 };

struct Joint {
  int id, method, type;
  double lower, upper, effort_limit;
  Pid pid;
  bool has_handle = false;
  b2mjJointLimits lim{};
  double prev_cmd;  // limit-handle state (NaN for position handles until the first call, 0 for velocity saturation)
  double position = 1.0, velocity = 0.0, effort = 1.0;                      // :132-134
  double effort_command = 0.0, position_command = 0.0, velocity_command = 0.0;  // :135-137
  double last_position_command = 0.0;
};

}  // namespace

struct OrcRobotHW {
  std::vector<Joint> joints;
  bool last_e_stop_active = false;
  int literal_indexing = 0;
};

namespace {

// joint_limits_interface: one handle per joint, chosen by the hardware interface (:411-445)
void enforceLimits(Joint& j, double period) {
  if (!j.has_handle) return;
  const b2mjJointLimits& L = j.lim;
  const int iface = (j.method == POSITION_PID) ? POSITION : (j.method == VELOCITY_PID) ? VELOCITY : j.method;
  if (iface == POSITION) {
    if (!L.has_soft_limits) {  // PositionJointSaturationHandle
      if (std::isnan(j.prev_cmd)) j.prev_cmd = j.position;
      const double lo_lim = L.has_position_limits ? L.min_position : -std::numeric_limits<double>::max();
      const double hi_lim = L.has_position_limits ? L.max_position : std::numeric_limits<double>::max();
      double min_pos, max_pos;
      if (L.has_velocity_limits) {
        const double delta_pos = L.max_velocity * period;
        min_pos = std::fmax(j.prev_cmd - delta_pos, lo_lim);
        max_pos = std::fmin(j.prev_cmd + delta_pos, hi_lim);
      } else {
        min_pos = lo_lim;
        max_pos = hi_lim;
      }
      const double cmd = saturate(j.position_command, min_pos, max_pos);
      j.position_command = cmd;
      j.prev_cmd = cmd;
    } else {  // PositionJointSoftLimitsHandle
      if (std::isnan(j.prev_cmd)) j.prev_cmd = j.position;
      const double pos = j.prev_cmd;
      double soft_min_vel, soft_max_vel;
      if (L.has_position_limits) {
        soft_min_vel = saturate(-L.k_position * (pos - L.soft_min_position), -L.max_velocity, L.max_velocity);
        soft_max_vel = saturate(-L.k_position * (pos - L.soft_max_position), -L.max_velocity, L.max_velocity);
      } else {
        soft_min_vel = -L.max_velocity;
        soft_max_vel = L.max_velocity;
      }
      double pos_low = pos + soft_min_vel * period, pos_high = pos + soft_max_vel * period;
      if (L.has_position_limits) {
        pos_low = std::fmax(pos_low, L.min_position);
        pos_high = std::fmin(pos_high, L.max_position);
      }
      j.position_command = saturate(j.position_command, pos_low, pos_high);
      j.prev_cmd = j.position_command;
    }
  } else if (iface == VELOCITY) {
    if (!L.has_soft_limits) {  // VelocityJointSaturationHandle
      double vel_low, vel_high;
      if (L.has_acceleration_limits) {
        vel_low = std::fmax(j.prev_cmd - L.max_acceleration * period, -L.max_velocity);
        vel_high = std::fmin(j.prev_cmd + L.max_acceleration * period, L.max_velocity);
      } else {
        vel_low = -L.max_velocity;
        vel_high = L.max_velocity;
      }
      j.velocity_command = saturate(j.velocity_command, vel_low, vel_high);
      j.prev_cmd = j.velocity_command;
    } else {  // VelocityJointSoftLimitsHandle
      double min_vel, max_vel;
      if (L.has_position_limits) {
        const double pos = j.position;
        min_vel = saturate(-L.k_position * (pos - L.soft_min_position), -L.max_velocity, L.max_velocity);
        max_vel = saturate(-L.k_position * (pos - L.soft_max_position), -L.max_velocity, L.max_velocity);
      } else {
        min_vel = -L.max_velocity;
        max_vel = L.max_velocity;
      }
      if (L.has_acceleration_limits) {
        const double vel = j.velocity;
        min_vel = std::fmax(vel - L.max_acceleration * period, min_vel);
        max_vel = std::fmin(vel + L.max_acceleration * period, max_vel);
      }
      j.velocity_command = saturate(j.velocity_command, min_vel, max_vel);
    }
  } else {  // EFFORT
    if (!L.has_soft_limits) {  // EffortJointSaturationHandle
      double min_eff = -L.max_effort, max_eff = L.max_effort;
      if (L.has_position_limits) {
        const double pos = j.position;
        if (pos < L.min_position) min_eff = 0.0;
        else if (pos > L.max_position) max_eff = 0.0;
      }
      const double vel = j.velocity;
      if (vel < -L.max_velocity) min_eff = 0.0;
      else if (vel > L.max_velocity) max_eff = 0.0;
      j.effort_command = saturate(j.effort_command, min_eff, max_eff);
    } else {  // EffortJointSoftLimitsHandle
      const double pos = j.position, vel = j.velocity;
      double soft_min_vel, soft_max_vel;
      if (L.has_position_limits) {
        soft_min_vel = saturate(-L.k_position * (pos - L.soft_min_position), -L.max_velocity, L.max_velocity);
        soft_max_vel = saturate(-L.k_position * (pos - L.soft_max_position), -L.max_velocity, L.max_velocity);
      } else {
        soft_min_vel = -L.max_velocity;
        soft_max_vel = L.max_velocity;
      }
      const double soft_min_eff = saturate(-L.k_velocity * (vel - soft_min_vel), -L.max_effort, L.max_effort);
      const double soft_max_eff = saturate(-L.k_velocity * (vel - soft_max_vel), -L.max_effort, L.max_effort);
      j.effort_command = saturate(j.effort_command, soft_min_eff, soft_max_eff);
    }
  }
}

}  // namespace

extern "C" {

// pid: [nj][6] = p, i, d, i_max, i_min, antiwindup; limits: [nj] or NULL (no handle registered)
OrcRobotHW* orc_hw_create(const b2mjModel* m, int nj, const int* joint_id, const int* method, const int* type,
                          const double* lower, const double* upper, const double* effort_limit, const double* pid,
                          const b2mjJointLimits* limits, int literal_indexing) {
  OrcRobotHW* hw = new OrcRobotHW();
  hw->literal_indexing = literal_indexing;
  hw->joints.resize(nj);
  for (int k = 0; k < nj; k++) {
    Joint& j = hw->joints[k];
    j.id = joint_id[k];
    j.method = method[k];
    j.type = type[k];
    j.lower = lower ? lower[k] : -std::numeric_limits<double>::max();
    j.upper = upper ? upper[k] : std::numeric_limits<double>::max();
    j.effort_limit = effort_limit ? effort_limit[k] : std::numeric_limits<double>::max();
    if (pid) {
      const double* g = pid + 6 * k;
      j.pid.p = g[0]; j.pid.i = g[1]; j.pid.d = g[2]; j.pid.i_max = g[3]; j.pid.i_min = g[4];
      j.pid.antiwindup = g[5] != 0;
    }
    if (limits) { j.has_handle = true; j.lim = limits[k]; }
    const int iface = (j.method == POSITION_PID) ? POSITION : (j.method == VELOCITY_PID) ? VELOCITY : j.method;
    j.prev_cmd = iface == POSITION ? std::numeric_limits<double>::quiet_NaN() : 0.0;
  }
  (void)m;
  return hw;
}
void orc_hw_free(OrcRobotHW* hw) { delete hw; }

// readSim (:230-246)
void orc_hw_read(OrcRobotHW* hw, const b2mjModel* m, const OrcData* d) {
  for (Joint& j : hw->joints) {
    const double position = d->qpos[m->jnt_qposadr[j.id]];  // getJointData (:333-338)
    const double velocity = d->qvel[m->jnt_dofadr[j.id]];
    const double effort = d->qfrc_applied[m->jnt_dofadr[j.id]];
    if (j.type == PRISMATIC) j.position = position;
    else j.position += shortest_angular_distance(j.position, position);
    j.velocity = velocity;
    j.effort = effort;
  }
}

// writeSim (:248-326).  cmd[k] is the command of joint k's own interface (effort, position or velocity).
void orc_hw_write(OrcRobotHW* hw, const b2mjModel* m, OrcData* d, const double* cmd, int e_stop, double period) {
  const int nj = (int)hw->joints.size();
  for (int k = 0; k < nj; k++) {
    Joint& j = hw->joints[k];
    const int iface = (j.method == POSITION_PID) ? POSITION : (j.method == VELOCITY_PID) ? VELOCITY : j.method;
    if (iface == EFFORT) j.effort_command = cmd[k];
    else if (iface == POSITION) j.position_command = cmd[k];
    else j.velocity_command = cmd[k];
  }
  if (e_stop) {  // :250-259: position-controlled joints hold the command of the moment the stop became active
    if (!hw->last_e_stop_active) {
      for (Joint& j : hw->joints) j.last_position_command = j.position_command;
      hw->last_e_stop_active = true;
    }
    for (Joint& j : hw->joints) j.position_command = j.last_position_command;
  } else {
    hw->last_e_stop_active = false;
  }
  for (Joint& j : hw->joints) enforceLimits(j, period);  // :262-267 (each joint is in at most one interface)
  for (int k = 0; k < nj; k++) {
    Joint& j = hw->joints[k];
    const int dofadr = hw->literal_indexing ? m->jnt_dofadr[k] : m->jnt_dofadr[j.id];
    const int qposadr = hw->literal_indexing ? m->jnt_dofadr[k] : m->jnt_qposadr[j.id];
    switch (j.method) {
      case EFFORT: d->qfrc_applied[dofadr] = e_stop ? 0 : j.effort_command; break;
      case POSITION:
        d->qpos[qposadr] = j.position_command;
        d->qvel[dofadr] = 0.;
        d->qfrc_applied[dofadr] = 0.;
        break;
      case POSITION_PID: {
        double error;
        switch (j.type) {
          case REVOLUTE:
            shortest_angular_distance_with_limits(j.position, j.position_command, j.lower, j.upper, error);
            break;
          case CONTINUOUS: error = shortest_angular_distance(j.position, j.position_command); break;
          default: error = j.position_command - j.position;
        }
        const double effort = saturate(j.pid.computeCommand(error, period), -j.effort_limit, j.effort_limit);
        d->qfrc_applied[dofadr] = effort;
        break;
      }
      case VELOCITY:
        d->qvel[dofadr] = e_stop ? 0. : j.velocity_command;
        d->qfrc_applied[dofadr] = 0.;
        break;
      case VELOCITY_PID: {
        const double error = e_stop ? -j.velocity : j.velocity_command - j.velocity;
        const double effort = saturate(j.pid.computeCommand(error, period), -j.effort_limit, j.effort_limit);
        d->qfrc_applied[dofadr] = effort;
        break;
      }
    }
  }
}

// exposed for the known-answer tests of the restated third-party helpers (tests/test_plugins_cpu.py)
int orc_angles_shortest_with_limits(double from, double to, double left, double right, double* out) {
  return shortest_angular_distance_with_limits(from, to, left, right, *out) ? 1 : 0;
}
double orc_angles_normalize(double a) { return normalize_angle(a); }
double orc_pid_run(const double* gains6, const double* errors, int n, double dt, double* out_cmds) {
  Pid pid;
  pid.p = gains6[0]; pid.i = gains6[1]; pid.d = gains6[2]; pid.i_max = gains6[3]; pid.i_min = gains6[4];
  pid.antiwindup = gains6[5] != 0;
  double last = 0;
  for (int k = 0; k < n; k++) { last = pid.computeCommand(errors[k], dt); if (out_cmds) out_cmds[k] = last; }
  return last;
}

void orc_hw_state(const OrcRobotHW* hw, double* pos, double* vel, double* eff) {
  for (size_t k = 0; k < hw->joints.size(); k++) {
    if (pos) pos[k] = hw->joints[k].position;
    if (vel) vel[k] = hw->joints[k].velocity;
    if (eff) eff[k] = hw->joints[k].effort;
  }
}

// ---------------------------------------------------------------- sensors: lastStageCallback (:175-437)
// normals: one standard-normal draw per noisy dimension, consumed in the order the reference calls
// noise_dist(rand_generator) (x, y, z of sensor 0, then sensor 1, ...); the caller supplies them so that the test can
// feed the device kernel's counter-based stream.  flag / mean / sigma: per sensor, SensorNoiseModel semantics
// (mean / sigma packed over the flagged dimensions only, :217-243).  values / gt: [nsensordata] doubles holding the
// float32-rounded message fields (noisy quaternions stay double: tf2::toMsg).
void orc_sensor_readout(const b2mjModel* m, const OrcData* d, const int* flag, const double* mean, const double* sigma,
                        const double* normals, double* values, double* gt) {
  int nn = 0;
  for (int n = 0; n < m->nsensor; n++) {
    const int adr = m->sensor_adr[n], type = m->sensor_type[n], dim = m->sensor_dim[n];
    const double cutoff = m->sensor_cutoff[n] > 0 ? m->sensor_cutoff[n] : 1;
    const int is_set = flag ? flag[n] : 0;
    const double* mu = mean + 3 * n;
    const double* sg = sigma + 3 * n;
    int noise_idx = 0;
    const bool quat = type == B2MJ_SENS_BALLQUAT || type == B2MJ_SENS_FRAMEQUAT;
    if (quat) {
      double q[4];
      for (int k = 0; k < 4; k++) {
        q[k] = (double)static_cast<float>(d->sensordata[adr + k] / cutoff);
        if (gt) gt[adr + k] = q[k];
        values[adr + k] = q[k];
      }
      if (is_set == 0) continue;
      // tf2::fromMsg -> normalize
      double w = q[0], x = q[1], y = q[2], z = q[3];
      // tf2::Quaternion::normalize(): *this /= length()  ==  *this *= 1 / length()
      double inv = 1.0 / std::sqrt(x * x + y * y + z * z + w * w);
      x *= inv; y *= inv; z *= inv; w *= inv;
      double rpy[3];
      for (int k = 0; k < 3; k++) {
        if (is_set & (1 << k)) {
          rpy[k] = normals[nn++] * sg[noise_idx] + mu[noise_idx];
          if (k < 2) noise_idx += 1;
        } else {
          rpy[k] = 0;
        }
      }
      // tf2::Quaternion::setRPY(roll, pitch, yaw)
      const double hy = rpy[2] * 0.5, hp = rpy[1] * 0.5, hr = rpy[0] * 0.5;
      const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr),
                   sr = std::sin(hr);
      double rx = sr * cp * cy - cr * sp * sy, ry = cr * sp * cy + sr * cp * sy, rz = cr * cp * sy - sr * sp * cy,
             rw = cr * cp * cy + sr * sp * sy;
      inv = 1.0 / std::sqrt(rx * rx + ry * ry + rz * rz + rw * rw);
      rx *= inv; ry *= inv; rz *= inv; rw *= inv;
      // q_rot * q_orig (tf2 operator*), then normalize
      double ox = rw * x + rx * w + ry * z - rz * y;
      double oy = rw * y + ry * w + rz * x - rx * z;
      double oz = rw * z + rz * w + rx * y - ry * x;
      double ow = rw * w - rx * x - ry * y - rz * z;
      inv = 1.0 / std::sqrt(ox * ox + oy * oy + oz * oz + ow * ow);
      values[adr] = ow * inv; values[adr + 1] = ox * inv; values[adr + 2] = oy * inv; values[adr + 3] = oz * inv;
      continue;
    }
    for (int k = 0; k < dim; k++) {
      const double x = (double)static_cast<float>(d->sensordata[adr + k] / cutoff);
      if (gt) gt[adr + k] = x;
      values[adr + k] = x;
    }
    if (is_set == 0) continue;
    if (dim == 1) {  // scalar sensors: one draw whatever the flag bits say (:369-371)
      const double noise = normals[nn++] * sg[0] + mu[0];
      values[adr] = (double)static_cast<float>(d->sensordata[adr] + noise / cutoff);
      continue;
    }
    for (int k = 0; k < 3 && k < dim; k++) {
      double noise = 0;
      if (is_set & (1 << k)) {
        noise = normals[nn++] * sg[noise_idx] + mu[noise_idx];
        if (k < 2) noise_idx += 1;
      }
      values[adr + k] = (double)static_cast<float>(d->sensordata[adr + k] + noise / cutoff);
    }
  }
}

// CPU baseline driver with the full state record and the actuator-write path (bench.py cpu_baseline / --impl reference
// for the configs that go through mujoco_ros_control): per env and step, readSim every `hw_control_every` steps, writeSim
// every step (mujoco_ros_control_plugin.cpp:176-193), mj_step, last-stage sensor readout.  Threads over envs.
double orc_rollout_ex(const b2mjModel* m, const OrcRolloutArgs* a) {
  const int nthreads = a->nthreads < 1 ? 1 : a->nthreads;
  std::vector<OrcData*> datas(nthreads);
  for (auto& p : datas) p = orc_make_data(m);
  const int nj = a->hw_njoint;
  auto worker = [&](int tid) {
    OrcData* d = datas[tid];
    std::vector<double> vals(m->nsensordata > 0 ? m->nsensordata : 1), gt(vals.size());
    std::vector<int> zflag(m->nsensor > 0 ? m->nsensor : 1, 0);
    std::vector<double> zms(3 * zflag.size(), 0.0);
    for (int e = tid; e < a->nenv; e += nthreads) {
      orc_reset_data(m, d);
      std::memcpy(d->qpos, a->qpos + (size_t)e * m->nq, sizeof(double) * m->nq);
      std::memcpy(d->qvel, a->qvel + (size_t)e * m->nv, sizeof(double) * m->nv);
      if (a->act && m->na) std::memcpy(d->act, a->act + (size_t)e * m->na, sizeof(double) * m->na);
      if (a->warm) std::memcpy(d->qacc_warmstart, a->warm + (size_t)e * m->nv, sizeof(double) * m->nv);
      if (a->time) d->time[0] = a->time[e];
      OrcRobotHW* hw = nj ? orc_hw_create(m, nj, a->hw_joint_id, a->hw_mode, a->hw_kind, a->hw_lower, a->hw_upper,
                                           a->hw_effort, a->hw_pid6, a->hw_limits, 0)
                          : nullptr;
      for (int s = 0; s < a->nsteps; s++) {
        if (a->ctrl && m->nu) std::memcpy(d->ctrl, a->ctrl + ((size_t)s * a->nenv + e) * m->nu, sizeof(double) * m->nu);
        if (hw) {
          if (s % (a->hw_control_every > 0 ? a->hw_control_every : 1) == 0) orc_hw_read(hw, m, d);
          orc_hw_write(hw, m, d, a->hw_cmd + ((size_t)s * a->nenv + e) * nj, 0, m->opt.timestep);
        }
        orc_step(m, d);
        if (a->sensor_out && m->nsensordata) {
          orc_sensor_readout(m, d, zflag.data(), zms.data(), zms.data(), zms.data(), vals.data(), gt.data());
          for (int k = 0; k < m->nsensordata; k++) a->sensor_out[(size_t)e * m->nsensordata + k] = (float)vals[k];
        }
      }
      if (hw) orc_hw_free(hw);
      std::memcpy(a->qpos + (size_t)e * m->nq, d->qpos, sizeof(double) * m->nq);
      std::memcpy(a->qvel + (size_t)e * m->nv, d->qvel, sizeof(double) * m->nv);
      if (a->act && m->na) std::memcpy(a->act + (size_t)e * m->na, d->act, sizeof(double) * m->na);
      if (a->warm) std::memcpy(a->warm + (size_t)e * m->nv, d->qacc_warmstart, sizeof(double) * m->nv);
      if (a->time) a->time[e] = d->time[0];
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (nthreads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (auto p : datas) orc_free_data(p);
  return secs;
}

}  // extern "C"
