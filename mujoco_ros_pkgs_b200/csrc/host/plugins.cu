// plugins.cu — batched device versions of the two plugin data paths either side of the step.
//
//  * robot_hw  : DefaultRobotHWSim::readSim / writeSim (reference mujoco_ros_control/src/
//                default_robot_hw_sim.cpp:230-246, :248-326): 5 control modes, PID (control_toolbox
//                semantics), effort clamp, e-stop hold.  Deviation, documented in SURVEY App. D: state is
//                indexed by the joint's MuJoCo id (the reference indexes jnt_dofadr with the transmission
//                index, :273-321, which is only right when transmissions are in joint order).
//  * sensor_readout : MujocoRosSensorsPlugin::lastStageCallback arithmetic (reference
//                mujoco_ros_sensors/src/mujoco_sensor_handler_plugin.cpp:175-437): value =
//                float(sensordata/cutoff), noisy value = float(sensordata + noise/cutoff) (the reference's
//                precedence, :241), quaternion noise composed as rpy2quat(noise) * normalize(q).  Noise
//                comes from a counter-based Philox4x32-10 stream keyed by (seed; env, sensor, dim, readout
//                count) instead of mt19937(random_device) (:94 of the header) so runs are reproducible.
//  * allgather_publish : the one exchange step on the path: NCCL all-gather of a field slab.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "b2mj.h"
#include "handle.h"
#include "model/model_core.h"

namespace b2mj {

#define CUDA_OK(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                          \
      return B2MJ_ECUDA;                                                                      \
    }                                                                                         \
  } while (0)

struct RobotHWState {
  int njoint = 0;
  int *joint_id = nullptr, *mode = nullptr, *kind = nullptr, *qposadr = nullptr, *dofadr = nullptr;
  double *effort_limit = nullptr, *pid = nullptr, *lower = nullptr, *upper = nullptr;
  b2mjJointLimits* limits = nullptr;  // [njoint] or null: no joint_limits_interface handle registered
  // per (env, joint) state
  double *pos = nullptr, *vel = nullptr, *eff = nullptr, *i_error = nullptr, *last_error = nullptr, *d_error = nullptr,
         *hold_cmd = nullptr, *prev_cmd = nullptr;
  double* cmd = nullptr;  // staging for host commands
  int last_e_stop = 0;
};

struct SensorReadoutState {
  int nmodel = 0;
  // per sensor (dense, nsensor entries)
  double *mean = nullptr, *sigma = nullptr;  // [nsensor][3]
  int* flag = nullptr;                       // [nsensor]
  uint64_t seed = 0;
  uint64_t count = 0;                        // readouts so far (advances the noise stream)
  double *values = nullptr, *gt = nullptr;   // device staging [nenv][nsensordata]
};

// ---------------------------------------------------------------- robot hw kernels
// ros/angles 1.9.13 (angles.h), the functions writeSim / readSim call (default_robot_hw_sim.cpp:242,287-295)
#define B2MJ_PI 3.14159265358979323846
__device__ __forceinline__ double normalize_angle(double angle) {
  const double result = fmod(angle + B2MJ_PI, 2.0 * B2MJ_PI);
  if (result <= 0.0) return result + B2MJ_PI;
  return result - B2MJ_PI;
}
__device__ __forceinline__ double shortest_angular_distance(double from, double to) { return normalize_angle(to - from); }
__device__ __forceinline__ double two_pi_complement(double angle) {
  if (angle > 2 * B2MJ_PI || angle < -2.0 * B2MJ_PI) angle = fmod(angle, 2.0 * B2MJ_PI);
  if (angle < 0) return 2 * B2MJ_PI + angle;
  else if (angle > 0) return -2 * B2MJ_PI + angle;
  return 2 * B2MJ_PI;
}
__device__ bool find_min_max_delta(double from, double left_limit, double right_limit, double& result_min_delta,
                                   double& result_max_delta) {
  double delta[4];
  delta[0] = shortest_angular_distance(from, left_limit);
  delta[1] = shortest_angular_distance(from, right_limit);
  delta[2] = two_pi_complement(delta[0]);
  delta[3] = two_pi_complement(delta[1]);
  if (delta[0] == 0) {
    result_min_delta = delta[0];
    result_max_delta = fmax(delta[1], delta[3]);
    return true;
  }
  if (delta[1] == 0) {
    result_max_delta = delta[1];
    result_min_delta = fmin(delta[0], delta[2]);
    return true;
  }
  double delta_min = delta[0], delta_min_2pi = delta[2];
  if (delta[2] < delta_min) { delta_min = delta[2]; delta_min_2pi = delta[0]; }
  double delta_max = delta[1], delta_max_2pi = delta[3];
  if (delta[3] > delta_max) { delta_max = delta[3]; delta_max_2pi = delta[1]; }
  if ((delta_min <= delta_max_2pi) || (delta_max >= delta_min_2pi)) {
    result_min_delta = delta_max_2pi;
    result_max_delta = delta_min_2pi;
    return left_limit == -B2MJ_PI && right_limit == B2MJ_PI;
  }
  result_min_delta = delta_min;
  result_max_delta = delta_max;
  return true;
}
__device__ double shortest_angular_distance_with_limits(double from, double to, double left_limit, double right_limit) {
  double min_delta = -2 * B2MJ_PI, max_delta = 2 * B2MJ_PI, min_delta_to = -2 * B2MJ_PI, max_delta_to = 2 * B2MJ_PI;
  const bool flag = find_min_max_delta(from, left_limit, right_limit, min_delta, max_delta);
  const double delta = shortest_angular_distance(from, to);
  const double delta_mod_2pi = two_pi_complement(delta);
  if (flag) {  // from position is within the limits
    if (delta >= min_delta && delta <= max_delta) return delta;
    if (delta_mod_2pi >= min_delta && delta_mod_2pi <= max_delta) return delta_mod_2pi;
    find_min_max_delta(to, left_limit, right_limit, min_delta_to, max_delta_to);
    if (fabs(min_delta_to) < fabs(max_delta_to)) return fmax(delta, delta_mod_2pi);
    if (fabs(min_delta_to) > fabs(max_delta_to)) return fmin(delta, delta_mod_2pi);
    return fabs(delta) < fabs(delta_mod_2pi) ? delta : delta_mod_2pi;
  }
  find_min_max_delta(to, left_limit, right_limit, min_delta_to, max_delta_to);
  if (fabs(min_delta) < fabs(max_delta)) return fmin(delta, delta_mod_2pi);
  if (fabs(min_delta) > fabs(max_delta)) return fmax(delta, delta_mod_2pi);
  return fabs(delta) < fabs(delta_mod_2pi) ? delta : delta_mod_2pi;
}
__device__ __forceinline__ double saturate_d(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// joint_limits_interface (ros_control 0.19): the one handle a joint owns, chosen by its hardware interface
// (default_robot_hw_sim.cpp:411-445); enforced on the command before writeSim's mode switch (:262-267)
__device__ double enforce_limits(const b2mjJointLimits& L, int iface, double c, double pos, double vel, double period,
                                 double* prev_cmd) {
  if (iface == B2MJ_CTRL_POSITION) {
    double prev = *prev_cmd;
    if (isnan(prev)) prev = pos;
    if (!L.has_soft_limits) {  // PositionJointSaturationHandle
      const double lo_lim = L.has_position_limits ? L.min_position : -1.7976931348623157e308;
      const double hi_lim = L.has_position_limits ? L.max_position : 1.7976931348623157e308;
      double min_pos = lo_lim, max_pos = hi_lim;
      if (L.has_velocity_limits) {
        const double delta_pos = L.max_velocity * period;
        min_pos = fmax(prev - delta_pos, lo_lim);
        max_pos = fmin(prev + delta_pos, hi_lim);
      }
      c = saturate_d(c, min_pos, max_pos);
    } else {  // PositionJointSoftLimitsHandle
      double soft_min_vel = -L.max_velocity, soft_max_vel = L.max_velocity;
      if (L.has_position_limits) {
        soft_min_vel = saturate_d(-L.k_position * (prev - L.soft_min_position), -L.max_velocity, L.max_velocity);
        soft_max_vel = saturate_d(-L.k_position * (prev - L.soft_max_position), -L.max_velocity, L.max_velocity);
      }
      double pos_low = prev + soft_min_vel * period, pos_high = prev + soft_max_vel * period;
      if (L.has_position_limits) {
        pos_low = fmax(pos_low, L.min_position);
        pos_high = fmin(pos_high, L.max_position);
      }
      c = saturate_d(c, pos_low, pos_high);
    }
    *prev_cmd = c;
  } else if (iface == B2MJ_CTRL_VELOCITY) {
    if (!L.has_soft_limits) {  // VelocityJointSaturationHandle
      double vel_low = -L.max_velocity, vel_high = L.max_velocity;
      if (L.has_acceleration_limits) {
        const double prev = isnan(*prev_cmd) ? 0.0 : *prev_cmd;  // prev_cmd_ starts at 0
        vel_low = fmax(prev - L.max_acceleration * period, -L.max_velocity);
        vel_high = fmin(prev + L.max_acceleration * period, L.max_velocity);
      }
      c = saturate_d(c, vel_low, vel_high);
      *prev_cmd = c;
    } else {  // VelocityJointSoftLimitsHandle
      double min_vel = -L.max_velocity, max_vel = L.max_velocity;
      if (L.has_position_limits) {
        min_vel = saturate_d(-L.k_position * (pos - L.soft_min_position), -L.max_velocity, L.max_velocity);
        max_vel = saturate_d(-L.k_position * (pos - L.soft_max_position), -L.max_velocity, L.max_velocity);
      }
      if (L.has_acceleration_limits) {
        min_vel = fmax(vel - L.max_acceleration * period, min_vel);
        max_vel = fmin(vel + L.max_acceleration * period, max_vel);
      }
      c = saturate_d(c, min_vel, max_vel);
    }
  } else {
    if (!L.has_soft_limits) {  // EffortJointSaturationHandle
      double min_eff = -L.max_effort, max_eff = L.max_effort;
      if (L.has_position_limits) {
        if (pos < L.min_position) min_eff = 0.0;
        else if (pos > L.max_position) max_eff = 0.0;
      }
      if (vel < -L.max_velocity) min_eff = 0.0;
      else if (vel > L.max_velocity) max_eff = 0.0;
      c = saturate_d(c, min_eff, max_eff);
    } else {  // EffortJointSoftLimitsHandle
      double soft_min_vel = -L.max_velocity, soft_max_vel = L.max_velocity;
      if (L.has_position_limits) {
        soft_min_vel = saturate_d(-L.k_position * (pos - L.soft_min_position), -L.max_velocity, L.max_velocity);
        soft_max_vel = saturate_d(-L.k_position * (pos - L.soft_max_position), -L.max_velocity, L.max_velocity);
      }
      const double soft_min_eff = saturate_d(-L.k_velocity * (vel - soft_min_vel), -L.max_effort, L.max_effort);
      const double soft_max_eff = saturate_d(-L.k_velocity * (vel - soft_max_vel), -L.max_effort, L.max_effort);
      c = saturate_d(c, soft_min_eff, soft_max_eff);
    }
  }
  return c;
}

// One thread per (env, joint).  do_read: readSim (:230-246) refreshes the joint-state interface values from the state
// record; do_write: writeSim (:248-326) consumes them.  The reference reads only when a control period has elapsed but
// writes every step (mujoco_ros_control_plugin.cpp:176-193), so between updates the PID runs on the state of the last
// read -- the two halves are therefore separate launches, not one fused read-modify-write.
__global__ void robot_hw_kernel(double* rec, int pitch, int rec_qpos, int rec_qvel, int rec_qfrc, int nenv, int nj,
                                const int* mode, const int* kind, const int* qposadr, const int* dofadr,
                                const double* effort_limit, const double* pid, const double* lower, const double* upper,
                                const b2mjJointLimits* limits, double* pos, double* vel, double* eff, double* i_error,
                                double* last_error, double* d_error, double* hold_cmd, double* prev_cmd, const double* cmd,
                                int e_stop, int e_stop_rising, double period, int do_read, int do_write) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nenv * nj) return;
  const int env = idx / nj, j = idx - env * nj;
  double* r = rec + (size_t)env * pitch;
  if (do_read) {
    const double position = r[rec_qpos + qposadr[j]], velocity = r[rec_qvel + dofadr[j]], effort = r[rec_qfrc + dofadr[j]];
    if (kind[j] == 2) pos[idx] = position;
    else pos[idx] += shortest_angular_distance(pos[idx], position);
    vel[idx] = velocity;
    eff[idx] = effort;
  }
  if (!do_write) return;
  double c = cmd[idx];
  const int md = mode[j];
  const int iface = md == B2MJ_CTRL_POSITION_PID ? B2MJ_CTRL_POSITION : md == B2MJ_CTRL_VELOCITY_PID ? B2MJ_CTRL_VELOCITY : md;
  if (iface == B2MJ_CTRL_POSITION && e_stop) {
    if (e_stop_rising) hold_cmd[idx] = c;
    c = hold_cmd[idx];
  }
  if (limits) c = enforce_limits(limits[j], iface, c, pos[idx], vel[idx], period, prev_cmd + idx);
  switch (md) {
    case B2MJ_CTRL_EFFORT: r[rec_qfrc + dofadr[j]] = e_stop ? 0.0 : c; break;
    case B2MJ_CTRL_POSITION:
      r[rec_qpos + qposadr[j]] = c;
      r[rec_qvel + dofadr[j]] = 0;
      r[rec_qfrc + dofadr[j]] = 0;
      break;
    case B2MJ_CTRL_VELOCITY:
      r[rec_qvel + dofadr[j]] = e_stop ? 0.0 : c;
      r[rec_qfrc + dofadr[j]] = 0;
      break;
    case B2MJ_CTRL_POSITION_PID:
    case B2MJ_CTRL_VELOCITY_PID: {
      double error;
      if (md == B2MJ_CTRL_POSITION_PID) {
        if (kind[j] == 0) error = shortest_angular_distance_with_limits(pos[idx], c, lower[j], upper[j]);
        else if (kind[j] == 1) error = shortest_angular_distance(pos[idx], c);
        else error = c - pos[idx];
      } else {
        error = e_stop ? -vel[idx] : c - vel[idx];
      }
      // control_toolbox::Pid::computeCommand(error, dt) (1.19.0); gains: p, i, d, i_max, i_min, antiwindup
      const double* g = pid + 6 * j;
      double out = 0;
      if (!(period == 0.0 || isnan(error) || isinf(error))) {
        double error_dot = d_error[idx];
        if (period > 0.0) {
          error_dot = (error - last_error[idx]) / period;
          last_error[idx] = error;
        }
        d_error[idx] = error_dot;
        if (!(isnan(error_dot) || isinf(error_dot))) {
          const double p_term = g[0] * error;
          double ie = i_error[idx] + period * error;
          const bool antiwindup = g[5] != 0;
          if (antiwindup && g[1] != 0) {
            const double a = g[4] / g[1], b = g[3] / g[1];
            ie = saturate_d(ie, fmin(a, b), fmax(a, b));
          }
          i_error[idx] = ie;
          double i_term = g[1] * ie;
          if (!antiwindup) i_term = saturate_d(i_term, g[4], g[3]);
          const double d_term = g[2] * error_dot;
          out = p_term + i_term + d_term;
        }
      }
      const double lim = effort_limit[j];
      r[rec_qfrc + dofadr[j]] = saturate_d(out, -lim, lim);
      break;
    }
  }
}

__global__ void fill_kernel(double* p, double v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------- sensor readout kernels
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned* out) {
  for (int r = 0; r < 10; r++) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// standard normal from one Philox block (Box-Muller on two 32-bit uniforms)
__device__ __forceinline__ double philox_normal(unsigned long long seed, unsigned env, unsigned sensor, unsigned dim,
                                                unsigned long long count) {
  unsigned o[4];
  philox4x32_10(env, sensor * 4u + dim, (unsigned)count, (unsigned)(count >> 32), (unsigned)seed, (unsigned)(seed >> 32), o);
  const double u1 = ((double)o[0] + 0.5) * (1.0 / 4294967296.0), u2 = ((double)o[1] + 0.5) * (1.0 / 4294967296.0);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

__global__ void sensor_readout_kernel(const double* rec, int pitch, int rec_sd, int nenv, int nsensor, int nsd,
                                      const int* s_type, const int* s_adr, const int* s_dim, const double* s_cutoff,
                                      const double* mean, const double* sigma, const int* flag, unsigned long long seed,
                                      unsigned long long count, double* values, double* gt) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nenv * nsensor) return;
  const int env = idx / nsensor, n = idx - env * nsensor;
  const double* sd = rec + (size_t)env * pitch + rec_sd;
  const int adr = s_adr[n], dim = s_dim[n], type = s_type[n];
  const double cutoff = s_cutoff[n] > 0 ? s_cutoff[n] : 1.0;
  double* v = values + (size_t)env * nsd + adr;
  double* g = gt ? gt + (size_t)env * nsd + adr : nullptr;
  const int fl = flag ? flag[n] : 0;
  // ground truth / noise-free value: float(sensordata / cutoff)
  for (int k = 0; k < dim; k++) {
    const double x = (double)(float)(sd[adr + k] / cutoff);
    if (g) g[k] = x;
    v[k] = x;
  }
  if (!fl) return;
  const bool quat = (type == B2MJ_SENS_BALLQUAT || type == B2MJ_SENS_FRAMEQUAT);
  if (!quat && dim == 1) {
    // scalar sensors draw once whichever flag bit is set and use mean[0] / sigma[0] (:369-371)
    const double noise = philox_normal(seed, (unsigned)env, (unsigned)n, 0u, count) * sigma[3 * n] + mean[3 * n];
    v[0] = (double)(float)(sd[adr] + noise / cutoff);
    return;
  }
  double noise[3] = {0, 0, 0};
  int ni = 0;
  for (int k = 0; k < 3; k++) {
    if (fl & (1 << k)) {
      noise[k] = philox_normal(seed, (unsigned)env, (unsigned)n, (unsigned)k, count) * sigma[3 * n + ni] + mean[3 * n + ni];
      if (k < 2) ni++;
    }
  }
  if (!quat) {
    for (int k = 0; k < dim && k < 3; k++) v[k] = (double)(float)(sd[adr + k] + noise[k] / cutoff);
  } else {
    // q = (setRPY(noise) * normalize(q_msg)).normalize(), sensor order (w,x,y,z); tf2 normalises by multiplying with
    // the reciprocal length and composes the product in this term order (tf2/LinearMath/Quaternion.h)
    double w = v[0], x = v[1], y = v[2], z = v[3];
    double inv = 1.0 / sqrt(x * x + y * y + z * z + w * w);
    x *= inv; y *= inv; z *= inv; w *= inv;
    const double hr = noise[0] * 0.5, hp = noise[1] * 0.5, hy = noise[2] * 0.5;
    const double cr = cos(hr), sr = sin(hr), cp = cos(hp), sp = sin(hp), cy = cos(hy), sy = sin(hy);
    double rx = sr * cp * cy - cr * sp * sy, ry = cr * sp * cy + sr * cp * sy, rz = cr * cp * sy - sr * sp * cy,
           rw = cr * cp * cy + sr * sp * sy;
    inv = 1.0 / sqrt(rx * rx + ry * ry + rz * rz + rw * rw);
    rx *= inv; ry *= inv; rz *= inv; rw *= inv;
    const double ox = rw * x + rx * w + ry * z - rz * y;
    const double oy = rw * y + ry * w + rz * x - rx * z;
    const double oz = rw * z + rz * w + rx * y - ry * x;
    const double ow = rw * w - rx * x - ry * y - rz * z;
    inv = 1.0 / sqrt(ox * ox + oy * oy + oz * oz + ow * ow);
    v[0] = ow * inv; v[1] = ox * inv; v[2] = oy * inv; v[3] = oz * inv;
  }
}

void handle_free_plugins(Handle* h) {
  if (h->robot_hw) {
    RobotHWState* s = h->robot_hw;
    cudaFree(s->joint_id); cudaFree(s->mode); cudaFree(s->kind); cudaFree(s->qposadr); cudaFree(s->dofadr);
    cudaFree(s->effort_limit); cudaFree(s->pid); cudaFree(s->lower); cudaFree(s->upper); cudaFree(s->limits);
    cudaFree(s->pos); cudaFree(s->vel); cudaFree(s->eff); cudaFree(s->i_error); cudaFree(s->last_error);
    cudaFree(s->d_error); cudaFree(s->hold_cmd); cudaFree(s->prev_cmd); cudaFree(s->cmd);
    delete s;
    h->robot_hw = nullptr;
  }
  if (h->sensor_ro) {
    SensorReadoutState* s = h->sensor_ro;
    cudaFree(s->mean); cudaFree(s->sigma); cudaFree(s->flag); cudaFree(s->values); cudaFree(s->gt);
    delete s;
    h->sensor_ro = nullptr;
  }
}

// joint-state interface values and controller state as DefaultRobotHWSim::initSim leaves them
// (default_robot_hw_sim.cpp:132-137: position 1.0, velocity 0, effort 1.0; Pid zeroed; limit handles fresh:
// position handles start with prev_cmd = NaN "take the current position", velocity handles with 0)
static void robot_hw_init_state(Handle* h) {
  RobotHWState* s = h->robot_hw;
  const size_t n = (size_t)h->nenv * s->njoint;
  const size_t bytes = n * sizeof(double);
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 1184);
  fill_kernel<<<blocks, 256, 0, h->stream>>>(s->pos, 1.0, n);
  fill_kernel<<<blocks, 256, 0, h->stream>>>(s->eff, 1.0, n);
  cudaMemsetAsync(s->vel, 0, bytes, h->stream);
  cudaMemsetAsync(s->i_error, 0, bytes, h->stream);
  cudaMemsetAsync(s->last_error, 0, bytes, h->stream);
  cudaMemsetAsync(s->d_error, 0, bytes, h->stream);
  cudaMemsetAsync(s->hold_cmd, 0, bytes, h->stream);
  // prev_cmd: NaN = "unset" (position handles then take the current position, velocity handles 0)
  fill_kernel<<<blocks, 256, 0, h->stream>>>(s->prev_cmd, nan(""), n);
  s->last_e_stop = 0;
}

void handle_reset_plugins(Handle* h, const uint8_t* env_mask) {
  // DefaultRobotHWSim has no reset hook of its own; its state restarts when the plugin is reloaded with the env.
  // Here a full reset puts the controller state back to its post-initSim values so that rollouts are reproducible.
  if (h->robot_hw && !env_mask) {
    robot_hw_init_state(h);
  }
  if (h->sensor_ro && !env_mask) h->sensor_ro->count = 0;
}

template <typename T>
static int upload(T** dst, const T* src, size_t n) {
  CUDA_OK(cudaMalloc(dst, std::max<size_t>(1, n) * sizeof(T)));
  if (n) CUDA_OK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace b2mj

using namespace b2mj;

extern "C" {

int b2mj_robot_hw_configure(b2mj_handle* hh, const b2mjRobotHW* cfg) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !cfg || cfg->njoint <= 0 || !cfg->joint_id || !cfg->control_mode) {
    set_error("b2mj_robot_hw_configure: bad argument");
    return B2MJ_EINVAL;
  }
  const b2mjModel* m = h->model;
  const int nj = cfg->njoint;
  std::vector<int> qa(nj), da(nj), kind(nj);
  std::vector<double> lim(nj), pid(6 * nj, 0.0), lo(nj), hi(nj);
  for (int j = 0; j < nj; j++) {
    const int id = cfg->joint_id[j];
    if (id < 0 || id >= m->njnt || (m->jnt_type[id] != B2MJ_JNT_HINGE && m->jnt_type[id] != B2MJ_JNT_SLIDE)) {
      set_error("b2mj_robot_hw_configure: joint " + std::to_string(j) + " is not a hinge/slide joint of the model");
      return B2MJ_EINVAL;
    }
    const int md = cfg->control_mode[j];
    if (md < B2MJ_CTRL_EFFORT || md > B2MJ_CTRL_VELOCITY_PID) {
      set_error("b2mj_robot_hw_configure: unknown control mode");
      return B2MJ_EINVAL;
    }
    const b2mjJointLimits* L = cfg->limits ? cfg->limits + j : nullptr;
    if (L) {
      // the joint_limits_interface handle constructors throw on these (ros_control joint_limits_interface.h), which
      // makes the reference's plugin fail to load
      const int iface = md == B2MJ_CTRL_POSITION_PID ? B2MJ_CTRL_POSITION : md == B2MJ_CTRL_VELOCITY_PID ? B2MJ_CTRL_VELOCITY : md;
      const bool need_vel = iface != B2MJ_CTRL_POSITION || L->has_soft_limits;
      if ((need_vel && !L->has_velocity_limits) || (iface == B2MJ_CTRL_EFFORT && !L->has_effort_limits)) {
        set_error("b2mj_robot_hw_configure: joint " + std::to_string(j) +
                  " lacks the velocity / effort limits its limit handle requires");
        return B2MJ_EINVAL;
      }
    }
    qa[j] = m->jnt_qposadr[id];
    da[j] = m->jnt_dofadr[id];
    kind[j] = cfg->joint_kind ? cfg->joint_kind[j] : (m->jnt_type[id] == B2MJ_JNT_SLIDE ? 2 : (m->jnt_limited[id] ? 0 : 1));
    // registerJointLimits (:340-446): effort / position limits default to "none" = +-max double
    lim[j] = cfg->effort_limit ? cfg->effort_limit[j] : (L && L->has_effort_limits ? L->max_effort : 1.7976931348623157e308);
    if (cfg->pid_gains) std::memcpy(&pid[6 * j], cfg->pid_gains + 5 * j, 5 * sizeof(double));
    pid[6 * j + 5] = cfg->pid_antiwindup ? (double)cfg->pid_antiwindup[j] : 0.0;
    lo[j] = cfg->lower_limit ? cfg->lower_limit[j] : (L && L->has_position_limits ? L->min_position : m->jnt_range[2 * id]);
    hi[j] = cfg->upper_limit ? cfg->upper_limit[j] : (L && L->has_position_limits ? L->max_position : m->jnt_range[2 * id + 1]);
  }
  CUDA_OK(cudaSetDevice(h->device));
  if (h->robot_hw) {
    SensorReadoutState* keep = h->sensor_ro;
    h->sensor_ro = nullptr;
    handle_free_plugins(h);
    h->sensor_ro = keep;
  }
  RobotHWState* s = new RobotHWState();
  h->robot_hw = s;
  s->njoint = nj;
  int rc = 0;
  rc |= upload(&s->joint_id, cfg->joint_id, nj);
  rc |= upload(&s->mode, cfg->control_mode, nj);
  rc |= upload(&s->kind, kind.data(), nj);
  rc |= upload(&s->qposadr, qa.data(), nj);
  rc |= upload(&s->dofadr, da.data(), nj);
  rc |= upload(&s->effort_limit, lim.data(), nj);
  rc |= upload(&s->pid, pid.data(), 6 * nj);
  rc |= upload(&s->lower, lo.data(), nj);
  rc |= upload(&s->upper, hi.data(), nj);
  if (cfg->limits) rc |= upload(&s->limits, cfg->limits, nj);
  if (rc) return B2MJ_ECUDA;
  const size_t n = (size_t)h->nenv * nj;
  double** per[] = {&s->pos, &s->vel, &s->eff, &s->i_error, &s->last_error, &s->d_error, &s->hold_cmd, &s->prev_cmd, &s->cmd};
  for (double** p : per) {
    CUDA_OK(cudaMalloc(p, n * sizeof(double)));
    CUDA_OK(cudaMemset(*p, 0, n * sizeof(double)));
  }
  robot_hw_init_state(h);
  CUDA_OK(cudaGetLastError());
  return 0;
}

static int robot_hw_run(Handle* h, const double* cmd_dev, int e_stop, double period, int do_read, int do_write) {
  RobotHWState* s = h->robot_hw;
  const b2k::DevModel& d = h->dm;
  const int n = h->nenv * s->njoint;
  const int rising = (do_write && e_stop && !s->last_e_stop) ? 1 : 0;
  robot_hw_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(h->rec, d.rec_pitch, d.rec_qpos, d.rec_qvel, d.rec_qfrc_applied,
                                                          h->nenv, s->njoint, s->mode, s->kind, s->qposadr, s->dofadr,
                                                          s->effort_limit, s->pid, s->lower, s->upper, s->limits, s->pos,
                                                          s->vel, s->eff, s->i_error, s->last_error, s->d_error,
                                                          s->hold_cmd, s->prev_cmd, cmd_dev, e_stop, rising, period,
                                                          do_read, do_write);
  CUDA_OK(cudaGetLastError());
  if (do_write) s->last_e_stop = e_stop ? 1 : 0;
  h->launches++;
  return 0;
}

int b2mj_robot_hw_write(b2mj_handle* hh, const double* cmd, int is_device, int e_stop, double period) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !cmd) return B2MJ_EINVAL;
  if (!h->robot_hw) { set_error("b2mj_robot_hw_write: call b2mj_robot_hw_configure first"); return B2MJ_ESTATE; }
  CUDA_OK(cudaSetDevice(h->device));
  const double* src = cmd;
  if (!is_device) {
    CUDA_OK(cudaMemcpyAsync(h->robot_hw->cmd, cmd, (size_t)h->nenv * h->robot_hw->njoint * sizeof(double),
                            cudaMemcpyHostToDevice, h->stream));
    src = h->robot_hw->cmd;
  }
  return robot_hw_run(h, src, e_stop, period, 0, 1);
}

int b2mj_robot_hw_read(b2mj_handle* hh, double* pos, double* vel, double* eff) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  if (!h->robot_hw) { set_error("b2mj_robot_hw_read: call b2mj_robot_hw_configure first"); return B2MJ_ESTATE; }
  CUDA_OK(cudaSetDevice(h->device));
  if (int rc = robot_hw_run(h, nullptr, 0, 0, 1, 0)) return rc;
  if (!pos && !vel && !eff) return 0;  // device-resident controllers: state refreshed, nothing copied, no sync
  const size_t bytes = (size_t)h->nenv * h->robot_hw->njoint * sizeof(double);
  if (pos) CUDA_OK(cudaMemcpyAsync(pos, h->robot_hw->pos, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (vel) CUDA_OK(cudaMemcpyAsync(vel, h->robot_hw->vel, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (eff) CUDA_OK(cudaMemcpyAsync(eff, h->robot_hw->eff, bytes, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int b2mj_robot_hw_state_ptrs(b2mj_handle* hh, double** dev_pos, double** dev_vel, double** dev_eff) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  if (!h->robot_hw) { set_error("b2mj_robot_hw_state_ptrs: call b2mj_robot_hw_configure first"); return B2MJ_ESTATE; }
  if (dev_pos) *dev_pos = h->robot_hw->pos;
  if (dev_vel) *dev_vel = h->robot_hw->vel;
  if (dev_eff) *dev_eff = h->robot_hw->eff;
  return 0;
}

int b2mj_sensor_configure_noise(b2mj_handle* hh, const b2mjSensorNoise* models, int nmodels, uint64_t seed) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || (nmodels > 0 && !models)) return B2MJ_EINVAL;
  const b2mjModel* m = h->model;
  std::vector<double> mean(3 * std::max(1, m->nsensor), 0.0), sigma(3 * std::max(1, m->nsensor), 0.0);
  std::vector<int> flag(std::max(1, m->nsensor), 0);
  for (int i = 0; i < nmodels; i++) {
    const int id = models[i].sensor_id;
    if (id < 0 || id >= m->nsensor) { set_error("b2mj_sensor_configure_noise: bad sensor id"); return B2MJ_EINVAL; }
    for (int k = 0; k < 3; k++) { mean[3 * id + k] = models[i].mean[k]; sigma[3 * id + k] = models[i].sigma[k]; }
    flag[id] = models[i].set_flag;
  }
  CUDA_OK(cudaSetDevice(h->device));
  if (!h->sensor_ro) h->sensor_ro = new SensorReadoutState();
  SensorReadoutState* s = h->sensor_ro;
  cudaFree(s->mean); cudaFree(s->sigma); cudaFree(s->flag);
  s->mean = s->sigma = nullptr; s->flag = nullptr;
  if (upload(&s->mean, mean.data(), mean.size()) || upload(&s->sigma, sigma.data(), sigma.size()) ||
      upload(&s->flag, flag.data(), flag.size()))
    return B2MJ_ECUDA;
  s->seed = seed;
  s->count = 0;
  s->nmodel = nmodels;
  return 0;
}

static int sensor_readout_launch(Handle* h, bool want_gt) {
  const b2mjModel* m = h->model;
  if (!h->sensor_ro) h->sensor_ro = new SensorReadoutState();
  SensorReadoutState* s = h->sensor_ro;
  const size_t n = (size_t)h->nenv * m->nsensordata;
  if (!s->values) {
    CUDA_OK(cudaMalloc(&s->values, n * sizeof(double)));
    CUDA_OK(cudaMalloc(&s->gt, n * sizeof(double)));
  }
  const b2k::DevModel& d = h->dm;
  const int total = h->nenv * m->nsensor;
  sensor_readout_kernel<<<(total + 127) / 128, 128, 0, h->stream>>>(
      h->rec, d.rec_pitch, d.rec_sensordata, h->nenv, m->nsensor, m->nsensordata, d.sensor_type, d.sensor_adr, d.sensor_dim,
      d.sensor_cutoff, s->mean, s->sigma, s->flag, (unsigned long long)s->seed, (unsigned long long)s->count, s->values,
      want_gt ? s->gt : nullptr);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  s->count++;
  return 0;
}

int b2mj_sensor_readout(b2mj_handle* hh, double* values, double* gt) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !values) return B2MJ_EINVAL;
  const b2mjModel* m = h->model;
  if (m->nsensordata == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  if (int rc = sensor_readout_launch(h, gt != nullptr)) return rc;
  SensorReadoutState* s = h->sensor_ro;
  const size_t n = (size_t)h->nenv * m->nsensordata;
  CUDA_OK(cudaMemcpyAsync(values, s->values, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (gt) CUDA_OK(cudaMemcpyAsync(gt, s->gt, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int b2mj_sensor_readout_device(b2mj_handle* hh, double** dev_values, double** dev_gt) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !dev_values) return B2MJ_EINVAL;
  if (h->model->nsensordata == 0) { *dev_values = nullptr; if (dev_gt) *dev_gt = nullptr; return 0; }
  CUDA_OK(cudaSetDevice(h->device));
  if (int rc = sensor_readout_launch(h, dev_gt != nullptr)) return rc;
  *dev_values = h->sensor_ro->values;
  if (dev_gt) *dev_gt = h->sensor_ro->gt;
  return 0;
}

// NCCL is resolved at run time from the process (torch ships libnccl); no link-time dependency.
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
static nccl_allgather_fn resolve_allgather() {
  static nccl_allgather_fn fn = nullptr;
  if (!fn) {
    fn = (nccl_allgather_fn)dlsym(RTLD_DEFAULT, "ncclAllGather");
    if (!fn) {
      void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (lib) fn = (nccl_allgather_fn)dlsym(lib, "ncclAllGather");
    }
  }
  return fn;
}

// one thread per (env, element of the packed row): rows gathered from the strided record / side arrays
struct PackSrc { const double* base; int pitch, count, dst_off; };
#define B2MJ_MAX_PUBLISH_FIELDS 8
struct PackArgs { PackSrc src[B2MJ_MAX_PUBLISH_FIELDS]; int nsrc, row, nenv; };
__global__ void publish_pack_kernel(PackArgs a, double* __restrict__ slab) {
  const long long total = (long long)a.nenv * a.row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int env = (int)(i / a.row), k = (int)(i - (long long)env * a.row);
    int s = 0;
    while (s + 1 < a.nsrc && k >= a.src[s + 1].dst_off) s++;
    slab[i] = a.src[s].base[(size_t)env * a.src[s].pitch + (k - a.src[s].dst_off)];
  }
}

int b2mj_publish_pack(b2mj_handle* hh, const b2mj_field* fields, int nfields, double** dev_slab, int* count_per_env) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !fields || nfields <= 0 || nfields > B2MJ_MAX_PUBLISH_FIELDS || !dev_slab) {
    set_error("b2mj_publish_pack: bad argument (1.." + std::to_string(B2MJ_MAX_PUBLISH_FIELDS) + " fields)");
    return B2MJ_EINVAL;
  }
  PackArgs a{};
  int row = 0;
  for (int i = 0; i < nfields; i++) {
    void* ptr; size_t pitch;
    if (int rc = b2mj_device_ptr(hh, fields[i], &ptr, &pitch)) return rc;
    int is_int = 0;
    const int n = b2mj_field_size(h->model, fields[i], &is_int);
    if (is_int) { set_error("b2mj_publish_pack: float64 fields only"); return B2MJ_EINVAL; }
    if (n == 0) continue;
    a.src[a.nsrc++] = PackSrc{(const double*)ptr, (int)pitch, n, row};
    row += n;
  }
  if (count_per_env) *count_per_env = row;
  if (row == 0) { *dev_slab = nullptr; return 0; }
  a.row = row;
  a.nenv = h->nenv;
  CUDA_OK(cudaSetDevice(h->device));
  const size_t cnt = (size_t)h->nenv * row;
  if (h->publish_slab_n < cnt) {
    CUDA_OK(cudaStreamSynchronize(h->stream));  // an earlier gather may still read the old slab
    cudaFree(h->publish_slab);
    h->publish_slab = nullptr;
    h->publish_slab_n = 0;
    CUDA_OK(cudaMalloc(&h->publish_slab, cnt * sizeof(double)));
    h->publish_slab_n = cnt;
  }
  const int blocks = (int)std::min<size_t>((cnt + 255) / 256, 148 * 8);
  publish_pack_kernel<<<blocks, 256, 0, h->stream>>>(a, h->publish_slab);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  *dev_slab = h->publish_slab;
  return 0;
}

int b2mj_allgather_publish_multi(b2mj_handle* hh, const b2mj_field* fields, int nfields, void* nccl_comm, void* dev_dst_all) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !nccl_comm || !dev_dst_all) return B2MJ_EINVAL;
  nccl_allgather_fn fn = resolve_allgather();
  if (!fn) { set_error("b2mj_allgather_publish: ncclAllGather not found in the process"); return B2MJ_EUNSUPPORTED; }
  double* slab = nullptr;
  int row = 0;
  if (int rc = b2mj_publish_pack(hh, fields, nfields, &slab, &row)) return rc;
  if (row == 0) return 0;
  const int rc = fn(slab, dev_dst_all, (size_t)h->nenv * row, /*ncclDouble*/ 8, nccl_comm, h->stream);
  if (rc != 0) { set_error("ncclAllGather failed with code " + std::to_string(rc)); return B2MJ_ECUDA; }
  h->launches++;
  return 0;
}

int b2mj_allgather_publish(b2mj_handle* hh, b2mj_field f, void* nccl_comm, void* dev_dst_all) {
  return b2mj_allgather_publish_multi(hh, &f, 1, nccl_comm, dev_dst_all);
}

}  // extern "C"

// ---- fused step + publish over peer memory (include/b2mj.h) ----
struct b2mj::FusedPublish {
  int world = 0, rank = 0, count = 0;
  b2k::PubArgs args{};           // host image; slab pointers rewritten per sequence parity
  b2k::PubArgs* dev_args[2] = {nullptr, nullptr};
  unsigned char* local = nullptr;                      // one allocation: slab[0] | slab[1] | flags
  unsigned char* peer[B2K_PUB_MAX_RANKS] = {nullptr};  // base of every rank's allocation in this address space
  bool opened[B2K_PUB_MAX_RANKS] = {false};
  size_t slab_bytes = 0;
  int seq = 0;
  bool connected = false, waits_issued = false;
  int* timed_out = nullptr;  // mapped host memory, set by the wait kernel when a peer never shows up
  double* slab_of(int r, int parity) const { return reinterpret_cast<double*>(peer[r] + (size_t)parity * slab_bytes); }
  int* flags_of(int r) const { return reinterpret_cast<int*>(peer[r] + 2 * slab_bytes); }
};

__global__ void publish_wait_kernel(const int* flags, int nranks, int seq, long long timeout_cycles, int* timed_out) {
  const int r = threadIdx.x;
  if (r < nranks) {
    const volatile int* f = flags + r;
    const long long t0 = clock64();
    while (*f < seq) {
      __nanosleep(200);
      // a peer that died must not hang this GPU: give up after the timeout and let the caller see it
      if (clock64() - t0 > timeout_cycles) { if (timed_out) *reinterpret_cast<volatile int*>(timed_out) = 1; break; }
    }
  }
  __threadfence_system();
}

void b2mj::handle_free_fused_publish(Handle* h) {
  FusedPublish* p = h->fused_pub;
  if (!p) return;
  for (int r = 0; r < p->world; r++)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->peer[r]);
  cudaFree(p->local);
  if (p->timed_out) cudaFreeHost(p->timed_out);
  cudaFree(p->dev_args[0]);
  cudaFree(p->dev_args[1]);
  delete p;
  h->fused_pub = nullptr;
}

extern "C" {

int b2mj_publish_fused_create(b2mj_handle* hh, int world, int rank, const b2mj_field* fields, int nfields,
                              unsigned char* ipc_handle_out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || world < 1 || world > B2K_PUB_MAX_RANKS || rank < 0 || rank >= world || !fields || nfields < 1 ||
      nfields > B2K_PUB_MAX_FIELDS || !ipc_handle_out) {
    set_error("b2mj_publish_fused_create: bad argument (world <= " + std::to_string(B2K_PUB_MAX_RANKS) + ", 1.." +
              std::to_string(B2K_PUB_MAX_FIELDS) + " fields)");
    return B2MJ_EINVAL;
  }
  CUDA_OK(cudaSetDevice(h->device));
  handle_free_fused_publish(h);
  FusedPublish* p = new FusedPublish();
  p->world = world;
  p->rank = rank;
  const b2k::DevModel& d = h->dm;
  int count = 0;
  p->args.nfields = 0;
  for (int i = 0; i < nfields; i++) {
    int off = -1;
    switch (fields[i]) {
      case B2MJ_F_CTRL: off = d.rec_ctrl; break;
      case B2MJ_F_QFRC_APPLIED: off = d.rec_qfrc_applied; break;
      case B2MJ_F_QPOS: off = d.rec_qpos; break;
      case B2MJ_F_QVEL: off = d.rec_qvel; break;
      case B2MJ_F_ACT: off = d.rec_act; break;
      case B2MJ_F_QACC_WARMSTART: off = d.rec_warm; break;
      case B2MJ_F_TIME: off = d.rec_time; break;
      case B2MJ_F_QACC: off = d.rec_qacc; break;
      case B2MJ_F_SENSORDATA: off = d.rec_sensordata; break;
      case B2MJ_F_ACT_DOT: off = d.rec_act_dot; break;
      default: break;
    }
    if (off < 0) {
      delete p;
      set_error("b2mj_publish_fused_create: only fields of the state record can be published from inside the step kernel");
      return B2MJ_EINVAL;
    }
    const int n = d.fsize[fields[i]];
    if (n == 0) continue;
    p->args.foff[p->args.nfields] = off;
    p->args.fcnt[p->args.nfields] = n;
    p->args.nfields++;
    count += n;
  }
  if (count == 0) { delete p; set_error("b2mj_publish_fused_create: nothing to publish"); return B2MJ_EINVAL; }
  p->count = count;
  p->args.nranks = world;
  p->args.rank = rank;
  p->args.count = count;
  p->slab_bytes = (((size_t)world * h->nenv * count * sizeof(double)) + 255) & ~(size_t)255;
  const size_t total = 2 * p->slab_bytes + 256;
  if (cudaMalloc(&p->local, total) != cudaSuccess) { delete p; set_error("b2mj_publish_fused_create: out of device memory"); return B2MJ_ECUDA; }
  cudaMemset(p->local, 0, total);
  if (cudaHostAlloc(reinterpret_cast<void**>(&p->timed_out), sizeof(int), cudaHostAllocMapped) == cudaSuccess) *p->timed_out = 0;
  else { cudaGetLastError(); p->timed_out = nullptr; }
  cudaMalloc(&p->dev_args[0], sizeof(b2k::PubArgs));
  cudaMalloc(&p->dev_args[1], sizeof(b2k::PubArgs));
  cudaIpcMemHandle_t hd;
  if (cudaIpcGetMemHandle(&hd, p->local) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(p->local); cudaFree(p->dev_args[0]); cudaFree(p->dev_args[1]);
    delete p;
    set_error("b2mj_publish_fused_create: cudaIpcGetMemHandle failed");
    return B2MJ_ECUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::memcpy(ipc_handle_out, &hd, 64);
  p->peer[rank] = p->local;
  h->fused_pub = p;
  return 0;
}

int b2mj_publish_fused_connect(b2mj_handle* hh, const unsigned char* all_handles) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !h->fused_pub || !all_handles) { set_error("b2mj_publish_fused_connect: create first"); return B2MJ_EINVAL; }
  FusedPublish* p = h->fused_pub;
  CUDA_OK(cudaSetDevice(h->device));
  for (int r = 0; r < p->world; r++) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, all_handles + 64 * (size_t)r, 64);
    void* ptr = nullptr;
    const cudaError_t err = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
      cudaGetLastError();
      set_error(std::string("b2mj_publish_fused_connect: cannot open the slab of rank ") + std::to_string(r) + ": " +
                cudaGetErrorString(err));
      return B2MJ_ECUDA;
    }
    p->peer[r] = static_cast<unsigned char*>(ptr);
    p->opened[r] = true;
  }
  for (int par = 0; par < 2; par++) {
    b2k::PubArgs a = p->args;
    for (int r = 0; r < p->world; r++) { a.slab[r] = p->slab_of(r, par); a.flags[r] = p->flags_of(r); }
    a.done = reinterpret_cast<unsigned*>(p->local + 2 * p->slab_bytes + 128);
    CUDA_OK(cudaMemcpy(p->dev_args[par], &a, sizeof(a), cudaMemcpyHostToDevice));
  }
  p->connected = true;
  return 0;
}

int b2mj_step_publish(b2mj_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !h->fused_pub || !h->fused_pub->connected) { set_error("b2mj_step_publish: create + connect first"); return B2MJ_EINVAL; }
  FusedPublish* p = h->fused_pub;
  CUDA_OK(cudaSetDevice(h->device));
  p->seq++;
  h->in_split_step = 0;
  h->launch_pub = p->dev_args[p->seq & 1];
  h->launch_pub_seq = p->seq;
  const int rc = handle_launch(h, b2k::MODE_STEP, 1);
  h->launch_pub = nullptr;
  return rc;
}

int b2mj_publish_fused_wait(b2mj_handle* hh, double** dev_gathered, int* count_per_env) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !h->fused_pub || !h->fused_pub->connected) { set_error("b2mj_publish_fused_wait: create + connect first"); return B2MJ_EINVAL; }
  FusedPublish* p = h->fused_pub;
  CUDA_OK(cudaSetDevice(h->device));
  // the timeout marker lives in mapped host memory: the host reads it without synchronising the stream
  if (p->timed_out && *reinterpret_cast<volatile int*>(p->timed_out)) {
    set_error("b2mj_publish_fused_wait: a peer did not publish within the timeout (its process died or never stepped)");
    return B2MJ_ECUDA;
  }
  publish_wait_kernel<<<1, 32, 0, h->stream>>>(p->flags_of(p->rank), p->world, p->seq, /* ~5 s */ 10000000000LL, p->timed_out);
  p->waits_issued = true;
  CUDA_OK(cudaGetLastError());
  h->launches++;
  if (dev_gathered) *dev_gathered = p->slab_of(p->rank, p->seq & 1);
  if (count_per_env) *count_per_env = p->count;
  return 0;
}

}  // extern "C"
