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
  // per (env, joint) state
  double *pos = nullptr, *vel = nullptr, *eff = nullptr, *i_error = nullptr, *last_error = nullptr, *hold_cmd = nullptr;
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
__device__ __forceinline__ double normalize_angle(double a) {
  const double two_pi = 6.283185307179586476925286766559;
  a = fmod(a + 3.14159265358979323846, two_pi);
  if (a < 0) a += two_pi;
  return a - 3.14159265358979323846;
}
__device__ __forceinline__ double shortest_angular_distance(double from, double to) { return normalize_angle(to - from); }

__global__ void robot_hw_kernel(double* rec, int pitch, int rec_qpos, int rec_qvel, int rec_qfrc, int nenv, int nj,
                                const int* mode, const int* kind, const int* qposadr, const int* dofadr,
                                const double* effort_limit, const double* pid, const double* lower, const double* upper,
                                double* pos, double* vel, double* eff, double* i_error, double* last_error, double* hold_cmd,
                                const double* cmd, int e_stop, int e_stop_rising, double period, int do_write) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nenv * nj) return;
  const int env = idx / nj, j = idx - env * nj;
  double* r = rec + (size_t)env * pitch;
  // readSim
  const double position = r[rec_qpos + qposadr[j]], velocity = r[rec_qvel + dofadr[j]], effort = r[rec_qfrc + dofadr[j]];
  if (kind[j] == 2) pos[idx] = position;
  else pos[idx] += shortest_angular_distance(pos[idx], position);
  vel[idx] = velocity;
  eff[idx] = effort;
  if (!do_write) return;
  // writeSim
  double c = cmd[idx];
  const int md = mode[j];
  if (md == B2MJ_CTRL_POSITION || md == B2MJ_CTRL_POSITION_PID) {
    if (e_stop) {
      if (e_stop_rising) hold_cmd[idx] = c;
      c = hold_cmd[idx];
    }
  }
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
        if (kind[j] == 0) {
          // revolute with limits: shortest distance that stays inside [lower, upper] when possible
          const double d1 = shortest_angular_distance(pos[idx], c);
          const double d2 = d1 > 0 ? d1 - 6.283185307179586476925286766559 : d1 + 6.283185307179586476925286766559;
          const double t1 = pos[idx] + d1, t2 = pos[idx] + d2;
          if (t1 >= lower[j] && t1 <= upper[j]) error = d1;
          else if (t2 >= lower[j] && t2 <= upper[j]) error = d2;
          else error = d1;
        } else if (kind[j] == 1) {
          error = shortest_angular_distance(pos[idx], c);
        } else {
          error = c - pos[idx];
        }
      } else {
        error = e_stop ? -vel[idx] : c - vel[idx];
      }
      // control_toolbox::Pid::computeCommand(error, dt)
      const double* g = pid + 5 * j;
      double out = 0;
      if (period > 0 && !isnan(error) && !isinf(error)) {
        const double error_dot = (error - last_error[idx]) / period;
        last_error[idx] = error;
        i_error[idx] += period * error;
        double i_term = g[1] * i_error[idx];
        i_term = fmax(g[4], fmin(i_term, g[3]));
        out = g[0] * error + i_term + g[2] * error_dot;
      }
      const double lim = effort_limit[j];
      r[rec_qfrc + dofadr[j]] = fmax(-lim, fmin(out, lim));
      break;
    }
  }
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
  double noise[3] = {0, 0, 0};
  int ni = 0;
  for (int k = 0; k < 3; k++) {
    if (fl & (1 << k)) {
      noise[k] = philox_normal(seed, (unsigned)env, (unsigned)n, (unsigned)k, count) * sigma[3 * n + ni] + mean[3 * n + ni];
      ni++;
    }
  }
  if (!quat) {
    for (int k = 0; k < dim && k < 3; k++) v[k] = (double)(float)(sd[adr + k] + noise[k] / cutoff);
  } else {
    // q = (setRPY(noise) * normalize(q_msg)).normalize(), sensor order (w,x,y,z)
    double w = v[0], x = v[1], y = v[2], z = v[3];
    double nrm = sqrt(w * w + x * x + y * y + z * z);
    if (nrm > 0) { w /= nrm; x /= nrm; y /= nrm; z /= nrm; }
    const double hr = noise[0] * 0.5, hp = noise[1] * 0.5, hy = noise[2] * 0.5;
    const double cr = cos(hr), sr = sin(hr), cp = cos(hp), sp = sin(hp), cy = cos(hy), sy = sin(hy);
    double rx = sr * cp * cy - cr * sp * sy, ry = cr * sp * cy + sr * cp * sy, rz = cr * cp * sy - sr * sp * cy,
           rw = cr * cp * cy + sr * sp * sy;
    nrm = sqrt(rw * rw + rx * rx + ry * ry + rz * rz);
    rw /= nrm; rx /= nrm; ry /= nrm; rz /= nrm;
    double ow = rw * w - rx * x - ry * y - rz * z;
    double ox = rw * x + rx * w + ry * z - rz * y;
    double oy = rw * y - rx * z + ry * w + rz * x;
    double oz = rw * z + rx * y - ry * x + rz * w;
    nrm = sqrt(ow * ow + ox * ox + oy * oy + oz * oz);
    v[0] = ow / nrm; v[1] = ox / nrm; v[2] = oy / nrm; v[3] = oz / nrm;
  }
}

void handle_free_plugins(Handle* h) {
  if (h->robot_hw) {
    RobotHWState* s = h->robot_hw;
    cudaFree(s->joint_id); cudaFree(s->mode); cudaFree(s->kind); cudaFree(s->qposadr); cudaFree(s->dofadr);
    cudaFree(s->effort_limit); cudaFree(s->pid); cudaFree(s->lower); cudaFree(s->upper);
    cudaFree(s->pos); cudaFree(s->vel); cudaFree(s->eff); cudaFree(s->i_error); cudaFree(s->last_error);
    cudaFree(s->hold_cmd); cudaFree(s->cmd);
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

void handle_reset_plugins(Handle* h, const uint8_t* env_mask) {
  // DefaultRobotHWSim has no reset hook of its own; PID integrators restart with the plugin reload.
  // Here a full reset clears the controller state so that rollouts are reproducible.
  if (h->robot_hw && !env_mask) {
    RobotHWState* s = h->robot_hw;
    const size_t n = (size_t)h->nenv * s->njoint * sizeof(double);
    cudaMemsetAsync(s->pos, 0, n, h->stream);
    cudaMemsetAsync(s->i_error, 0, n, h->stream);
    cudaMemsetAsync(s->last_error, 0, n, h->stream);
    cudaMemsetAsync(s->hold_cmd, 0, n, h->stream);
    s->last_e_stop = 0;
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
  std::vector<double> lim(nj), pid(5 * nj, 0.0), lo(nj), hi(nj);
  for (int j = 0; j < nj; j++) {
    const int id = cfg->joint_id[j];
    if (id < 0 || id >= m->njnt || (m->jnt_type[id] != B2MJ_JNT_HINGE && m->jnt_type[id] != B2MJ_JNT_SLIDE)) {
      set_error("b2mj_robot_hw_configure: joint " + std::to_string(j) + " is not a hinge/slide joint of the model");
      return B2MJ_EINVAL;
    }
    if (cfg->control_mode[j] < B2MJ_CTRL_EFFORT || cfg->control_mode[j] > B2MJ_CTRL_VELOCITY_PID) {
      set_error("b2mj_robot_hw_configure: unknown control mode");
      return B2MJ_EINVAL;
    }
    qa[j] = m->jnt_qposadr[id];
    da[j] = m->jnt_dofadr[id];
    kind[j] = cfg->joint_kind ? cfg->joint_kind[j] : (m->jnt_type[id] == B2MJ_JNT_SLIDE ? 2 : (m->jnt_limited[id] ? 0 : 1));
    lim[j] = cfg->effort_limit ? cfg->effort_limit[j] : 1e300;
    if (cfg->pid_gains) std::memcpy(&pid[5 * j], cfg->pid_gains + 5 * j, 5 * sizeof(double));
    lo[j] = cfg->lower_limit ? cfg->lower_limit[j] : m->jnt_range[2 * id];
    hi[j] = cfg->upper_limit ? cfg->upper_limit[j] : m->jnt_range[2 * id + 1];
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
  rc |= upload(&s->pid, pid.data(), 5 * nj);
  rc |= upload(&s->lower, lo.data(), nj);
  rc |= upload(&s->upper, hi.data(), nj);
  if (rc) return B2MJ_ECUDA;
  const size_t n = (size_t)h->nenv * nj;
  double** per[] = {&s->pos, &s->vel, &s->eff, &s->i_error, &s->last_error, &s->hold_cmd, &s->cmd};
  for (double** p : per) {
    CUDA_OK(cudaMalloc(p, n * sizeof(double)));
    CUDA_OK(cudaMemset(*p, 0, n * sizeof(double)));
  }
  return 0;
}

static int robot_hw_run(Handle* h, const double* cmd_dev, int e_stop, double period, int do_write) {
  RobotHWState* s = h->robot_hw;
  const b2k::DevModel& d = h->dm;
  const int n = h->nenv * s->njoint;
  const int rising = (e_stop && !s->last_e_stop) ? 1 : 0;
  robot_hw_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(h->rec, d.rec_pitch, d.rec_qpos, d.rec_qvel, d.rec_qfrc_applied,
                                                          h->nenv, s->njoint, s->mode, s->kind, s->qposadr, s->dofadr,
                                                          s->effort_limit, s->pid, s->lower, s->upper, s->pos, s->vel, s->eff,
                                                          s->i_error, s->last_error, s->hold_cmd, cmd_dev, e_stop, rising,
                                                          period, do_write);
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
  return robot_hw_run(h, src, e_stop, period, 1);
}

int b2mj_robot_hw_read(b2mj_handle* hh, double* pos, double* vel, double* eff) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  if (!h->robot_hw) { set_error("b2mj_robot_hw_read: call b2mj_robot_hw_configure first"); return B2MJ_ESTATE; }
  CUDA_OK(cudaSetDevice(h->device));
  if (int rc = robot_hw_run(h, nullptr, 0, 0, 0)) return rc;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  const size_t bytes = (size_t)h->nenv * h->robot_hw->njoint * sizeof(double);
  if (pos) CUDA_OK(cudaMemcpy(pos, h->robot_hw->pos, bytes, cudaMemcpyDeviceToHost));
  if (vel) CUDA_OK(cudaMemcpy(vel, h->robot_hw->vel, bytes, cudaMemcpyDeviceToHost));
  if (eff) CUDA_OK(cudaMemcpy(eff, h->robot_hw->eff, bytes, cudaMemcpyDeviceToHost));
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

int b2mj_sensor_readout(b2mj_handle* hh, double* values, double* gt) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !values) return B2MJ_EINVAL;
  const b2mjModel* m = h->model;
  if (m->nsensordata == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
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
      gt ? s->gt : nullptr);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  s->count++;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaMemcpy(values, s->values, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (gt) CUDA_OK(cudaMemcpy(gt, s->gt, n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

// NCCL is resolved at run time from the process (torch ships libnccl); no link-time dependency.
int b2mj_allgather_publish(b2mj_handle* hh, b2mj_field f, void* nccl_comm, void* dev_dst_all) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !nccl_comm || !dev_dst_all) return B2MJ_EINVAL;
  typedef int (*allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
  static allgather_fn fn = nullptr;
  if (!fn) {
    fn = (allgather_fn)dlsym(RTLD_DEFAULT, "ncclAllGather");
    if (!fn) {
      void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (lib) fn = (allgather_fn)dlsym(lib, "ncclAllGather");
    }
    if (!fn) { set_error("b2mj_allgather_publish: ncclAllGather not found in the process"); return B2MJ_EUNSUPPORTED; }
  }
  void* ptr; size_t pitch;
  if (int rc = b2mj_device_ptr(hh, f, &ptr, &pitch)) return rc;
  int is_int = 0;
  const int n = b2mj_field_size(h->model, f, &is_int);
  if (is_int) { set_error("b2mj_allgather_publish: float64 fields only"); return B2MJ_EINVAL; }
  CUDA_OK(cudaSetDevice(h->device));
  // pack the strided field into a contiguous slab at this rank's slot, then gather in place
  // (rank slot unknown here: use a private staging slab and an out-of-place gather)
  const size_t cnt = (size_t)h->nenv * n;
  if (h->publish_slab_n < cnt) {
    CUDA_OK(cudaStreamSynchronize(h->stream));  // an earlier gather may still read the old slab
    cudaFree(h->publish_slab);
    h->publish_slab = nullptr;
    h->publish_slab_n = 0;
    CUDA_OK(cudaMalloc(&h->publish_slab, cnt * sizeof(double)));
    h->publish_slab_n = cnt;
  }
  double* slab = h->publish_slab;
  CUDA_OK(cudaMemcpy2DAsync(slab, n * sizeof(double), ptr, pitch * sizeof(double), n * sizeof(double), h->nenv,
                            cudaMemcpyDeviceToDevice, h->stream));
  const int rc = fn(slab, dev_dst_all, cnt, /*ncclDouble*/ 8, nccl_comm, h->stream);
  if (rc != 0) { set_error("ncclAllGather failed with code " + std::to_string(rc)); return B2MJ_ECUDA; }
  h->launches++;
  return 0;
}

}  // extern "C"
