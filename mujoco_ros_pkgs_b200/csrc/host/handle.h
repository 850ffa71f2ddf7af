// handle.h — internal definition of b2mj_handle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "b2mj.h"
#include "kernels/dev_model.h"

namespace b2mj {

struct RobotHWState;
struct FusedPublish;
struct SensorReadoutState;

struct Handle {
  int device = 0;
  int nenv = 0;
  b2mjModel* model = nullptr;     // host clone
  b2k::DevModel dm{};             // device view (pointers into model_blob) + layouts
  void* model_blob = nullptr;
  size_t model_blob_bytes = 0;
  cudaStream_t stream = nullptr;

  double* rec = nullptr;          // [nenv][rec_pitch] state records
  double* rec_init = nullptr;     // reset template
  double* garena_d = nullptr;     // [nenv][arena_g_doubles]
  int* garena_i = nullptr;        // [nenv][arena_g_ints]
  int* warning = nullptr;         // [nenv][8]
  int* stats = nullptr;           // [nenv][4]
  double* xfrc = nullptr;         // [nenv][6*nbody]
  double* mocap = nullptr;        // [nenv][7*nmocap]
  double* mocap_init = nullptr;
  double* rec_key = nullptr;      // keyframe reset templates (b2mj_reset_keyframe), allocated on first use
  double* mocap_key = nullptr;
  unsigned char* mask_dev = nullptr;
  int* perm = nullptr;                 // [nenv] launch slot -> env, heaviest first (refreshed after every step launch)
  int perm_valid = 0;
  // asynchronous order refresh of single-step launches (handle_launch): the order kernel runs on a side stream under the
  // NEXT step launch, which uses the order computed one launch earlier
  int* perm_async = nullptr;           // [2][nenv]
  int* cost = nullptr;                 // [2][nenv] per-env residency written by the step kernel
  cudaStream_t order_stream = nullptr;
  cudaEvent_t order_step_done[2] = {nullptr, nullptr}, order_done[2] = {nullptr, nullptr};
  uint64_t order_n = 0;                // order kernels issued so far; kernel j reads cost[j & 1], writes perm_async[j & 1]
  int* sched = nullptr;                // [1 + nenv] ticket counter + per-env chunk progress (persistent rollout)
  unsigned long long* prof = nullptr;  // [PROF_COUNT] stage cycle totals (b2mj_stage_profile), null = off

  int warps_per_cta = 4;
  int resident_envs = 0;          // envs the whole GPU keeps resident at the step launch shape (make_layout)
  int last_warps_per_cta = 0;     // shape of the most recent step / rollout launch (b2mj_launch_info reports it)
  int rollout_warps_per_cta = 0;  // CTA shape of the static fused rollout (lock-stepped stages), 0 = same as steps
  size_t smem_bytes = 0;
  size_t smem_target_bytes = 0;   // 0 = default policy
  int force_warps_per_cta = 0;
  int arena_in_smem = 0;
  int keep_intermediates = 0;
  int dump_valid = 0;
  int in_split_step = 0;
  int rk_stage = 0;               // split RK4 step: sub-step whose second half the next b2mj_step_end runs
  uint64_t launches = 0;

  // per-env model variants (b2mj_set_env_models): variant blob, env -> variant index, device view pointing at variant 0
  void* env_blob = nullptr;
  int* env_model_idx = nullptr;
  int n_env_models = 0;
  b2k::DevModel dm_env{};

  double* publish_slab = nullptr;       // staging of b2mj_allgather_publish (per handle: device- and stream-local)
  size_t publish_slab_n = 0;

  FusedPublish* fused_pub = nullptr;    // b2mj_publish_fused_*: peer slabs, flags, sequence
  const b2k::PubArgs* launch_pub = nullptr;  // set for the duration of a b2mj_step_publish launch
  int launch_pub_seq = 0;

  RobotHWState* robot_hw = nullptr;
  SensorReadoutState* sensor_ro = nullptr;
};

int handle_launch(Handle* h, int mode, int nsteps, const double* ctrl_seq = nullptr, double* traj_qpos = nullptr,
                  double* traj_qvel = nullptr, double* traj_sensor = nullptr, int chunk = 0);
void handle_free_plugins(Handle* h);
void handle_free_fused_publish(Handle* h);
void handle_reset_plugins(Handle* h, const uint8_t* env_mask);

}  // namespace b2mj
