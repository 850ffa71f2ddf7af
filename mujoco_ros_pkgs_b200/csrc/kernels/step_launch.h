// step_launch.h — host-callable launchers of the CUDA kernels (implemented in the .cu files).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "dev_model.h"

#define B2K_MAX_THREADS 512 /* widest CTA: the lock-stepped rollout shape (half of an SM's resident envs per CTA) */
#define B2K_MIN_CTAS 1      /* 512 threads x 128 registers = the whole register file */
#define B2K_NEWTON_MAX_NV 128 /* Newton / CG: cholSolve_warp keeps x in B2K_CHOL_SLOTS = 128 / 32 registers per lane */
#define B2K_TEAM_MIN_NV 64   /* Newton models at least this wide run in team mode (kernels/team.cuh) */
#define B2K_TEAM_WARPS 8     /* warps per env in team mode */
#define B2K_STEP_THREADS 128 /* widest CTA of a per-step launch (small CTAs free their slots as envs finish) */

extern "C" {
int b2k_launch_step(const b2k::DevModel* m, const b2k::LaunchArgs* a, int warps_per_cta, size_t smem_bytes,
                    cudaStream_t stream);
#ifndef B2K_PER_ENV_MODEL
/* the same kernel compiled with per-env model variants (step_kernel.cu with -DB2K_PER_ENV_MODEL, namespace b2k_em);
 * DevModel / LaunchArgs have one layout in both builds */
int b2k_em_launch_step(const b2k::DevModel* m, const b2k::LaunchArgs* a, int warps_per_cta, size_t smem_bytes,
                       cudaStream_t stream);
#endif
int b2k_step_kernel_attrs(int* regs, int* static_smem, int* max_threads);
int b2k_occupancy(int threads, size_t smem_bytes, int* ctas_per_sm);
int b2k_launch_order(const int* stats, const int* cost, int nenv, int* perm, int legacy, cudaStream_t stream);
}
