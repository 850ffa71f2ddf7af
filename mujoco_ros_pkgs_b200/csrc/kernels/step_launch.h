// step_launch.h — host-callable launchers of the CUDA kernels (implemented in the .cu files).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "dev_model.h"

#define B2K_MAX_THREADS 128
#define B2K_MIN_CTAS 4   /* 4 x 128 threads = 16 warps per SM -> at most 128 registers per thread */

extern "C" {
int b2k_launch_step(const b2k::DevModel* m, const b2k::LaunchArgs* a, int warps_per_cta, size_t smem_bytes,
                    cudaStream_t stream);
int b2k_step_kernel_attrs(int* regs, int* static_smem, int* max_threads);
int b2k_occupancy(int threads, size_t smem_bytes, int* ctas_per_sm);
}
