// ubench.cu — measured FP64 peak of the device the handle runs on, for the roofline bench.py reports.
//
// The batched step is FP64 vector arithmetic (no tensor-core contraction): the only compute roofline that can bind it is
// the DFMA issue rate ("measure it before quoting FP64 fractions").  The kernel
// below keeps 8 independent DFMA chains per thread (latency 8.8 cycles measured, tools/ubench/lat.cu) on every SM with
// full occupancy, so the FP64 pipe is the only limiter; the result is the denominator of roofline.fp64 in bench.py.
#include <cuda_runtime.h>

#include "b2mj.h"

namespace {
__global__ void __launch_bounds__(1024, 2) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[0] = s;  // never true: keeps the chains live
}
}  // namespace

extern "C" int b2mj_ubench_dfma(int device, double* tflops, double* dfma_per_clk_per_sm) {
  if (!tflops) return B2MJ_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return B2MJ_ECUDA;
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double)) != cudaSuccess) return B2MJ_ECUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4096, blocks = sms * 2, threads = 1024;
  dfma_peak_kernel<<<blocks, threads>>>(out, 64, 0.999999, 1e-9);  // warm-up
  float best_ms = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return B2MJ_ECUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best_ms) best_ms = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  const double fmas = (double)blocks * threads * (double)iters * 64.0;
  *tflops = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
  if (dfma_per_clk_per_sm) *dfma_per_clk_per_sm = fmas / (best_ms * 1e-3) / ((double)khz * 1e3) / sms;
  return 0;
}
