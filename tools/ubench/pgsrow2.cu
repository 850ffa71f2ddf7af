// PGS row chain (no memory) with 1, 4, 8, 16 warps resident on one SM: does co-residency inflate the chain?
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 100
__global__ void k(double* out, long long* cyc, int nefc, unsigned mask_arg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double iA = 0.5, Aii = 2.0, lo = 0.0, up = 1e30;
  double f = 0.1 * lane, r = 0.01 * lane - 0.1;
  const unsigned mask = mask_arg;  // runtime mask like e.mask
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
    double improvement = 0;
#pragma unroll 1
    for (int i = 0; i < nefc; i++) {
      const double ai = Aii;
      double fn = f - r * iA; fn = fn < lo ? lo : fn; fn = fn > up ? up : fn;
      double delta = fn - f; double change = delta * (0.5 * delta * Aii + r);
      const bool reject = change > 1e-10; delta = reject ? 0.0 : delta; change = reject ? 0.0 : change;
      const double d = __shfl_sync(mask, delta, i, 32); improvement -= __shfl_sync(mask, change, i, 32);
      r += ai * d; if (lane == i) f = reject ? f : fn;
    }
    if (improvement > 1e300) break;
  }
  long long t1 = clock64();
  if (lane == 0) cyc[warp] = t1 - t0;
  out[threadIdx.x] = f + r;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 64);
  for (int nw : {1, 4, 8, 16}) {
    for (int rep = 0; rep < 2; rep++) { k<<<1, 32 * nw>>>(out, cyc, 21, 0xffffffffu); cudaDeviceSynchronize(); }
    long long hc[64]; cudaMemcpy(hc, cyc, sizeof(long long) * nw, cudaMemcpyDeviceToHost);
    double s = 0; for (int w = 0; w < nw; w++) s += hc[w];
    printf("%2d warps on the SM: %.1f cycles/row (runtime mask)\n", nw, s / nw / (ITERS * 21.0));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
