// cycles per PGS row update when the candidate change travels through shared memory (predicated STS by the owner lane,
// __syncwarp, broadcast LDS by all lanes) instead of a 64-bit SHFL.IDX (two shuffles behind a BRA.DIV)
#include <cstdio>
#include <cuda_runtime.h>
#define ROWS 21
#define ITERS 200
__global__ void k(double* out, long long* cyc, const double* AR, int nefc) {
  __shared__ double slot[2];
  const int lane = threadIdx.x;
  const int me = lane < nefc ? lane : nefc - 1;
  const double* rowc = AR + nefc * nefc;
  const double iA = rowc[4 * me], Aii = rowc[4 * me + 1], lo = rowc[4 * me + 2];
  const bool pos = !(lo < 0);
  const double hA = 0.5 * Aii;
  double f = 0.1 * lane, r = 0.01 * lane - 0.1;
  long long t0, t1; int c = 0;
  // L1: full row (commit, improvement, guard bookkeeping), broadcast through shared memory, two alternating slots
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
      double improvement = 0; int worst = (int)0x80000000;
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        const double x = f - r * iA;
        const double fn = (pos && __double2hiint(x) < 0) ? 0.0 : x;
        const double delta = fn - f;
        if (lane == i) slot[i & 1] = delta;
        __syncwarp();
        const double d = slot[i & 1];
        const double change = delta * (hA * delta + r);
        r += Aii * d;
        if (lane == i) { f = fn; improvement -= change; worst = max(worst, __double2hiint(change)); }
      }
      if (improvement > 1e300 || worst == 12345) break;
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // L2: chain only through shared memory
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        const double x = f - r * iA;
        const double fn = (pos && __double2hiint(x) < 0) ? 0.0 : x;
        const double delta = fn - f;
        if (lane == i) slot[i & 1] = delta;
        __syncwarp();
        r += Aii * slot[i & 1];
      }
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // L3: like L1, two rows per loop trip
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
      double improvement = 0; int worst = (int)0x80000000;
#pragma unroll 1
      for (int i = 0; i + 2 <= nefc; i += 2) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const double x = f - r * iA;
          const double fn = (pos && __double2hiint(x) < 0) ? 0.0 : x;
          const double delta = fn - f;
          if (lane == i + u) slot[u] = delta;
          __syncwarp();
          const double d = slot[u];
          const double change = delta * (hA * delta + r);
          r += Aii * d;
          if (lane == i + u) { f = fn; improvement -= change; worst = max(worst, __double2hiint(change)); }
        }
      }
      if (improvement > 1e300 || worst == 12345) break;
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // L4: the OWNER publishes the new force candidate x (one FMA earlier), every lane clamps and forms the change itself
  //     from a copy of f_i broadcast the sweep before:  FMA, STS, sync, LDS, clamp, ADD, FMA
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        const double x = f - r * iA;
        if (lane == i) { slot[0] = x; slot[1] = f; }
        __syncwarp();
        const double xi = slot[0], fi = slot[1];
        const double fn = (__double2hiint(xi) < 0) ? 0.0 : xi;
        r += Aii * (fn - fi);
        if (lane == i) f = fn;
        __syncwarp();
      }
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  out[lane] = f + r;
}
int main() {
  const int nefc = ROWS;
  double* AR; double* out; long long* cyc;
  cudaMalloc(&AR, (nefc * nefc + 4 * nefc) * 8); cudaMalloc(&out, 256); cudaMalloc(&cyc, 64);
  double* h = new double[nefc * nefc + 4 * nefc];
  for (int i = 0; i < nefc * nefc; i++) h[i] = (i % (nefc + 1) == 0) ? 2.0 : 0.01;
  for (int i = 0; i < nefc; i++) { h[nefc * nefc + 4 * i] = 0.5; h[nefc * nefc + 4 * i + 1] = 2.0; h[nefc * nefc + 4 * i + 2] = 0; h[nefc * nefc + 4 * i + 3] = 1e30; }
  cudaMemcpy(AR, h, (nefc * nefc + 4 * nefc) * 8, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(out, cyc, AR, nefc); cudaDeviceSynchronize(); }
  long long hc[8]; cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
  const char* names[] = {"L1 smem broadcast, full row", "L2 smem broadcast, chain", "L3 smem, two rows per trip", "L4 publish x, clamp after"};
  for (int i = 0; i < 4; i++) printf("%-30s %7.1f cycles/row\n", names[i], (double)hc[i] / (ITERS * nefc));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
