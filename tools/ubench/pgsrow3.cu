// cycles per PGS row update, one warp on an idle SM: the shipped owner-computes row (guard off the serial chain, sign-bit
// clamp for (0, inf) rows) and variants that bound what is left on the chain
#include <cstdio>
#include <cuda_runtime.h>
#define ROWS 21
#define ITERS 200
__global__ void k(double* out, long long* cyc, const double* AR, int nefc) {
  const int lane = threadIdx.x;
  const int me = lane < nefc ? lane : nefc - 1;
  const double* rowc = AR + nefc * nefc;
  const double iA = rowc[4 * me], Aii = rowc[4 * me + 1], lo = rowc[4 * me + 2];
  const bool pos = !(lo < 0);
  const double hA = 0.5 * Aii;
  double f = 0.1 * lane, r = 0.01 * lane - 0.1;
  long long t0, t1; int c = 0;
  // F: shipped fast row, no memory
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
      double improvement = 0; int worst = (int)0x80000000;
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        const double x = f - r * iA;
        const double fn = (pos && __double2hiint(x) < 0) ? 0.0 : x;
        const double delta = fn - f;
        const double d = __shfl_sync(0xffffffffu, delta, i);
        const double change = delta * (hA * delta + r);
        r += Aii * d;
        if (lane == i) { f = fn; improvement -= change; worst = max(worst, __double2hiint(change)); }
      }
      if (improvement > 1e300 || worst == 12345) break;
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // G: chain only: FMA, clamp, ADD, SHFL, FMA (no commit, no bookkeeping)
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        const double x = f - r * iA;
        const double fn = (pos && __double2hiint(x) < 0) ? 0.0 : x;
        const double delta = fn - f;
        const double d = __shfl_sync(0xffffffffu, delta, i);
        r += Aii * d;
      }
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // H: shuffle only the candidate force fn of lane i; every lane rebuilds delta from its copy of f_i?  (not possible:
  //    f_i lives on lane i) -> instead shuffle x and clamp after: FMA, SHFL, clamp, ADD(f_i broadcast earlier), FMA
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        const double x = f - r * iA;
        const double xi = __shfl_sync(0xffffffffu, x, i);
        const double fi = __shfl_sync(0xffffffffu, f, i);  // off chain: f_i was final one sweep ago
        const double fn = (__double2hiint(xi) < 0) ? 0.0 : xi;
        const double d = fn - fi;
        r += Aii * d;
        if (lane == i) f = fn;
      }
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // I: two rows unrolled per loop trip (loop overhead share)
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
      for (int i = 0; i + 2 <= nefc; i += 2) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const double x = f - r * iA;
          const double fn = (pos && __double2hiint(x) < 0) ? 0.0 : x;
          const double delta = fn - f;
          const double d = __shfl_sync(0xffffffffu, delta, i + u);
          r += Aii * d;
          if (lane == i + u) f = fn;
        }
      }
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // J: dependent DFMA chain only (2 per row) for reference
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
      for (int i = 0; i < nefc; i++) { const double x = f - r * iA; r += Aii * x; }
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // K: SHFL round trip only (1 double per row)
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
      for (int i = 0; i < nefc; i++) { r = __shfl_sync(0xffffffffu, r, i) + 1e-300; }
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
  const char* names[] = {"F shipped fast row", "G chain only", "H shuffle x, clamp after", "I two rows per trip", "J 2 dependent DFMA", "K SHFL + DADD"};
  for (int i = 0; i < 6; i++) printf("%-28s %7.1f cycles/row\n", names[i], (double)hc[i] / (ITERS * nefc));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
