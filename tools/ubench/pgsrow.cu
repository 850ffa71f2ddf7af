// cycles per PGS row update for variants of the owner-computes chain, one warp on an idle SM
#include <cstdio>
#include <cuda_runtime.h>
#define ROWS 21
#define ITERS 200
__global__ void k(double* out, long long* cyc, const double* AR, int nefc) {
  const int lane = threadIdx.x;
  const int me = lane < nefc ? lane : nefc - 1;
  const double* rowc = AR + nefc * nefc;
  const double iA = rowc[4 * me], Aii = rowc[4 * me + 1], lo = rowc[4 * me + 2], up = rowc[4 * me + 3];
  double f = 0.1 * lane, r = 0.01 * lane - 0.1;
  const double* col = AR + me;
  long long t0, t1; int c = 0;
  // A: full chain + ring with moves
  {
    double p0 = col[0], p1 = col[nefc], p2 = col[2 * nefc], p3 = col[3 * nefc]; int nxt = 4;
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
      double improvement = 0;
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        const double ai = p0; p0 = p1; p1 = p2; p2 = p3; p3 = col[nxt * nefc]; nxt = nxt + 1 == nefc ? 0 : nxt + 1;
        double fn = f - r * iA; fn = fn < lo ? lo : fn; fn = fn > up ? up : fn;
        double delta = fn - f; double change = delta * (0.5 * delta * Aii + r);
        const bool reject = change > 1e-10; delta = reject ? 0.0 : delta; change = reject ? 0.0 : change;
        const double d = __shfl_sync(0xffffffffu, delta, i); improvement -= __shfl_sync(0xffffffffu, change, i);
        r += ai * d; if (lane == i) f = reject ? f : fn;
      }
      if (improvement > 1e300) break;
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // B: no memory at all (ai constant)
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
      double improvement = 0;
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        const double ai = Aii;
        double fn = f - r * iA; fn = fn < lo ? lo : fn; fn = fn > up ? up : fn;
        double delta = fn - f; double change = delta * (0.5 * delta * Aii + r);
        const bool reject = change > 1e-10; delta = reject ? 0.0 : delta; change = reject ? 0.0 : change;
        const double d = __shfl_sync(0xffffffffu, delta, i); improvement -= __shfl_sync(0xffffffffu, change, i);
        r += ai * d; if (lane == i) f = reject ? f : fn;
      }
      if (improvement > 1e300) break;
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // C: no memory, no cost guard, no improvement shuffle (minimal chain)
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        double fn = f - r * iA; fn = fn < lo ? lo : fn; fn = fn > up ? up : fn;
        double delta = fn - f;
        const double d = __shfl_sync(0xffffffffu, delta, i);
        r += Aii * d; if (lane == i) f = fn;
      }
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // D: like B but fmin/fmax clamps
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
      double improvement = 0;
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        double fn = fmin(fmax(f - r * iA, lo), up);
        double delta = fn - f; double change = delta * (0.5 * delta * Aii + r);
        const bool reject = change > 1e-10; delta = reject ? 0.0 : delta; change = reject ? 0.0 : change;
        const double d = __shfl_sync(0xffffffffu, delta, i); improvement -= __shfl_sync(0xffffffffu, change, i);
        r += Aii * d; if (lane == i) f = reject ? f : fn;
      }
      if (improvement > 1e300) break;
    }
    t1 = clock64(); cyc[c++] = t1 - t0;
  }
  // E: guard decided after the broadcast (raw delta / change shuffled, every lane applies the guard)
  {
    t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
      double improvement = 0;
#pragma unroll 1
      for (int i = 0; i < nefc; i++) {
        double fn = f - r * iA; fn = fn < lo ? lo : fn; fn = fn > up ? up : fn;
        double delta = fn - f; double change = delta * (0.5 * delta * Aii + r);
        double d = __shfl_sync(0xffffffffu, delta, i); double ch = __shfl_sync(0xffffffffu, change, i);
        const bool reject = ch > 1e-10; d = reject ? 0.0 : d; ch = reject ? 0.0 : ch; improvement -= ch;
        r += Aii * d; if (lane == i) f = reject ? f : fn;
      }
      if (improvement > 1e300) break;
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
  const char* names[] = {"A ring+moves (L1)", "B no memory", "C minimal chain", "D fmin/fmax clamp", "E guard after broadcast"};
  for (int i = 0; i < 5; i++) printf("%-26s %7.1f cycles/row\n", names[i], (double)hc[i] / (ITERS * nefc));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
