// latency microbenchmarks for the fused-step cost model: one warp, dependent chains
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void k(double* out, long long* cyc, const double* g, const int* gi, int stride) {
  __shared__ double sm[1024];
  __shared__ int smi[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) { sm[i] = 1.0 + 1e-9 * i; smi[i] = (i * 7 + 1) & 1023; }
  __syncwarp();
  double x = g[threadIdx.x], y = 1.000001;
  long long t0, t1; int c = 0;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
  t1 = clock64(); cyc[c++] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x + y;
  t1 = clock64(); cyc[c++] = t1 - t0;
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x * y;
  t1 = clock64(); cyc[c++] = t1 - t0;
  // double shuffle chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  t1 = clock64(); cyc[c++] = t1 - t0;
  // LDS pointer chase (int)
  int p = threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) p = smi[p];
  t1 = clock64(); cyc[c++] = t1 - t0;
  // LDS double + DADD chain (index depends on value)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) { x = x + sm[p]; p = (p + (x > 1e300)) & 1023; }
  t1 = clock64(); cyc[c++] = t1 - t0;
  // global pointer chase, small footprint (L1 hit)
  int q = threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) q = gi[q];
  t1 = clock64(); cyc[c++] = t1 - t0;
  // global pointer chase, large stride (L2 hit)
  int q2 = threadIdx.x * stride;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) q2 = __ldcg(gi + (1 << 20) + q2);
  t1 = clock64(); cyc[c++] = t1 - t0;
  // double division chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) x = 1.0 / (x + 1.5);
  t1 = clock64(); cyc[c++] = t1 - t0;
  // sqrt chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) x = sqrt(x + 1.5);
  t1 = clock64(); cyc[c++] = t1 - t0;
  // sincos chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) { double s, co; sincos(x, &s, &co); x = s + co; }
  t1 = clock64(); cyc[c++] = t1 - t0;
  // int IMAD chain
  int a = threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = a * 3 + 1;
  t1 = clock64(); cyc[c++] = t1 - t0;
  // DSETP + select chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x < y ? y : x * 1.0000001;
  t1 = clock64(); cyc[c++] = t1 - t0;
  // syncwarp chain + smem store/load roundtrip (lane i writes, lane i+1 reads)
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); cyc[c++] = t1 - t0;
  // global store/load roundtrip between lanes (same warp) via L1/L2
  double* gg = const_cast<double*>(g) + 4096;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 256; i++) { gg[threadIdx.x] = x; __syncwarp(); x = gg[(threadIdx.x + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); cyc[c++] = (t1 - t0) * (N / 256);
  out[threadIdx.x] = x + p + q + q2 + a;
}
int main() {
  double *g, *out; int* gi; long long* cyc;
  cudaMalloc(&g, 1 << 20); cudaMalloc(&out, 4096); cudaMalloc(&gi, (1 << 22) * 4 + (1 << 22)); cudaMalloc(&cyc, 64 * 8);
  int* h = new int[1 << 21];
  for (int i = 0; i < (1 << 20); i++) h[i] = (i * 7 + 1) & 255;            // small footprint chase
  const int stride = 64;
  for (int i = 0; i < (1 << 20); i++) h[(1 << 20) + i] = (i * 1027 + 64 * 33) & ((1 << 20) - 1) & ~63;  // big-stride chase
  cudaMemcpy(gi, h, (1 << 21) * 4, cudaMemcpyHostToDevice);
  double hd[32]; for (int i = 0; i < 32; i++) hd[i] = 1.0 + i * 1e-3;
  cudaMemcpy(g, hd, sizeof hd, cudaMemcpyHostToDevice);
  const char* names[] = {"DFMA", "DADD", "DMUL", "SHFL.f64", "LDS int chase", "LDS f64+DADD", "LDG L1 chase", "LDG.CG L2 chase",
                         "DDIV(1/x)+DADD", "DSQRT+DADD", "sincos+DADD", "IMAD", "DSETP+SEL+DMUL", "STS+sync+LDS+DADD", "STG+sync+LDG+DADD"};
  for (int rep = 0; rep < 2; rep++) {
    k<<<1, 32>>>(out, cyc, g, gi, stride);
    cudaDeviceSynchronize();
  }
  long long hc[64]; cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 15; i++) printf("%-22s %7.1f cycles/iter\n", names[i], (double)hc[i] / N);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
