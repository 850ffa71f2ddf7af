// env_ctx.cuh — per-warp environment context: one env per warp, arena in shared memory.
#pragma once
#include "dev_math.cuh"
#include "dev_model.h"

namespace b2k {

struct Env {
  const DevModel& m;
  double* sd;  // shared arena (doubles); starts with the image of the HBM state record
  int* si;     // shared arena (ints)
  double* gd;  // this env's global arena (doubles)
  int* gi;     // this env's global arena (ints)
  int lane;

  __device__ __forceinline__ double* D(int f) const {
    const int o = m.off_s[f];
    return o >= 0 ? sd + o : gd + m.off_g[f];
  }
  __device__ __forceinline__ int* I(int f) const {
    const int o = m.off_s[f];
    return o >= 0 ? si + o : gi + m.off_g[f];
  }
  __device__ __forceinline__ double* X(int xf) const {
    const int o = m.xoff_s[xf];
    return o >= 0 ? sd + o : gd + m.xoff_g[xf];
  }
};

#define FORL(i, n) for (int i = e.lane; i < (n); i += 32)
#define WSYNC() __syncwarp()

// ---- TMA 1-D bulk copy + mbarrier wrappers (sm_90+/sm_100a PTX) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace b2k
