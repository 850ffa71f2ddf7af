// env_ctx.cuh — per-warp environment context: one env per warp, arena in shared memory.
#pragma once
#include "dev_math.cuh"
#include "dev_model.h"

namespace b2k {

// The device view of the model of the handle that is currently stepping lives in constant memory:
// sizes, options, layout offsets and array pointers become constant-bank operands of the instructions
// that use them.  b2k_launch_step re-uploads it when a different handle (or an edited model) launches.
__constant__ DevModel c_dm;

// dynamic shared memory of the step kernel: [nwarp mbarriers, 16 B each][nwarp env blocks: doubles | ints].
// Declared at namespace scope so every accessor forms its address from the shared symbol itself and
// the compiler emits LDS / STS (not generic LD / ST) for arena traffic.
extern __shared__ __align__(16) unsigned char b2k_smem[];

#ifdef B2K_PER_ENV_MODEL
// byte offset of this warp's model variant: written by the env's main warp into the spare half of its 16-byte
// mbarrier slot when the env is picked up (step_kernel.cu), read back by every model array access
__device__ __forceinline__ long long env_model_offset() {
  const int tw = c_dm.team_warps;
  const int slot = ((int)threadIdx.x / 32) / tw;
  return *reinterpret_cast<const long long*>(b2k_smem + (tw > 1 ? 64 : 16) * slot + 8);
}
#endif

struct Env {
  unsigned sbd;  // byte offset in b2k_smem of this env's shared arena (doubles; starts with the record image)
  unsigned sbi;  // byte offset of the int part
  double* gd;    // this env's HBM/L2 arena (doubles)
  int* gi;       // this env's HBM/L2 arena (ints)
  int lane;      // lane within the env's group: 0 .. B2K_G-1
  unsigned mask; // the group's lanes within the warp (operand of every *_sync collective)
  int dump;      // the arena will be read back (keep_intermediates / split step): also produce API-only fields

  __device__ __forceinline__ double* sd() const { return reinterpret_cast<double*>(b2k_smem + sbd); }
  __device__ __forceinline__ int* si() const { return reinterpret_cast<int*>(b2k_smem + sbi); }
  // fields that are always shared-memory resident (everything the host never demotes)
  __device__ __forceinline__ double* D(int f) const {
    return reinterpret_cast<double*>(b2k_smem + sbd + 8u * (unsigned)c_dm.off_s[f]);
  }
  __device__ __forceinline__ int* I(int f) const {
    return reinterpret_cast<int*>(b2k_smem + sbi + 4u * (unsigned)c_dm.off_s[f]);
  }
  __device__ __forceinline__ double* X(int xf) const {
    return reinterpret_cast<double*>(b2k_smem + sbd + 8u * (unsigned)c_dm.xoff_s[xf]);
  }
  // fields the host may demote to the env's HBM/L2 arena when they are large (handle.cu::make_layout)
  __device__ __forceinline__ double* DG(int f) const {
    const int o = c_dm.off_s[f];
    return o >= 0 ? sd() + o : gd + c_dm.off_g[f];
  }
  __device__ __forceinline__ int* IG(int f) const {
    const int o = c_dm.off_s[f];
    return o >= 0 ? si() + o : gi + c_dm.off_g[f];
  }
  __device__ __forceinline__ double* XG(int xf) const {
    const int o = c_dm.xoff_s[xf];
    return o >= 0 ? sd() + o : gd + c_dm.xoff_g[xf];
  }
};

// Pair table of a row-major lower triangle of order nv: entry p -> (a, b <= a) packed a | b << 8 (XF_TRI).  The
// factorisations (sparse L'DL of qM, Newton Cholesky) hand the entries of their triangular trailing blocks to lanes
// through it: no integer division, no wasted upper-triangle slots.  Built once per work item.
__device__ __forceinline__ const unsigned short* triTable(const Env e) {
  return c_dm.xsize[XF_TRI] > 0 ? reinterpret_cast<const unsigned short*>(e.X(XF_TRI)) : nullptr;
}
__device__ __noinline__ void triTableBuild(const Env e) {
  if (c_dm.xsize[XF_TRI] <= 0) return;
  unsigned short* tri = reinterpret_cast<unsigned short*>(e.X(XF_TRI));
  const int n = c_dm.nv, ntri = n * (n + 1) / 2;
  _Pragma("unroll 1") for (int p = e.lane; p < ntri; p += B2K_G) {
    int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while (a * (a + 1) / 2 > p) a--;
    while ((a + 1) * (a + 2) / 2 <= p) a++;
    tri[p] = (unsigned short)(a | ((p - a * (a + 1) / 2) << 8));
  }
  __syncwarp(e.mask);
}

// The megakernel is instruction-fetch bound (profiles/): run-time-bounded loops stay rolled so their bodies are
// re-executed from the instruction cache instead of being fetched as straight-line copies.
#define B2K_NOUNROLL _Pragma("unroll 1")
// Row dot products whose operand streams from the env's L2 arena (efc_J): four loads in flight per trip, ONE accumulator,
// so the summation order -- and the result -- is that of the rolled loop while the exposed L2 latency drops fourfold.
#define B2K_UNROLL4 _Pragma("unroll 4")

// lane-strided loop; never unrolled: trip counts are 1-2 for the models this kernel targets and the
// megakernel is instruction-cache bound, so code size matters more than loop overhead
#define FORL(i, n) _Pragma("unroll 1") for (int i = e.lane; i < (n); i += B2K_G)
#define WSYNC() __syncwarp(e.mask)

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
__device__ __forceinline__ void bulk_commit_wait_all() {  // waits for the global writes themselves
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace b2k
