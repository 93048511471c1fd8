// Shared helpers for the centernet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <utility>
#include <stdlib.h>
#include "../../include/centernet_b200.h"

namespace cnb {

// ---- host-side error plumbing ------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define CNB_CHECK_ARG(cond, ...)                       \
  do {                                                 \
    if (!(cond)) {                                     \
      ::cnb::set_error(__VA_ARGS__);                   \
      return CNB_ERR_INVALID;                          \
    }                                                  \
  } while (0)

#define CNB_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::cnb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                       __LINE__);                                                          \
      return CNB_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define CNB_LAUNCH_CHECK()                                                                 \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::cnb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),         \
                       __FILE__, __LINE__);                                                \
      return CNB_ERR_CUDA;                                                                 \
    }                                                                                      \
    ::cnb::count_launch();                                                                 \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE attribute: `static PerDeviceOnce once;
// if (once.need()) { ...set attributes...; once.mark(); }` configures each kernel once on every device a process uses
// (setting it twice from two racing threads is harmless).
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
  bool need() const { return !((done.load(std::memory_order_acquire) >> dev()) & 1ull); }
  void mark() { done.fetch_or(1ull << dev(), std::memory_order_release); }
};
// SM count of the current device (cached per device)
inline int sm_count() {
  static std::atomic<int> cache[64];
  const int d = PerDeviceOnce::dev();
  int n = cache[d].load(std::memory_order_relaxed);
  if (n <= 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
    cache[d].store(n, std::memory_order_relaxed);
  }
  return n;
}

// kernel launch with the programmatic-stream-serialization attribute (CNB_PDL=0: plain launch)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool on = [] { const char* e = getenv("CNB_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- device helpers ----------------------------------------------------------------------------
typedef unsigned long long u64;
typedef unsigned int u32;

__device__ __forceinline__ u32 smem_u32(const void* p) {
  return (u32)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(void* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, u32 parity) {
  u32 ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a lost arrival traps (-> CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(void* bar, u32 parity) {
  u32 spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 27)) __trap();
  }
}
// The same with a suspend-time hint, for the warp-specialised kernels: the hardware parks the waiting thread
// until the phase completes (or the hint expires) instead of returning at once, so an idle role does not burn
// issue slots that the working warps of the same SM sub-partition need.
__device__ __forceinline__ bool mbar_try_wait_parked(void* bar, u32 parity) {
  u32 ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parked(void* bar, u32 parity) {
  if (mbar_try_wait_parked(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait_parked(bar, parity)) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 4000000000ull) __trap();   // 4 s: a lost arrival becomes a CUDA error, not a hang
  }
}
// The same on a precomputed shared-memory address: single-warp role loops that touch a barrier per K block keep the
// addresses as loop-carried state (the generic-to-shared conversion of `&bar[i]` costs a special-register read and a few
// dependent uniform instructions each time, which a lone warp cannot hide).
__device__ __forceinline__ bool mbar_try_wait_parked_a(u32 addr, u32 parity) {
  u32 ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parked_a(u32 addr, u32 parity) {
  if (mbar_try_wait_parked_a(addr, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait_parked_a(addr, parity)) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 4000000000ull) __trap();   // 4 s: a lost arrival becomes a CUDA error, not a hang
  }
}
// Non-blocking poll (mbarrier.test_wait never suspends the thread): for single-thread roles whose wake-up latency
// after a suspended try_wait would sit on the critical path.
__device__ __forceinline__ bool mbar_test_wait(void* bar, u32 parity) {
  u32 ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_spin(void* bar, u32 parity) {
  u32 spins = 0;
  while (!mbar_test_wait(bar, parity)) {
    if (++spins > (1u << 28)) __trap();
  }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0, 16B-aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, u32 bytes, void* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// One elected lane of a converged warp.  The single-thread instructions of the tensor-core kernels (tcgen05.mma,
// tcgen05.commit, TMA loads, expect_tx) are issued as `if (elect_one()) ...` from loops that ALL lanes of the role's
// warp execute: operands computed in warp-uniform control flow live in uniform registers, which is what UTCHMMA /
// UTMALDG take.  Inside an `if (lane == 0)` region the compiler cannot prove uniformity and wraps every such
// instruction in a R2UR.BROADCAST waterfall loop (~150 clocks per tcgen05.mma, measured).
__device__ __forceinline__ bool elect_one() {
  u32 pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p mov.u32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}

// Programmatic dependent launch: a kernel launched with launch_pdl() may start (prologue: barrier init, TMEM
// allocation, constants) while the previous kernel in the stream drains; pdl_wait() blocks until that kernel has
// completed and its writes are visible -- it must precede every access to activations.  Both are no-ops for a
// normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

}  // namespace cnb
