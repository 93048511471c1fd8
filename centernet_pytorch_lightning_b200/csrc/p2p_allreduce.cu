// Gradient all-reduce over NVLink peer memory, without NCCL on the data path (training, BASELINE config 3).
//
// What it replaces: the DistributedDataParallel bucket all-reduce Lightning gives the reference (SURVEY.md section 2
// rows 17-18; CenterNet/centernet.py:70-80 + pl.Trainer).  Every rank's flat gradient buffer lives in symmetric memory
// (torch.distributed._symmetric_memory: the same allocation mapped into every peer), so a bucket is reduced by ONE kernel
// per rank that talks to its peers directly:
//   1. publish  : "my bucket is written" -> a release store (system scope) into every peer's signal pad;
//   2. wait     : until every peer has published this bucket for this step (acquire loads of the LOCAL pad);
//   3. reduce   : this rank owns 1/world of the bucket: it sums that shard over all peers' buffers with 16-byte peer
//                 loads and stores the sum back into every peer's buffer (two-shot all-reduce: reduce-scatter and
//                 all-gather in one pass over the shard); with NVSwitch multicast (multimem) the sum is formed IN THE
//                 SWITCH by multimem.ld_reduce and broadcast by multimem.st;
//   4. done     : the last CTA of the kernel publishes "my shard is everywhere" to every peer; cnb_p2p_wait (one tiny
//                 kernel before Adam) waits for all shards of all buckets.
// The kernel uses no shared memory and few CTAs, so it co-resides with the backward GEMMs that are still running; it is
// launched per bucket the moment the bucket's last gradient kernel has been enqueued (trainer.FlatTrainer).
#include "cnb_common.cuh"

namespace cnb {
namespace {

constexpr int MAXW = 8;
constexpr int MAXSLOT = 64;

struct P2PArgs {
  float* buf[MAXW];          // the flat gradient buffer as mapped for each rank
  u32* sig[MAXW];            // signal pads: [MAXSLOT][2][MAXW] u32 per rank
  float* mc;                 // multicast mapping of the buffer (nullptr: no NVLS)
  u32* counter;              // [MAXSLOT] local: CTAs finished
  const u32* epoch;          // device scalar: the step number (so that a captured graph stays valid)
  long long off, n;          // bucket = [off, off + n) floats, n % (4 * world) == 0 is arranged by the host
  int rank, world, slot;
};

__device__ __forceinline__ void st_release_sys(u32* p, u32 v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u32 ld_acquire_sys(const u32* p) {
  u32 v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int sig_idx(int slot, int phase, int src) { return (slot * 2 + phase) * MAXW + src; }

__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const P2PArgs a) {
  const u32 ep = *a.epoch;
  // 1. publish (one CTA) and 2. wait (every CTA polls its own copy of the local pad)
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.sig[threadIdx.x] + sig_idx(a.slot, 0, a.rank), ep);
  }
  if (threadIdx.x < a.world) {
    const u32* p = a.sig[a.rank] + sig_idx(a.slot, 0, threadIdx.x);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(p) < ep) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) __trap();   // 20 s: a lost peer becomes a CUDA error, not a hang
    }
  }
  __syncthreads();
  // 3. this rank's shard
  const long long shard = a.n / a.world;
  const long long s0 = a.off + (long long)a.rank * shard;
  const long long nv = shard / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const long long e = s0 + 4 * i;
    if (a.mc) {
      float4 v;
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "l"(a.mc + e)
                   : "memory");
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a.mc + e), "f"(v.x), "f"(v.y), "f"(v.z),
                   "f"(v.w)
                   : "memory");
    } else {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < MAXW; ++r) {
        if (r >= a.world) break;
        const float4 v = *reinterpret_cast<const float4*>(a.buf[r] + e);   // rank order: every rank forms the same sum
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
#pragma unroll
      for (int r = 0; r < MAXW; ++r) {
        if (r >= a.world) break;
        *reinterpret_cast<float4*>(a.buf[r] + e) = acc;
      }
    }
  }
  // 4. done: the last CTA tells every peer
  __threadfence_system();
  __syncthreads();
  __shared__ u32 s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(a.counter + a.slot, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) a.counter[a.slot] = 0;
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(a.sig[threadIdx.x] + sig_idx(a.slot, 1, a.rank), ep);
    }
  }
}

// blocks the stream until every peer's shard of slots [0, nslots) has arrived for this step
__global__ void p2p_wait_kernel(const u32* __restrict__ sig_local, const u32* __restrict__ epoch, int nslots, int world) {
  const u32 ep = *epoch;
  for (int i = threadIdx.x; i < nslots * world; i += blockDim.x) {
    const u32* p = sig_local + sig_idx(i / world, 1, i % world);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(p) < ep) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) __trap();
    }
  }
}

__global__ void p2p_bump_kernel(u32* epoch) { *epoch += 1; }

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" size_t cnb_p2p_signal_bytes(void) { return (size_t)MAXSLOT * 2 * MAXW * sizeof(u32); }

extern "C" int cnb_p2p_next_step(unsigned int* epoch_dev, cnb_stream_t stream) {
  CNB_CHECK_ARG(epoch_dev, "p2p_next_step: null pointer");
  p2p_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(epoch_dev);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_p2p_allreduce(float* const* bufs, unsigned int* const* sigs, float* multicast_or_null,
                                 unsigned int* counters, const unsigned int* epoch_dev, long long off, long long n,
                                 int rank, int world, int slot, int ctas, cnb_stream_t stream) {
  CNB_CHECK_ARG(bufs && sigs && counters && epoch_dev, "p2p_allreduce: null pointer");
  CNB_CHECK_ARG(world >= 1 && world <= MAXW && rank >= 0 && rank < world && slot >= 0 && slot < MAXSLOT, "p2p_allreduce: bad rank/world/slot");
  CNB_CHECK_ARG(n > 0 && off % 4 == 0 && n % (4 * world) == 0, "p2p_allreduce: bucket offset / length must be multiples of 4 / 4*world floats");
  P2PArgs a;
  for (int r = 0; r < MAXW; ++r) {
    a.buf[r] = r < world ? bufs[r] : nullptr;
    a.sig[r] = r < world ? sigs[r] : nullptr;
  }
  a.mc = multicast_or_null;
  a.counter = counters;
  a.epoch = epoch_dev;
  a.off = off; a.n = n; a.rank = rank; a.world = world; a.slot = slot;
  if (ctas < 1) ctas = 16;
  p2p_allreduce_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(a);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_p2p_wait(const unsigned int* sig_local, const unsigned int* epoch_dev, int nslots, int world,
                            cnb_stream_t stream) {
  CNB_CHECK_ARG(sig_local && epoch_dev && nslots >= 1 && nslots <= MAXSLOT && world >= 1 && world <= MAXW, "p2p_wait: bad argument");
  p2p_wait_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sig_local, epoch_dev, nslots, world);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
