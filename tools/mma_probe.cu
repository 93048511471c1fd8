// tcgen05.mma issue-rate probe: one CTA per SM, one thread issues `groups` K blocks of four M=128 / K=16 MMAs back to back
// (operands are whatever shared / tensor memory holds: only timing matters) and the clocks per MMA are reported for
// N in {16..256}, A from shared memory (SS) or tensor memory (TS), 1/2/4 accumulators in rotation, with or without one
// tcgen05.commit per K block, and 1 or 4 issuing warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I centernet_pytorch_lightning_b200/csrc tools/mma_probe.cu -o tools/_bin/mma_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "umma.cuh"

using namespace cnb;

__device__ __forceinline__ void mma_ts(u32 tmem_d, u32 tmem_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct PArgs {
  int N, ts, nacc, commit_each, issuers, groups, acc_per_kk, planes;
  long long* out;   // [grid] clocks
};

__global__ void __launch_bounds__(256, 1) probe_kernel(const PArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) u64 s_done[8];
  __shared__ __align__(8) u64 s_sink[8];
  __shared__ u32 s_tmem;
  __shared__ long long s_clk[8];
  const int tid = threadIdx.x, warp = tid >> 5;
  const u32 smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<u32*>(smem_dyn + (smem_base - smem_u32(smem_dyn)))[i] = 0x3c003c00u;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&s_done[i], 1);
      mbar_init(&s_sink[i], 1u << 20);   // never completes: absorbs the per-K-block commits
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = s_tmem;
  const u32 idesc = make_idesc_bf16(128, a.N);
  const u64 da = make_sdesc(smem_base, 16, 1024, 2);               // A: 128 rows x 64 bf16, SWIZZLE_128B
  const u64 db = make_sdesc(smem_base + 16 * 1024, 16, 1024, 2);   // B: N rows x 64 bf16
  // planes: A as in conv_rows -- no swizzle, planes of 16-byte pixels (2080 bytes each), a core matrix = 8 consecutive
  // pixels of one plane, LBO = one plane, window started (g % 3) pixels into the plane
  const u32 plane = 2080;
  const u32 a_col0 = 384;                                          // TS: A stages in columns 384..511 (4 x 32)
  if (warp < a.issuers) {
    // issuer w uses accumulators w*nacc .. w*nacc + nacc - 1 (N columns each; wraps inside 384 columns)
    const u32 per = (u32)a.N;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      u32 k = 0;
      for (int g = 0; g < a.groups; ++g) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk, ++k) {
          const u32 ai = a.acc_per_kk ? (k % (u32)a.nacc) : ((u32)g % (u32)a.nacc);
          const u32 d = tmem_base + (((u32)warp * (u32)a.nacc + ai) * per) % 384u;
          if (a.ts) mma_ts(d, tmem_base + a_col0 + (u32)((g & 3) * 32 + 8 * kk), db + (u64)(2 * kk), idesc, 1u);
          else if (a.planes) umma_bf16(d, make_sdesc(smem_base + (u32)(2 * kk) * plane + 16u * (u32)(g % 3), plane, 128, 0), db + (u64)(2 * kk), idesc, 1u);
          else umma_bf16(d, da + (u64)(2 * kk), db + (u64)(2 * kk), idesc, 1u);
        }
        if (a.commit_each) umma_commit(&s_sink[warp]);
      }
      umma_commit(&s_done[warp]);
      t1 = clock64();   // issue time only
      s_clk[warp] = t1 - t0;
    }
    __syncwarp();
    mbar_wait(&s_done[warp], 0);
    tc_fence_after();
    if (elect_one()) {
      const long long t2 = clock64();
      if (warp == 0) {
        a.out[2 * blockIdx.x] = t2 - t0;          // until the last MMA has completed
        a.out[2 * blockIdx.x + 1] = s_clk[0];     // until the last MMA was issued
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  int dev = 0, sms = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 2 * sms);
  long long* h = (long long*)malloc(sizeof(long long) * 2 * sms);
  const int smem = 50 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int groups = 2048;
  printf("clocks per M=128 K=16 MMA (median over %d SMs; issue = until the last one was issued)\n", sms);
  printf("%4s %3s %5s %7s %7s %8s | %9s %9s\n", "N", "A", "nacc", "perkk", "commit", "issuers", "clk/MMA", "issue/MMA");
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int ts = 0; ts < 2; ++ts)
    for (int ni = 0; ni < 5; ++ni)
      for (int cfg = 0; cfg < 10; ++cfg) {
        // cfg: 0 one accumulator; 1 two accumulators per K block; 2 four per K=16 step; 3 = 0 + commit per K block;
        //      4 four issuers, own accumulators, commits; 5 = 2 with commits; 6 = four issuers, no commits;
        //      7 eight issuers; 8 two issuers; 9 eight issuers, conv_rows operand layout (SS only)
        PArgs a;
        a.N = Ns[ni];
        a.ts = ts;
        a.nacc = cfg == 1 ? 2 : (cfg == 2 || cfg == 5) ? 4 : 1;
        a.acc_per_kk = (cfg == 2 || cfg == 5) ? 1 : 0;
        a.commit_each = (cfg == 3 || cfg == 4 || cfg == 5) ? 1 : 0;
        a.issuers = (cfg == 4 || cfg == 6) ? 4 : (cfg == 7 || cfg == 9) ? 8 : cfg == 8 ? 2 : 1;
        a.planes = 0;
        a.groups = groups;
        a.out = out;
        if (cfg == 9) { if (ts) continue; a.planes = 1; }
        if ((long long)a.nacc * a.issuers * a.N > 384 && a.nacc * a.issuers > 1) continue;
        for (int rep = 0; rep < 2; ++rep) probe_kernel<<<sms, 256, smem>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("N=%d ts=%d cfg=%d: %s\n", a.N, ts, cfg, cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(h, out, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
        // median over SMs
        long long best[2];
        for (int w = 0; w < 2; ++w) {
          long long* v = (long long*)malloc(sizeof(long long) * sms);
          for (int i = 0; i < sms; ++i) v[i] = h[2 * i + w];
          for (int i = 0; i < sms; ++i)
            for (int j = i + 1; j < sms; ++j)
              if (v[j] < v[i]) { long long t = v[i]; v[i] = v[j]; v[j] = t; }
          best[w] = v[sms / 2];
          free(v);
        }
        const double n_mma = 4.0 * groups * a.issuers;
        printf("%4d %3s %5d %7d %7d %8d | %9.1f %9.1f\n", a.N, ts ? "TS" : a.planes ? "SSp" : "SS", a.nacc, a.acc_per_kk, a.commit_each, a.issuers,
               best[0] / n_mma, best[1] / (4.0 * groups));
      }
  return 0;
}
