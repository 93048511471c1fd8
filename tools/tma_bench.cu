// Microbenchmark: per-SM throughput of TMA loads into a shared-memory ring, im2col mode vs tiled mode.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bench tma_bench.cu -lcuda && ./tma_bench
// Tensor: NHWC bf16 [32][128][128][64] (67 MB).  Each CTA walks "tiles" of 128 pixels x 9 taps x 64 ch like the
// conv kernel's producer, one elected thread issues, the same thread retires (no MMA), S stages in flight.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef unsigned int u32;
typedef unsigned long long u64;

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, u32 parity) {
  u32 ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void* bar, u32 parity) {
  u32 spins = 0;
  while (!mbar_try_wait(bar, parity)) if (++spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_im2col(u32 dst, const void* tm, int c, int w, int h, int n, unsigned short ow,
                                           unsigned short oh, void* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst), "l"((u64)tm), "r"(smem_u32(bar)), "r"(c),
               "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh) : "memory");
}
__device__ __forceinline__ void tma_tile4d(u32 dst, const void* tm, int c, int w, int h, int n, void* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"((u64)tm), "r"(smem_u32(bar)), "r"(c), "r"(w),
               "r"(h), "r"(n) : "memory");
}

constexpr int S = 8;
constexpr int H = 128, W = 128, C = 64, B = 32;

// mode 0: im2col 128 px; mode 1: tiled box {64,16,8,1}; mode 2: tiled box {64,128,1,1}
__global__ void __launch_bounds__(128, 1) bench(const __grid_constant__ CUtensorMap tm, int mode, int tiles_per_cta,
                                                 long long* cycles_out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) u64 full[S];
  const u32 base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int total_steps = tiles_per_cta * 9;
  long long t0 = clock64();
  int issued = 0, retired = 0;
  while (retired < total_steps) {
    while (issued < total_steps && issued - retired < S) {
      const int s = issued % S;
      const int tile = blockIdx.x * tiles_per_cta + issued / 9;     // 128-pixel tile id
      const int tap = issued % 9, kh = tap / 3, kw = tap % 3;
      const int m0 = tile * 128;
      const int n = (m0 / (H * W)) % B, rem = m0 % (H * W);
      mbar_expect_tx(&full[s], 128 * 128);
      if (mode == 0) {
        const int oy = rem / W, ox = rem % W;
        tma_im2col(base + s * 16384, &tm, 0, ox - 1, oy - 1, n, (unsigned short)kw, (unsigned short)kh, &full[s]);
      } else if (mode == 1) {
        const int t = rem / 128;                       // 8x16 block index inside the image: 8 blocks across, 16 down
        const int by = t / 8, bx = t % 8;
        tma_tile4d(base + s * 16384, &tm, 0, bx * 16 + kw - 1, by * 8 + kh - 1, n, &full[s]);
      } else {
        const int oy = rem / W;
        tma_tile4d(base + s * 16384, &tm, 0, kw - 1, oy + kh - 1, n, &full[s]);
      }
      ++issued;
    }
    mbar_wait(&full[retired % S], (retired / S) & 1);
    ++retired;
  }
  long long t1 = clock64();
  cycles_out[blockIdx.x] = t1 - t0;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main() {
  __nv_bfloat16* x;
  const size_t n = (size_t)B * H * W * C;
  CK(cudaMalloc(&x, n * 2));
  CK(cudaMemset(x, 0, n * 2));
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * 8));
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, S * 16384 + 1024));
  for (int mode = 0; mode < 3; ++mode) {
    CUtensorMap tm;
    CUresult r;
    if (mode == 0) {
      int lower[2] = {-1, -1}, upper[2] = {-1, -1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      r = cuTensorMapEncodeIm2col(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, lower, upper, 64, 128, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint32_t box1[4] = {64, 16, 8, 1}, box2[4] = {64, 128, 1, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, mode == 1 ? box1 : box2, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("encode failed mode %d: %d\n", mode, (int)r); continue; }
    for (int grid : {1, 148}) {
      const int tiles_per_cta = 27;
      for (int rep = 0; rep < 2; ++rep) {
        bench<<<grid, 128, S * 16384 + 1024>>>(tm, mode, tiles_per_cta, cyc);
        CK(cudaDeviceSynchronize());
      }
      long long h[148];
      CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)tiles_per_cta * 9 * 16384;
      printf("mode %d grid %3d: %lld cycles for %d loads of 16 KB -> %.1f B/cycle/SM, %.0f cycles/load\n", mode, grid, mx,
             tiles_per_cta * 9, bytes / mx, (double)mx / (tiles_per_cta * 9));
    }
  }
  return 0;
}
