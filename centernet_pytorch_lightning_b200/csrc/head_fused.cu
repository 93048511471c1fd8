// CenterHead in ONE kernel: conv3x3(64 -> 256, bias) -> ReLU -> conv1x1(256 -> C_out, bias) [-> sigmoid], every head.
//
// Replaces CenterNet/models/heads.py:4-25 (HeadConv.fc = Conv2d 3x3 + ReLU + Conv2d 1x1) for all heads of
// heads.py:28-43 (CenterHead).  The 256-channel intermediate -- [B,256,128,128] per head, 16.8 MB per image and head
// in fp32, the largest activation of the whole network -- never leaves the SM:
//   * GEMM 1 (implicit 3x3 conv): A = TMA-im2col tiles of the 64-channel feature map (128 pixels x 64 ch per tap), B =
//     the head's packed 3x3 filter (256 x 64 per tap), both in shared memory, accumulator D1[128 x 256] fp32 in TMEM
//     (two accumulators: the epilogue of tile i overlaps the main loop of tile i+1);
//   * epilogue 1: tcgen05.ld D1 -> + bias -> ReLU -> bf16 pairs -> tcgen05.st back into the first 128 columns of the
//     SAME TMEM accumulator (the fp32 values are dead by then);
//   * GEMM 2 (1x1 conv): tcgen05.mma with the A operand read FROM TENSOR MEMORY (those 128 columns = 128 pixels x 256
//     bf16) and B = the head's 1x1 filter resident in shared memory, D2[128 x C_out] into columns 128.. of the same
//     accumulator; issued by the MMA warp between K blocks of the next tile's main loop;
//   * epilogue 2: tcgen05.ld D2 -> + bias [-> sigmoid] -> NCHW fp32 head map (what ctdet_decode / the losses consume).
// Tiles run head-major (all pixel tiles of head 0, then head 1, ...), so the 1x1 filter in shared memory changes only
// nheads times per CTA.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..9 = epilogue (two per TMEM lane quarter).
#include "umma.cuh"
#include "tma_host.h"

namespace cnb {
namespace {

constexpr int BM = 128;
constexpr int NEPI = 8;
constexpr int NTHREADS = (2 + NEPI) * 32;
constexpr int STAGES = 3;
constexpr int MAXH = 8;
constexpr int MID = 256;                       // head_conv
constexpr int C2MAX = 96;                      // largest padded C_out handled
constexpr u32 A_BYTES = BM * 64 * 2;           // 16 KB
constexpr u32 B_BYTES = MID * 64 * 2;          // 32 KB
constexpr u32 STAGE_BYTES = A_BYTES + B_BYTES;
constexpr u32 W1_BYTES = 4 * C2MAX * 128;      // 4 K blocks x 96 rows x 128 B = 48 KB
constexpr u32 ACC_COLS = 256;                  // per accumulator; D2 lives at columns [128, 128 + C2MAX)

struct HArgs {
  int B, H, W, x_cstride, x_coffset;
  int M, m_tiles, nheads, total_tiles;
  const float* bias3;          // [nheads * 256]
  const float* bias1[MAXH];    // [c_out]
  float* y[MAXH];              // [B, c_out, H, W] fp32
  int c2[MAXH], c2pad[MAXH], act[MAXH], w1_row0[MAXH];
  int swap_halves;             // CNB_HEAD_SWAP (bring-up aid): exchange the two bf16 halves of a packed TMEM word
};

__device__ __forceinline__ void tmem_st32(u32 taddr, const u32 (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows x 16 bf16 = 8 packed columns) is read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(u32 tmem_d, u32 tmem_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
head_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmW1, const HArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) u64 s_full[STAGES], s_empty[STAGES];
  __shared__ __align__(8) u64 s_tfull[2], s_tempty[2], s_midfull[2], s_outfull[2];
  __shared__ __align__(8) u64 s_w1full, s_w1empty;
  __shared__ u32 s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const u32 smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const u32 w1_base = smem_base + STAGES * STAGE_BYTES;
  float* s_bias3 = reinterpret_cast<float*>(smem_dyn + (smem_base - smem_u32(smem_dyn)) + STAGES * STAGE_BYTES + W1_BYTES);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_tfull[i], 1);
      mbar_init(&s_tempty[i], NEPI);
      mbar_init(&s_midfull[i], NEPI);
      mbar_init(&s_outfull[i], 1);
    }
    mbar_init(&s_w1full, 1);
    mbar_init(&s_w1empty, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmW1);
  }
  if (warp == 1) tmem_alloc(&s_tmem, 512);
  pdl_launch_dependents();
  pdl_wait();
  for (int i = tid; i < a.nheads * MID; i += NTHREADS) s_bias3[i] = a.bias3 ? a.bias3[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = s_tmem;
  const int HW = a.H * a.W;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;

  if (warp == 0) {
    // =============================== TMA producer ==========================================================
    u32 s = 0, ph = 0, w1n = 0;
    int cur_head = -1;
    for (int t = cta; t < a.total_tiles; t += G) {
      const int head = t / a.m_tiles, m_tile = t - head * a.m_tiles;
      if (head != cur_head) {   // the 1x1 filter of the new head: free once the last GEMM 2 of the old head has completed
        mbar_wait_parked(&s_w1empty, (w1n & 1u) ^ 1u);
        const int rows = a.c2pad[head];
        if (elect_one()) mbar_expect_tx(&s_w1full, (u32)(4 * rows * 128));
        for (int kb = 0; kb < 4; ++kb)
          for (int r = 0; r < rows; r += 16)
            if (elect_one())
              tma_load_2d(w1_base + (u32)(kb * rows * 128 + r * 128), &tmW1, kb * 64, a.w1_row0[head] + r, &s_w1full);
        cur_head = head;
        ++w1n;
      }
      const int m0 = m_tile * BM;
      const int n = m0 / HW;
      const int rem = m0 - n * HW;
      const int oy = rem / a.W, ox = rem - oy * a.W;
      int kh = 0, kw = 0;
      for (int kb = 0; kb < 9; ++kb) {
        mbar_wait_parked(&s_empty[s], ph ^ 1u);
        const u32 sa = smem_base + s * STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&s_full[s], STAGE_BYTES);
          tma_load_im2col_4d(sa, &tmA, 0, ox - 1, oy - 1, n, (unsigned short)kw, (unsigned short)kh, &s_full[s]);
          tma_load_2d(sa + A_BYTES, &tmB, kb * 64, head * MID, &s_full[s]);
        }
        if (++kw == 3) {
          kw = 0;
          ++kh;
        }
        if (++s == STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ============================================================
    u32 s = 0, ph = 0, i = 0, w1n = 0;
    const u64 da0 = make_sdesc(smem_base, 16, 1024, 2);
    const u64 db0 = make_sdesc(smem_base + A_BYTES, 16, 1024, 2);
    const u32 idesc1 = make_idesc_bf16(BM, MID);
    bool pending = false;
    u32 p_buf = 0, p_ph = 0;
    int p_head = -1, w1_head = -1;
    bool p_last_of_head = false;
    auto gemm2 = [&]() {   // D2 = relu(D1 + b)[bf16, in TMEM] * W1^T for the tile whose intermediate is complete
      if (p_head != w1_head) {
        mbar_wait_parked(&s_w1full, w1n & 1u);
        w1_head = p_head;
        ++w1n;
      }
      tc_fence_after();
      const int rows = a.c2pad[p_head];
      const u32 idesc2 = make_idesc_bf16(BM, rows);
      const u32 acc = tmem_base + p_buf * ACC_COLS;
      if (elect_one()) {
#pragma unroll 4
        for (int kk = 0; kk < 16; ++kk) {
          const u64 db = make_sdesc(w1_base + (u32)((kk >> 2) * rows * 128 + (kk & 3) * 32), 16, 1024, 2);
          umma_bf16_ts(acc + 128, acc + (u32)(kk * 8), db, idesc2, kk ? 1u : 0u);
        }
        umma_commit(&s_outfull[p_buf]);
        if (p_last_of_head) umma_commit(&s_w1empty);
      }
      __syncwarp();
      pending = false;
    };
    for (int t = cta; t < a.total_tiles; t += G, ++i) {
      const int head = t / a.m_tiles;
      const u32 buf = i & 1u, bph = (i >> 1) & 1u;
      if (pending && head != p_head) {
        // head boundary: the producer loads the new head's 1x1 filter (and only then this tile's K blocks) once the old
        // head's last GEMM 2 has completed -- which is issued from THIS warp, so it must not wait for those K blocks first
        mbar_wait_parked(&s_midfull[p_buf], p_ph);
        gemm2();
      }
      mbar_wait_parked(&s_tempty[buf], bph ^ 1u);
      tc_fence_after();
      const u32 tmem_d = tmem_base + buf * ACC_COLS;
      for (int kb = 0; kb < 9; ++kb) {
        mbar_wait_parked(&s_full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const u64 off = (u64)((s * STAGE_BYTES) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem_d, da0 + off + (u64)(2 * kk), db0 + off + (u64)(2 * kk), idesc1, (kb | kk) ? 1u : 0u);
          umma_commit(&s_empty[s]);
        }
        __syncwarp();
        if (++s == STAGES) {
          s = 0;
          ph ^= 1u;
        }
        // warp-uniform poll (the phase only ever completes: if any lane saw it, it is complete)
        if (pending && __any_sync(0xffffffffu, mbar_test_wait(&s_midfull[p_buf], p_ph))) gemm2();
      }
      if (elect_one()) umma_commit(&s_tfull[buf]);
      __syncwarp();
      if (pending) {
        mbar_wait_parked(&s_midfull[p_buf], p_ph);
        gemm2();
      }
      pending = true;
      p_buf = buf;
      p_ph = bph;
      p_head = head;
      const int tn = t + G;
      p_last_of_head = tn >= a.total_tiles || tn / a.m_tiles != head;
    }
    if (pending) {
      mbar_wait_parked(&s_midfull[p_buf], p_ph);
      gemm2();
    }
  } else {
    // =============================== epilogue ==============================================================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    u32 i = 0;
    for (int t = cta; t < a.total_tiles; t += G, ++i) {
      const int head = t / a.m_tiles, m_tile = t - head * a.m_tiles;
      const u32 buf = i & 1u, bph = (i >> 1) & 1u;
      const int m = m_tile * BM + 32 * q + lane;
      const u32 taddr = tmem_base + buf * ACC_COLS + ((u32)(32 * q) << 16);
      mbar_wait_parked(&s_tfull[buf], bph);
      tc_fence_after();
      // ---- epilogue 1: two 64-column blocks per warp -> bias, ReLU, bf16 pairs
      u32 pk[2][32];
#pragma unroll
      for (int bi = 0; bi < 2; ++bi) {
        const int blk = half + 2 * bi;
        u32 v[64];
        tmem_ld64_nowait(taddr + (u32)(blk * 64), v);
        tmem_ld_wait();
        const float* bp = s_bias3 + head * MID + blk * 64;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float lo = fmaxf(__uint_as_float(v[2 * j]) + bp[2 * j], 0.f);
          const float hi = fmaxf(__uint_as_float(v[2 * j + 1]) + bp[2 * j + 1], 0.f);
          pk[bi][j] = a.swap_halves ? pack_bf16x2(hi, lo) : pack_bf16x2(lo, hi);
        }
      }
      // both warps of the quarter have read all four fp32 blocks before any of them is overwritten
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      tmem_st32(taddr + (u32)(half * 32), pk[0]);
      tmem_st32(taddr + (u32)((half + 2) * 32), pk[1]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_midfull[buf]);
      // ---- epilogue 2
      mbar_wait_parked(&s_outfull[buf], bph);
      tc_fence_after();
      const int c2 = a.c2[head], ngr = a.c2pad[head] / 16;
      const int act = a.act[head];
      const float* b1 = a.bias1[head];
      int on = 0, opix = 0;
      if (m < a.M) {
        on = m / HW;
        opix = m - on * HW;
      }
      for (int g = half; g < ngr; g += 2) {
        u32 v[16];
        tmem_ld16_nowait(taddr + 128u + (u32)(g * 16), v);
        tmem_ld_wait();
        if (m < a.M) {
          float* yp = a.y[head] + ((size_t)on * c2 + g * 16) * HW + opix;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = g * 16 + j;
            if (c < c2) {
              float f = __uint_as_float(v[j]) + (b1 ? __ldg(b1 + c) : 0.f);
              if (act == 1) f = fmaxf(f, 0.f);
              else if (act == 2) f = __fdividef(1.f, 1.f + __expf(-f));
              *yp = f;
            }
            yp += HW;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_tempty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" int cnb_head_fused_supported(const cnb_head_desc* d) {
  if (!d || d->Ci != 64 || d->head_conv != MID || d->nheads < 1 || d->nheads > MAXH) return 0;
  for (int h = 0; h < d->nheads; ++h)
    if (d->c_out[h] < 1 || (d->c_out[h] + 15) / 16 * 16 > C2MAX) return 0;
  return 1;
}

extern "C" int cnb_head_fused_fprop(const cnb_head_desc* d, const void* x, const void* w3pk, const float* bias3,
                                    const void* w1pk, const float* const* bias1, float* const* y, cnb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CNB_CHECK_ARG(d && x && w3pk && w1pk && y, "head_fused_fprop: null pointer");
  CNB_CHECK_ARG(cnb_head_fused_supported(d), "head_fused_fprop: needs Ci = 64, head_conv = 256, <= %d heads of <= %d channels", MAXH, C2MAX);
  TmaDriver& drv = tma_driver();
  if (!drv.ok) {
    set_error("head_fused_fprop: cuTensorMapEncode{Tiled,Im2col} entry points unavailable");
    return CNB_ERR_CUDA;
  }
  HArgs a;
  a.B = d->B; a.H = d->H; a.W = d->W; a.x_cstride = d->x_cstride; a.x_coffset = d->x_coffset;
  const long long M = (long long)d->B * d->H * d->W;
  a.M = (int)M;
  a.m_tiles = (int)((M + BM - 1) / BM);
  a.nheads = d->nheads;
  a.total_tiles = a.m_tiles * a.nheads;
  a.bias3 = bias3;
  int row0 = 0;
  for (int h = 0; h < MAXH; ++h) {
    const bool on = h < d->nheads;
    a.c2[h] = on ? d->c_out[h] : 0;
    a.c2pad[h] = on ? (d->c_out[h] + 15) / 16 * 16 : 0;
    a.act[h] = on ? d->act[h] : 0;
    a.bias1[h] = (on && bias1) ? bias1[h] : nullptr;
    a.y[h] = on ? y[h] : nullptr;
    a.w1_row0[h] = row0;
    row0 += a.c2pad[h];
    CNB_CHECK_ARG(!on || a.y[h], "head_fused_fprop: null output for head %d", h);
  }
  static const int env_swap = [] { const char* e = getenv("CNB_HEAD_SWAP"); return e ? atoi(e) : 0; }();
  a.swap_halves = env_swap;

  CUtensorMap tmA, tmB, tmW1;
  {
    const __nv_bfloat16* base = (const __nv_bfloat16*)x + d->x_coffset;
    cuuint64_t dims[4] = {64, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_cstride * 2, (cuuint64_t)d->W * d->x_cstride * 2,
                             (cuuint64_t)d->H * d->W * d->x_cstride * 2};
    int lower[2] = {-1, -1}, upper[2] = {-1, -1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = drv.im2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, lower, upper, 64u,
                            (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("head_fused_fprop: cuTensorMapEncodeIm2col failed (%d)", (int)r);
      return CNB_ERR_CUDA;
    }
    const unsigned long long bytes = (unsigned long long)d->B * d->H * d->W * d->x_cstride * 2;
    if (drv.driver_version <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&tmA)[1] &= ~(1ull << 21);
  }
  {
    cuuint64_t dims[2] = {576, (cuuint64_t)(d->nheads * MID)};
    cuuint64_t strides[1] = {576 * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)MID}, estr[2] = {1, 1};
    CUresult r = drv.tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w3pk, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("head_fused_fprop: cuTensorMapEncodeTiled(W3) failed (%d)", (int)r);
      return CNB_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)MID, (cuuint64_t)row0};
    cuuint64_t strides[1] = {(cuuint64_t)MID * 2};
    cuuint32_t box[2] = {64u, 16u}, estr[2] = {1, 1};
    CUresult r = drv.tiled(&tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w1pk, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("head_fused_fprop: cuTensorMapEncodeTiled(W1) failed (%d)", (int)r);
      return CNB_ERR_CUDA;
    }
  }
  const size_t smem = (size_t)STAGES * STAGE_BYTES + W1_BYTES + (size_t)d->nheads * MID * 4 + 1024;
  static PerDeviceOnce once;
  if (once.need()) {
    CNB_CUDA(cudaFuncSetAttribute(head_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    once.mark();
  }
  const int nsm = sm_count();
  const int grid = a.total_tiles < nsm ? a.total_tiles : nsm;
  CNB_CUDA(launch_pdl(head_fused_kernel, dim3(grid), dim3(NTHREADS), smem, st, tmA, tmB, tmW1, a));
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
