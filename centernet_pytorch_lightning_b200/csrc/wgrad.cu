// Weight gradient of a convolution on the 5th-gen tensor cores (training path).
//
// Replaces what autograd + cuDNN give the reference for free on every nn.Conv2d of the hot path
// (CenterNet/centernet.py:70-80 training_step -> loss.backward(); layers: models/backbones/pose_dla_dcn.py:28-68,
// :165-188, :281-285, :351-370, models/heads.py:4-25, and the two GEMMs inside DCN.dcn_v2.DCN):
//   dW[co, kh, kw, ci] = sum over (n, oy, ox) of dY[n, oy, ox, co] * X[n, oy*s - p + kh, ox*s - p + kw, ci]
//
// GEMM view: D[M' = Co, N' = (tap, ci)] = A'[K' = pixels, M']^T * B'[K' = pixels, N'] -- the reduction runs over output
// PIXELS, so both operands are "MN-major" for tcgen05.mma (the contraction index is the slow one in memory):
//   * A' = a 128-pixel x 64-channel tile of dY (NHWC bf16), loaded by a tiled TMA map, 128B swizzle;
//   * B' = the matching 128-pixel x slab tile of X shifted by the filter tap, loaded by the SAME im2col TMA map the
//     forward kernel uses (conv_tma.cu): no im2col buffer, halo zero-filled by the hardware;
//   the very shared-memory tiles that are K-major operands of the forward GEMM (row = pixel, 128 B of channels) are
//   MN-major operands here -- only the descriptors change (a_major = b_major = 1; LBO = distance between 64-channel
//   tiles, SBO = 8 pixels = 1024 B) -- so one MMA covers N' = up to 4 (tap, slab) tiles = 256 columns.
// Work split: units = (Co tile, group of <= 256 N' columns, 128-pixel tile); persistent CTAs own contiguous unit
// ranges (combo-major), accumulate in TMEM over their pixel range and flush each finished (Co tile, group) segment
// with vector reductions (red.global.add.v4.f32) into the fp32 accumulation buffer dW_acc[Co][KH*KWp*Ci] (caller
// zero-fills; K order = the packed-weight order (kh, kw, ci)).  Warp roles: 0 = TMA producer, 1 = MMA issuer,
// 2..5 = flush.
#include "umma.cuh"
#include "tma_host.h"

namespace cnb {
namespace {

constexpr int PT = 128;                    // pixels per K' tile
constexpr int NTHREADS = 6 * 32;
constexpr int MAX_STAGES = 4;
constexpr u32 A_TILE = PT * 64 * 2;        // 16 KB: 128 pixels x 64 channels of dY

struct WArgs {
  cnb_conv_desc d;
  float* dw;
  int Kacc;                // row length of dw: KH * kwp * Ci
  int M, p_tiles;
  int slabW, nslab, kwp;   // N'-tile width (channels), slabs per tap, filter width the weights are packed with
  int NT, G, NG, MT, combos;
  long long units;
  int stages;
  u32 b_tile, stage_bytes;
  u32 b_layout, b_lbo, b_sbo, b_kstep16;
  u32 tmem_cols;
  int swap_lbo_sbo;        // CNB_WGRAD_SWAP (bring-up aid): exchange the two descriptor strides of both operands
};

__host__ __device__ __forceinline__ u32 make_idesc_bf16_mn(int M, int N) {
  return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy, const WArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) u64 s_full[MAX_STAGES];
  __shared__ __align__(8) u64 s_empty[MAX_STAGES];
  __shared__ __align__(8) u64 s_tfull, s_tempty;
  __shared__ u32 s_tmem;

  const cnb_conv_desc& d = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const u32 smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    mbar_init(&s_tfull, 1);
    mbar_init(&s_tempty, 4);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDy);
  }
  if (warp == 1) tmem_alloc(&s_tmem, a.tmem_cols);
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = s_tmem;
  const int HoWo = d.Ho * d.Wo;

  const long long u_begin = (long long)blockIdx.x * a.units / gridDim.x;
  const long long u_end = (long long)(blockIdx.x + 1) * a.units / gridDim.x;
  const int combo0 = (int)(u_begin / a.p_tiles);
  const int pt0 = (int)(u_begin - (long long)combo0 * a.p_tiles);

  if (warp == 0) {
    // =============================== TMA producer ==========================================================
    u32 s = 0, ph = 0;
    int combo = combo0, pt = pt0;
    for (long long u = u_begin; u < u_end; ++u) {
      const int mt = combo / a.NG, ng = combo - mt * a.NG;
      const int m0 = pt * PT;
      const int n = m0 / HoWo;
      const int rem = m0 - n * HoWo;
      const int oy = rem / d.Wo, ox = rem - oy * d.Wo;
      const int w0 = ox * d.stride - d.pad, h0 = oy * d.stride - d.pad;
      const int g_cnt = min(a.G, a.NT - ng * a.G);
      const int mw = min(2, (d.Co - mt * 128 + 63) >> 6);
      mbar_wait_parked(&s_empty[s], ph ^ 1u);
      const u32 sa = smem_base + s * a.stage_bytes;
      if (elect_one()) mbar_expect_tx(&s_full[s], (u32)mw * A_TILE + (u32)g_cnt * a.b_tile);
      for (int i = 0; i < mw; ++i)
        if (elect_one()) tma_load_2d(sa + (u32)i * A_TILE, &tmDy, mt * 128 + i * 64, m0, &s_full[s]);
      int t = ng * a.G;
      int tap = t / a.nslab, slab = t - tap * a.nslab;
      int kh = tap / a.kwp, kw = tap - kh * a.kwp;
      u32 db = sa + 2u * A_TILE;
      for (int j = 0; j < g_cnt; ++j) {
        if (elect_one())   // a tap of the zero-padded filter column (kw >= KW) repeats the last real tap; its "gradient"
                           // lands in a column of dW_acc that the unpack kernel never reads
          tma_load_im2col_4d(db, &tmX, slab * a.slabW, w0, h0, n, (unsigned short)(min(kw, d.KW - 1) * d.dil),
                             (unsigned short)(kh * d.dil), &s_full[s]);
        db += a.b_tile;
        if (++slab == a.nslab) {
          slab = 0;
          if (++kw == a.kwp) {
            kw = 0;
            ++kh;
          }
        }
      }
      if (++s == (u32)a.stages) {
        s = 0;
        ph ^= 1u;
      }
      if (++pt == a.p_tiles) {
        pt = 0;
        ++combo;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ============================================================
    u32 s = 0, ph = 0, seg = 0;
    int combo = combo0, pt = pt0;
    const u32 a_lbo = a.swap_lbo_sbo ? 1024u : A_TILE, a_sbo = a.swap_lbo_sbo ? A_TILE : 1024u;
    const u32 b_lbo = a.swap_lbo_sbo ? a.b_sbo : a.b_lbo, b_sbo = a.swap_lbo_sbo ? a.b_lbo : a.b_sbo;
    const u64 da0 = make_sdesc(smem_base, a_lbo, a_sbo, 2);
    const u64 db0 = make_sdesc(smem_base + 2u * A_TILE, b_lbo, b_sbo, a.b_layout);
    const u32 stage16 = a.stage_bytes >> 4;
    u32 soff16 = 0;
    u32 accumulate = 0;
    bool seg_open = false;
    for (long long u = u_begin; u < u_end; ++u) {
      const int mt = combo / a.NG, ng = combo - mt * a.NG;
      (void)mt;
      const int g_cnt = min(a.G, a.NT - ng * a.G);
      const u32 idesc = make_idesc_bf16_mn(128, g_cnt * a.slabW);
      if (!seg_open) {   // first unit of a (Co tile, group) segment: the accumulator must have been flushed
        mbar_wait_parked(&s_tempty, (seg & 1u) ^ 1u);
        tc_fence_after();
        accumulate = 0;
        seg_open = true;
      }
      mbar_wait_parked(&s_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        u64 da = da0 + (u64)soff16, db = db0 + (u64)soff16;
#pragma unroll
        for (int k16 = 0; k16 < PT / 16; ++k16) {   // 16 pixels per instruction
          umma_bf16(tmem_base, da, db, idesc, accumulate);
          accumulate = 1;
          da += (u64)(2048u >> 4);
          db += (u64)a.b_kstep16;
        }
        umma_commit(&s_empty[s]);
      }
      __syncwarp();
      accumulate = 1;
      soff16 += stage16;
      if (++s == (u32)a.stages) {
        s = 0;
        ph ^= 1u;
        soff16 = 0;
      }
      const bool last = (pt + 1 == a.p_tiles) || (u + 1 == u_end);
      if (last) {
        if (elect_one()) umma_commit(&s_tfull);
        __syncwarp();
        seg_open = false;
        ++seg;
      }
      if (++pt == a.p_tiles) {
        pt = 0;
        ++combo;
      }
    }
  } else {
    // =============================== flush: TMEM -> red.global.add into dW_acc ===============================
    const int q = warp & 3;
    u32 seg = 0;
    int combo = combo0, pt = pt0;
    for (long long u = u_begin; u < u_end; ++u) {
      const bool last = (pt + 1 == a.p_tiles) || (u + 1 == u_end);
      if (last) {
        const int mt = combo / a.NG, ng = combo - mt * a.NG;
        const int g_cnt = min(a.G, a.NT - ng * a.G);
        const int ncols = g_cnt * a.slabW;
        const int co = mt * 128 + 32 * q + lane;
        mbar_wait_parked(&s_tfull, seg & 1u);
        tc_fence_after();
        const u32 taddr = tmem_base + ((u32)(32 * q) << 16);
        float* row = a.dw + (size_t)co * a.Kacc + (size_t)ng * a.G * a.slabW;
        for (int c0 = 0; c0 < ncols; c0 += 16) {
          u32 v[16];
          tmem_ld16_nowait(taddr + (u32)c0, v);
          tmem_ld_wait();
          if (co < d.Co) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              red_add_v4(row + c0 + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                         __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_tempty);
        ++seg;
      }
      if (++pt == a.p_tiles) {
        pt = 0;
        ++combo;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// dW_acc[Co][(kh*kwp + kw)*Ci_pad + ci] fp32 -> dW[Co][Ci][KH][KW] fp32 (PyTorch layout), optionally accumulating
__global__ void wgrad_unpack_kernel(const float* __restrict__ acc, float* __restrict__ dw, int Co, int Ci, int Ci_pad,
                                    int KH, int KW, int kwp, int accumulate, float scale) {
  const long long total = (long long)Co * Ci * KH * KW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kw = (int)(i % KW);
    long long r = i / KW;
    const int kh = (int)(r % KH);
    r /= KH;
    const int ci = (int)(r % Ci);
    const int co = (int)(r / Ci);
    const float v = acc[(size_t)co * KH * kwp * Ci_pad + (size_t)(kh * kwp + kw) * Ci_pad + ci] * scale;
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" size_t cnb_conv_wgrad_acc_elems(int Co, int Ci_pad, int KH, int KWp) {
  return (size_t)Co * KH * KWp * Ci_pad;
}

extern "C" int cnb_conv2d_wgrad(const cnb_conv_desc* d, const void* x, const void* dy, int dy_cstride, int dy_coffset,
                                float* dw_acc, cnb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CNB_CHECK_ARG(d && x && dy && dw_acc, "conv2d_wgrad: null pointer");
  TmaDriver& drv = tma_driver();
  if (!drv.ok) {
    set_error("conv2d_wgrad: cuTensorMapEncode{Tiled,Im2col} entry points unavailable");
    return CNB_ERR_CUDA;
  }
  const int Ci = d->Ci;
  CNB_CHECK_ARG(Ci == 8 || Ci == 16 || Ci == 32 || Ci % 64 == 0, "conv2d_wgrad: Ci=%d must be 8, 16, 32 or a multiple of 64", Ci);
  CNB_CHECK_ARG(d->pad_w1 == 0, "conv2d_wgrad: rectangular filters are not supported");
  CNB_CHECK_ARG((dy_cstride & 7) == 0 && (dy_coffset & 7) == 0, "conv2d_wgrad: dY channel stride/offset must be multiples of 8");
  const long long M = (long long)d->B * d->Ho * d->Wo;
  WArgs a;
  a.d = *d;
  a.dw = dw_acc;
  a.kwp = d->w_kw > 0 ? d->w_kw : d->KW;
  a.Kacc = d->KH * a.kwp * Ci;
  a.M = (int)M;
  a.p_tiles = (int)((M + PT - 1) / PT);
  a.slabW = Ci >= 64 ? 64 : Ci;
  a.nslab = Ci >= 64 ? Ci / 64 : 1;
  a.NT = d->KH * a.kwp * a.nslab;
  const int gmax = 256 / a.slabW;
  const int ngroups = (a.NT + gmax - 1) / gmax;
  a.G = (a.NT + ngroups - 1) / ngroups;            // balanced groups of <= 256 columns
  if (a.slabW == 8) a.G = round_up(a.G, 2);        // N' must be a multiple of 16
  a.NG = (a.NT + a.G - 1) / a.G;
  CNB_CHECK_ARG(a.slabW != 8 || a.NT % 2 == 0, "conv2d_wgrad: 8-channel input needs an even number of packed taps");
  a.MT = (d->Co + 127) / 128;
  a.combos = a.MT * a.NG;
  a.units = (long long)a.combos * a.p_tiles;
  a.b_tile = (u32)(PT * a.slabW * 2);
  CUtensorMapSwizzle swz;
  switch (a.slabW) {
    case 64: a.b_layout = 2; a.b_lbo = a.b_tile; a.b_sbo = 1024; a.b_kstep16 = 2048 >> 4; swz = CU_TENSOR_MAP_SWIZZLE_128B; break;
    case 32: a.b_layout = 4; a.b_lbo = a.b_tile; a.b_sbo = 512; a.b_kstep16 = 1024 >> 4; swz = CU_TENSOR_MAP_SWIZZLE_64B; break;
    case 16: a.b_layout = 6; a.b_lbo = a.b_tile; a.b_sbo = 256; a.b_kstep16 = 512 >> 4; swz = CU_TENSOR_MAP_SWIZZLE_32B; break;
    default:  // no swizzle: 8 pixels x 16 B contiguous core matrices; LBO = next 8 pixels, SBO = next 8 channels (tile)
      a.b_layout = 0; a.b_lbo = 128; a.b_sbo = a.b_tile; a.b_kstep16 = 256 >> 4; swz = CU_TENSOR_MAP_SWIZZLE_NONE; break;
  }
  a.stage_bytes = 2u * A_TILE + (u32)a.G * a.b_tile;
  a.stage_bytes = (a.stage_bytes + 1023u) & ~1023u;
  a.stages = (int)((200 * 1024) / a.stage_bytes);
  if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
  CNB_CHECK_ARG(a.stages >= 2, "conv2d_wgrad: stage does not fit in shared memory");
  a.tmem_cols = 32;
  while ((int)a.tmem_cols < a.G * a.slabW) a.tmem_cols <<= 1;
  static const int env_swap = [] { const char* e = getenv("CNB_WGRAD_SWAP"); return e ? atoi(e) : 0; }();
  a.swap_lbo_sbo = env_swap;
  const size_t smem = (size_t)a.stages * a.stage_bytes + 1024;

  CUtensorMap tmX, tmDy;
  {
    const __nv_bfloat16* base = (const __nv_bfloat16*)x + d->x_coffset;
    cuuint64_t dims[4] = {(cuuint64_t)Ci, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_cstride * 2, (cuuint64_t)d->Wi * d->x_cstride * 2,
                             (cuuint64_t)d->Hi * d->Wi * d->x_cstride * 2};
    int lower[2] = {-d->pad, -d->pad};
    int upper[2] = {d->pad - (d->KW - 1) * d->dil, d->pad - (d->KH - 1) * d->dil};
    CNB_CHECK_ARG(lower[0] >= -128 && upper[0] >= -128 && upper[0] <= 127 && (d->KW - 1) * d->dil <= 255 &&
                      (d->KH - 1) * d->dil <= 255,
                  "conv2d_wgrad: padding / filter extent outside the TMA im2col range");
    cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    CUresult r = drv.im2col(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, lower, upper,
                            (cuuint32_t)a.slabW, (cuuint32_t)PT, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv2d_wgrad: cuTensorMapEncodeIm2col failed (%d) Ci=%d cstride=%d %dx%d k%d s%d", (int)r, Ci,
                d->x_cstride, d->Hi, d->Wi, d->KH, d->stride);
      return CNB_ERR_CUDA;
    }
    const unsigned long long bytes = (unsigned long long)d->B * d->Hi * d->Wi * d->x_cstride * 2;
    if (drv.driver_version <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&tmX)[1] &= ~(1ull << 21);
  }
  {
    const __nv_bfloat16* base = (const __nv_bfloat16*)dy + dy_coffset;
    cuuint64_t dims[2] = {(cuuint64_t)d->Co, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)dy_cstride * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)PT};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = drv.tiled(&tmDy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv2d_wgrad: cuTensorMapEncodeTiled(dY) failed (%d) Co=%d cstride=%d M=%lld", (int)r, d->Co, dy_cstride, M);
      return CNB_ERR_CUDA;
    }
  }
  static PerDeviceOnce once;
  if (once.need()) {
    CNB_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    once.mark();
  }
  const int nsm = sm_count();
  const int grid = a.units < nsm ? (int)a.units : nsm;
  CNB_CUDA(launch_pdl(conv_wgrad_kernel, dim3(grid), dim3(NTHREADS), smem, st, tmX, tmDy, a));
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_conv_unpack_wgrad(const float* dw_acc, float* dw, int Co, int Ci, int Ci_pad, int KH, int KW,
                                     int KWp, int accumulate, float scale, cnb_stream_t stream) {
  CNB_CHECK_ARG(dw_acc && dw && Co >= 1 && Ci >= 1 && Ci_pad >= Ci && KH >= 1 && KW >= 1 && KWp >= KW,
                "conv_unpack_wgrad: bad argument");
  const long long total = (long long)Co * Ci * KH * KW;
  const int grid = (int)((total + 255) / 256 > 2368 ? 2368 : (total + 255) / 256);
  wgrad_unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dw_acc, dw, Co, Ci, Ci_pad, KH, KW, KWp, accumulate, scale);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
