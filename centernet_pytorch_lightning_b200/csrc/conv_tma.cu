// Implicit-GEMM convolution, TMA-fed and warp-specialised, on the 5th-gen tensor cores.
//
// Replaces (reference file:line under CenterNet/models/): the nn.Conv2d + nn.BatchNorm2d(eval) [+ residual]
// [+ ReLU] chains of backbones/pose_dla_dcn.py:28-68 (BasicBlock), :165-188 (Root), :351-370 (conv levels),
// :281-285 (stem), backbones/resnet_dcn.py:29-128 / msra_resnet.py:25-100 (BasicBlock, Bottleneck),
// heads.py:4-25 (HeadConv) and the conv_offset_mask conv of DCN.dcn_v2.DCN (pose_dla_dcn.py:441).
//
// GEMM view: D[M = B*Ho*Wo, N = Co] = A[M, K = KH*KW*Ci] * W[N, K]^T, bf16 operands, fp32 accumulation in TMEM.
//   * A is never materialised: the TMA engine walks the NHWC activation in *im2col mode* (one call per
//     (filter tap, <=64-channel slab): 128 output pixels x slab channels, halo / image borders zero-filled by
//     the hardware) straight into the K-major swizzled shared-memory layout tcgen05.mma consumes.
//   * W tiles come from the packed [Co_pad][Kpad] bf16 matrix with a tiled tensor map, same swizzle.
//   * Persistent CTAs (one per SM) loop over 128 x BN output tiles.  Warp roles: warp 0 = TMA producer,
//     warp 1 = MMA issuer (one elected lane), warps 2-9 = epilogue (two per TMEM lane quarter).  Rings: smem stages (full/empty
//     mbarriers) and TWO TMEM accumulators (tmem_full/tmem_empty), so the epilogue of tile i overlaps the
//     main loop of tile i+1.
//   * Epilogue: tcgen05.ld (lane = output pixel) -> scale/shift (folded BN or bias) -> (+residual) ->
//     ReLU/sigmoid -> NHWC bf16 (optionally a channel slice of a concat buffer) | NCHW fp32 | NHWC fp32.
#include "umma.cuh"
#include "tma_host.h"
#include <mutex>

namespace cnb {
namespace {

constexpr int BM = 128;
constexpr int NEPI = 8;                      // epilogue warps: two per TMEM lane quarter, alternating 16-column groups
constexpr int NTHREADS = (2 + NEPI) * 32;   // 320
constexpr int MAX_STAGES = 8;

struct TArgs {
  cnb_conv_desc d;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* res;
  void* y;
  int M, m_tiles, n_tiles, total_tiles;
  int BN;          // N tile (multiple of 16, <= 256)
  int slabW;       // channels per slab (64, 32, 16 or 8)
  int ntaps;       // KH*KW
  int nslabs;      // slabs along K (taps x channel chunks), padded to a multiple of the MMA K granularity
  int g;           // slabs per pipeline stage
  int nsteps;      // pipeline steps per tile
  int stages;
  u32 a_slab_bytes, b_slab_bytes, a_bytes, stage_bytes;
  u32 layout_type; // smem descriptor layout type for this slab width
  u32 sbo;         // bytes between 8-row groups
  u32 tmem_cols;
  u32 acc_stride;  // TMEM columns between the two accumulators
  u32 idesc;
  int nscale;      // n_tiles * BN
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) u64 s_full[MAX_STAGES];
  __shared__ __align__(8) u64 s_empty[MAX_STAGES];
  __shared__ __align__(8) u64 s_tfull[2];
  __shared__ __align__(8) u64 s_tempty[2];
  __shared__ u32 s_tmem;

  const cnb_conv_desc& d = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const u32 smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  // per-output-channel scale/shift for all N tiles live behind the stage ring
  float* s_scale = reinterpret_cast<float*>(smem_dyn + (smem_base - smem_u32(smem_dyn)) + (size_t)a.stages * a.stage_bytes);
  float* s_shift = s_scale + a.nscale;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_tfull[i], 1);
      mbar_init(&s_tempty[i], NEPI);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(&s_tmem, a.tmem_cols);
  for (int i = tid; i < a.nscale; i += NTHREADS) {
    s_scale[i] = (i < d.Co && a.scale) ? a.scale[i] : 1.f;
    s_shift[i] = (i < d.Co && a.shift) ? a.shift[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = s_tmem;
  const int HoWo = d.Ho * d.Wo;

  if (warp == 0) {
    // =============================== TMA producer =========================================================
    if (lane == 0) {
      u32 it = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
        const int m0 = m_tile * BM;
        const int n = m0 / HoWo;
        const int rem = m0 - n * HoWo;
        const int oy = rem / d.Wo, ox = rem - oy * d.Wo;
        const int w0 = ox * d.stride - d.pad, h0 = oy * d.stride - d.pad;
        const int n0 = n_tile * a.BN;
        for (int ks = 0; ks < a.nsteps; ++ks, ++it) {
          const u32 s = it % (u32)a.stages, ph = (it / (u32)a.stages) & 1u;
          mbar_wait_parked(&s_empty[s], ph ^ 1u);
          const int sl0 = ks * a.g;
          const int nsl = min(a.g, a.nslabs - sl0);
          mbar_expect_tx(&s_full[s], (u32)nsl * (a.a_slab_bytes + a.b_slab_bytes));
          const u32 sa = smem_base + s * a.stage_bytes;
          const u32 sb = sa + a.a_bytes;
          for (int j = 0; j < nsl; ++j) {
            const int sl = sl0 + j;
            const int k0 = sl * a.slabW;                 // K index of the slab in the packed weights
            int tap = k0 / d.Ci;
            const int c0 = k0 - tap * d.Ci;
            if (tap >= a.ntaps) tap = a.ntaps - 1;       // K padding (zero weights): any finite activations do
            const int kh = tap / d.KW, kw = tap - kh * d.KW;
            tma_load_im2col_4d(sa + (u32)j * a.a_slab_bytes, &tmA, c0, w0, h0, n, (unsigned short)(kw * d.dil),
                               (unsigned short)(kh * d.dil), &s_full[s]);
            tma_load_2d(sb + (u32)j * a.b_slab_bytes, &tmB, k0, n0, &s_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===========================================================
    if (lane == 0) {
      u32 it = 0, t = 0;
      const int mma_per_slab = a.slabW >= 16 ? a.slabW / 16 : 1;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++t) {
        const u32 acc = t & 1u, acc_ph = (t >> 1) & 1u;
        mbar_wait_parked(&s_tempty[acc], acc_ph ^ 1u);
        tc_fence_after();
        const u32 tmem_d = tmem_base + acc * a.acc_stride;
        u32 accumulate = 0;
        for (int ks = 0; ks < a.nsteps; ++ks, ++it) {
          const u32 s = it % (u32)a.stages, ph = (it / (u32)a.stages) & 1u;
          mbar_wait_parked(&s_full[s], ph);
          tc_fence_after();
          const u32 sa = smem_base + s * a.stage_bytes;
          const u32 sb = sa + a.a_bytes;
          const int nsl = min(a.g, a.nslabs - ks * a.g);
          if (a.slabW >= 16) {
            for (int j = 0; j < nsl; ++j) {
              const u64 da = make_sdesc(sa + (u32)j * a.a_slab_bytes, 16, a.sbo, a.layout_type);
              const u64 db = make_sdesc(sb + (u32)j * a.b_slab_bytes, 16, a.sbo, a.layout_type);
              for (int kk = 0; kk < mma_per_slab; ++kk) {   // +32 bytes of K inside the swizzle atom
                umma_bf16(tmem_d, da + (u64)(2 * kk), db + (u64)(2 * kk), a.idesc, accumulate);
                accumulate = 1;
              }
            }
          } else {
            // 8-channel slabs (16-byte rows, no swizzle): one K=16 instruction spans two slabs (LBO = slab size)
            for (int j = 0; j < nsl; j += 2) {
              const u64 da = make_sdesc(sa + (u32)j * a.a_slab_bytes, a.a_slab_bytes, 128, 0);
              const u64 db = make_sdesc(sb + (u32)j * a.b_slab_bytes, a.b_slab_bytes, 128, 0);
              umma_bf16(tmem_d, da, db, a.idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(&s_empty[s]);      // stage is free once these MMAs have read it
        }
        umma_commit(&s_tfull[acc]);      // accumulator complete
      }
    }
  } else {
    // =============================== epilogue (warps 2..5) ================================================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which of the quarter's two warps: even / odd column groups
    u32 t = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++t) {
      const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
      const u32 acc = t & 1u, acc_ph = (t >> 1) & 1u;
      const int m = m_tile * BM + 32 * q + lane;
      const int n0 = n_tile * a.BN;
      int on = 0, opix = 0;
      if (d.out_nchw_f32 == 1 && m < a.M) {
        on = m / HoWo;
        opix = m - on * HoWo;
      }
      mbar_wait_parked(&s_tfull[acc], acc_ph);
      tc_fence_after();
      const u32 taddr = tmem_base + acc * a.acc_stride + ((u32)(32 * q) << 16);
      const int ngroups = a.BN / 16;
      for (int g = half; g < ngroups; g += 2) {
        u32 v[16];
        tmem_ld16_nowait(taddr + (u32)(g * 16), v);
        tmem_ld_wait();
        const int co0 = n0 + g * 16;
        if (m < a.M && co0 < d.Co) epilogue_store(a, s_scale, s_shift, v, m, co0, co0, HoWo, on, opix);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ---- host side (tensor-map encode entry points: tma_host.h) ------------------------------------------------
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

TmaDriver& tma_driver() {
  static TmaDriver drv;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaDriverEntryPointQueryResult q1, q2;
    void *p1 = nullptr, *p2 = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p1, cudaEnableDefault, &q1) != cudaSuccess ||
        q1 != cudaDriverEntryPointSuccess)
      return;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p2, cudaEnableDefault, &q2) != cudaSuccess ||
        q2 != cudaDriverEntryPointSuccess)
      return;
    drv.tiled = (EncodeTiledFn)p1;
    drv.im2col = (EncodeIm2colFn)p2;
    cudaDriverGetVersion(&drv.driver_version);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&drv.num_sms, cudaDevAttrMultiProcessorCount, dev);
    drv.ok = drv.num_sms > 0;
  });
  return drv;
}

// Returns CNB_OK, or CNB_ERR_INVALID with the reason in cnb_last_error when this geometry is not covered
// (callers treat that as an error: there is no fallback path for plain convolutions).
int conv_tma_run(const cnb_conv_desc* d, const void* x, const void* wpk, const float* scale, const float* shift,
                 const void* res, void* y, cudaStream_t st) {
  TmaDriver& drv = tma_driver();
  if (!drv.ok) {
    set_error("conv: cuTensorMapEncode{Tiled,Im2col} entry points unavailable (driver too old?)");
    return CNB_ERR_CUDA;
  }
  const int Ci = d->Ci;
  CNB_CHECK_ARG(Ci == 8 || Ci == 16 || Ci == 32 || Ci % 64 == 0,
                "conv: Ci=%d must be 8, 16, 32 or a multiple of 64", Ci);
  const long long M = (long long)d->B * d->Ho * d->Wo;
  TArgs a;
  a.d = *d;
  a.scale = scale;
  a.shift = shift;
  a.res = (const __nv_bfloat16*)res;
  a.y = y;
  a.M = (int)M;
  a.m_tiles = (int)((M + BM - 1) / BM);
  const int Co_pad = round_up(d->Co, 16);
  a.n_tiles = (Co_pad + 255) / 256;
  a.BN = round_up((Co_pad + a.n_tiles - 1) / a.n_tiles, 16);
  a.total_tiles = a.m_tiles * a.n_tiles;
  a.slabW = Ci >= 64 ? 64 : Ci;
  a.ntaps = d->KH * d->KW;
  const int Ktot = a.ntaps * Ci;
  const int Kpad = round_up(Ktot, 64);           // layout of cnb_conv_pack_weights
  a.g = 64 / a.slabW;
  // K padding: 16-element MMA granularity (8-channel slabs pair up); the packed weights are zero there
  a.nslabs = round_up(Ktot, 16) / a.slabW;
  if (a.slabW == 8) a.nslabs = round_up(a.nslabs, 2);
  CNB_CHECK_ARG(a.nslabs * a.slabW <= Kpad, "conv: internal K padding error");
  a.nsteps = (a.nslabs + a.g - 1) / a.g;
  a.a_slab_bytes = (u32)(BM * a.slabW * 2);
  a.b_slab_bytes = (u32)(a.BN * a.slabW * 2);
  a.a_bytes = (u32)(BM * 64 * 2);
  a.stage_bytes = a.a_bytes + (u32)(a.BN * 64 * 2);
  CUtensorMapSwizzle swz;
  switch (a.slabW) {
    case 64: a.layout_type = 2; a.sbo = 1024; swz = CU_TENSOR_MAP_SWIZZLE_128B; break;
    case 32: a.layout_type = 4; a.sbo = 512; swz = CU_TENSOR_MAP_SWIZZLE_64B; break;
    case 16: a.layout_type = 6; a.sbo = 256; swz = CU_TENSOR_MAP_SWIZZLE_32B; break;
    default: a.layout_type = 0; a.sbo = 128; swz = CU_TENSOR_MAP_SWIZZLE_NONE; break;
  }
  a.nscale = a.n_tiles * a.BN;
  const size_t budget = 200 * 1024 - (size_t)a.nscale * 8;
  a.stages = (int)(budget / a.stage_bytes);
  if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
  CNB_CHECK_ARG(a.stages >= 2, "conv: tile does not fit in shared memory");
  a.acc_stride = (u32)round_up(a.BN, 32);
  a.tmem_cols = 32;
  while (a.tmem_cols < 2 * a.acc_stride) a.tmem_cols <<= 1;
  a.idesc = make_idesc_bf16(BM, a.BN);
  const size_t smem = (size_t)a.stages * a.stage_bytes + (size_t)a.nscale * 8 + 1024;

  // ---- tensor maps ------------------------------------------------------------------------------------------
  CUtensorMap tmA, tmB;
  {
    const __nv_bfloat16* base = (const __nv_bfloat16*)x + d->x_coffset;
    cuuint64_t dims[4] = {(cuuint64_t)Ci, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_cstride * 2, (cuuint64_t)d->Wi * d->x_cstride * 2,
                             (cuuint64_t)d->Hi * d->Wi * d->x_cstride * 2};
    int lower[2] = {-d->pad, -d->pad};
    int upper[2] = {d->pad - (d->KW - 1) * d->dil, d->pad - (d->KH - 1) * d->dil};
    CNB_CHECK_ARG(lower[0] >= -128 && upper[0] >= -128 && upper[0] <= 127 && (d->KW - 1) * d->dil <= 255 &&
                      (d->KH - 1) * d->dil <= 255,
                  "conv: padding / filter extent outside the TMA im2col range");
    cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    CUresult r = drv.im2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, lower, upper,
                            (cuuint32_t)a.slabW, (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv: cuTensorMapEncodeIm2col failed (%d) Ci=%d cstride=%d %dx%d k%d s%d", (int)r, Ci, d->x_cstride,
                d->Hi, d->Wi, d->KH, d->stride);
      return CNB_ERR_CUDA;
    }
    // driver bug workaround used by CUTLASS (copy_traits_sm90_im2col.hpp): small tensors, driver <= 13.1
    const unsigned long long bytes = (unsigned long long)d->B * d->Hi * d->Wi * d->x_cstride * 2;
    if (drv.driver_version <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&tmA)[1] &= ~(1ull << 21);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)Co_pad};
    cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)a.slabW, (cuuint32_t)a.BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = drv.tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wpk, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv: cuTensorMapEncodeTiled failed (%d) Kpad=%d Co_pad=%d BN=%d", (int)r, Kpad, Co_pad, a.BN);
      return CNB_ERR_CUDA;
    }
  }
  static bool configured = false;
  if (!configured) {
    CNB_CUDA(cudaFuncSetAttribute(conv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  const int grid = a.total_tiles < drv.num_sms ? a.total_tiles : drv.num_sms;
  conv_tma_kernel<<<grid, NTHREADS, smem, st>>>(tmA, tmB, a);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

}  // namespace cnb
