// Implicit-GEMM convolution and DCNv2 on the 5th-gen tensor cores (tcgen05.mma, accumulator in TMEM): the general
// cp.async-gather kernel, and the dispatch of cnb_conv2d_fprop / cnb_dcnv2_fprop to the specialised kernels
// (conv_rows.cu: wide thin layers; conv_tma.cu: Ci >= 32 TMA im2col; dcn_ws.cu: warp-specialised DCNv2).  This kernel
// keeps every remaining geometry correct (thin inputs at widths that are not a multiple of 128, Ci % 64 != 0 DCN,
// CNB_CONV_IMPL=v1 / CNB_DCN_IMPL=v1 for A/B runs).
//
// Replaces (reference file:line under CenterNet/models/):
//   nn.Conv2d + nn.BatchNorm2d(eval) [+ residual add] [+ nn.ReLU] chains of
//     backbones/pose_dla_dcn.py:28-68 (BasicBlock), :165-188 (Root: cat + 1x1), :351-370 (conv levels),
//     :281-285 (stem, after channel padding), heads.py:4-25 (HeadConv 3x3+ReLU, 1x1),
//   DCN.dcn_v2.DCN forward (external tteepe/DCNv2; call sites pose_dla_dcn.py:441-449,
//     resnet_dcn.py:202-210) incl. the conv_offset_mask conv (an ordinary conv through this kernel).
//
// GEMM view:  D[M = B*Ho*Wo pixels, N = Co] = A[M, K = KH*KW*Ci] * W[N, K]^T, bf16 operands, fp32 accum.
//   A tile (128 x 64 bf16) is never materialised in HBM: producer threads gather 16-byte channel
//   chunks of the NHWC activation straight into the canonical K-major SWIZZLE_128B shared-memory
//   layout with zero-filling cp.async (halo / K tail), or -- DCN mode -- bilinearly sample, modulate
//   and convert in registers and st.shared the chunk ("column tile written straight into SMEM").
//   One elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) per 32 bytes of K;
//   tcgen05.commit -> mbarrier recycles the stage.  Epilogue: tcgen05.ld 32x32b -> scale/shift (folded
//   BN or bias) -> (+residual) -> ReLU/sigmoid -> bf16 NHWC (optionally into a concat slice) or fp32.
#include "cnb_common.cuh"

namespace cnb {
int conv_tma_run(const cnb_conv_desc* d, const void* x, const void* wpk, const float* scale, const float* shift,
                 const void* res, void* y, cudaStream_t st);   // conv_tma.cu
bool conv_rows_supported(const cnb_conv_desc* d);                // conv_rows.cu
int conv_rows_run(const cnb_conv_desc* d, const void* x, const void* wpk, const float* scale, const float* shift,
                  const void* res, void* y, cudaStream_t st);
bool dcn_fp_supported(const cnb_conv_desc* d, int om_cstride);                  // dcn_fp.cu
int dcn_fp_run(const cnb_conv_desc* d, const void* x, const float* om, int om_cstride, const void* wpk,
               const float* scale, const float* shift, void* y, cudaStream_t st);
bool conv_fp_supported(const cnb_conv_desc* d);                                 // dcn_fp.cu (plain mode)
int conv_fp_run(const cnb_conv_desc* d, const void* x, const void* wpk, const float* scale, const float* shift,
                const void* res, void* y, cudaStream_t st);
bool dcn_ws_supported(const cnb_conv_desc* d, int om_cstride);                  // dcn_ws.cu
int dcn_ws_run(const cnb_conv_desc* d, const void* x, const float* om, int om_cstride, const void* wpk,
               const float* scale, const float* shift, void* y, cudaStream_t st);
namespace {

constexpr int CT = 256;   // threads per CTA (8 warps)
constexpr int BM = 128;   // pixels per tile  (UMMA M)
constexpr int BK = 64;    // K elements per stage (one 128-byte swizzle atom)

// ---- tcgen05 / TMEM PTX wrappers -----------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(u32* smem_dst, u32 ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(u32 taddr, u32 ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(void* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(u32 taddr, u32 (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row groups 1024 B apart (SBO = 64),
// LBO unused (1), descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ u64 make_sdesc(u32 smem_addr) {
  return (u64)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void cp_async16(u32 dst, const void* src, u32 src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct ConvArgs {
  cnb_conv_desc d;
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;    // packed [Co_pad][Kpad]
  const float* scale;
  const float* shift;
  const __nv_bfloat16* res;
  void* y;
  const float* om;           // DCN: [B,H,W,om_cstride] fp32 (27 used)
  int om_cstride;
  int M;                     // B*Ho*Wo
  int Ktot, Kpad, nkb;
  int KWp;                   // filter width of the packed weights (>= KW; extra columns are zero)
  int BN;                    // N tile (multiple of 16, <= 256)
  int Co_pad;
  u32 tmem_cols;
  u32 idesc;
};

__device__ __forceinline__ u32 pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<u32*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(u32 v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}

// =================================================================================================
template <int STAGES, bool DCN>
__global__ void __launch_bounds__(CT) conv_umma_kernel(const ConvArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) u64 s_mma_done[STAGES];
  __shared__ __align__(8) u64 s_acc_full;
  __shared__ u32 s_tmem;
  __shared__ float s_scale[256], s_shift[256];

  const cnb_conv_desc& d = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * a.BN;
  const int BN = a.BN;

  const u32 smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const u32 a_bytes = BM * BK * 2;           // 16 KB
  const u32 b_bytes = (u32)BN * BK * 2;      // BN * 128 B (BN % 8 == 0 -> multiple of 1024)
  const u32 stage_bytes = a_bytes + b_bytes;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&s_mma_done[s], 1);
    mbar_init(&s_acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem, a.tmem_cols);
  for (int i = tid; i < BN; i += CT) {
    const int co = n0 + i;
    s_scale[i] = (co < d.Co && a.scale) ? a.scale[co] : 1.f;
    s_shift[i] = (co < d.Co && a.shift) ? a.shift[co] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_d = s_tmem;

  // ---- per-thread gather geometry: 4 fixed rows (tid/8 + 32*i), fixed 16-byte chunk column tid%8 ----
  const int cchunk = tid & 7;
  const int HoWo = d.Ho * d.Wo;
  int r_n[4], r_iy0[4], r_ix0[4];
  bool r_ok[4];
  const int padx = d.pad_w1 > 0 ? d.pad_w1 - 1 : d.pad;   // rectangular filters: separate horizontal padding
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + (tid >> 3) + 32 * i;
    r_ok[i] = m < a.M;
    const int mm = r_ok[i] ? m : 0;
    const int n = mm / HoWo;
    const int rem = mm - n * HoWo;
    const int oy = rem / d.Wo, ox = rem - oy * d.Wo;
    r_n[i] = n;
    r_iy0[i] = oy * d.stride - d.pad;
    r_ix0[i] = ox * d.stride - padx;
  }

  auto load_stage = [&](int kb) {
    const int s = kb % STAGES;
    const u32 sa = smem_base + (u32)s * stage_bytes;
    const u32 sb = sa + a_bytes;
    // ---- B tile: BN rows x 8 chunks of the packed weights ------------------------------------
    for (int q = tid; q < BN * 8; q += CT) {
      const int row = q >> 3, c = q & 7;
      const int co = n0 + row;
      const __nv_bfloat16* src = a.w + (size_t)(co < a.Co_pad ? co : 0) * a.Kpad + kb * BK + c * 8;
      cp_async16(sb + row * 128 + ((c ^ (row & 7)) << 4), src, co < a.Co_pad ? 16u : 0u);
    }
    // ---- A tile ---------------------------------------------------------------------------------
    const int kg = kb * BK + cchunk * 8;   // first K index of this thread's chunk
    const int tap = kg / d.Ci;
    const int ci = kg - tap * d.Ci;
    const int kh = tap / a.KWp, kw = tap - kh * a.KWp;
    const bool kvalid = kg < a.Ktot;
    if (!DCN) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = (tid >> 3) + 32 * i;
        const int iy = r_iy0[i] + kh * d.dil, ix = r_ix0[i] + kw * d.dil;
        const bool ok = kvalid && r_ok[i] && iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi;
        const __nv_bfloat16* src =
            ok ? a.x + ((size_t)(r_n[i] * d.Hi + iy) * d.Wi + ix) * d.x_cstride + d.x_coffset + ci : a.x;
        cp_async16(sa + row * 128 + ((cchunk ^ (row & 7)) << 4), src, ok ? 16u : 0u);
      }
    } else {
      // modulated bilinear sampling (DCNv2 / torchvision deform_conv2d semantics), 8 channels/thread
      uint4 q[4][4];
      float wgt[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int oy = r_iy0[i] + d.pad, ox = r_ix0[i] + padx;   // stride 1
        const float* omp = a.om + ((size_t)(r_n[i] * d.Hi + oy) * d.Wi + ox) * a.om_cstride;
        float dy = 0.f, dx = 0.f, mk = 0.f;
        if (r_ok[i] && kvalid) {
          dy = __ldg(omp + 2 * tap);
          dx = __ldg(omp + 2 * tap + 1);
          mk = 1.f / (1.f + __expf(-__ldg(omp + 18 + tap)));
        }
        const float py = (float)(r_iy0[i] + kh * d.dil) + dy;
        const float px = (float)(r_ix0[i] + kw * d.dil) + dx;
        const bool inside = r_ok[i] && kvalid && py > -1.f && px > -1.f && py < (float)d.Hi && px < (float)d.Wi;
        const int y0 = (int)floorf(py), x0 = (int)floorf(px);
        const float ly = py - (float)y0, lx = px - (float)x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const bool v00 = inside && y0 >= 0 && x0 >= 0;
        const bool v01 = inside && y0 >= 0 && x0 + 1 <= d.Wi - 1;
        const bool v10 = inside && y0 + 1 <= d.Hi - 1 && x0 >= 0;
        const bool v11 = inside && y0 + 1 <= d.Hi - 1 && x0 + 1 <= d.Wi - 1;
        wgt[i][0] = v00 ? hy * hx * mk : 0.f;
        wgt[i][1] = v01 ? hy * lx * mk : 0.f;
        wgt[i][2] = v10 ? ly * hx * mk : 0.f;
        wgt[i][3] = v11 ? ly * lx * mk : 0.f;
        const __nv_bfloat16* base = a.x + (size_t)r_n[i] * d.Hi * d.Wi * d.x_cstride + d.x_coffset + ci;
        const uint4 z = make_uint4(0, 0, 0, 0);
        q[i][0] = v00 ? __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y0 * d.Wi + x0) * d.x_cstride)) : z;
        q[i][1] = v01 ? __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y0 * d.Wi + x0 + 1) * d.x_cstride)) : z;
        q[i][2] = v10 ? __ldg(reinterpret_cast<const uint4*>(base + ((size_t)(y0 + 1) * d.Wi + x0) * d.x_cstride)) : z;
        q[i][3] = v11 ? __ldg(reinterpret_cast<const uint4*>(base + ((size_t)(y0 + 1) * d.Wi + x0 + 1) * d.x_cstride)) : z;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = (tid >> 3) + 32 * i;
        u32 o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const u32 c0 = (&q[i][0].x)[e], c1 = (&q[i][1].x)[e], c2 = (&q[i][2].x)[e], c3 = (&q[i][3].x)[e];
          const float2 f0 = unpack_bf16x2(c0), f1 = unpack_bf16x2(c1), f2 = unpack_bf16x2(c2),
                       f3 = unpack_bf16x2(c3);
          const float lo = wgt[i][0] * f0.x + wgt[i][1] * f1.x + wgt[i][2] * f2.x + wgt[i][3] * f3.x;
          const float hi = wgt[i][0] * f0.y + wgt[i][1] * f1.y + wgt[i][2] * f2.y + wgt[i][3] * f3.y;
          o[e] = pack_bf16x2(lo, hi);
        }
        const u32 dst = sa + row * 128 + ((cchunk ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                     "r"(o[3])
                     : "memory");
      }
    }
  };

  // ---- software pipeline ---------------------------------------------------------------------------
  const int nkb = a.nkb;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkb) load_stage(s);
    cp_async_commit();
  }
  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<STAGES - 2>();   // this thread's copies for stage kb have landed
    fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const u32 sa = smem_base + (u32)(kb % STAGES) * stage_bytes;
      const u64 da = make_sdesc(sa), db = make_sdesc(sa + a_bytes);
#pragma unroll
      for (int j = 0; j < BK / 16; ++j)   // advance 32 bytes (encoded +2) along K inside the swizzle atom
        umma_bf16(tmem_d, da + (u64)(2 * j), db + (u64)(2 * j), a.idesc, (kb | j) ? 1u : 0u);
      umma_commit(&s_mma_done[kb % STAGES]);
      if (kb == nkb - 1) umma_commit(&s_acc_full);
    }
    // refill the stage consumed by MMA(kb-1) once that MMA has retired
    const int nxt = kb + STAGES - 1;
    if (nxt < nkb) {
      if (kb >= 1) mbar_wait(&s_mma_done[(kb - 1) % STAGES], (u32)(((kb - 1) / STAGES) & 1));
      load_stage(nxt);
    }
    cp_async_commit();
  }

  // ---- epilogue ------------------------------------------------------------------------------------
  mbar_wait(&s_acc_full, 0);
  tc_fence_after();
  const int row = 32 * (warp & 3) + lane;
  const int m = m0 + row;
  const int ngroups = BN / 16;
  const int g_begin = (warp < 4) ? 0 : (ngroups + 1) / 2;
  const int g_end = (warp < 4) ? (ngroups + 1) / 2 : ngroups;
  int on = 0, opix = 0;
  if (d.out_nchw_f32 == 1 && m < a.M) {
    on = m / HoWo;
    opix = m - on * HoWo;
  }
  for (int g = g_begin; g < g_end; ++g) {
    u32 v[16];
    tmem_ld16(tmem_d + ((u32)(32 * (warp & 3)) << 16) + (u32)(g * 16), v);
    if (m >= a.M) continue;
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * s_scale[g * 16 + j] + s_shift[g * 16 + j];
    const int co0 = n0 + g * 16;
    if (a.res) {
      const uint4* rp = reinterpret_cast<const uint4*>(a.res + (size_t)m * d.res_cstride + d.res_coffset + co0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (co0 + 8 * h < d.Co) {
          const uint4 r = __ldg(rp + h);
          const float2 p0 = unpack_bf16x2(r.x), p1 = unpack_bf16x2(r.y), p2 = unpack_bf16x2(r.z),
                       p3 = unpack_bf16x2(r.w);
          f[8 * h + 0] += p0.x; f[8 * h + 1] += p0.y; f[8 * h + 2] += p1.x; f[8 * h + 3] += p1.y;
          f[8 * h + 4] += p2.x; f[8 * h + 5] += p2.y; f[8 * h + 6] += p3.x; f[8 * h + 7] += p3.y;
        }
      }
    }
    if (d.act == 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
    } else if (d.act == 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = 1.f / (1.f + __expf(-f[j]));
    }
    if (d.out_nchw_f32 == 0) {          // NHWC bf16 (optionally a channel slice of a concat buffer)
      __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y) + (size_t)m * d.y_cstride + d.y_coffset + co0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (co0 + 8 * h < d.Co) {
          uint4 o;
          o.x = pack_bf16x2(f[8 * h + 0], f[8 * h + 1]);
          o.y = pack_bf16x2(f[8 * h + 2], f[8 * h + 3]);
          o.z = pack_bf16x2(f[8 * h + 4], f[8 * h + 5]);
          o.w = pack_bf16x2(f[8 * h + 6], f[8 * h + 7]);
          *reinterpret_cast<uint4*>(yp + 8 * h) = o;
        }
      }
    } else if (d.out_nchw_f32 == 1) {   // NCHW fp32 (head maps for decode / losses)
      float* yp = reinterpret_cast<float*>(a.y) + ((size_t)on * d.Co + co0) * HoWo + opix;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (co0 + j < d.Co) yp[(size_t)j * HoWo] = f[j];
    } else {                            // NHWC fp32 (offset/mask maps feeding the DCN sampler)
      float* yp = reinterpret_cast<float*>(a.y) + (size_t)m * d.y_cstride + d.y_coffset + co0;
#pragma unroll
      for (int h = 0; h < 4; ++h)
        if (co0 + 4 * h < d.y_cstride)
          *reinterpret_cast<float4*>(yp + 4 * h) = make_float4(f[4 * h], f[4 * h + 1], f[4 * h + 2], f[4 * h + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, a.tmem_cols);
}

// ---- weight packing: [Co,Ci,KH,KW] fp32 -> [Co_pad][Kpad] bf16, K order (kh,kw,ci), zero padded ----
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Co,
                                    int Ci, int Ci_pad, int KH, int KW, int Co_pad, int Kpad) {
  const long long total = (long long)Co_pad * Kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / Kpad), k = (int)(i % Kpad);
    float v = 0.f;
    if (co < Co && k < KH * KW * Ci_pad) {
      const int tap = k / Ci_pad, ci = k % Ci_pad;
      if (ci < Ci) v = w[((size_t)co * Ci + ci) * KH * KW + tap];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct Plan {
  int Co_pad, Ktot, Kpad, nkb, BN, ntiles_n, stages;
  u32 tmem_cols;
  size_t smem;
};

static Plan make_plan(const cnb_conv_desc& d) {
  Plan p;
  p.Co_pad = round_up(d.Co, 16);
  p.Ktot = d.KH * (d.w_kw > 0 ? d.w_kw : d.KW) * d.Ci;
  p.Kpad = round_up(p.Ktot, BK);
  p.nkb = p.Kpad / BK;
  p.BN = p.Co_pad <= 128 ? p.Co_pad : 128;
  p.ntiles_n = (p.Co_pad + p.BN - 1) / p.BN;
  p.tmem_cols = 32;
  while ((int)p.tmem_cols < p.BN) p.tmem_cols <<= 1;
  p.stages = 4;
  p.smem = (size_t)p.stages * (BM * BK * 2 + (size_t)p.BN * BK * 2) + 1024;
  return p;
}

template <int STAGES, bool DCN>
static cudaError_t launch_conv(const ConvArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<STAGES, DCN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    once.mark();
  }
  conv_umma_kernel<STAGES, DCN><<<grid, CT, smem, st>>>(a);
  return cudaGetLastError();
}

static int run_conv(const cnb_conv_desc* d, const void* x, const float* om, int om_cstride, const void* wpk,
                    const float* scale, const float* shift, const void* res, void* y, bool dcn,
                    cudaStream_t st) {
  CNB_CHECK_ARG(d && x && wpk && y, "conv: null pointer");
  CNB_CHECK_ARG(d->B >= 1 && d->Hi >= 1 && d->Wi >= 1 && d->Ho >= 1 && d->Wo >= 1, "conv: bad geometry");
  CNB_CHECK_ARG(d->Ci >= 8 && d->Ci % 8 == 0, "conv: Ci=%d must be a multiple of 8 (pad the input)", d->Ci);
  CNB_CHECK_ARG(d->x_cstride % 8 == 0 && d->x_coffset % 8 == 0 && d->x_cstride >= d->Ci,
                "conv: x channel stride/offset must be multiples of 8");
  CNB_CHECK_ARG(d->Co >= 1 && d->KH >= 1 && d->KW >= 1 && d->stride >= 1 && d->dil >= 1, "conv: bad filter");
  CNB_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wpk & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
                    ((uintptr_t)res & 15) == 0,
                "conv: pointers must be 16-byte aligned");
  if (d->out_nchw_f32 == 0) {
    CNB_CHECK_ARG(d->y_cstride % 8 == 0 && d->y_coffset % 8 == 0 && d->Co % 8 == 0,
                  "conv: NHWC bf16 output needs Co, y_cstride, y_coffset multiples of 8");
  } else if (d->out_nchw_f32 == 2) {
    CNB_CHECK_ARG(d->y_cstride % 4 == 0 && d->y_coffset % 16 == 0 && d->y_cstride >= ((d->Co + 15) / 16) * 16,
                  "conv: NHWC fp32 output needs y_cstride >= Co rounded up to 16");
  }
  if (res) CNB_CHECK_ARG(d->res_cstride % 8 == 0 && d->res_coffset % 8 == 0 && d->out_nchw_f32 == 0,
                         "conv: residual needs NHWC bf16 output and 8-aligned channel stride/offset");
  const long long M = (long long)d->B * d->Ho * d->Wo;
  CNB_CHECK_ARG(M < (1ll << 31) - BM, "conv: too many output pixels");
  if (dcn) {
    CNB_CHECK_ARG(om && om_cstride >= 27, "dcnv2: offset/mask map required");
    CNB_CHECK_ARG(d->KH == 3 && d->KW == 3 && d->stride == 1 && d->pad == 1 && d->dil == 1 &&
                      d->Ho == d->Hi && d->Wo == d->Wi,
                  "dcnv2: only 3x3 / stride 1 / pad 1 / dil 1 (the reference's configuration)");
    CNB_CHECK_ARG(d->Ci % 8 == 0, "dcnv2: Ci must be a multiple of 8");
    CNB_CHECK_ARG(d->out_nchw_f32 == 0 && ((uintptr_t)om & 15) == 0, "dcnv2: NHWC bf16 output, 16-byte aligned om");
    // sampling footprint staged in shared memory (dcn_fp.cu); CNB_DCN_IMPL=ws keeps the global-gather kernel
    if (dcn_fp_supported(d, om_cstride)) return dcn_fp_run(d, x, om, om_cstride, wpk, scale, shift, y, st);
    // warp-specialised sampler kernel (dcn_ws.cu); CNB_DCN_IMPL=v1 keeps the gather kernel below for A/B runs
    if (dcn_ws_supported(d, om_cstride)) return dcn_ws_run(d, x, om, om_cstride, wpk, scale, shift, y, st);
  }
  if (!dcn) {
    // plain convolutions: TMA-im2col warp-specialised kernel (conv_tma.cu); CNB_CONV_IMPL=v1 keeps the
    // cp.async gather kernel below for A/B comparisons
    static const bool use_v1 = [] { const char* e = getenv("CNB_CONV_IMPL"); return e && e[0] == 'v' && e[1] == '1'; }();
    CNB_CHECK_ARG(d->w_kw == 0 || d->w_kw >= d->KW, "conv: w_kw=%d smaller than KW=%d", d->w_kw, d->KW);
    // 3x3 / stride 1 with Ci % 64 == 0 and N <= 128: input box in shared memory, A operand through tensor memory (dcn_fp.cu)
    if (!use_v1 && conv_fp_supported(d)) return conv_fp_run(d, x, wpk, scale, shift, res, y, st);
    // wide thin layers (W_out % 128 == 0, Ci <= 64, KxK): row-window kernel, every input pixel fetched once
    if (!use_v1 && conv_rows_supported(d)) return conv_rows_run(d, x, wpk, scale, shift, res, y, st);
    if (!use_v1 && (d->Ci >= 32) && (d->w_kw == 0 || d->w_kw == d->KW) && d->KH == d->KW && d->pad_w1 == 0)
      return conv_tma_run(d, x, wpk, scale, shift, res, y, st);
  }
  const Plan p = make_plan(*d);
  ConvArgs a;
  a.d = *d;
  a.x = (const __nv_bfloat16*)x;
  a.w = (const __nv_bfloat16*)wpk;
  a.scale = scale;
  a.shift = shift;
  a.res = (const __nv_bfloat16*)res;
  a.y = y;
  a.om = om;
  a.om_cstride = om_cstride;
  a.M = (int)M;
  a.Ktot = p.Ktot;
  a.KWp = d->w_kw > 0 ? d->w_kw : d->KW;
  a.Kpad = p.Kpad;
  a.nkb = p.nkb;
  a.BN = p.BN;
  a.Co_pad = p.Co_pad;
  a.tmem_cols = p.tmem_cols;
  // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both, N>>3 @17, M>>4 @24
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(p.BN >> 3) << 17) | ((u32)(BM >> 4) << 24);
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)p.ntiles_n);
  cudaError_t e = dcn ? launch_conv<4, true>(a, grid, p.smem, st) : launch_conv<4, false>(a, grid, p.smem, st);
  if (e != cudaSuccess) {
    set_error("conv launch failed: %s", cudaGetErrorString(e));
    return CNB_ERR_CUDA;
  }
  count_launch();
  return CNB_OK;
}

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" size_t cnb_conv_packed_weight_bytes(int Co, int Ci, int KH, int KW) {
  if (Co < 1 || Ci < 1 || KH < 1 || KW < 1) return 0;
  const int Ci_pad = round_up(Ci, 8);
  return (size_t)round_up(Co, 16) * round_up(KH * KW * Ci_pad, BK) * sizeof(__nv_bfloat16);
}

extern "C" int cnb_conv_pack_weights(const float* w, void* wpk, int Co, int Ci, int KH, int KW,
                                     cnb_stream_t stream) {
  CNB_CHECK_ARG(w && wpk && Co >= 1 && Ci >= 1 && KH >= 1 && KW >= 1, "conv_pack_weights: bad argument");
  const int Ci_pad = round_up(Ci, 8);
  const int Co_pad = round_up(Co, 16), Kpad = round_up(KH * KW * Ci_pad, BK);
  const long long total = (long long)Co_pad * Kpad;
  const int grid = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  pack_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wpk, Co, Ci, Ci_pad, KH, KW,
                                                              Co_pad, Kpad);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_conv2d_fprop(const cnb_conv_desc* d, const void* x, const void* wpk, const float* scale,
                                const float* shift, const void* res, void* y, cnb_stream_t stream) {
  return run_conv(d, x, nullptr, 0, wpk, scale, shift, res, y, false, (cudaStream_t)stream);
}

extern "C" int cnb_dcnv2_fprop(const cnb_conv_desc* d, const void* x, const float* om, int om_cstride,
                               const void* wpk, const float* scale, const float* shift, void* y,
                               cnb_stream_t stream) {
  return run_conv(d, x, om, om_cstride, wpk, scale, shift, nullptr, y, true, (cudaStream_t)stream);
}
