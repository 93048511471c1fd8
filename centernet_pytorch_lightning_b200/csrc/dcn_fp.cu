// Modulated deformable convolution v2 (DCNv2) with the sampling FOOTPRINT staged in shared memory.
//
// Replaces DCN.dcn_v2.DCN forward (external tteepe/DCNv2; call sites CenterNet/models/backbones/
// pose_dla_dcn.py:441-449 and resnet_dcn.py:202-210): 3x3, stride 1, pad 1, dil 1, deformable_groups 1,
//   y[n,co,p] = b[co] + sum_{ci,k} W[co,ci,k] * sigmoid(m_k(p)) * bilinear(x[n,ci], p + tap_k + offset_k(p))
// with the (+ BatchNorm(eval) + ReLU) of DeformConv (pose_dla_dcn.py:435-454) folded into the epilogue.
//
// Why a second kernel next to dcn_ws.cu: there the 16 sampler warps gather the 4 x 9 corners of every pixel straight
// from global memory through L1 -- 64 KB of 32-byte gathers per 128-pixel K block.  ncu (profiles/r02i_prof_dcn64.txt):
// the L1 load/store data pipe is the busiest unit (64 %: ~810 wavefronts per K block for the gathers, 1.6 per 128 bytes,
// because a multi-line load replays at ~2 clocks per line), the hit rate is 72 % (the ~80 KB footprint of a tile row does
// not fit next to 180 KB of shared memory) and every warp issue is followed by 5.5 stall cycles on those loads.  Here
// a tile is a 2-D block of pixels (8 rows x 16 columns, or 16 x 8 for narrow maps) and the box of input pixels its
// samples can reach with offsets up to +-R -- (TH + 2R + 3) x (TW + 2R + 3) pixels of one 64-channel slab, zero-filled
// outside the image by the TMA engine -- is loaded ONCE per (tile, slab) with one tensor-map box copy, double buffered.
// The samplers then read corners from shared memory: one conflict-free wavefront per 128 bytes, fixed 30-clock latency,
// no tag lookups, no misses; L2 traffic falls from ~9x to ~3x the input.  A (pixel, tap) whose corners leave the box
// (offsets beyond R) is flagged in the table and sampled from global memory through the same generic-address loads,
// so the result does not depend on R.
//
// Shared-memory bandwidth is then what bounds the sampler (ncu, first version of this kernel: ~1180 bank wavefronts per K
// block -- 560 corner/table loads, 160 operand stores, 190 operand reads by the tensor core, 120 TMA writes -- at 0.7 per
// clock), so the sampled operand does not go through shared memory at all: a sampler thread owns one pixel ROW of the
// tile, blends its 64 channels chunk by chunk and writes them with tcgen05.st into a ring of A stages in TENSOR MEMORY
// (32 columns per K block, up to 12 stages next to the two accumulators); tcgen05.mma reads A from there.  One thread
// per row also means one table entry and one set of corner addresses per 128 output bytes instead of per 32.  The box
// is stored with the 128-byte swizzle (16-byte chunk index XOR pixel index mod 8), so the 8 lanes of a shared-memory
// wavefront -- 8 neighbouring pixels reading the same logical chunk of 8 different box pixels -- hit 8 different bank
// groups whenever the sampled pixels are distinct mod 8 (always for smooth offsets).
//
// GEMM view: D[128 pixels, Co] = A[128, 9*Ci] * W^T, one K block = one tap of one 64-channel slab.  Roles: 16 sampler
// warps in 4 groups (one warp per TMEM lane quarter) that take the K blocks of the CTA's stream round-robin; 4 setup warps
// (offsets/masks -> table of bilinear weights + box or global pixel index); one MMA warp (tcgen05.mma with the A operand
// in tensor memory, weight tile by TMA into a shared-memory ring, elected lane); 4 epilogue warps; one loader warp (boxes).
#include "umma.cuh"
#include "tma_host.h"
#include <stdlib.h>

namespace cnb {
namespace {

constexpr int BM = 128;
constexpr int NPROD_WARPS = 16;

constexpr int NSETUP_WARPS = 4;
constexpr int NSETUP = NSETUP_WARPS * 32;
constexpr int W_SETUP0 = NPROD_WARPS;                // warps 16..19
constexpr int W_MMA = W_SETUP0 + NSETUP_WARPS;       // warp 20
constexpr int W_EPI0 = W_MMA + 1;                    // warps 21..24 (TMEM lane quarters 1,2,3,0)
constexpr int W_LOAD = W_EPI0 + 4;                   // warp 25: footprint boxes
constexpr int NTHREADS = (W_LOAD + 1) * 32;          // 832
constexpr int NG = 4;                                // sampler groups: K block k of the CTA's stream belongs to group k % NG
                                                     // (5 groups = 960 threads at 64 registers measured slower: 64->64@128x128 149 -> 162 us)
constexpr int WPG = NPROD_WARPS / NG;                // 4 warps per group = the 4 TMEM lane quarters
constexpr int MAX_STAGES = 12;
constexpr int MAX_B = 18;
constexpr int MAX_FP = 6;          // boxes in flight (plain mode)
constexpr int A_COLS = 32;                           // TMEM columns of one K block of A: 64 bf16 per row
constexpr int NTAB = BM * 9;                         // (pixel, tap) entries per tile
constexpr int OM_CS = 32;                            // channel stride of the offset/mask map this kernel takes

struct FArgs {
  cnb_conv_desc d;
  const __nv_bfloat16* x;
  const float* om;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* res;   // unused (kept for the shared epilogue)
  void* y;
  int m_tiles;     // B * tiles_y * tiles_x
  int tiles_x, tiles_y;
  int tw_shift;    // tile = (128 >> tw_shift) rows x (1 << tw_shift) columns
  int R;           // offsets up to +-R stay inside the staged box
  int FW, FH;      // box: (TW + 2R + 3, rounded up to a multiple of 8) x (TH + 2R + 3) pixels
  u32 fp_bytes;    // FW * FH * 128
  u32 fp_stride;   // fp_bytes rounded up to 1 KB
  int nfp;         // boxes in the ring: 2 (DCN mode: shared memory is full), up to MAX_FP in plain mode
  int BN;          // = Co rounded up to 16, <= 256
  int nacc;        // accumulators in tensor memory: 2 (epilogue of tile i overlaps tile i+1), 1 for BN > 128
  int nkb;         // 9 * Ci/64
  u32 bstage;      // bytes of one weight-tile slot (b_bytes rounded up to 1 KB)
  int stages;      // A ring depth: stage i = tensor-memory columns a_col0 + 32 i
  int nb;          // weight-tile slots in shared memory; K block j of the CTA's stream uses slot j % nb
  int b_resident;  // nb >= nkb: the 9 * Ci/64 weight tiles are loaded once and stay
  u32 b_bytes, tmem_cols, acc_stride, part_stride, a_col0, idesc;
  int kps;         // K blocks per A stage: 1, or 3 in plain mode with N <= 64 (a group samples a whole filter row per handshake)
  int plain;       // 1: plain 3x3 / stride 1 / pad 1 convolution through the same machinery (no offsets, no table, R = 0)
  int debug;       // timing experiments (-DCNB_DCN_EXPERIMENTS): 1 no corner loads, 2 no blend, 4 no tcgen05.st, 8 no table read
};

// packed fp32x2 helpers (Blackwell FFMA2): a 64-bit register holds (low, high) floats
__device__ __forceinline__ u64 dup2(float w) {
  u64 r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(w));
  return r;
}
__device__ __forceinline__ u64 pair_from_bf16x2(u32 v) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(v << 16), "r"(v & 0xffff0000u));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ void tma_load_tile_4d(u32 dst_smem, const void* tmap, int c, int w, int h, int n, void* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(dst_smem),
      "l"(reinterpret_cast<u64>(tmap)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}
__device__ __forceinline__ uint4 lds128(u32 addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 ldgen128(u64 addr) {   // generic address: shared window or global memory
  uint4 v;
  asm volatile("ld.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr));
  return v;
}

__device__ __forceinline__ void tmem_st4(u32 taddr, const uint4 v) {   // 32 lanes x 4 columns: lane = tile row
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(u32 taddr, const uint4 a, const uint4 b, const uint4 c, const uint4 d) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w),
      "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
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

// 4-corner blend of one 16-byte chunk (8 channels): fp32 (packed FFMA2), one rounding to bf16
template <int BLEND_BF16>   // 0: fp32 blend; 1: bf16x2 blend; 2: fp32 accumulation of bf16 x bf16(weight) products
__device__ __forceinline__ uint4 blend4(const float4 w, const uint4 c0, const uint4 c1, const uint4 c2, const uint4 c3) {
  const u32 q0[4] = {c0.x, c0.y, c0.z, c0.w}, q1[4] = {c1.x, c1.y, c1.z, c1.w}, q2[4] = {c2.x, c2.y, c2.z, c2.w},
            q3[4] = {c3.x, c3.y, c3.z, c3.w};
  u32 o[4];
  if constexpr (BLEND_BF16 == 2) {
    // Mixed-precision FMA (SASS FHFMA.BF16: bf16 x bf16 + f32 -> f32, operand halves selected for free): no unpack
    // instructions at all -- 4 per element instead of 6.  The products are exact and the sums fp32; the only difference
    // to the fp32 blend is that the four bilinear weights are rounded to bf16 (relative 2^-9, i.e. a sampling-position
    // error below 0.002 pixel).
    const unsigned short w0 = __bfloat16_as_ushort(__float2bfloat16_rn(w.x)), w1 = __bfloat16_as_ushort(__float2bfloat16_rn(w.y)),
                         w2 = __bfloat16_as_ushort(__float2bfloat16_rn(w.z)), w3 = __bfloat16_as_ushort(__float2bfloat16_rn(w.w));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float lo, hi;
      asm("{\n\t.reg .b16 a0, a1, b0, b1, c0, c1, d0, d1;\n\t"
          "mov.b32 {a0, a1}, %2;\n\tmov.b32 {b0, b1}, %3;\n\tmov.b32 {c0, c1}, %4;\n\tmov.b32 {d0, d1}, %5;\n\t"
          "fma.rn.f32.bf16 %0, a0, %6, 0f00000000;\n\tfma.rn.f32.bf16 %1, a1, %6, 0f00000000;\n\t"
          "fma.rn.f32.bf16 %0, b0, %7, %0;\n\tfma.rn.f32.bf16 %1, b1, %7, %1;\n\t"
          "fma.rn.f32.bf16 %0, c0, %8, %0;\n\tfma.rn.f32.bf16 %1, c1, %8, %1;\n\t"
          "fma.rn.f32.bf16 %0, d0, %9, %0;\n\tfma.rn.f32.bf16 %1, d1, %9, %1;\n\t}"
          : "=&f"(lo), "=&f"(hi)
          : "r"(q0[e]), "r"(q1[e]), "r"(q2[e]), "r"(q3[e]), "h"(w0), "h"(w1), "h"(w2), "h"(w3));
      o[e] = pack_bf16x2(lo, hi);
    }
  } else if constexpr (BLEND_BF16 == 1) {
    const u32 wh0 = pack_bf16x2(w.x, w.x), wh1 = pack_bf16x2(w.y, w.y), wh2 = pack_bf16x2(w.z, w.z),
              wh3 = pack_bf16x2(w.w, w.w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      u32 acc;
      asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(acc) : "r"(wh0), "r"(q0[e]));
      asm("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc) : "r"(wh1), "r"(q1[e]));
      asm("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc) : "r"(wh2), "r"(q2[e]));
      asm("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc) : "r"(wh3), "r"(q3[e]));
      o[e] = acc;
    }
  } else {
    const u64 w0 = dup2(w.x), w1 = dup2(w.y), w2 = dup2(w.z), w3 = dup2(w.w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // bf16 -> fp32 is a 16-bit shift: low element = v << 16, high element = v & 0xffff0000
      u64 acc = mul2(w0, pair_from_bf16x2(q0[e]));
      acc = fma2(w1, pair_from_bf16x2(q1[e]), acc);
      acc = fma2(w2, pair_from_bf16x2(q2[e]), acc);
      acc = fma2(w3, pair_from_bf16x2(q3[e]), acc);
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
      o[e] = pack_bf16x2(lo, hi);
    }
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

struct TileXY {
  int n, y0, x0;   // image, first row, first column of the tile
};
__device__ __forceinline__ TileXY tile_xy(const FArgs& a, int tile) {
  const int per = a.tiles_x * a.tiles_y;
  TileXY t;
  t.n = tile / per;
  const int rem = tile - t.n * per;
  const int ty = rem / a.tiles_x;
  t.y0 = ty * (BM >> a.tw_shift);
  t.x0 = (rem - ty * a.tiles_x) << a.tw_shift;
  return t;
}

#ifdef CNB_DCN_EXPERIMENTS
#define FP_DBG(a) ((a).debug)
#define PLAIN_WAIT(bar, par) do { if (FP_DBG(a) & 128) mbar_wait_spin(bar, par); else mbar_wait_parked(bar, par); } while (0)
#else
#define FP_DBG(a) 0
#define PLAIN_WAIT(bar, par) mbar_wait_parked(bar, par)
#endif

template <int BLEND_BF16>
__global__ void __launch_bounds__(NTHREADS, 1)
dcn_fp_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmX,
              const __grid_constant__ CUtensorMap tmO, const FArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) u64 s_full[MAX_STAGES];
  __shared__ __align__(8) u64 s_empty[MAX_STAGES];
  __shared__ __align__(8) u64 s_bfull[MAX_B];
  __shared__ __align__(8) u64 s_bempty[MAX_B];
  __shared__ __align__(8) u64 s_tfull[2];
  __shared__ __align__(8) u64 s_tempty[2];
  __shared__ __align__(8) u64 s_tabfull[2];
  __shared__ __align__(8) u64 s_tabempty[2];
  __shared__ __align__(8) u64 s_fpfull[MAX_FP];
  __shared__ __align__(8) u64 s_fpempty[MAX_FP];
  __shared__ __align__(8) u64 s_omfull;
  __shared__ u32 s_tmem;

  const cnb_conv_desc& d = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const u32 smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const u32 fp_s = smem_base + (u32)a.nb * a.bstage;                                 // two footprint boxes
  unsigned char* after_fp = smem_al + (size_t)a.nb * a.bstage + (size_t)a.nfp * a.fp_stride;
  const int ntab2 = a.plain ? 0 : 2 * NTAB;                                         // (no table / offsets in plain mode)
  float4* s_tabw = reinterpret_cast<float4*>(after_fp);                              // [2][NTAB]
  u32* s_tabb = reinterpret_cast<u32*>(s_tabw + ntab2);                              // [2][NTAB]
  float* s_om = reinterpret_cast<float*>(s_tabb + ntab2);                            // [BM][OM_CS]
  float* s_scale = s_om + (a.plain ? 0 : BM * OM_CS);
  float* s_shift = s_scale + a.BN;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&s_full[s], WPG);       // the four warps of the stage's sampler group
      mbar_init(&s_empty[s], 1);
    }
    for (int s = 0; s < a.nb; ++s) {
      mbar_init(&s_bfull[s], 1);
      mbar_init(&s_bempty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_tfull[i], a.plain ? NG : 1);     // plain mode: one commit per issuing group
      mbar_init(&s_tempty[i], a.plain ? 8 : 4);   // one arrival per epilogue warp (plain mode: the setup warps join)
      mbar_init(&s_tabfull[i], NSETUP);           // every setup thread releases its own table rows
      mbar_init(&s_tabempty[i], NPROD_WARPS * 32);  // every sampler thread: it is the thread that read them
    }
    for (int i = 0; i < a.nfp; ++i) {
      mbar_init(&s_fpfull[i], 1);
      mbar_init(&s_fpempty[i], NPROD_WARPS);
    }
    mbar_init(&s_omfull, 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == W_LOAD && lane == 0) {
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmO);
  }
  if (warp == W_MMA) tmem_alloc(&s_tmem, a.tmem_cols);
  pdl_launch_dependents();
  pdl_wait();   // global memory (x, offsets, scale/shift) is read only after the previous kernel has completed
  for (int i = tid; i < a.BN; i += NTHREADS) {
    s_scale[i] = (i < d.Co && a.scale) ? a.scale[i] : 1.f;
    s_shift[i] = (i < d.Co && a.shift) ? a.shift[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = s_tmem;
  const int nslabs = d.Ci >> 6;
  const int TW = 1 << a.tw_shift;
  // contiguous tile range per CTA
  const int tile_begin = (int)((long long)blockIdx.x * a.m_tiles / gridDim.x);
  const int tile_end = (int)((long long)(blockIdx.x + 1) * a.m_tiles / gridDim.x);

  // TMEM -> scale/shift (bias or folded BN) -> (+residual) -> ReLU -> NHWC bf16 | NHWC fp32; warp `half` of `nhalves` per TMEM
  // lane quarter takes every nhalves-th 16-column group
  auto run_epilogue = [&](int half, int nhalves) {
    const int q = warp & 3;
    u32 t = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const u32 acc = a.nacc == 2 ? (t & 1u) : 0u, acc_ph = (a.nacc == 2 ? (t >> 1) : t) & 1u;
      const TileXY tc = tile_xy(a, tile);
      const int r = 32 * q + lane;
      const int oy = tc.y0 + (r >> a.tw_shift), ox = tc.x0 + (r & (TW - 1));
      const int m = (tc.n * d.Hi + oy) * d.Wi + ox;
      mbar_wait_parked(&s_tfull[acc], acc_ph);
      tc_fence_after();
      const u32 taddr = tmem_base + acc * a.acc_stride + ((u32)(32 * q) << 16);
      const int ngroups = a.BN / 16;
      for (int g = half; g < ngroups; g += nhalves) {
        u32 v[16];
        tmem_ld16_nowait(taddr + (u32)(g * 16), v);
        tmem_ld_wait();
        const int co0 = g * 16;
        if (a.plain) {   // one partial sum per issuing group, added in group order
          for (int pg = 1; pg < NG; ++pg) {
            u32 w[16];
            tmem_ld16_nowait(taddr + (u32)pg * a.part_stride + (u32)(g * 16), w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
          }
        }
        if (oy < d.Hi && co0 < d.Co && !(FP_DBG(a) & 64)) epilogue_store(a, s_scale, s_shift, v, m, co0, co0, d.Hi * d.Wi, 0, 0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_tempty[acc]);
    }
  };

  if (warp < NPROD_WARPS) {
    // =============================== samplers ===============================================================
    // Group g = warp / 4 takes the K blocks k = g, g + 4, ... of the CTA's stream (tile-major, slab-major, tap-minor);
    // warp q = warp % 4 of the group owns TMEM lanes 32q..32q+31, i.e. a thread owns tile row 32q + lane: one table
    // entry, four corner addresses, then 8 chunks of (4 shared-memory loads, fp32 blend, tcgen05.st of 4 columns).  The
    // loads of chunk c+1 are issued before chunk c is blended.
    const int grp = warp >> 2;
    const int row = ((warp & 3) << 5) + lane;
    const u32 cs2 = (u32)d.x_cstride * 2u;             // bytes per pixel of the global tensor
    const u32 rowb = (u32)d.Wi * cs2;
    const u64 xg = reinterpret_cast<u64>(a.x + d.x_coffset);
    const u32 ta_row = tmem_base + ((u32)((warp & 3) << 5) << 16) + a.a_col0;   // this warp's lanes, first A stage
    u32 s = (u32)grp % (u32)a.stages, ph = ((u32)grp / (u32)a.stages) & 1u, t = 0, u = 0;
    u32 fb = 0, fph = 0;   // box of the current (tile, slab) and its barrier parity
    int tap = grp;
    if (a.plain) {
      // Plain 3x3 convolution: the A row of tap (kh, kw) is box pixel (ty + kh, tx + kw) as it is -- 8 shared-memory
      // loads and two 16-column tcgen05.st per row and K block; the box is read 9 times from shared memory instead of 9
      // times from L2 (TMA im2col).  With no arithmetic in the sampler the kernel is bound by the lone MMA warp's
      // bookkeeping (~500 clocks per K block, measured), so in this mode there is none: the four warps of a group meet
      // at a named barrier and the group's first warp issues the K block's MMAs itself -- four issuers in parallel.
      // MMAs of different issuers reach the tensor core in any order, so each group accumulates into its OWN partial
      // accumulator (one thread issues all of a partial's MMAs, in program order: the result is reproducible bit for bit)
      // and the epilogue adds the four partials in group order.
      const int ty = row >> a.tw_shift, tx = row & (TW - 1);
      const bool issuer = (warp & 3) == 0;
      const u64 db0 = make_sdesc(smem_base, 16, 1024, 2);
      const u32 bstage16 = a.bstage >> 4;
      const bool two_acc = a.nacc == 2;
      const u32 wrapb = a.b_resident ? (u32)a.nkb : (u32)a.nb;
      u32 sb = (u32)grp % wrapb, useb = (u32)grp / wrapb;   // weight-tile slot of this group's next K block, and its use count
      for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
        const u32 acc = two_acc ? (t & 1u) : 0u, acc_ph = (two_acc ? (t >> 1) : t) & 1u;
        const u32 tmem_d = tmem_base + acc * a.acc_stride + (u32)grp * a.part_stride;
        bool first = true;                                 // first K block of this group in this tile
        for (int slab = 0; slab < nslabs; ++slab, ++u) {
          const u32 fpb = fp_s + fb * a.fp_stride;
          PLAIN_WAIT(&s_fpfull[fb], fph);
          for (; tap < 9; tap += NG) {
            const int kh = (tap * 11) >> 5, kw = tap - 3 * kh;     // tap / 3 for tap < 9
            const u32 pix = (u32)((ty + kh) * a.FW + tx + kw);
            const u32 e = (fpb + pix * 128u) | ((pix & 7u) << 4);
            PLAIN_WAIT(&s_empty[s], ph ^ 1u);        // the MMAs that read this A stage last time have completed
            tc_fence_after();
            const u32 ta = ta_row + s * A_COLS;
            if (FP_DBG(a) & 1) {   // experiment: no shared-memory reads
              const uint4 z = make_uint4(e, e, e, e);
              if (!(FP_DBG(a) & 2)) {
                tmem_st16(ta, z, z, z, z);
                tmem_st16(ta + 16u, z, z, z, z);
              }
            } else {
              const uint4 q0 = lds128(e), q1 = lds128(e ^ 16u), q2 = lds128(e ^ 32u), q3 = lds128(e ^ 48u);
              const uint4 q4 = lds128(e ^ 64u), q5 = lds128(e ^ 80u), q6 = lds128(e ^ 96u), q7 = lds128(e ^ 112u);
              tmem_st16(ta, q0, q1, q2, q3);
              tmem_st16(ta + 16u, q4, q5, q6, q7);
            }
            tmem_st_wait();
            tc_fence_before();
            // the group's 128 rows are in tensor memory (named barrier 2 + group, as an immediate: a register id makes
            // ptxas reserve all 16 barriers)
            if (grp == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
            else if (grp == 1) asm volatile("bar.sync 3, 128;" ::: "memory");
            else if (grp == 2) asm volatile("bar.sync 4, 128;" ::: "memory");
            else asm volatile("bar.sync 5, 128;" ::: "memory");
            if (issuer) {
              if (first) PLAIN_WAIT(&s_tempty[acc], acc_ph ^ 1u);  // the accumulator is drained
              if (!a.b_resident || useb == 0) PLAIN_WAIT(&s_bfull[sb], useb & 1u);
              tc_fence_after();
              const bool last = slab == nslabs - 1 && tap + NG >= 9;     // this group's last K block of the tile
              if (elect_one()) {
                const u64 db = db0 + (u64)(sb * bstage16);
                const u32 tam = tmem_base + a.a_col0 + s * A_COLS;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  if (!(FP_DBG(a) & 4)) umma_bf16_ts(tmem_d, tam + (u32)(8 * kk), db + (u64)(2 * kk), a.idesc, (first && kk == 0) ? 0u : 1u);
                umma_commit(&s_empty[s]);
                if (!a.b_resident) umma_commit(&s_bempty[sb]);
                if (last) umma_commit(&s_tfull[acc]);
              }
              __syncwarp();
            }
            first = false;
            s += (u32)NG;
            if (s >= (u32)a.stages) {
              s -= (u32)a.stages;
              ph ^= 1u;
            }
            sb += (u32)NG;
            while (sb >= wrapb) {
              sb -= wrapb;
              ++useb;
            }
          }
          tap -= 9;
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_fpempty[fb]);
          if (++fb == (u32)a.nfp) {
            fb = 0;
            fph ^= 1u;
          }
        }
      }
    } else
    for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const u32 tb = t & 1u;
      mbar_wait_parked(&s_tabfull[tb], (t >> 1) & 1u);
      const float4* tw = s_tabw + tb * NTAB + row * 9;
      const u32* tbs = s_tabb + tb * NTAB + row * 9;
      for (int slab = 0; slab < nslabs; ++slab, ++u) {
        // this group's taps of the (tile, slab) box: tap, tap + 4, ... < 9
        const u32 fpb = fp_s + fb * a.fp_stride;
        mbar_wait_parked(&s_fpfull[fb], fph);
        float4 w_next = tw[tap];          // the table entry of the next K block is read one K block ahead
        u32 b_next = tbs[tap];
        for (; tap < 9; tap += NG) {
          const float4 w = (FP_DBG(a) & 8) ? make_float4(0.25f, 0.25f, 0.25f, 0.25f) : w_next;
          const u32 b = (FP_DBG(a) & 8) ? (u32)(row + tap) : b_next;
          if (tap + NG < 9) {
            w_next = tw[tap + NG];
            b_next = tbs[tap + NG];
            if (b_next >> 31) {   // a far sample of the next K block: pull its four corner lines into L1 now (the generic
                                  // loads below would otherwise each wait for L2)
              const u64 pb = xg + (u64)(slab * 128) + (u64)(b_next & 0x1FFFFFFFu) * cs2;
              const u32 pd1 = ((b_next >> 29) & 1u) ? cs2 : 0u, pd2 = ((b_next >> 30) & 1u) ? rowb : 0u;
              asm volatile("prefetch.global.L1 [%0];" ::"l"(pb));
              asm volatile("prefetch.global.L1 [%0];" ::"l"(pb + pd1));
              asm volatile("prefetch.global.L1 [%0];" ::"l"(pb + pd2));
              asm volatile("prefetch.global.L1 [%0];" ::"l"(pb + pd2 + pd1));
            }
          }
          if (FP_DBG(a) & 128) mbar_wait_spin(&s_empty[s], ph ^ 1u); else mbar_wait_parked(&s_empty[s], ph ^ 1u);
          tc_fence_after();
          const u32 pix = b & 0x1FFFFFFFu;
          const u32 ta = ta_row + s * A_COLS;
          if (__all_sync(0xffffffffu, (int)(b >> 31) == 0)) {
            // every lane's corners are inside the staged box.  Corner k is box pixel pk; its logical chunk c sits at
            // 16-byte slot c ^ (pk & 7) of the pixel's 128 bytes: e_k = pixel address | swizzle, load address = e_k ^ 16c
            const u32 p1 = pix + 1u, p2 = pix + (u32)a.FW, p3 = p2 + 1u;
            const u32 e0 = (fpb + pix * 128u) | ((pix & 7u) << 4), e1 = (fpb + p1 * 128u) | ((p1 & 7u) << 4),
                      e2 = (fpb + p2 * 128u) | ((p2 & 7u) << 4), e3 = (fpb + p3 * 128u) | ((p3 & 7u) << 4);
            uint4 q0[4], q1[4];
            const int dbg = FP_DBG(a);
            auto ld = [&](u32 adr) { return (dbg & 1) ? make_uint4(adr, adr + 1, adr + 2, adr + 3) : lds128(adr); };
            auto bl = [&](const uint4 (&q)[4]) {
              return (dbg & 2) ? make_uint4(q[0].x ^ q[1].y, q[2].z ^ q[3].w, q[0].w, q[1].x) : blend4<BLEND_BF16>(w, q[0], q[1], q[2], q[3]);
            };
            auto st = [&](u32 adr, const uint4 v) {
              if (!(dbg & 4)) tmem_st4(adr, v);
              else if (v.x == 0x12345678u && v.w == 0x9abcdef0u) tmem_st4(adr, v);
            };
            q0[0] = ld(e0); q0[1] = ld(e1); q0[2] = ld(e2); q0[3] = ld(e3);
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              const u32 x1 = (u32)(c + 1) << 4;
              q1[0] = ld(e0 ^ x1); q1[1] = ld(e1 ^ x1); q1[2] = ld(e2 ^ x1); q1[3] = ld(e3 ^ x1);
              st(ta + (u32)(4 * c), bl(q0));
              if (c + 2 < 8) {
                const u32 x2 = (u32)(c + 2) << 4;
                q0[0] = ld(e0 ^ x2); q0[1] = ld(e1 ^ x2); q0[2] = ld(e2 ^ x2); q0[3] = ld(e3 ^ x2);
              }
              st(ta + (u32)(4 * c + 4), bl(q1));
            }
          } else {
            // some (pixel, tap) of this warp reaches beyond the box: generic loads, each lane from its own window
            const bool g = (b >> 31) != 0;
            const u64 base = g ? xg + (u64)(slab * 128) + (u64)pix * cs2 : (u64)__cvta_shared_to_generic(fpb) + (u64)pix * 128u;
            const u32 d1 = g ? (((b >> 29) & 1u) ? cs2 : 0u) : 128u;
            const u32 d2 = g ? (((b >> 30) & 1u) ? rowb : 0u) : (u32)a.FW * 128u;
            const u64 adr[4] = {base, base + d1, base + d2, base + d2 + d1};
            const u32 p1 = pix + 1u, p2 = pix + (u32)a.FW, p3 = p2 + 1u;
            const u32 sw[4] = {g ? 0u : (pix & 7u) << 4, g ? 0u : (p1 & 7u) << 4, g ? 0u : (p2 & 7u) << 4, g ? 0u : (p3 & 7u) << 4};
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
              uint4 q[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) q[k] = ldgen128(adr[k] + (((u32)c << 4) ^ sw[k]));
              tmem_st4(ta + (u32)(4 * c), blend4<BLEND_BF16>(w, q[0], q[1], q[2], q[3]));
            }
          }
          tmem_st_wait();                 // the row's 32 columns are in tensor memory
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_full[s]);
          s += (u32)NG;                   // next K block of this group: NG further in the stream and in the ring
          if (s >= (u32)a.stages) {
            s -= (u32)a.stages;
            ph ^= 1u;
          }
        }
        tap -= 9;
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_fpempty[fb]);   // the group leaves this (tile, slab) box
        if (++fb == (u32)a.nfp) {
          fb = 0;
          fph ^= 1u;
        }
      }
      mbar_arrive(&s_tabempty[tb]);                   // ... and this tile (per thread: each read its own table row)
    }
  } else if (warp < W_MMA) {
    // =============================== setup: (pixel, tap) -> weights + corner index ============================
    // A setup thread owns one tile row: it pulls the row's 27 offset/mask values into registers (7 x 16 bytes; the box of
    // offsets arrives by ONE tensor-map copy with the 128-byte swizzle, so the 8 lanes of a wavefront read 8 different
    // bank groups), hands the staging buffer back at once -- the copy for tile t+1 flies while the table of tile t is
    // computed -- and then works out its 9 taps as 9 independent instruction streams.  (First version: one (pixel, tap)
    // item per thread and step, the buffer refilled after the table was written: a serial chain of load latency +
    // lone-warp arithmetic of ~10000 clocks per tile, which bounded the whole kernel.)
    const int r = tid - W_SETUP0 * 32;                 // tile row 0..127
    if (a.plain) {
      run_epilogue(1, 2);   // nothing to set up: these warps (TMEM lane quarters 0..3) take every other column group
    } else {
    if (r == 0 && tile_begin < tile_end) {
      const TileXY tc = tile_xy(a, tile_begin);
      mbar_expect_tx(&s_omfull, BM * OM_CS * 4u);
      tma_load_tile_4d(smem_u32(s_om), &tmO, 0, tc.x0, tc.y0, tc.n, &s_omfull);
    }
    const u32 om_row = smem_u32(s_om) + (u32)r * 128u;
    const u32 om_sw = (u32)(r & 7) << 4;
    u32 t = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const u32 tb = t & 1u;
      mbar_wait_parked(&s_omfull, t & 1u);
      float om[28];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const uint4 v = lds128(om_row + (((u32)j << 4) ^ om_sw));
        om[4 * j] = __uint_as_float(v.x); om[4 * j + 1] = __uint_as_float(v.y);
        om[4 * j + 2] = __uint_as_float(v.z); om[4 * j + 3] = __uint_as_float(v.w);
      }
      // All 128 rows are in registers -- once the loads have actually RETURNED: the copy that refills the buffer is an
      // async-proxy write, and a barrier alone does not hold it behind shared-memory loads that are still in flight
      // (nothing has consumed their registers yet).  Without this fence one run in ~500 computed a tile row of table
      // entries from the NEXT tile's offsets (found by tests/test_conv_gpu.py::test_dcnv2_repeats_are_bit_identical).
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(NSETUP) : "memory");
      if (r == 0 && tile + 1 < tile_end) {
        const TileXY tn = tile_xy(a, tile + 1);
        mbar_expect_tx(&s_omfull, BM * OM_CS * 4u);
        tma_load_tile_4d(smem_u32(s_om), &tmO, 0, tn.x0, tn.y0, tn.n, &s_omfull);
      }
      const TileXY tc = tile_xy(a, tile);
      const int bx0 = tc.x0 - 1 - a.R, by0 = tc.y0 - 1 - a.R;   // box origin (may be negative: zero-filled)
      const int oy = tc.y0 + (r >> a.tw_shift), ox = tc.x0 + (r & (TW - 1));
      const u32 gbase = (u32)(tc.n * d.Hi) * (u32)d.Wi;
      mbar_wait_parked(&s_tabempty[tb], ((t >> 1) & 1u) ^ 1u);
      float4* ow = s_tabw + tb * NTAB + r * 9;
      u32* ob = s_tabb + tb * NTAB + r * 9;
      const float fH = (float)d.Hi, fW = (float)d.Wi;
      const float fy = (float)(oy - 1), fx = (float)(ox - 1);
      const int boff = -by0 * a.FW - bx0;                // box pixel index = y0 * FW + x0 + boff
      u32 far = 0;                                       // taps of this row that leave the box (rare)
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int kh = tap / 3, kw = tap - 3 * kh;
        const float mk = __fdividef(1.f, 1.f + __expf(-om[18 + tap]));
        const float py = fy + (float)kh + om[2 * tap];
        const float px = fx + (float)kw + om[2 * tap + 1];
        const float fy0 = floorf(py), fx0 = floorf(px);
        const float ly = py - fy0, lx = px - fx0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const int y0 = (int)fy0, x0 = (int)fx0;
        const bool valid = (oy < d.Hi) && py > -1.f && px > -1.f && py < fH && px < fW;
        // inside the box all four corners are read as they are: corners outside the image hit the zero fill
        const bool inbox = y0 >= by0 && x0 >= bx0 && y0 + 1 < by0 + a.FH && x0 + 1 < bx0 + a.FW;
        const float m = valid ? mk : 0.f;
        const float wy0 = hy * m, wy1 = ly * m;
        ow[tap] = make_float4(wy0 * hx, wy0 * lx, wy1 * hx, wy1 * lx);
        ob[tap] = (valid && inbox) ? (u32)(y0 * a.FW + x0 + boff) : 0u;
        far |= (valid && !inbox) ? (1u << tap) : 0u;
      }
      if (FP_DBG(a) & 16) far = 0;
      if (__any_sync(0xffffffffu, far != 0)) {
        // samples beyond the box: clamped global pixel + validity-masked weights, flagged for the generic-load path
        for (int tap = 0; tap < 9; ++tap) {
          if (!((far >> tap) & 1u)) continue;
          const int kh = tap / 3, kw = tap - 3 * kh;
          float dyv = 0.f, dxv = 0.f, mv = 0.f;
#pragma unroll
          for (int k = 0; k < 9; ++k)
            if (k == tap) { dyv = om[2 * k]; dxv = om[2 * k + 1]; mv = om[18 + k]; }
          const float mk = __fdividef(1.f, 1.f + __expf(-mv));
          const float py = fy + (float)kh + dyv, px = fx + (float)kw + dxv;
          const int y0 = (int)floorf(py), x0 = (int)floorf(px);
          const float ly = py - (float)y0, lx = px - (float)x0;
          const float hy = 1.f - ly, hx = 1.f - lx;
          const bool vy0 = y0 >= 0, vy1 = y0 + 1 <= d.Hi - 1, vx0 = x0 >= 0, vx1 = x0 + 1 <= d.Wi - 1;
          float4 w;
          w.x = (vy0 && vx0) ? hy * hx * mk : 0.f;
          w.y = (vy0 && vx1) ? hy * lx * mk : 0.f;
          w.z = (vy1 && vx0) ? ly * hx * mk : 0.f;
          w.w = (vy1 && vx1) ? ly * lx * mk : 0.f;
          const int y0c = max(y0, 0), y1c = min(y0 + 1, d.Hi - 1);
          const int x0c = max(x0, 0), x1c = min(x0 + 1, d.Wi - 1);
          ow[tap] = w;
          ob[tap] = 0x80000000u | ((u32)(y1c - y0c) << 30) | ((u32)(x1c - x0c) << 29) | (gbase + (u32)(y0c * d.Wi + x0c));
        }
      }
      mbar_arrive(&s_tabfull[tb]);
    }
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer ==============================================================
    // A lone warp executes ~1 dependent instruction per 5-8 clocks, and in plain mode this loop is what bounds the kernel
    // (first version: ~120 instructions per K block = 650 clocks): barrier addresses, the weight-tile descriptor and the
    // A-stage address are loop-carried state advanced by adds, everything else is hoisted.
    if (!a.plain) {   // (plain mode: the sampler groups issue their own MMAs)
    const u32 full0 = smem_u32(&s_full[0]), empty0 = smem_u32(&s_empty[0]);
    const u32 bfull0 = smem_u32(&s_bfull[0]), bempty0 = smem_u32(&s_bempty[0]);
    const u32 tfull0 = smem_u32(&s_tfull[0]), tempty0 = smem_u32(&s_tempty[0]);
    const u32 s_wrap = (u32)a.stages * 8u, sb_wrap = (u32)(a.b_resident ? a.nkb : a.nb) * 8u;
    const bool resident = a.b_resident != 0;
    const u32 idesc = a.idesc, ta0 = tmem_base + a.a_col0, acc_stride = a.acc_stride;
    const bool two_acc = a.nacc == 2;
    const int nkb = a.nkb, kps = a.kps;
    const u64 db0 = make_sdesc(smem_base, 16, 1024, 2);
    const u64 bstage16 = (u64)(a.bstage >> 4);
    u32 s8 = 0, ph = 0, sb8 = 0, phb = 0, ta = ta0, t = 0;
    u64 db = db0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const u32 acc = two_acc ? (t & 1u) : 0u, acc_ph = (two_acc ? (t >> 1) : t) & 1u;
      mbar_wait_parked_a(tempty0 + acc * 8u, acc_ph ^ 1u);
      tc_fence_after();
      const u32 tmem_d = tmem_base + acc * acc_stride;
      const bool wait_b = !resident || t == 0;
      u32 accumulate = 0;
      for (int kb = 0; kb < nkb; kb += kps) {
        mbar_wait_parked_a(full0 + s8, ph);                  // the sampled rows of kps K blocks
        for (int j = 0; j < kps; ++j) {
          if (wait_b) mbar_wait_parked_a(bfull0 + sb8, phb);   // the weight tile (prefetched far ahead)
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)   // K = 16: 8 packed columns of A, +32 bytes of B inside the swizzle atom
              if (!(FP_DBG(a) & 32)) umma_bf16_ts(tmem_d, ta + (u32)(8 * kk), db + (u64)(2 * kk), idesc, kk == 0 ? accumulate : 1u);
            if (j == kps - 1) umma_commit_a(empty0 + s8);
            if (!resident) umma_commit_a(bempty0 + sb8);
          }
          __syncwarp();
          accumulate = 1;
          ta += A_COLS;
          sb8 += 8u;
          db += bstage16;
          if (sb8 == sb_wrap) {
            sb8 = 0;
            db = db0;
            phb ^= 1u;
          }
        }
        s8 += 8u;
        if (s8 == s_wrap) {
          s8 = 0;
          ta = ta0;
          ph ^= 1u;
        }
      }
      if (elect_one()) umma_commit_a(tfull0 + acc * 8u);
      __syncwarp();
    }
    }
  } else if (warp < W_LOAD) {
    // =============================== epilogue ================================================================
    run_epilogue(0, a.plain ? 2 : 1);
  } else if (warp == W_LOAD) {
    // =============================== loader: one box per (tile, slab), two in flight ===========================
    // Boxes and weight tiles share the SM's TMA queue; each is requested as early as its slot allows (a weight tile
    // requested only when its K block is sampled arrives ~1500 clocks later, behind 40 KB boxes: measured as a
    // ~850-clock floor per K block).  Non-blocking polls, whichever slot is free goes next.
    const int ntiles = tile_end - tile_begin;
    const long long u_total = (long long)ntiles * nslabs;
    const long long j_total = a.b_resident ? (ntiles > 0 ? a.nkb : 0) : (long long)ntiles * a.nkb;
    long long u = 0, j = 0;
    int utile = tile_begin, uslab = 0, jtap = 0, jslab = 0;
    u32 jslot = 0, juse = 0, ufb = 0, uuse = 0;
    while (u < u_total || j < j_total) {
      bool progressed = false;
      if (j < j_total && (juse == 0 || mbar_test_wait(&s_bempty[jslot], (juse - 1u) & 1u))) {
        if (elect_one()) {
          mbar_expect_tx(&s_bfull[jslot], a.b_bytes);
          tma_load_2d(smem_base + jslot * a.bstage, &tmB, jtap * d.Ci + jslab * 64, 0, &s_bfull[jslot]);
        }
        __syncwarp();
        if (++jtap == 9) {
          jtap = 0;
          if (++jslab == nslabs) jslab = 0;
        }
        if (++jslot == (u32)a.nb) {
          jslot = 0;
          ++juse;
        }
        ++j;
        progressed = true;
      }
      if (u < u_total && (uuse == 0 || mbar_test_wait(&s_fpempty[ufb], (uuse - 1u) & 1u))) {
        const TileXY tc = tile_xy(a, utile);
        const u32 fb = ufb;
        if (elect_one()) {
          if (FP_DBG(a) & 256) {
            mbar_arrive(&s_fpfull[fb]);   // experiment: no box load
          } else {
            mbar_expect_tx(&s_fpfull[fb], a.fp_bytes);
            tma_load_tile_4d(fp_s + fb * a.fp_stride, &tmX, uslab * 64, tc.x0 - 1 - a.R, tc.y0 - 1 - a.R, tc.n, &s_fpfull[fb]);
          }
        }
        __syncwarp();
        if (++uslab == nslabs) {
          uslab = 0;
          ++utile;
        }
        if (++ufb == (u32)a.nfp) {
          ufb = 0;
          ++uuse;
        }
        ++u;
        progressed = true;
      }
      if (!progressed && !(FP_DBG(a) & 512)) __nanosleep(100);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, a.tmem_cols);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

static int box_width(int TW, int R, bool plain) {
  static const bool exact = [] { const char* e = getenv("CNB_DCN_BOXW"); return e && e[0] == 'e'; }();   // A/B: exact width
  const int w = plain ? TW + 2 : TW + 2 * R + 3;
  return exact ? w : (w + 7) / 8 * 8;
}

struct Plan {
  int tw_shift, R, stages, nb, BN, kps, nfp;
  u32 bstage, fp_bytes, fp_stride;
  size_t smem;
};

// Shared memory: tables + offset staging (fixed), two boxes of reach R, nb weight-tile slots.  All 9 * Ci/64 weight tiles
// resident (no reloads at all) is worth one pixel of reach; otherwise >= 4 slots, then the largest reach.  The A ring lives
// in the tensor-memory columns the two accumulators leave free (>= NG stages: a group must not lap the ring).
bool make_plan(const cnb_conv_desc* d, Plan* p, bool plain = false) {
  static const int env_r = [] { const char* e = getenv("CNB_DCN_REACH"); return e ? atoi(e) : 0; }();
  static const int env_stages = [] { const char* e = getenv("CNB_DCN_STAGES"); return e ? atoi(e) : 0; }();
  static const int env_nb = [] { const char* e = getenv("CNB_DCN_NB"); return e ? atoi(e) : 0; }();
  static const int env_nfp = [] { const char* e = getenv("CNB_CONV_FP_BOXES"); return e ? atoi(e) : 0; }();
  p->tw_shift = d->Wi % 16 == 0 ? 4 : 3;
  const int TW = 1 << p->tw_shift, TH = BM >> p->tw_shift;
  p->BN = round_up(d->Co, 16);
  p->bstage = ((u32)p->BN * 128u + 1023u) & ~1023u;
  const int nkb = 9 * (d->Ci / 64);
  // K blocks per A stage: in plain mode with N <= 64 a sampler group fills a whole filter row per handshake
  const int kps = 1;
  p->kps = kps;
  // plain mode: NG partial accumulators per tile (double buffered while they fit)
  if (plain && round_up(p->BN, 32) > 64) return false;
  const int acc_cols = plain ? (p->BN > 32 ? 1 : 2) * NG * round_up(p->BN, 32) : (p->BN > 128 ? 1 : 2) * round_up(p->BN, 32);
  int st = (512 - acc_cols) / (A_COLS * kps);
  if (st > MAX_STAGES) st = MAX_STAGES;
  if (env_stages > 0 && st > env_stages) st = env_stages;
  // a stage must always come back to the SAME group (depth a multiple of NG): a group is then ordered behind its own
  // previous use of the stage, and no group can run two rounds ahead of another one's pending commit on a stage (the
  // parity wait cannot tell round r+1 from round r-1)
  st -= st % NG;
  if (st < NG) return false;
  p->stages = st;
  const size_t fixed = (plain ? 0 : (size_t)2 * NTAB * (sizeof(float4) + sizeof(u32)) + (size_t)BM * OM_CS * 4) +
                       (size_t)p->BN * 8 + 1024;
  const size_t budget = 226 * 1024;
  static const int pref[6][2] = {{3, -1}, {2, -1}, {3, 4}, {2, 4}, {2, 2}, {1, 2}};   // (R, minimum slots; -1 = all tiles resident)
  static const int pref_plain[1][2] = {{0, 2}};
  const int (*cands)[2] = plain ? pref_plain : pref;
  const int ncand = plain ? 1 : 6;
  for (int ci = 0; ci < ncand; ++ci) {
    const int* c = cands[ci];
    const int R = plain ? 0 : (env_r > 0 ? env_r : c[0]);
    // box width rounded up to a multiple of 8 pixels: a row step then leaves the swizzle (box pixel index mod 8) unchanged,
    // so lanes that sample distinct columns mod 8 never share a bank group however their rows jitter (with 23 columns a
    // neighbour one row down and one column right collided: measured 2-way conflicts on the network's offset fields)
    const int FW = box_width(TW, R, plain), FH = plain ? TH + 2 : TH + 2 * R + 3;
    const u32 fpb = (u32)FW * FH * 128u;
    const u32 fps = (fpb + 1023u) & ~1023u;
    if (fixed + 2 * (size_t)fps + 2 * (size_t)p->bstage > budget) continue;
    int nb = (int)((budget - fixed - 2 * (size_t)fps) / p->bstage);
    if (nb > nkb) nb = nkb;
    if (nb > MAX_B) nb = MAX_B;
    if (env_nb > 0 && nb > env_nb) nb = env_nb;
    const int need = env_nb > 0 ? 2 : (c[1] < 0 ? nkb : c[1]);
    if (nb >= need && nb >= 2) {
      p->R = R; p->nb = nb; p->fp_bytes = fpb; p->fp_stride = fps;
      // plain mode: CNB_CONV_FP_BOXES > 2 puts what the weights leave into a deeper box ring (measured: no gain, the
      // samplers' own instruction stream bounds this mode, profiles/r02s3f_plain_conv_stage_skipping.txt)
      int nfp = 2;
      if (plain && env_nfp > 2) {
        nfp = (int)((budget - fixed - (size_t)nb * p->bstage) / fps);
        if (nfp > MAX_FP) nfp = MAX_FP;
        if (nfp > env_nfp) nfp = env_nfp;
        if (nfp < 2) nfp = 2;
      }
      p->nfp = nfp;
      p->smem = fixed + (size_t)nfp * fps + (size_t)nb * p->bstage;
      return true;
    }
  }
  return false;
}

}  // namespace

// Geometries this kernel covers (the rest stays on the global-gather kernel of dcn_ws.cu).  CNB_DCN_IMPL=ws keeps that
// kernel for A/B runs.
bool dcn_fp_supported(const cnb_conv_desc* d, int om_cstride) {
  static const bool off = [] { const char* e = getenv("CNB_DCN_IMPL"); return e && (e[0] == 'v' || e[0] == 'w'); }();
  Plan p;
  return !off && om_cstride == OM_CS && d->Ci % 64 == 0 && d->Co % 8 == 0 && round_up(d->Co, 16) <= 256 && d->Wi % 8 == 0 &&
         (long long)d->B * d->Hi * d->Wi < (1ll << 29) &&
         d->x_cstride % 8 == 0 && d->x_coffset % 8 == 0 && make_plan(d, &p);
}

static int fp_run(const cnb_conv_desc* d, const void* x, const float* om, const void* wpk, const float* scale,
                  const float* shift, const void* res, void* y, cudaStream_t st, bool plain) {
  TmaDriver& drv = tma_driver();
  if (!drv.ok) {
    set_error("dcnv2: cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    return CNB_ERR_CUDA;
  }
  Plan p;
  if (!make_plan(d, &p, plain)) {
    set_error("dcnv2: no shared-memory plan for Co=%d", d->Co);
    return CNB_ERR_INVALID;
  }
  FArgs a;
  a.d = *d;
  a.x = (const __nv_bfloat16*)x;
  a.om = om;
  a.scale = scale;
  a.shift = shift;
  a.res = (const __nv_bfloat16*)res;
  a.y = y;
  a.plain = plain ? 1 : 0;
  a.tw_shift = p.tw_shift;
  const int TW = 1 << p.tw_shift, TH = BM >> p.tw_shift;
  a.tiles_x = d->Wi / TW;
  a.tiles_y = (d->Hi + TH - 1) / TH;
  a.m_tiles = d->B * a.tiles_x * a.tiles_y;
  a.R = p.R;
  a.FW = box_width(TW, p.R, plain);
  a.FH = plain ? TH + 2 : TH + 2 * p.R + 3;
  a.fp_bytes = p.fp_bytes;
  a.fp_stride = p.fp_stride;
  a.nfp = p.nfp;
  a.BN = p.BN;
  a.nkb = 9 * (d->Ci / 64);
  a.b_bytes = (u32)a.BN * 128u;
  a.bstage = p.bstage;
  a.stages = p.stages;
  a.kps = p.kps;
  a.nb = p.nb;
  a.b_resident = p.nb >= a.nkb ? 1 : 0;
  a.part_stride = (u32)round_up(a.BN, 32);
  a.acc_stride = plain ? (u32)NG * a.part_stride : a.part_stride;
  a.nacc = plain ? (a.BN > 32 ? 1 : 2) : (a.BN > 128 ? 1 : 2);
  a.a_col0 = (u32)a.nacc * a.acc_stride;
  a.tmem_cols = 512;   // one CTA per SM: two accumulators + the A ring
  a.idesc = make_idesc_bf16(BM, a.BN);

  CUtensorMap tmB, tmX, tmO;
  {
    const int Kpad = 9 * d->Ci;   // Ci % 64 == 0: already a multiple of the packing granularity
    cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)a.BN};
    cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)a.BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = drv.tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wpk, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("dcnv2: cuTensorMapEncodeTiled (weights) failed (%d) Kpad=%d BN=%d", (int)r, Kpad, a.BN);
      return CNB_ERR_CUDA;
    }
  }
  {
    // the input as {C, W, H, N}; one box = 64 channels x FW x FH pixels, 128 bytes per pixel in shared memory with the
    // 128-byte swizzle (chunk index XOR box pixel index mod 8), out-of-image pixels zero-filled
    const cuuint64_t cs2 = (cuuint64_t)d->x_cstride * 2;
    cuuint64_t dims[4] = {(cuuint64_t)d->Ci, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {cs2, cs2 * d->Wi, cs2 * d->Wi * d->Hi};
    cuuint32_t box[4] = {64u, (cuuint32_t)a.FW, (cuuint32_t)a.FH, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = drv.tiled(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)((const __nv_bfloat16*)x + d->x_coffset), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("dcnv2: cuTensorMapEncodeTiled (input box %dx%d) failed (%d)", a.FW, a.FH, (int)r);
      return CNB_ERR_CUDA;
    }
  }
  if (plain) {
    tmO = tmX;   // unused
  } else {
    // the offset/mask map as {32 floats, W, H, N}: one box = the tile's TW x TH pixels, 128 bytes each, 128-byte swizzle;
    // rows below the image are zero-filled (their table entries are never used)
    cuuint64_t dims[4] = {(cuuint64_t)OM_CS, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)OM_CS * 4, (cuuint64_t)OM_CS * 4 * d->Wi, (cuuint64_t)OM_CS * 4 * d->Wi * d->Hi};
    cuuint32_t box[4] = {(cuuint32_t)OM_CS, (cuuint32_t)TW, (cuuint32_t)TH, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = drv.tiled(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)om, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("dcnv2: cuTensorMapEncodeTiled (offset map) failed (%d)", (int)r);
      return CNB_ERR_CUDA;
    }
  }
  // CNB_DCN_BLEND: fp32 (default) | wbf16 (bilinear weights rounded to bf16, mixed-precision FMAs) | bf16 (bf16x2 blend)
  static const int blend_mode = [] { const char* e = getenv("CNB_DCN_BLEND"); return !e ? 0 : e[0] == 'w' ? 2 : e[0] == 'b' ? 1 : 0; }();
  static const int env_dbg = [] { const char* e = getenv("CNB_DCN_DEBUG"); return e ? atoi(e) : 0; }();
  a.debug = env_dbg;
  static PerDeviceOnce once;
  if (once.need()) {
    CNB_CUDA(cudaFuncSetAttribute(dcn_fp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    CNB_CUDA(cudaFuncSetAttribute(dcn_fp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    CNB_CUDA(cudaFuncSetAttribute(dcn_fp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    once.mark();
  }
  const int grid = a.m_tiles < drv.num_sms ? a.m_tiles : drv.num_sms;
  if (blend_mode == 1 && !plain)
    CNB_CUDA(launch_pdl(dcn_fp_kernel<1>, dim3(grid), dim3(NTHREADS), p.smem, st, tmB, tmX, tmO, a));
  else if (blend_mode == 2 && !plain)
    CNB_CUDA(launch_pdl(dcn_fp_kernel<2>, dim3(grid), dim3(NTHREADS), p.smem, st, tmB, tmX, tmO, a));
  else
    CNB_CUDA(launch_pdl(dcn_fp_kernel<0>, dim3(grid), dim3(NTHREADS), p.smem, st, tmB, tmX, tmO, a));
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

int dcn_fp_run(const cnb_conv_desc* d, const void* x, const float* om, int om_cstride, const void* wpk,
               const float* scale, const float* shift, void* y, cudaStream_t st) {
  (void)om_cstride;
  return fp_run(d, x, om, wpk, scale, shift, nullptr, y, st, false);
}

// Plain 3x3 / stride 1 / pad 1 convolutions with Ci % 64 == 0 and N <= 64 through the same kernel (`plain` mode: input
// box in shared memory read 9 times by the sampler threads, A operand in tensor memory, the sampler groups issue their own
// MMAs, each into its own partial accumulator; the epilogue adds the partials in group order, so the result is
// bit-reproducible -- tools/conv_race_hunt.py).  Measured against conv_rows / conv_tma (B=32, us,
// profiles/r02s3a_plain_conv_partials.txt): 128->27@64x64 33.8 vs 37.9, 128->64@64x64 39.3 vs 41.9; 64->27@128x128 64.4 vs
// 58.9 -- only the layers whose im2col traffic or tile count hurt the other kernels gain (Ci >= 128, at least 4 tiles per
// SM).  What bounds the mode is the sampler groups' own instruction stream (profiles/r02s3f_plain_conv_stage_skipping.txt).
bool conv_fp_supported(const cnb_conv_desc* d) {
  // CNB_CONV_FP: 0 = never, 1 = every geometry the mode covers, 2 = the policy below with N <= 64; unset = the policy:
  // deep thin layers (Ci >= 128, N <= 32: the offset/mask convolutions at 64x64 and up) with >= 4 tiles per SM, where it
  // measured faster than the TMA-im2col kernel
  static const int env = [] { const char* e = getenv("CNB_CONV_FP"); return e ? atoi(e) : -1; }();
  if (env == 0) return false;
  Plan p;
  const bool ok = d->KH == 3 && d->KW == 3 && d->stride == 1 && d->pad == 1 && d->dil == 1 && d->pad_w1 == 0 &&
                  (d->w_kw == 0 || d->w_kw == 3) && d->Ho == d->Hi && d->Wo == d->Wi && d->Ci % 64 == 0 &&
                  round_up(d->Co, 16) <= 64 && d->Wi % 8 == 0 && d->out_nchw_f32 != 1 && d->x_cstride % 8 == 0 &&
                  d->x_coffset % 8 == 0 && (long long)d->B * d->Hi * d->Wi < (1ll << 29) && make_plan(d, &p, true);
  if (!ok || env == 1) return ok;
  const long long m_tiles = ((long long)d->B * d->Hi * d->Wi + BM - 1) / BM;
  return d->Ci >= 128 && round_up(d->Co, 16) <= (env == 2 ? 64 : 32) && m_tiles >= 4LL * sm_count();
}

int conv_fp_run(const cnb_conv_desc* d, const void* x, const void* wpk, const float* scale, const float* shift,
                const void* res, void* y, cudaStream_t st) {
  return fp_run(d, x, nullptr, wpk, scale, shift, res, y, st, true);
}

}  // namespace cnb
