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
// GEMM view and roles are those of dcn_ws.cu: D[128 pixels, Co] = A[128, 9*Ci] * W^T, A sampled on the fly into a ring of
// K-major SWIZZLE_128B stages (one K block = one tap of one 64-channel slab per stage); the sampler warps form groups
// that take the K blocks round-robin; setup warps turn offsets/masks into a table of bilinear weights + box (or global)
// pixel index; one MMA warp (tcgen05.mma, elected lane); 4 epilogue warps; one loader warp (footprint boxes).
#include "umma.cuh"
#include "tma_host.h"
#include <stdlib.h>

namespace cnb {
namespace {

constexpr int BM = 128;
constexpr int NPROD_WARPS = 16;
constexpr int NPROD = NPROD_WARPS * 32;
constexpr int NSETUP_WARPS = 4;
constexpr int NSETUP = NSETUP_WARPS * 32;
constexpr int W_SETUP0 = NPROD_WARPS;                // warps 16..19
constexpr int W_MMA = W_SETUP0 + NSETUP_WARPS;       // warp 20
constexpr int W_EPI0 = W_MMA + 1;                    // warps 21..24 (TMEM lane quarters 1,2,3,0)
constexpr int W_LOAD = W_EPI0 + 4;                   // warp 25: footprint boxes
constexpr int NWARPS = 28;                           // warps 26, 27 idle: whole warpgroups for setmaxnreg
constexpr int NTHREADS = NWARPS * 32;                // 896 threads x 72 registers at launch
constexpr int NG = 2;                                // sampler groups: K block k of the CTA's stream belongs to group k % NG
constexpr int WPG = NPROD_WARPS / NG;                // warps per group
constexpr int ROW_STRIDE = BM / NG;                  // a thread owns the rows r0 and r0 + 64
constexpr int REGS_SAMPLER = 96, REGS_OTHER = 40;    // 16 x 96 + 12 x 40 = 28 x 72
constexpr int MAX_STAGES = 6;
constexpr int NTAB = BM * 9;                         // (pixel, tap) entries per tile
constexpr int OM_CS = 32;                            // channel stride of the offset/mask map this kernel takes
constexpr u32 A_BYTES = BM * 64 * 2;                 // one K block of the sampled operand: 128 rows x 128 B

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
  int FW, FH;      // box: (TW + 2R + 3) x (TH + 2R + 3) pixels
  u32 fp_bytes;    // FW * FH * 128
  u32 fp_stride;   // fp_bytes rounded up to 1 KB
  int BN;          // = Co rounded up to 16, <= 128
  int nkb;         // 9 * Ci/64
  u32 bstage;      // bytes of the K block's weight tile inside its stage (b_bytes rounded up to 1 KB)
  int stages;
  u32 b_bytes, stage_bytes, tmem_cols, acc_stride, idesc;
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

// 4-corner blend of one 16-byte chunk (8 channels): fp32 (packed FFMA2), one rounding to bf16
template <bool BLEND_BF16>
__device__ __forceinline__ uint4 blend4(const float4 w, const uint4 c0, const uint4 c1, const uint4 c2, const uint4 c3) {
  const u32 q0[4] = {c0.x, c0.y, c0.z, c0.w}, q1[4] = {c1.x, c1.y, c1.z, c1.w}, q2[4] = {c2.x, c2.y, c2.z, c2.w},
            q3[4] = {c3.x, c3.y, c3.z, c3.w};
  u32 o[4];
  if constexpr (BLEND_BF16) {
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

template <bool BLEND_BF16>
__global__ void __launch_bounds__(NTHREADS, 1)
dcn_fp_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmX, const FArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) u64 s_full[MAX_STAGES];
  __shared__ __align__(8) u64 s_empty[MAX_STAGES];
  __shared__ __align__(8) u64 s_tfull[2];
  __shared__ __align__(8) u64 s_tempty[2];
  __shared__ __align__(8) u64 s_tabfull[2];
  __shared__ __align__(8) u64 s_tabempty[2];
  __shared__ __align__(8) u64 s_fpfull[2];
  __shared__ __align__(8) u64 s_fpempty[2];
  __shared__ __align__(8) u64 s_omfull;
  __shared__ u32 s_tmem;

  const cnb_conv_desc& d = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const u32 smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const u32 fp_s = smem_base + (u32)a.stages * a.stage_bytes;                        // two footprint boxes
  unsigned char* after_fp = smem_al + (size_t)a.stages * a.stage_bytes + 2 * (size_t)a.fp_stride;
  float4* s_tabw = reinterpret_cast<float4*>(after_fp);                              // [2][NTAB]
  u32* s_tabb = reinterpret_cast<u32*>(s_tabw + 2 * NTAB);                           // [2][NTAB]
  float* s_om = reinterpret_cast<float*>(s_tabb + 2 * NTAB);                         // [BM][OM_CS]
  float* s_scale = s_om + BM * OM_CS;
  float* s_shift = s_scale + a.BN;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&s_full[s], WPG + 1);   // the stage's sampler group + the weight tile's expect_tx
      mbar_init(&s_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_tfull[i], 1);
      mbar_init(&s_tempty[i], 4);
      mbar_init(&s_tabfull[i], NSETUP_WARPS);
      mbar_init(&s_tabempty[i], NPROD_WARPS);
      mbar_init(&s_fpfull[i], 1);
      mbar_init(&s_fpempty[i], NPROD_WARPS);
    }
    mbar_init(&s_omfull, 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmB);
  if (warp == W_LOAD && lane == 0) tma_prefetch_desc(&tmX);
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

  // Register file: the samplers keep two (pixel, tap) items in flight per thread (64 data registers); every other role
  // needs few.  Whole warpgroups trade registers: 12 warps shrink to 40, the 16 sampler warps grow to 96.
  if (warp < NPROD_WARPS) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SAMPLER));
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));
  }

  if (warp < NPROD_WARPS) {
    // =============================== samplers ===============================================================
    // Group g takes the K blocks k = g, g + 2, ... of the CTA's stream (tile-major, slab-major, tap-minor); inside a
    // group a thread owns the tile rows r0 and r0 + 64 and a PAIR of 16-byte channel chunks.  The four threads of a
    // row cover the 128 bytes of each corner pixel; odd rows read their two chunks in the opposite order, so the 8
    // lanes of one shared-memory wavefront (two rows) touch 8 different 16-byte bank groups whatever pixels they
    // sample.  Software pipeline: the table entry and the 8 corner loads of the NEXT (row, tap) item are issued before
    // the current item is blended (two register buffers, statically named A / B), so the shared-memory latency hides
    // behind ~110 blend instructions of the same warp instead of needing other warps to be ready.
    const int grp = warp / WPG;
    const int ltid = tid - grp * WPG * 32;
    const int cp = ltid & 3;
    const int row0 = ltid >> 2;                        // rows row0, row0 + 64: same parity
    const u32 odd = (u32)(row0 & 1);
    const u32 c_first = (u32)(2 * cp) + odd, c_second = (u32)(2 * cp + 1) - odd;
    const u32 cs2 = (u32)d.x_cstride * 2u;             // bytes per pixel of the global tensor
    const u32 rowb = (u32)d.Wi * cs2;
    const u32 boxrow = (u32)a.FW * 128u;
    const u64 xg = reinterpret_cast<u64>(a.x + d.x_coffset);
    const bool issuer = (warp == grp * WPG);
    // swizzled destinations of this thread's two chunks inside a stage, for its two rows ((row + 64) & 7 == row & 7)
    const u32 dst_lo0 = (u32)row0 * 128u + ((u32)((2 * cp) ^ (row0 & 7)) << 4);
    const u32 dst_hi0 = (u32)row0 * 128u + ((u32)((2 * cp + 1) ^ (row0 & 7)) << 4);

    struct Item {
      float4 w;
      uint4 q[8];      // corners 0..3 of chunk c_first, then of chunk c_second
      u32 b;
      bool pre;        // corners already loaded (every lane of the warp samples inside the staged box)
    };
    // table entry + (if the whole warp is inside the box) the 8 shared-memory loads of one item
    auto fetch = [&](Item& it, int row, int tap, u32 tb, u32 fpb) {
      it.w = s_tabw[tb * NTAB + row * 9 + tap];
      it.b = s_tabb[tb * NTAB + row * 9 + tap];
      it.pre = __all_sync(0xffffffffu, (int)(it.b >> 31) == 0) != 0;
      if (it.pre) {
        const u32 a0 = fpb + (it.b & 0x1FFFFFFFu) * 128u;
        const u32 adr[4] = {a0, a0 + 128u, a0 + boxrow, a0 + boxrow + 128u};
#pragma unroll
        for (int c = 0; c < 4; ++c) it.q[c] = lds128(adr[c] + (c_first << 4));
#pragma unroll
        for (int c = 0; c < 4; ++c) it.q[4 + c] = lds128(adr[c] + (c_second << 4));
      }
    };
    // blend + store into the stage; items with a corner outside the box load here, each lane from its own window
    auto finish = [&](Item& it, u32 sa_row, u32 fpb, u64 xgs) {
      if (!it.pre) {
        const bool g = (it.b >> 31) != 0;
        const u32 pix = it.b & 0x1FFFFFFFu;
        const u64 base = g ? xgs + (u64)pix * cs2 : (u64)__cvta_shared_to_generic(fpb) + (u64)pix * 128u;
        const u32 d1 = g ? (((it.b >> 29) & 1u) ? cs2 : 0u) : 128u;
        const u32 d2 = g ? (((it.b >> 30) & 1u) ? rowb : 0u) : boxrow;
        const u64 adr[4] = {base, base + d1, base + d2, base + d2 + d1};
#pragma unroll
        for (int c = 0; c < 4; ++c) it.q[c] = ldgen128(adr[c] + (c_first << 4));
#pragma unroll
        for (int c = 0; c < 4; ++c) it.q[4 + c] = ldgen128(adr[c] + (c_second << 4));
      }
      const uint4 oa = blend4<BLEND_BF16>(it.w, it.q[0], it.q[1], it.q[2], it.q[3]);
      const uint4 ob = blend4<BLEND_BF16>(it.w, it.q[4], it.q[5], it.q[6], it.q[7]);
      // even chunk first (the two rows of a wavefront then differ in bit 0 of the swizzled chunk index: conflict-free)
      const uint4 lo = odd ? ob : oa, hi = odd ? oa : ob;
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sa_row + dst_lo0), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sa_row + dst_hi0), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
    };

    u32 s = (u32)grp % (u32)a.stages, ph = ((u32)grp / (u32)a.stages) & 1u, t = 0, u = 0;
    int tap = grp;
    Item A, B;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const u32 tb = t & 1u;
      mbar_wait_parked(&s_tabfull[tb], (t >> 1) & 1u);
      for (int slab = 0; slab < nslabs; ++slab, ++u) {
        // this group's taps of the (tile, slab) box: tap, tap + 2, ... < 9
        const u32 fb = u & 1u;
        const u32 fpb = fp_s + fb * a.fp_stride;
        const u64 xgs = xg + (u64)(slab * 128);
        mbar_wait_parked(&s_fpfull[fb], (u >> 1) & 1u);
        fetch(A, row0, tap, tb, fpb);
        while (tap < 9) {
          mbar_wait_parked(&s_empty[s], ph ^ 1u);
          const u32 sa = smem_base + s * a.stage_bytes;
          if (issuer && elect_one()) {
            mbar_expect_tx(&s_full[s], a.b_bytes);
            tma_load_2d(sa + A_BYTES, &tmB, tap * d.Ci + slab * 64, 0, &s_full[s]);
          }
          fetch(B, row0 + ROW_STRIDE, tap, tb, fpb);
          finish(A, sa, fpb, xgs);
          const int tap_next = tap + NG;
          if (tap_next < 9) fetch(A, row0, tap_next, tb, fpb);
          finish(B, sa + (u32)ROW_STRIDE * 128u, fpb, xgs);
          fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_full[s]);
          s += (u32)NG;                   // next K block of this group: NG further in the stream and in the ring
          if (s >= (u32)a.stages) {
            s -= (u32)a.stages;
            ph ^= 1u;
          }
          tap = tap_next;
        }
        tap -= 9;
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_fpempty[fb]);   // the group leaves this (tile, slab) box
      }
      if (lane == 0) mbar_arrive(&s_tabempty[tb]);    // ... and this tile
    }
  } else if (warp < W_MMA) {
    // =============================== setup: (pixel, tap) -> weights + corner index ============================
    // The tile's offset/mask values (TH runs of TW pixels x 32 floats) are staged by 1-D bulk copies; the single
    // buffer is refilled for tile t+1 as soon as the table of tile t is written, long before it is needed.
    const int stid = tid - W_SETUP0 * 32;
    const int TH = BM >> a.tw_shift;
    auto issue_om = [&](int tile) {
      const TileXY tc = tile_xy(a, tile);
      const int rows = min(TH, d.Hi - tc.y0);
      const u32 run = (u32)TW * OM_CS * 4u;
      mbar_expect_tx(&s_omfull, (u32)rows * run);
      for (int r = 0; r < rows; ++r)
        bulk_g2s(s_om + r * TW * OM_CS, a.om + ((size_t)(tc.n * d.Hi + tc.y0 + r) * d.Wi + tc.x0) * OM_CS, run, &s_omfull);
    };
    if (stid == 0 && tile_begin < tile_end) issue_om(tile_begin);
    u32 t = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const u32 tb = t & 1u;
      mbar_wait_parked(&s_tabempty[tb], ((t >> 1) & 1u) ^ 1u);
      mbar_wait_parked(&s_omfull, t & 1u);
      const TileXY tc = tile_xy(a, tile);
      const int bx0 = tc.x0 - 1 - a.R, by0 = tc.y0 - 1 - a.R;   // box origin (may be negative: zero-filled)
#pragma unroll 3
      for (int item = stid; item < NTAB; item += NSETUP) {
        const int r = item / 9, tap = item - r * 9;
        const int oy = tc.y0 + (r >> a.tw_shift), ox = tc.x0 + (r & (TW - 1));
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        u32 b = 0;
        if (oy < d.Hi) {
          const int kh = tap / 3, kw = tap - 3 * kh;
          const float* omp = s_om + r * OM_CS;
          const float dy = omp[2 * tap];
          const float dx = omp[2 * tap + 1];
          const float mk = 1.f / (1.f + __expf(-omp[18 + tap]));
          const float py = (float)(oy - 1 + kh) + dy;
          const float px = (float)(ox - 1 + kw) + dx;
          if (py > -1.f && px > -1.f && py < (float)d.Hi && px < (float)d.Wi) {
            const int y0 = (int)floorf(py), x0 = (int)floorf(px);
            const float ly = py - (float)y0, lx = px - (float)x0;
            const float hy = 1.f - ly, hx = 1.f - lx;
            const int by = y0 - by0, bx = x0 - bx0;
            if (by >= 0 && bx >= 0 && by + 1 < a.FH && bx + 1 < a.FW) {
              // all four corners inside the box; corners outside the image read the zero fill
              w = make_float4(hy * hx * mk, hy * lx * mk, ly * hx * mk, ly * lx * mk);
              b = (u32)(by * a.FW + bx);
            } else {
              const bool vy0 = y0 >= 0, vy1 = y0 + 1 <= d.Hi - 1, vx0 = x0 >= 0, vx1 = x0 + 1 <= d.Wi - 1;
              w.x = (vy0 && vx0) ? hy * hx * mk : 0.f;
              w.y = (vy0 && vx1) ? hy * lx * mk : 0.f;
              w.z = (vy1 && vx0) ? ly * hx * mk : 0.f;
              w.w = (vy1 && vx1) ? ly * lx * mk : 0.f;
              const int y0c = max(y0, 0), y1c = min(y0 + 1, d.Hi - 1);
              const int x0c = max(x0, 0), x1c = min(x0 + 1, d.Wi - 1);
              b = 0x80000000u | ((u32)(y1c - y0c) << 30) | ((u32)(x1c - x0c) << 29) |
                  (u32)((tc.n * d.Hi + y0c) * d.Wi + x0c);
            }
          }
        }
        s_tabw[tb * NTAB + item] = w;
        s_tabb[tb * NTAB + item] = b;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_tabfull[tb]);
      // all setup warps are done with the om buffer: refill it for the next tile
      asm volatile("bar.sync 1, %0;" ::"n"(NSETUP) : "memory");
      if (stid == 0 && tile + 1 < tile_end) issue_om(tile + 1);
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer ==============================================================
    u32 s = 0, ph = 0, t = 0;
    const u64 da0 = make_sdesc(smem_base, 16, 1024, 2);
    const u64 db0 = make_sdesc(smem_base + A_BYTES, 16, 1024, 2);
    const u32 stage16 = a.stage_bytes >> 4;
    u32 soff16 = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const u32 acc = t & 1u, acc_ph = (t >> 1) & 1u;
      mbar_wait_parked(&s_tempty[acc], acc_ph ^ 1u);
      tc_fence_after();
      const u32 tmem_d = tmem_base + acc * a.acc_stride;
      u32 accumulate = 0;
      for (int kb = 0; kb < a.nkb; ++kb) {
        mbar_wait_parked(&s_full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const u64 da = da0 + (u64)soff16, db = db0 + (u64)soff16;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // +32 bytes of K inside the swizzle atom
            umma_bf16(tmem_d, da + (u64)(2 * kk), db + (u64)(2 * kk), a.idesc, kk == 0 ? accumulate : 1u);
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
      }
      if (elect_one()) umma_commit(&s_tfull[acc]);
      __syncwarp();
    }
  } else if (warp < W_LOAD) {
    // =============================== epilogue ================================================================
    const int q = warp & 3;
    u32 t = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const u32 acc = t & 1u, acc_ph = (t >> 1) & 1u;
      const TileXY tc = tile_xy(a, tile);
      const int r = 32 * q + lane;
      const int oy = tc.y0 + (r >> a.tw_shift), ox = tc.x0 + (r & (TW - 1));
      const int m = (tc.n * d.Hi + oy) * d.Wi + ox;
      mbar_wait_parked(&s_tfull[acc], acc_ph);
      tc_fence_after();
      const u32 taddr = tmem_base + acc * a.acc_stride + ((u32)(32 * q) << 16);
      const int ngroups = a.BN / 16;
      for (int g = 0; g < ngroups; ++g) {
        u32 v[16];
        tmem_ld16_nowait(taddr + (u32)(g * 16), v);
        tmem_ld_wait();
        const int co0 = g * 16;
        if (oy < d.Hi && co0 < d.Co) epilogue_store(a, s_scale, s_shift, v, m, co0, co0, d.Hi * d.Wi, 0, 0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_tempty[acc]);
    }
  } else if (warp == W_LOAD) {
    // =============================== loader: one box per (tile, slab), two in flight ===========================
    u32 u = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const TileXY tc = tile_xy(a, tile);
      for (int slab = 0; slab < nslabs; ++slab, ++u) {
        const u32 fb = u & 1u;
        mbar_wait_parked(&s_fpempty[fb], ((u >> 1) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&s_fpfull[fb], a.fp_bytes);
          tma_load_tile_4d(fp_s + fb * a.fp_stride, &tmX, slab * 64, tc.x0 - 1 - a.R, tc.y0 - 1 - a.R, tc.n, &s_fpfull[fb]);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, a.tmem_cols);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct Plan {
  int tw_shift, R, stages, BN;
  u32 bstage, fp_bytes, fp_stride;
  size_t smem;
};

// (reach R, ring depth) in order of preference: three stages before a third pixel of reach, reach 2 before depth
bool make_plan(const cnb_conv_desc* d, Plan* p) {
  static const int env_r = [] { const char* e = getenv("CNB_DCN_REACH"); return e ? atoi(e) : 0; }();
  static const int env_stages = [] { const char* e = getenv("CNB_DCN_STAGES"); return e ? atoi(e) : 0; }();
  p->tw_shift = d->Wi % 16 == 0 ? 4 : 3;
  const int TW = 1 << p->tw_shift, TH = BM >> p->tw_shift;
  p->BN = round_up(d->Co, 16);
  p->bstage = ((u32)p->BN * 128u + 1023u) & ~1023u;
  const size_t stage = A_BYTES + p->bstage;
  const size_t fixed = (size_t)2 * NTAB * (sizeof(float4) + sizeof(u32)) + (size_t)BM * OM_CS * 4 + (size_t)p->BN * 8 + 1024;
  const size_t budget = 226 * 1024;
  static const int pref[][2] = {{3, 3}, {2, 3}, {3, 2}, {2, 2}, {1, 3}, {1, 2}};
  for (const auto& c : pref) {
    const int R = env_r > 0 ? env_r : c[0];
    const int st = env_stages > 0 ? (env_stages > MAX_STAGES ? MAX_STAGES : env_stages) : c[1];
    const int FW = TW + 2 * R + 3, FH = TH + 2 * R + 3;
    const u32 fpb = (u32)FW * FH * 128u;
    const u32 fps = (fpb + 1023u) & ~1023u;
    const size_t smem = fixed + 2 * (size_t)fps + (size_t)st * stage;
    if (smem <= budget && st >= 2 && R >= 1) {
      p->R = R; p->stages = st; p->fp_bytes = fpb; p->fp_stride = fps; p->smem = smem;
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
  return !off && om_cstride == OM_CS && d->Ci % 64 == 0 && d->Co % 8 == 0 && round_up(d->Co, 16) <= 128 && d->Wi % 8 == 0 &&
         (long long)d->B * d->Hi * d->Wi < (1ll << 29) &&
         d->x_cstride % 8 == 0 && d->x_coffset % 8 == 0 && make_plan(d, &p);
}

int dcn_fp_run(const cnb_conv_desc* d, const void* x, const float* om, int om_cstride, const void* wpk,
               const float* scale, const float* shift, void* y, cudaStream_t st) {
  TmaDriver& drv = tma_driver();
  if (!drv.ok) {
    set_error("dcnv2: cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    return CNB_ERR_CUDA;
  }
  Plan p;
  if (!make_plan(d, &p)) {
    set_error("dcnv2: no shared-memory plan for Co=%d", d->Co);
    return CNB_ERR_INVALID;
  }
  FArgs a;
  a.d = *d;
  a.x = (const __nv_bfloat16*)x;
  a.om = om;
  a.scale = scale;
  a.shift = shift;
  a.res = nullptr;
  a.y = y;
  a.tw_shift = p.tw_shift;
  const int TW = 1 << p.tw_shift, TH = BM >> p.tw_shift;
  a.tiles_x = d->Wi / TW;
  a.tiles_y = (d->Hi + TH - 1) / TH;
  a.m_tiles = d->B * a.tiles_x * a.tiles_y;
  a.R = p.R;
  a.FW = TW + 2 * p.R + 3;
  a.FH = TH + 2 * p.R + 3;
  a.fp_bytes = p.fp_bytes;
  a.fp_stride = p.fp_stride;
  a.BN = p.BN;
  a.nkb = 9 * (d->Ci / 64);
  a.b_bytes = (u32)a.BN * 128u;
  a.bstage = p.bstage;
  a.stage_bytes = A_BYTES + a.bstage;
  a.stages = p.stages;
  a.acc_stride = (u32)round_up(a.BN, 32);
  a.tmem_cols = 32;
  while (a.tmem_cols < 2 * a.acc_stride) a.tmem_cols <<= 1;
  a.idesc = make_idesc_bf16(BM, a.BN);

  CUtensorMap tmB, tmX;
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
    // the input as {C, W, H, N}; one box = 64 channels x FW x FH pixels, 128 bytes per pixel in shared memory,
    // out-of-image pixels zero-filled
    const cuuint64_t cs2 = (cuuint64_t)d->x_cstride * 2;
    cuuint64_t dims[4] = {(cuuint64_t)d->Ci, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {cs2, cs2 * d->Wi, cs2 * d->Wi * d->Hi};
    cuuint32_t box[4] = {64u, (cuuint32_t)a.FW, (cuuint32_t)a.FH, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = drv.tiled(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)((const __nv_bfloat16*)x + d->x_coffset), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("dcnv2: cuTensorMapEncodeTiled (input box %dx%d) failed (%d)", a.FW, a.FH, (int)r);
      return CNB_ERR_CUDA;
    }
  }
  static const bool blend_bf16 = [] { const char* e = getenv("CNB_DCN_BLEND"); return e && e[0] == 'b'; }();
  static PerDeviceOnce once;
  if (once.need()) {
    CNB_CUDA(cudaFuncSetAttribute(dcn_fp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    CNB_CUDA(cudaFuncSetAttribute(dcn_fp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    once.mark();
  }
  const int grid = a.m_tiles < drv.num_sms ? a.m_tiles : drv.num_sms;
  if (blend_bf16)
    CNB_CUDA(launch_pdl(dcn_fp_kernel<true>, dim3(grid), dim3(NTHREADS), p.smem, st, tmB, tmX, a));
  else
    CNB_CUDA(launch_pdl(dcn_fp_kernel<false>, dim3(grid), dim3(NTHREADS), p.smem, st, tmB, tmX, a));
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

}  // namespace cnb
