// Modulated deformable convolution v2 (DCNv2), warp-specialised and persistent, on the 5th-gen tensor cores.
//
// Replaces DCN.dcn_v2.DCN forward (external tteepe/DCNv2; call sites CenterNet/models/backbones/
// pose_dla_dcn.py:441-449 and resnet_dcn.py:202-210): 3x3, stride 1, pad 1, dil 1, deformable_groups 1,
//   y[n,co,p] = b[co] + sum_{ci,k} W[co,ci,k] * sigmoid(m_k(p)) * bilinear(x[n,ci], p + tap_k + offset_k(p))
// with the (+ BatchNorm(eval) + ReLU) of DeformConv (pose_dla_dcn.py:435-454) folded into the epilogue.
//
// GEMM view: D[M = B*H*W pixels, N = Co] = A[M, K = 9*Ci] * W[N, K]^T where A is the *sampled* column matrix.
// A is never materialised in HBM.  Per 128-pixel tile the CTA's warps play four roles:
//   setup warps    : read the 27 offset/mask channels of the tile's pixels once, and turn every (pixel, tap)
//                    into 4 bilinear weights (mask and validity folded in) + a clamped corner address; the table
//                    lives in shared memory, double buffered, so tile i+1 is prepared while tile i is sampled;
//   sampler warps  : per (tap, 64-channel slab) K block gather the 4 corners of each pixel as 16-byte channel
//                    chunks through L1 (neighbouring pixels/taps share corners), blend in fp32, round to bf16 and
//                    st.shared the chunk straight into the K-major SWIZZLE_128B operand layout;
//   MMA warp       : one lane issues tcgen05.mma (M=128, N=Co, K=16) from the stage ring; the weight tile of the
//                    K block arrives by TMA on the same full barrier; tcgen05.commit recycles the stage;
//   epilogue warps : TMEM -> scale/shift (bias or folded BN) -> ReLU -> NHWC bf16; two TMEM accumulators so the
//                    epilogue of tile i overlaps the main loop of tile i+1.
// K blocks run slab-major / tap-minor: the 9 taps of one slab re-touch the same ~(rows+2) x W x 128 B footprint,
// which stays in L1.  The kernel is bound by the sampler (L1 wavefronts + fp32 blend issue), not by the MMA.
#include "umma.cuh"
#include "tma_host.h"
#include <stdlib.h>

namespace cnb {
namespace {

constexpr int BM = 128;
constexpr int NPROD_WARPS = 16;                      // sampler warps: 512 threads = 128 pixel rows x 4 pairs of 8-channel chunks
constexpr int NPROD = NPROD_WARPS * 32;
constexpr int NSETUP_WARPS = 4;
constexpr int NSETUP = NSETUP_WARPS * 32;
constexpr int W_SETUP0 = NPROD_WARPS;                // warps 16..19
constexpr int W_MMA = W_SETUP0 + NSETUP_WARPS;       // warp 20
constexpr int W_EPI0 = W_MMA + 1;                    // warps 21..24 (TMEM lane quarters 1,2,3,0)
constexpr int NTHREADS = (W_EPI0 + 4) * 32;          // 800
static_assert(NPROD == BM * 4, "one sampler thread per (pixel row, 16-channel chunk pair)");
constexpr int MAX_STAGES = 6;
constexpr int NTAB = BM * 9;                         // (pixel, tap) entries per tile
constexpr int OM_CS = 32;                            // channel stride of the offset/mask map this kernel takes
constexpr u32 A_BYTES = BM * 64 * 2;                 // one K block of the sampled operand: 128 rows x 128 B

struct DArgs {
  cnb_conv_desc d;
  const __nv_bfloat16* x;
  const float* om;
  int om_cstride;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* res;   // unused (kept for the shared epilogue)
  void* y;
  int M, m_tiles;
  int BN;          // = Co rounded up to 16, <= 256
  int nkb;         // 9 * Ci/64
  int ngroups;     // sampler groups (1, 2 or 4): K block k of the CTA's stream is sampled by group k % ngroups
  u32 bstage;      // bytes of the K block's weight tile inside its stage (b_bytes rounded up to 1 KB)
  int stages;
  u32 b_bytes, stage_bytes, tmem_cols, acc_stride, idesc;
  long long* trace;  // CNB_DCN_TRACE: clock stamps of CTA 0 (see dcn_ws_run)
  int debug;       // CNB_DCN_DEBUG (timing experiments): 1 no corner loads, 2 no blend/store, 4 no MMAs, 8 no proxy fence, 16 no table math
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

// BLEND_BF16: the 4-corner blend runs as packed bf16x2 FMAs straight on the loaded pairs (weights rounded to
// bf16, three more bf16 roundings than the fp32 blend) -- about half the sampler's instructions.  Off unless
// CNB_DCN_BLEND=bf16; the default blends in fp32 and rounds once.
// The timing experiments (CNB_DCN_DEBUG stage skipping, CNB_DCN_TRACE clock stamps) are compiled in only with
// -DCNB_DCN_EXPERIMENTS: their predicates sit in the latency-bound role loops.
#ifdef CNB_DCN_EXPERIMENTS
#define DCN_DBG(a) ((a).debug)
#define DCN_STAMP(slot, idx) \
  do { if (a.trace && blockIdx.x == 0 && (idx) < 512) a.trace[(slot) * 512 + (idx)] = clock64(); } while (0)
#else
#define DCN_DBG(a) 0
#define DCN_STAMP(slot, idx) do { } while (0)
#endif

template <bool BLEND_BF16>
__global__ void __launch_bounds__(NTHREADS, 1)
dcn_ws_kernel(const __grid_constant__ CUtensorMap tmB, const DArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) u64 s_full[MAX_STAGES];
  __shared__ __align__(8) u64 s_empty[MAX_STAGES];
  __shared__ __align__(8) u64 s_tfull[2];
  __shared__ __align__(8) u64 s_tempty[2];
  __shared__ __align__(8) u64 s_tabfull[2];
  __shared__ __align__(8) u64 s_tabempty[2];
  __shared__ __align__(8) u64 s_omfull[2];
  __shared__ u32 s_tmem;

  const cnb_conv_desc& d = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const u32 smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  float4* s_tabw = reinterpret_cast<float4*>(smem_al + (size_t)a.stages * a.stage_bytes);   // [2][NTAB]
  u32* s_tabb = reinterpret_cast<u32*>(s_tabw + 2 * NTAB);                                  // [2][NTAB]
  float* s_om = reinterpret_cast<float*>(s_tabb + 2 * NTAB);                                // [2][BM][OM_CS]
  float* s_scale = s_om + 2 * BM * OM_CS;
  float* s_shift = s_scale + a.BN;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&s_full[s], NPROD_WARPS / a.ngroups + 1);   // one arrival per sampler warp of the stage's group + the weight tile's expect_tx
      mbar_init(&s_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_tfull[i], 1);
      mbar_init(&s_tempty[i], 4);
      mbar_init(&s_tabfull[i], NSETUP);     // every setup thread releases the entries it wrote ...
      mbar_init(&s_tabempty[i], NPROD);     // ... and every sampler thread the rows it read
      mbar_init(&s_omfull[i], 1);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmB);
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
  const int HW = d.Hi * d.Wi;
  // contiguous tile range per CTA: consecutive tiles are neighbouring image rows, whose sampling footprints
  // overlap in L1 / L2
  const int tile_begin = (int)((long long)blockIdx.x * a.m_tiles / gridDim.x);
  const int tile_end = (int)((long long)(blockIdx.x + 1) * a.m_tiles / gridDim.x);

  if (warp < NPROD_WARPS) {
    // =============================== samplers ===============================================================
    // The 16 sampler warps form `ngroups` independent groups; K block k of the CTA's stream (tile-major, then
    // slab-major / tap-minor) belongs to group k % ngroups and to ring stage k % stages, so the groups run out of
    // phase: while one group waits for its gathers the other blends (one lock-stepped group leaves L1 idle while it
    // blends and the issue slots idle while it loads).  Inside a group a thread owns the pixel rows
    // r0, r0 + 128/ngroups, ... and a PAIR of 16-byte channel chunks (4 threads cover the 128-byte line of one
    // corner: one L1 wavefront per pixel); each corner is one 256-bit load when the tensor allows it.
    const int ng = a.ngroups;
    const int wpg = NPROD_WARPS / ng;                 // warps per group
    const int grp = warp / wpg;
    const int ltid = tid - grp * wpg * 32;
    const int cp = ltid & 3;                          // chunk pair inside the 64-channel slab: chunks 2cp, 2cp+1
    const int row0 = ltid >> 2;
    const int row_stride = BM / ng;
    const u32 cs = (u32)d.x_cstride;
    const u32 row_step = (u32)d.Wi * cs;
    const bool wide_ld = ((d.x_cstride | d.x_coffset) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 31) == 0;
    const bool issuer = (warp == grp * wpg);
    const int nslabs = d.Ci >> 6;
    u32 s = (u32)grp % (u32)a.stages, ph = ((u32)grp / (u32)a.stages) & 1u, t = 0;
    int slab = 0, tap = grp;                          // ngroups <= 4 < 9
    bool have_tab = false;
    const long long kb_end = (long long)(tile_end - tile_begin) * a.nkb;
    for (long long kb = grp; kb < kb_end; kb += ng) {
      const u32 tb = t & 1u;
      if (!have_tab) {
        mbar_wait_parked(&s_tabfull[tb], (t >> 1) & 1u);
        have_tab = true;
      }
      mbar_wait_parked(&s_empty[s], ph ^ 1u);
      if (ltid == 0 && grp < 2) DCN_STAMP(2 * grp, (int)(kb / ng));
      const u32 sa = smem_base + s * a.stage_bytes;
      if (issuer && elect_one()) {   // (elected: uniform operands, no broadcast loop)
        mbar_expect_tx(&s_full[s], a.b_bytes);
        tma_load_2d(sa + A_BYTES, &tmB, tap * d.Ci + slab * 64, 0, &s_full[s]);
      }
      const __nv_bfloat16* xs = a.x + d.x_coffset + slab * 64 + cp * 16;
      for (int row = row0; row < BM; row += row_stride) {
        const u32 dst_lo = (u32)row * 128u + ((u32)((2 * cp) ^ (row & 7)) << 4);
        const u32 dst_hi = (u32)row * 128u + ((u32)((2 * cp + 1) ^ (row & 7)) << 4);
        u32 q[4][8];
        const float4 w = s_tabw[tb * NTAB + row * 9 + tap];
        const u32 b = s_tabb[tb * NTAB + row * 9 + tap];
        const u32 o00 = b & 0x3FFFFFFFu;                       // element offset of the clamped (y0, x0) pixel
        const u32 o01 = o00 + (((b >> 30) & 1u) ? cs : 0u);
        const u32 ddy = (b >> 31) ? row_step : 0u;
        const u32 off[4] = {o00, o01, o00 + ddy, o01 + ddy};
        if (DCN_DBG(a) & 1) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int e = 0; e < 8; ++e) q[c][e] = off[c] + e;
        } else if (wide_ld) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(q[c][0]), "=r"(q[c][1]), "=r"(q[c][2]), "=r"(q[c][3]), "=r"(q[c][4]), "=r"(q[c][5]),
                           "=r"(q[c][6]), "=r"(q[c][7])
                         : "l"(xs + off[c]));
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 lo = __ldg(reinterpret_cast<const uint4*>(xs + off[c]));
            const uint4 hi = __ldg(reinterpret_cast<const uint4*>(xs + off[c] + 8));
            q[c][0] = lo.x; q[c][1] = lo.y; q[c][2] = lo.z; q[c][3] = lo.w;
            q[c][4] = hi.x; q[c][5] = hi.y; q[c][6] = hi.z; q[c][7] = hi.w;
          }
        }
        if (!(DCN_DBG(a) & 2)) {
          u32 o[8];
          if constexpr (BLEND_BF16) {
            const u32 wh0 = pack_bf16x2(w.x, w.x), wh1 = pack_bf16x2(w.y, w.y), wh2 = pack_bf16x2(w.z, w.z),
                      wh3 = pack_bf16x2(w.w, w.w);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              u32 acc;
              asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(acc) : "r"(wh0), "r"(q[0][e]));
              asm("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc) : "r"(wh1), "r"(q[1][e]));
              asm("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc) : "r"(wh2), "r"(q[2][e]));
              asm("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc) : "r"(wh3), "r"(q[3][e]));
              o[e] = acc;
            }
          } else {
            const u64 w0 = dup2(w.x), w1 = dup2(w.y), w2 = dup2(w.z), w3 = dup2(w.w);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              // bf16 -> fp32 is a 16-bit shift: low element = v << 16, high element = v & 0xffff0000;
              // the (low, high) pair is blended with one packed fp32x2 FMA per corner
              u64 acc = mul2(w0, pair_from_bf16x2(q[0][e]));
              acc = fma2(w1, pair_from_bf16x2(q[1][e]), acc);
              acc = fma2(w2, pair_from_bf16x2(q[2][e]), acc);
              acc = fma2(w3, pair_from_bf16x2(q[3][e]), acc);
              float lo, hi;
              asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
              o[e] = pack_bf16x2(lo, hi);
            }
          }
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sa + dst_lo), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                       "r"(o[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sa + dst_hi), "r"(o[4]), "r"(o[5]), "r"(o[6]),
                       "r"(o[7])
                       : "memory");
        }
      }   // rows of this thread
      if (!(DCN_DBG(a) & 8)) fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full[s]);
      if (ltid == 0 && grp < 2) DCN_STAMP(2 * grp + 1, (int)(kb / ng));
      // next K block of this group: ng further in the stream and in the ring
      s += (u32)ng;
      while (s >= (u32)a.stages) {
        s -= (u32)a.stages;
        ph ^= 1u;
      }
      tap += ng;
      if (tap >= 9) {
        tap -= 9;
        if (++slab == nslabs) {   // the group leaves this tile
          slab = 0;
          __syncwarp();
          mbar_arrive(&s_tabempty[tb]);
          ++t;
          have_tab = false;
        }
      }
    }
  } else if (warp < W_MMA) {
    // =============================== setup: (pixel, tap) -> weights + corner address =========================
    // The tile's 128 x 32 offset/mask floats are one contiguous 16 KB run of the NHWC fp32 map: a 1-D bulk copy
    // (double buffered, issued two tiles ahead) stages them in shared memory, so the table math never waits on
    // global memory.
    const int stid = tid - W_SETUP0 * 32;
    auto issue_om = [&](int tile, u32 buf) {
      const int m0 = tile * BM;
      const int rows = min(BM, a.M - m0);
      const u32 bytes = (u32)rows * OM_CS * 4u;
      mbar_expect_tx(&s_omfull[buf], bytes);
      bulk_g2s(s_om + buf * (BM * OM_CS), a.om + (size_t)m0 * OM_CS, bytes, &s_omfull[buf]);
    };
    if (stid == 0) {
      if (tile_begin < tile_end) issue_om(tile_begin, 0);
      if (tile_begin + 1 < tile_end) issue_om(tile_begin + 1, 1);
    }
    u32 t = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const u32 tb = t & 1u;
      mbar_wait_parked(&s_tabempty[tb], ((t >> 1) & 1u) ^ 1u);
      mbar_wait_parked(&s_omfull[tb], (t >> 1) & 1u);
      const float* oms = s_om + tb * (BM * OM_CS);
      const int m0 = tile * BM;
#pragma unroll 3
      for (int item = stid; item < NTAB; item += NSETUP) {
        const int r = item / 9, tap = item - r * 9;
        const int m = m0 + r;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        u32 b = 0;
        if (DCN_DBG(a) & 16) {
          w = make_float4(0.25f, 0.25f, 0.25f, 0.25f);
        } else if (m < a.M) {
          const int n = m / HW;
          const int rem = m - n * HW;
          const int oy = rem / d.Wi, ox = rem - oy * d.Wi;
          const int kh = tap / 3, kw = tap - 3 * kh;
          const float* omp = oms + r * OM_CS;
          const float dy = omp[2 * tap];
          const float dx = omp[2 * tap + 1];
          const float mk = 1.f / (1.f + __expf(-omp[18 + tap]));
          const float py = (float)(oy - 1 + kh) + dy;
          const float px = (float)(ox - 1 + kw) + dx;
          if (py > -1.f && px > -1.f && py < (float)d.Hi && px < (float)d.Wi) {
            const int y0 = (int)floorf(py), x0 = (int)floorf(px);
            const float ly = py - (float)y0, lx = px - (float)x0;
            const float hy = 1.f - ly, hx = 1.f - lx;
            const bool vy0 = y0 >= 0, vy1 = y0 + 1 <= d.Hi - 1, vx0 = x0 >= 0, vx1 = x0 + 1 <= d.Wi - 1;
            w.x = (vy0 && vx0) ? hy * hx * mk : 0.f;
            w.y = (vy0 && vx1) ? hy * lx * mk : 0.f;
            w.z = (vy1 && vx0) ? ly * hx * mk : 0.f;
            w.w = (vy1 && vx1) ? ly * lx * mk : 0.f;
            const int y0c = max(y0, 0), y1c = min(y0 + 1, d.Hi - 1);
            const int x0c = max(x0, 0), x1c = min(x0 + 1, d.Wi - 1);
            b = (u32)((n * d.Hi + y0c) * d.Wi + x0c) * (u32)d.x_cstride | ((u32)(x1c - x0c) << 30) |
                ((u32)(y1c - y0c) << 31);
          }
        }
        s_tabw[tb * NTAB + item] = w;
        s_tabb[tb * NTAB + item] = b;
      }
      __syncwarp();
      mbar_arrive(&s_tabfull[tb]);
      if (stid == 0) DCN_STAMP(5, (int)t);
      // all setup warps are done with this om buffer: refill it for tile + 2
      asm volatile("bar.sync 1, %0;" ::"n"(NSETUP) : "memory");
      if (stid == 0 && tile + 2 < tile_end) issue_om(tile + 2, tb);
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer ==============================================================
    {   // all lanes walk the loop (warp-uniform operands -> uniform registers); one elected lane issues
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
          if (lane == 0) DCN_STAMP(4, (int)(t * a.nkb) + kb);
          if (elect_one()) {
            const u64 da = da0 + (u64)soff16, db = db0 + (u64)soff16;
            if (!(DCN_DBG(a) & 4)) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)   // +32 bytes of K inside the swizzle atom
                umma_bf16(tmem_d, da + (u64)(2 * kk), db + (u64)(2 * kk), a.idesc, kk == 0 ? accumulate : 1u);
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
        }
        if (elect_one()) umma_commit(&s_tfull[acc]);
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue ================================================================
    const int q = warp & 3;
    u32 t = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const u32 acc = t & 1u, acc_ph = (t >> 1) & 1u;
      const int m = tile * BM + 32 * q + lane;
      mbar_wait_parked(&s_tfull[acc], acc_ph);
      tc_fence_after();
      if (warp == W_EPI0 && lane == 0) DCN_STAMP(6, (int)t);
      const u32 taddr = tmem_base + acc * a.acc_stride + ((u32)(32 * q) << 16);
      const int ngroups = a.BN / 16;
      for (int g = 0; g < ngroups; ++g) {
        u32 v[16];
        tmem_ld16_nowait(taddr + (u32)(g * 16), v);
        tmem_ld_wait();
        const int co0 = g * 16;
        if (m < a.M && co0 < d.Co) epilogue_store(a, s_scale, s_shift, v, m, co0, co0, HW, 0, 0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_tempty[acc]);
      if (warp == W_EPI0 && lane == 0) DCN_STAMP(7, (int)t);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, a.tmem_cols);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

// Geometries this kernel covers (the rest stays on the cp.async gather kernel of conv_umma.cu).
bool dcn_ws_supported(const cnb_conv_desc* d, int om_cstride) {
  static const bool off = [] { const char* e = getenv("CNB_DCN_IMPL"); return e && e[0] == 'v' && e[1] == '1'; }();
  return !off && om_cstride == OM_CS && d->Ci % 64 == 0 && d->Co % 8 == 0 && round_up(d->Co, 16) <= 256 &&
         (long long)d->B * d->Hi * d->Wi * d->x_cstride < (1ll << 30);
}

int dcn_ws_run(const cnb_conv_desc* d, const void* x, const float* om, int om_cstride, const void* wpk,
               const float* scale, const float* shift, void* y, cudaStream_t st) {
  TmaDriver& drv = tma_driver();
  if (!drv.ok) {
    set_error("dcnv2: cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    return CNB_ERR_CUDA;
  }
  DArgs a;
  a.d = *d;
  a.x = (const __nv_bfloat16*)x;
  a.om = om;
  a.om_cstride = om_cstride;
  a.scale = scale;
  a.shift = shift;
  a.res = nullptr;
  a.y = y;
  a.M = d->B * d->Hi * d->Wi;
  a.m_tiles = (a.M + BM - 1) / BM;
  a.BN = round_up(d->Co, 16);
  a.nkb = 9 * (d->Ci / 64);
  a.b_bytes = (u32)a.BN * 128u;
  a.bstage = (a.b_bytes + 1023u) & ~1023u;
  const size_t fixed = (size_t)2 * NTAB * (sizeof(float4) + sizeof(u32)) + (size_t)2 * BM * OM_CS * 4 +
                       (size_t)a.BN * 8 + 1024;
  // One ring stage = one K block (sampled operand 16 KB + weight tile); the sampler groups take the K blocks of the
  // CTA's stream round-robin (CNB_DCN_GROUPS / CNB_DCN_STAGES override).
  static const int env_groups = [] { const char* e = getenv("CNB_DCN_GROUPS"); return e ? atoi(e) : 0; }();
  static const int env_stages = [] { const char* e = getenv("CNB_DCN_STAGES"); return e ? atoi(e) : 0; }();
  a.ngroups = env_groups == 1 || env_groups == 2 || env_groups == 4 ? env_groups : 2;
  a.stage_bytes = A_BYTES + a.bstage;
  a.stages = env_stages > 0 ? env_stages : MAX_STAGES;
  if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
  while (a.stages > 2 && (size_t)a.stages * a.stage_bytes + fixed > 220 * 1024) --a.stages;
  while (a.ngroups > a.stages) a.ngroups >>= 1;   // a group must not lap the ring (mbarrier phase parity)
  // a ring depth that is a multiple of the group count gives every stage ONE owner group: the same threads rewrite the
  // same shared-memory rows each time round (their hand-off to the tensor core is tcgen05.commit -> mbarrier, which
  // compute-sanitizer racecheck cannot see and reported as a write-after-write hazard between the groups)
  if (a.stages % a.ngroups != 0 && a.stages - a.stages % a.ngroups >= 2) a.stages -= a.stages % a.ngroups;
  a.acc_stride = (u32)round_up(a.BN, 32);
  a.tmem_cols = 32;
  while (a.tmem_cols < 2 * a.acc_stride) a.tmem_cols <<= 1;
  a.idesc = make_idesc_bf16(BM, a.BN);
  const size_t smem = (size_t)a.stages * a.stage_bytes + fixed;

  CUtensorMap tmB;
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
      set_error("dcnv2: cuTensorMapEncodeTiled failed (%d) Kpad=%d BN=%d", (int)r, Kpad, a.BN);
      return CNB_ERR_CUDA;
    }
  }
  static const int env_dbg = [] { const char* e = getenv("CNB_DCN_DEBUG"); return e ? atoi(e) : 0; }();
  a.debug = env_dbg;
  static const bool env_trace = getenv("CNB_DCN_TRACE") != nullptr;
  static long long* trace_buf = nullptr;
  a.trace = nullptr;
  if (env_trace) {
    if (!trace_buf) cudaMalloc(&trace_buf, 8 * 512 * sizeof(long long));
    cudaMemset(trace_buf, 0, 8 * 512 * sizeof(long long));
    a.trace = trace_buf;
  }
  static const bool blend_bf16 = [] { const char* e = getenv("CNB_DCN_BLEND"); return e && e[0] == 'b'; }();
  static PerDeviceOnce once;
  if (once.need()) {
    CNB_CUDA(cudaFuncSetAttribute(dcn_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    CNB_CUDA(cudaFuncSetAttribute(dcn_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    once.mark();
  }
  const int grid = a.m_tiles < drv.num_sms ? a.m_tiles : drv.num_sms;
  if (blend_bf16)
    CNB_CUDA(launch_pdl(dcn_ws_kernel<true>, dim3(grid), dim3(NTHREADS), smem, st, tmB, a));
  else
    CNB_CUDA(launch_pdl(dcn_ws_kernel<false>, dim3(grid), dim3(NTHREADS), smem, st, tmB, a));
  CNB_LAUNCH_CHECK();
  if (env_trace) {   // debugging aid: clock stamps of CTA 0, printed relative to the first sampler stamp
    static int printed = 0;
    cudaStreamSynchronize(st);
    if (printed++ == 3) {
      static long long h[8 * 512];
      cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
      const long long t0 = h[0];
      fprintf(stderr, "dcn_ws trace Ci=%d Co=%d nkb=%d stages=%d groups=%d (clocks since the first K block)\n", d->Ci, d->Co,
              a.nkb, a.stages, a.ngroups);
      for (int k = 0; k < 45 && h[k]; ++k)
        fprintf(stderr, "k=%3d group0 got_stage %7lld arrived %7lld | group1 got_stage %7lld arrived %7lld | mma got K block %d %7lld, %d %7lld\n",
                k, h[k] - t0, h[512 + k] - t0, h[1024 + k] - t0, h[1536 + k] - t0, 2 * k, h[2048 + 2 * k] - t0, 2 * k + 1,
                h[2048 + 2 * k + 1] - t0);
      for (int t = 0; t < 6 && h[2560 + t]; ++t)
        fprintf(stderr, "tile %d table_done %7lld | epilogue start %7lld done %7lld\n", t, h[2560 + t] - t0, h[3072 + t] - t0,
                h[3584 + t] - t0);
    }
  }
  return CNB_OK;
}

}  // namespace cnb
