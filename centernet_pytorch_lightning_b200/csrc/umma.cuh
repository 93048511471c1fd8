// tcgen05 / TMEM / TMA PTX wrappers shared by the tensor-core kernels (sm_100a only).
#pragma once
#include "cnb_common.cuh"

namespace cnb {

// ---- TMEM allocation (one full warp) ------------------------------------------------------------------
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
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, single CTA, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(void* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_commit_a(u32 bar_addr) {   // ... on a precomputed shared-memory address
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
// the same, arriving on the barrier at this shared-memory offset in EVERY CTA of the cluster named by cta_mask
__device__ __forceinline__ void umma_commit_mc(void* bar, unsigned short cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (lane = TMEM lane = tile row)
__device__ __forceinline__ void tmem_ld16_nowait(u32 taddr, u32 (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// the same for 64 consecutive columns in ONE instruction.  Every tcgen05.ld, whatever its width, holds up the MMA
// pipe for a few tens of clocks (measured: the K blocks issued while an epilogue runs are ~40 clk slower per
// concurrent load), so epilogues that overlap a main loop read TMEM in as few, as wide loads as registers allow.
__device__ __forceinline__ void tmem_ld64_nowait(u32 taddr, u32 (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,"
      "%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
        "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
        "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
        "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA pairs (cta_group::2): two CTAs of a cluster on neighbouring SMs run ONE 256-row MMA ---------------------
// Each CTA holds its 128 rows of A and HALF of the B tile at the same shared-memory offsets and receives its 128 rows
// of D in its own tensor memory; only the leader (cluster rank 0) issues tcgen05.mma / tcgen05.commit.  Per SM a K block
// then costs the shared-memory traffic of A + B/2 instead of A + B -- the measured limiter of the wide-N layers.
__device__ __forceinline__ void tmem_alloc_pair(u32* smem_dst, u32 ncols) {   // the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(u32 taddr, u32 ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in both CTAs of the pair once the pair's MMAs have completed
__device__ __forceinline__ void umma_commit_pair(void* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((unsigned short)3)
      : "memory");
}
// shared-memory address of `bar` in the pair's leader CTA (its shared::cluster address with the peer bit cleared)
__device__ __forceinline__ u32 pair_leader_addr(const void* bar) { return smem_u32(bar) & 0xFEFFFFFFu; }
// TMA loads issued by either CTA of a pair: data into the issuing CTA's shared memory, transaction bytes onto the
// LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(u32 dst_smem, const void* tmap, int c0, int c1, void* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst_smem),
      "l"(reinterpret_cast<u64>(tmap)), "r"(pair_leader_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(u32 dst_smem, const void* tmap, int c, int w, int h, int n,
                                                        unsigned short off_w, unsigned short off_h, void* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst_smem),
      "l"(reinterpret_cast<u64>(tmap)), "r"(pair_leader_addr(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h)
      : "memory");
}
// mbarrier arrive on the barrier at this shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(void* bar, u32 rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}

// ---- shared-memory matrix descriptors (K-major operands) ----------------------------------------------
// layout_type: 2 = SWIZZLE_128B (rows of 128 B, 8-row groups 1024 B apart), 4 = SWIZZLE_64B (64 B rows, 512 B
// groups), 6 = SWIZZLE_32B (32 B rows, 256 B groups), 0 = no swizzle (core matrix = 8 rows x 16 B contiguous;
// SBO between 8-row groups, LBO between the two 8-element K halves of one K=16 instruction).
__device__ __forceinline__ u64 make_sdesc(u32 smem_addr, u32 lbo_bytes, u32 sbo_bytes, u32 layout_type) {
  return (u64)((smem_addr >> 4) & 0x3FFF) | ((u64)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((u64)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((u64)layout_type << 61);
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M x N tile
__host__ __device__ __forceinline__ u32 make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(N >> 3) << 17) | ((u32)(M >> 4) << 24);
}

// ---- TMA (tensor maps) ----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<u64>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(u32 dst_smem, const void* tmap, int c0, int c1, void* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst_smem),
      "l"(reinterpret_cast<u64>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// multicast: the tile lands at the same shared-memory offset of every CTA in cta_mask and completes tx bytes on
// the barrier at the same offset there
__device__ __forceinline__ void tma_load_2d_mc(u32 dst_smem, const void* tmap, int c0, int c1, void* bar,
                                               unsigned short cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst_smem),
      "l"(reinterpret_cast<u64>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ u32 cluster_ctarank() {
  u32 r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ u32 cluster_id_x() {
  u32 r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ u32 cluster_count_x() {
  u32 r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// im2col mode over an NHWC tensor {C, W, H, N}: (c, w, h, n) = first channel and the base pixel of the first
// row of the column tile (already shifted by the lower corner), (off_w, off_h) = filter tap offset.
__device__ __forceinline__ void tma_load_im2col_4d(u32 dst_smem, const void* tmap, int c, int w, int h, int n,
                                                   unsigned short off_w, unsigned short off_h, void* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst_smem),
      "l"(reinterpret_cast<u64>(tmap)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h)
      : "memory");
}

__device__ __forceinline__ u32 pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<u32*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(u32 v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}

// ---- shared epilogue of the tensor-core kernels ------------------------------------------------------------
// 16 consecutive output channels [co0, co0+16) of output pixel m: accumulator -> scale/shift (folded BN or
// bias) -> (+residual) -> ReLU/sigmoid -> NHWC bf16 (optionally a concat slice) | NCHW fp32 | NHWC fp32.
// Args needs: cnb_conv_desc d; const __nv_bfloat16* res; void* y.   cg0 indexes s_scale/s_shift.
template <class Args>
__device__ __forceinline__ void epilogue_store(const Args& a, const float* s_scale, const float* s_shift,
                                               u32 (&v)[16], int m, int cg0, int co0, int HoWo, int on,
                                               int opix) {
  const cnb_conv_desc& d = a.d;
  float f[16];
  {  // scale/shift as 8 LDS.128 (cg0 % 16 == 0, arrays 16-byte aligned) instead of 32 scalar loads: the shared-memory
     // pipe is also carrying the producers' cp.async traffic, and the scalar form made this line the epilogue's
     // dominant stall
    const float4* sc4 = reinterpret_cast<const float4*>(s_scale + cg0);
    const float4* sh4 = reinterpret_cast<const float4*>(s_shift + cg0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 sc = sc4[q], sh = sh4[q];
      f[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), sc.x, sh.x);
      f[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), sc.y, sh.y);
      f[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), sc.z, sh.z);
      f[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), sc.w, sh.w);
    }
  }
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
    for (int j = 0; j < 16; ++j) f[j] = __fdividef(1.f, 1.f + __expf(-f[j]));   // MUFU.EX2 + MUFU.RCP, ~2 ulp
  }
  if (d.out_nchw_f32 == 0) {          // NHWC bf16 (optionally a channel slice of a concat buffer)
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y) + (size_t)m * d.y_cstride + d.y_coffset + co0;
    u32 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
    if (co0 + 8 < d.Co && ((d.y_cstride | d.y_coffset) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 31) == 0) {
      // all 16 channels, 32-byte aligned: one 256-bit store = one whole sector per lane (two 128-bit stores put
      // two half-sector writes on the crossbar)
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(yp), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                   "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                   : "memory");
    } else {
      *reinterpret_cast<uint4*>(yp) = make_uint4(o[0], o[1], o[2], o[3]);
      if (co0 + 8 < d.Co) *reinterpret_cast<uint4*>(yp + 8) = make_uint4(o[4], o[5], o[6], o[7]);
    }
  } else if (d.out_nchw_f32 == 1) {   // NCHW fp32 (head maps for decode / losses): lanes = consecutive pixels
    float* yp = reinterpret_cast<float*>(a.y) + ((size_t)on * d.Co + co0) * HoWo + opix;
    const int nj = min(16, d.Co - co0);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < nj) *yp = f[j];
      yp += HoWo;
    }
  } else {                            // NHWC fp32 (offset/mask maps feeding the DCN sampler)
    float* yp = reinterpret_cast<float*>(a.y) + (size_t)m * d.y_cstride + d.y_coffset + co0;
    if (co0 + 16 <= d.y_cstride && (d.y_cstride & 7) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 31) == 0) {
#pragma unroll
      for (int h = 0; h < 2; ++h)   // 256-bit stores: whole sectors
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(yp + 8 * h), "f"(f[8 * h]),
                     "f"(f[8 * h + 1]), "f"(f[8 * h + 2]), "f"(f[8 * h + 3]), "f"(f[8 * h + 4]), "f"(f[8 * h + 5]),
                     "f"(f[8 * h + 6]), "f"(f[8 * h + 7])
                     : "memory");
    } else {
#pragma unroll
      for (int h = 0; h < 4; ++h)
        if (co0 + 4 * h < d.y_cstride)
          *reinterpret_cast<float4*>(yp + 4 * h) = make_float4(f[4 * h], f[4 * h + 1], f[4 * h + 2], f[4 * h + 3]);
    }
  }
}

}  // namespace cnb
