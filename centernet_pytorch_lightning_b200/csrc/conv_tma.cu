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
//   * Layers whose weight tile dominates the L2 -> SM traffic (Ci >= 64 slabs, wide N) run as thread-block CLUSTERS of
//     2 or 4 CTAs that work on neighbouring M tiles of the same N tile: every CTA fetches 1/c of each weight slab and
//     TMA-multicasts it into all c shared memories, so a K block costs 16 KB + BN*128/c bytes of L2 bandwidth per SM
//     instead of 16 KB + BN*128.  A stage is reusable once the MMA warps of ALL c CTAs have committed to it
//     (tcgen05.commit multicast onto every CTA's empty barrier).
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
  int csize;       // CTAs per cluster (1, 2 or 4): consecutive M tiles sharing one multicast weight tile
  int bn_share;    // weight rows each CTA of the cluster fetches per slab (BN / csize)
  long long* trace;  // CNB_TMA_TRACE: clock64 stamps of CTA 0's first 1024 K blocks (producer / MMA thread)
  int debug;       // CNB_TMA_DEBUG (timing experiments only): 1 no A loads, 2 no B loads, 4 no MMAs, 8 no stores, 16 no epilogue
};

// CL: launched as clusters of a.csize > 1 CTAs (multicast weight tiles); DBG: the CNB_TMA_DEBUG / CNB_TMA_TRACE timing
// experiments (kept out of the production loops); PAIR (implies CL, csize 2): the two CTAs run ONE 256-row
// tcgen05.mma.cta_group::2 per K step -- each holds its 128 rows of A and half of the weight tile, the leader issues.
template <bool CL, bool DBG, bool PAIR = false>
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
      mbar_init(&s_empty[s], (CL && !PAIR) ? (u32)a.csize : 1u);   // one tcgen05.commit per CTA of the cluster
                                                                   // (pair: the leader's, multicast to both)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_tfull[i], 1);
      mbar_init(&s_tempty[i], PAIR ? 2 * NEPI : NEPI);   // one arrival per epilogue warp (pair: of both CTAs, on the leader)
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (PAIR) cluster_sync_all();   // both CTAs are resident before the pair allocates tensor memory
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(&s_tmem, a.tmem_cols);
    else tmem_alloc(&s_tmem, a.tmem_cols);
  }
  pdl_launch_dependents();      // the next kernel may start its prologue as soon as SMs free up ...
  pdl_wait();                   // ... and this one reads global memory (scale/shift included: they may have been
                                // produced in-stream) only after its predecessor has completed
  for (int i = tid; i < a.nscale; i += NTHREADS) {
    s_scale[i] = (i < d.Co && a.scale) ? a.scale[i] : 1.f;
    s_shift[i] = (i < d.Co && a.shift) ? a.shift[i] : 0.f;
  }
  tc_fence_before();
  if (CL) cluster_sync_all();   // peers' barriers are initialised before anything is multicast to them
  else __syncthreads();
  tc_fence_after();
  const u32 tmem_base = s_tmem;
  const int HoWo = d.Ho * d.Wo;
  // tile walk: cluster `cid` of `ncl` takes (M group, N tile) pairs; CTA `crank` of the cluster owns M tile
  // group * csize + crank (possibly past the end: it still fetches its share of the weights)
  const int dbg = DBG ? a.debug : 0;
  long long* const trace = DBG ? a.trace : nullptr;
  const int csize = CL ? a.csize : 1;
  const int crank = CL ? (int)cluster_ctarank() : 0;
  const int cid = CL ? (int)cluster_id_x() : (int)blockIdx.x;
  const int ncl = CL ? (int)cluster_count_x() : (int)gridDim.x;
  const unsigned short cmask = (unsigned short)((1u << csize) - 1u);

  // Both single-thread roles below keep every per-K-block quantity (stage index, phase, shared-memory address,
  // filter tap, channel offset, descriptors) as loop-carried state updated with adds and compares.  A single
  // thread has no other warp to hide behind: the integer divisions this loop used to do per K block (stage =
  // it % stages, tap = k0 / Ci, ...) cost more than the K block's MMAs and its TMA loads together.
  if (warp == 0) {
    // =============================== TMA producer =========================================================
    {   // every lane walks the loop (uniform control flow); one elected lane issues
      const bool ldB = !(dbg & 2);
      const u32 b_share_off = CL ? (u32)(crank * a.bn_share * a.slabW * 2) : 0u;
      u32 s = 0, ph = 0;                   // stage and its phase
      u32 sa = smem_base;                  // A area of stage s
      int kcount = 0;
      for (int tile = cid; tile < a.total_tiles; tile += ncl) {
        const int m_group = tile / a.n_tiles, n_tile = tile - m_group * a.n_tiles;
        const int m_tile = m_group * csize + crank;
        const bool valid = m_tile < a.m_tiles;
        const int m0 = (valid ? m_tile : 0) * BM;
        const int n = m0 / HoWo;
        const int rem = m0 - n * HoWo;
        const int oy = rem / d.Wo, ox = rem - oy * d.Wo;
        const int w0 = ox * d.stride - d.pad, h0 = oy * d.stride - d.pad;
        const int n0 = n_tile * a.BN + (CL ? crank * a.bn_share : 0);
        const bool ldA = valid && !(dbg & 1);
        // pair: the leader's barrier collects the bytes of BOTH CTAs (its own tile is always valid)
        const bool peer_valid = PAIR && (m_group * 2 + 1 < a.m_tiles);
        const u32 slab_tx = PAIR ? a.a_slab_bytes * (peer_valid ? 2u : 1u) + 2u * a.b_slab_bytes
                                 : (ldA ? a.a_slab_bytes : 0u) + (ldB ? a.b_slab_bytes : 0u);
        int c0 = 0, kw = 0, kh = 0, k0 = 0;     // slab cursor: channel offset, filter tap, K index
        int left = a.nslabs;
        for (int ks = 0; ks < a.nsteps; ++ks) {
          if (dbg & 128) mbar_wait_spin(&s_empty[s], ph ^ 1u); else mbar_wait_parked(&s_empty[s], ph ^ 1u);
          if (DBG && trace && blockIdx.x == 0 && kcount < 1024 && lane == 0) trace[kcount] = clock64();
          ++kcount;
          const int nsl = min(a.g, left);
          left -= nsl;
          if (elect_one() && (!PAIR || crank == 0)) mbar_expect_tx(&s_full[s], (u32)nsl * slab_tx);
          u32 da = sa, db = sa + a.a_bytes + (PAIR ? 0u : b_share_off);
          for (int j = 0; j < nsl; ++j) {
            if (elect_one()) {
            if (PAIR) {   // own A rows and own half of the weight tile into own shared memory, bytes onto the leader's barrier
              if (ldA)
                tma_load_im2col_4d_pair(da, &tmA, c0, w0, h0, n, (unsigned short)((kh < d.KH ? kw : d.KW - 1) * d.dil),
                                        (unsigned short)((kh < d.KH ? kh : d.KH - 1) * d.dil), &s_full[s]);
              tma_load_2d_pair(db, &tmB, k0, n0, &s_full[s]);
            } else {
            if (ldA)   // past the last tap (K padding, zero weights) any finite activations do: repeat the last tap
              tma_load_im2col_4d(da, &tmA, c0, w0, h0, n, (unsigned short)((kh < d.KH ? kw : d.KW - 1) * d.dil),
                                 (unsigned short)((kh < d.KH ? kh : d.KH - 1) * d.dil), &s_full[s]);
            if (ldB) {
              if (!CL) tma_load_2d(db, &tmB, k0, n0, &s_full[s]);
              else tma_load_2d_mc(db, &tmB, k0, n0, &s_full[s], cmask);   // this CTA's rows, into every CTA
            }
            }
            }
            da += a.a_slab_bytes;
            db += a.b_slab_bytes;
            k0 += a.slabW;
            c0 += a.slabW;
            if (c0 >= d.Ci) {
              c0 = 0;
              if (++kw == d.KW) {
                kw = 0;
                ++kh;
              }
            }
          }
          sa += a.stage_bytes;
          if (++s == (u32)a.stages) {
            s = 0;
            ph ^= 1u;
            sa = smem_base;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===========================================================
    if (!PAIR || crank == 0) {   // every lane walks the loop; one elected lane (always the same one) issues MMAs and
                                 // commits; in a pair only the leader CTA issues (for both)
      u32 t = 0;
      const int mma_per_slab = a.slabW >= 16 ? a.slabW / 16 : 1;
      const bool wide = a.slabW >= 16;
      // descriptors of stage 0 / slab 0; stage s adds s * stage_bytes / 16, slab j adds j * slab_bytes / 16
      const u64 da0 = wide ? make_sdesc(smem_base, 16, a.sbo, a.layout_type) : make_sdesc(smem_base, a.a_slab_bytes, 128, 0);
      const u64 db0 = wide ? make_sdesc(smem_base + a.a_bytes, 16, a.sbo, a.layout_type)
                           : make_sdesc(smem_base + a.a_bytes, a.b_slab_bytes, 128, 0);
      const u32 stage16 = a.stage_bytes >> 4;
      const u32 aslab16 = (wide ? a.a_slab_bytes : 2u * a.a_slab_bytes) >> 4;   // narrow: one MMA spans two slabs
      const u32 bslab16 = (wide ? a.b_slab_bytes : 2u * a.b_slab_bytes) >> 4;
      u32 s = 0, ph = 0, soff16 = 0;
      int kcount = 0;
      for (int tile = cid; tile < a.total_tiles; tile += ncl, ++t) {
        const u32 acc = t & 1u, acc_ph = (t >> 1) & 1u;
        mbar_wait_parked(&s_tempty[acc], acc_ph ^ 1u);
        tc_fence_after();
        const u32 tmem_d = tmem_base + acc * a.acc_stride;
        u32 accumulate = 0;
        int left = a.nslabs;
        for (int ks = 0; ks < a.nsteps; ++ks) {
          if (dbg & 256) mbar_wait_spin(&s_full[s], ph); else mbar_wait_parked(&s_full[s], ph);
          if (!(dbg & 64)) tc_fence_after();
          if (DBG && trace && blockIdx.x == 0 && kcount < 1024 && lane == 0) trace[1024 + kcount] = clock64();
          const int nsl = min(a.g, left);
          left -= nsl;
          if (elect_one()) {
          if (!(dbg & 4)) {
            u64 da = da0 + (u64)soff16, db = db0 + (u64)soff16;
            const int nmma = wide ? nsl : (nsl + 1) >> 1;
            for (int j = 0; j < nmma; ++j) {
              for (int kk = 0; kk < mma_per_slab; ++kk) {   // +32 bytes of K inside the swizzle atom
                if (PAIR) umma_bf16_pair(tmem_d, da + (u64)(2 * kk), db + (u64)(2 * kk), a.idesc, accumulate);
                else umma_bf16(tmem_d, da + (u64)(2 * kk), db + (u64)(2 * kk), a.idesc, accumulate);
                accumulate = 1;
              }
              da += (u64)aslab16;
              db += (u64)bslab16;
            }
          }
          // stage is free once these MMAs have read it -- in every CTA of the cluster, whose producers multicast into it
          if (dbg & 32) mbar_arrive(&s_empty[s]);   // (timing experiment, only meaningful without MMAs)
          else if (PAIR) umma_commit_pair(&s_empty[s]);
          else if (!CL) umma_commit(&s_empty[s]);
          else umma_commit_mc(&s_empty[s], cmask);
          }
          __syncwarp();
          if (DBG && trace && blockIdx.x == 0 && kcount < 1024 && lane == 0) trace[2048 + kcount] = clock64();
          ++kcount;
          soff16 += stage16;
          if (++s == (u32)a.stages) {
            s = 0;
            ph ^= 1u;
            soff16 = 0;
          }
        }
        if (elect_one()) {   // accumulator complete (pair: in both CTAs' tensor memories)
          if (PAIR) umma_commit_pair(&s_tfull[acc]);
          else umma_commit(&s_tfull[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue (warps 2..5) ================================================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which of the quarter's two warps: even / odd column groups
    u32 t = 0;
    for (int tile = cid; tile < a.total_tiles; tile += ncl, ++t) {
      const int m_group = tile / a.n_tiles, n_tile = tile - m_group * a.n_tiles;
      const int m_tile = m_group * csize + crank;
      const u32 acc = t & 1u, acc_ph = (t >> 1) & 1u;
      const int m = m_tile * BM + 32 * q + lane;   // >= M for a CTA past the end of the last group
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
      // wide part: the quarter's two warps alternate over 64-column blocks, one tcgen05.ld each (see umma.cuh)
      const int nwide = a.BN / 64;
      for (int b = half; b < nwide; b += 2) {
        if (dbg & 16) continue;
        u32 v[64];
        tmem_ld64_nowait(taddr + (u32)(b * 64), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int co0 = n0 + b * 64 + j * 16;
          u32(&vj)[16] = *reinterpret_cast<u32(*)[16]>(&v[16 * j]);
          if (m < a.M && co0 < d.Co && !(dbg & 8)) epilogue_store(a, s_scale, s_shift, vj, m, co0, co0, HoWo, on, opix);
        }
      }
      for (int g = nwide * 4 + half; g < ngroups; g += 2) {   // the last BN % 64 columns, 16 at a time
        u32 v[16];
        if (dbg & 16) continue;
        tmem_ld16_nowait(taddr + (u32)(g * 16), v);
        tmem_ld_wait();
        const int co0 = n0 + g * 16;
        if (m < a.M && co0 < d.Co && !(dbg & 8)) epilogue_store(a, s_scale, s_shift, v, m, co0, co0, HoWo, on, opix);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_cluster(&s_tempty[acc], 0);   // the leader's MMA warp owns both accumulators
        else mbar_arrive(&s_tempty[acc]);
      }
    }
  }
  tc_fence_before();
  if (CL) cluster_sync_all();   // no CTA leaves while a peer can still signal its barriers
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair(tmem_base, a.tmem_cols);
    else tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ---- host side (tensor-map encode entry points: tma_host.h) ------------------------------------------------
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

TmaDriver& tma_driver() {
  static TmaDriver drv;
  static std::once_flag once;
  // cuTensorMapEncode* are DRIVER entry points and need a context current on the calling thread; a thread that has
  // only ever seen runtime calls that do not touch the device (autograd's backward worker before its first kernel)
  // has none yet and the encoders return CUDA_ERROR_INVALID_CONTEXT (measured).  cudaFree(0) binds the primary context.
  thread_local bool bound = false;
  if (!bound) {
    cudaFree(0);
    bound = true;
  }
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

static int max_active_clusters_pair(size_t smem) {   // CTA pairs of the cta_group::2 instantiation; cached per KB
  static std::mutex mu;
  static int cache[2] = {};
  std::lock_guard<std::mutex> lock(mu);
  const int kb = (int)((smem + 1023) / 1024);
  if (cache[0] == kb && cache[1] > 0) return cache[1];
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(128);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_tma_kernel<true, false, true>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  cache[0] = kb;
  cache[1] = n;
  return n;
}

// clusters of `c` CTAs (one CTA per SM at this shared-memory size) that can be resident at once; cached
static int max_active_clusters(int c, size_t smem) {
  static std::mutex mu;
  static int cache[5][2] = {};   // [c] -> {smem KB it was computed for, result}
  std::lock_guard<std::mutex> lock(mu);
  const int kb = (int)((smem + 1023) / 1024);
  if (cache[c][0] == kb && cache[c][1] > 0) return cache[c][1];
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(c * 64));
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)c;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_tma_kernel<true, false>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  cache[c][0] = kb;
  cache[c][1] = n;
  return n;
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
  a.slabW = Ci >= 64 ? 64 : Ci;
  a.ntaps = d->KH * d->KW;
  const int Ktot = a.ntaps * Ci;
  const int Kpad = round_up(Ktot, 64);           // layout of cnb_conv_pack_weights
  a.g = 64 / a.slabW;
  // K padding: 16-element MMA granularity (8-channel slabs pair up); the packed weights are zero there
  a.nslabs = round_up(Ktot, 16) / a.slabW;
  if (a.slabW == 8) a.nslabs = round_up(a.nslabs, 2);
  CNB_CHECK_ARG(a.nslabs * a.slabW <= Kpad, "conv: internal K padding error");
  // CTA pairs (tcgen05.mma.cta_group::2, M = 256): wide-N layers are bound by shared-memory bandwidth -- per K block the TMA
  // engine writes and the tensor core reads A (16 KB) + B (BN * 128 B) -- and a pair halves the B part per SM.
  // CNB_CONV_PAIR=0 disables, =1 forces it for every eligible geometry.
  static const int env_pair = [] { const char* e = getenv("CNB_CONV_PAIR"); return e ? atoi(e) : -1; }();
  static const int env_c = [] { const char* e = getenv("CNB_CONV_CLUSTER"); return e ? atoi(e) : 0; }();
  const bool pair = env_pair != 0 && env_c == 0 && a.slabW == 64 && a.BN % 16 == 0 && a.m_tiles >= 2 &&
                    (env_pair == 1 || a.BN >= 192);   // measured: BN = 256 +10-12 %, BN = 128 +-0
  a.a_slab_bytes = (u32)(BM * a.slabW * 2);
  a.b_slab_bytes = (u32)((pair ? a.BN / 2 : a.BN) * a.slabW * 2);
  a.nscale = a.n_tiles * a.BN;
  const size_t budget = 200 * 1024 - (size_t)a.nscale * 8;
  // K per pipeline stage: the two single-thread roles pay a few hundred clocks of bookkeeping per stage (barrier
  // wait, phase flip, commit), so a stage carries as many 64-wide K blocks as still leave 3 stages in flight
  // (CNB_TMA_KMULT overrides)
  {
    static const int env_km = [] { const char* e = getenv("CNB_TMA_KMULT"); return e ? atoi(e) : 0; }();
    const int g1 = a.g;
    for (int km = env_km > 0 ? env_km : 4; km >= 1; --km) {
      const size_t sb = (size_t)km * g1 * (a.a_slab_bytes + a.b_slab_bytes);
      const int nst = (a.nslabs + km * g1 - 1) / (km * g1);
      // at least three stages: with two, the ~1600-clock load latency of a refilled stage is exposed behind a
      // single stage's worth of MMAs (measured with CNB_TMA_TRACE)
      if (km == 1 || (sb * 3 <= budget && nst >= 2)) {
        a.g = km * g1;
        break;
      }
    }
  }
  a.nsteps = (a.nslabs + a.g - 1) / a.g;
  a.a_bytes = (u32)a.g * a.a_slab_bytes;
  a.stage_bytes = (u32)a.g * (a.a_slab_bytes + a.b_slab_bytes);
  CUtensorMapSwizzle swz;
  switch (a.slabW) {
    case 64: a.layout_type = 2; a.sbo = 1024; swz = CU_TENSOR_MAP_SWIZZLE_128B; break;
    case 32: a.layout_type = 4; a.sbo = 512; swz = CU_TENSOR_MAP_SWIZZLE_64B; break;
    case 16: a.layout_type = 6; a.sbo = 256; swz = CU_TENSOR_MAP_SWIZZLE_32B; break;
    default: a.layout_type = 0; a.sbo = 128; swz = CU_TENSOR_MAP_SWIZZLE_NONE; break;
  }
  a.stages = (int)(budget / a.stage_bytes);
  if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
  static const int env_stages = [] { const char* e = getenv("CNB_TMA_STAGES"); return e ? atoi(e) : 0; }();
  if (env_stages >= 2 && env_stages < a.stages) a.stages = env_stages;
  CNB_CHECK_ARG(a.stages >= 2, "conv: tile does not fit in shared memory");
  a.acc_stride = (u32)round_up(a.BN, 32);
  a.tmem_cols = 32;
  while (a.tmem_cols < 2 * a.acc_stride) a.tmem_cols <<= 1;
  a.idesc = make_idesc_bf16(pair ? 2 * BM : BM, a.BN);
  const size_t smem = (size_t)a.stages * a.stage_bytes + (size_t)a.nscale * 8 + 1024;
  static PerDeviceOnce once;
  if (once.need()) {
    CNB_CUDA(cudaFuncSetAttribute(conv_tma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CNB_CUDA(cudaFuncSetAttribute(conv_tma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CNB_CUDA(cudaFuncSetAttribute(conv_tma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CNB_CUDA(cudaFuncSetAttribute(conv_tma_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    once.mark();
  }

  // ---- clusters (opt-in: CNB_CONV_CLUSTER=2|4).  Multicasting the weight tile cuts the L2 -> SM bytes of a K block
  // from 16 KB + BN*128 to 16 KB + BN*128/c, but measured on B200 it buys nothing: the kernel is paced by the MMA
  // issue thread (tcgen05.mma with both operands in shared memory accepts one 128 x BN x 16 instruction per
  // ~BN clocks) plus its per-stage bookkeeping, not by the loads.  Default: no clusters.
  a.csize = 1;
  int nclusters = drv.num_sms;
  if (pair) {
    const int ncl = max_active_clusters_pair(smem);
    CNB_CHECK_ARG(ncl >= 1, "conv: no CTA pair fits (shared memory %zu bytes)", smem);
    a.csize = 2;
    nclusters = ncl;
  } else if ((env_c == 2 || env_c == 4) && a.slabW == 64 && a.BN % (8 * env_c) == 0 && a.m_tiles >= env_c) {
    const int ncl = max_active_clusters(env_c, smem);
    if (ncl >= 1) {
      a.csize = env_c;
      nclusters = ncl;
    }
  }
  static const bool verbose = getenv("CNB_CONV_VERBOSE") != nullptr;
  if (verbose)
    fprintf(stderr, "conv_tma: Ci=%d Co=%d %dx%d BN=%d m_tiles=%d -> cluster %d (%d clusters resident)\n", Ci, d->Co,
            d->Ho, d->Wo, a.BN, a.m_tiles, a.csize, nclusters);
  static const bool env_trace = getenv("CNB_TMA_TRACE") != nullptr;
  static long long* trace_buf = nullptr;
  a.trace = nullptr;
  if (env_trace) {
    if (!trace_buf) cudaMalloc(&trace_buf, 3 * 1024 * sizeof(long long));
    cudaMemset(trace_buf, 0, 3 * 1024 * sizeof(long long));
    a.trace = trace_buf;
  }
  static const int env_dbg = [] { const char* e = getenv("CNB_TMA_DEBUG"); return e ? atoi(e) : 0; }();
  a.debug = env_dbg;
  a.bn_share = a.BN / a.csize;
  a.total_tiles = (a.m_tiles + a.csize - 1) / a.csize * a.n_tiles;

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
    cuuint32_t box[2] = {(cuuint32_t)a.slabW, (cuuint32_t)a.bn_share};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = drv.tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wpk, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv: cuTensorMapEncodeTiled failed (%d) Kpad=%d Co_pad=%d BN=%d", (int)r, Kpad, Co_pad, a.BN);
      return CNB_ERR_CUDA;
    }
  }
  if (a.csize == 1) {
    const int grid = a.total_tiles < drv.num_sms ? a.total_tiles : drv.num_sms;
    if (a.debug || a.trace) conv_tma_kernel<false, true><<<grid, NTHREADS, smem, st>>>(tmA, tmB, a);
    else CNB_CUDA(launch_pdl(conv_tma_kernel<false, false>, dim3(grid), dim3(NTHREADS), smem, st, tmA, tmB, a));
  } else {
    const int ncl = a.total_tiles < nclusters ? a.total_tiles : nclusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ncl * a.csize));
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (pair) {   // programmatic dependent launch as for the single-CTA kernel
      static const bool pdl_on = [] { const char* e = getenv("CNB_PDL"); return !(e && e[0] == '0'); }();
      cudaLaunchAttribute attr2[2];
      attr2[0] = attr[0];
      attr2[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr2[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr2;
      cfg.numAttrs = pdl_on ? 2 : 1;
      CNB_CUDA(cudaLaunchKernelEx(&cfg, conv_tma_kernel<true, false, true>, tmA, tmB, a));
    } else {
      CNB_CUDA(cudaLaunchKernelEx(&cfg, conv_tma_kernel<true, false>, tmA, tmB, a));
    }
  }
  CNB_LAUNCH_CHECK();
  if (env_trace) {   // debugging aid: per-K-block clock stamps of CTA 0 (producer got the stage | MMA thread got the
                     // data | MMA thread committed), printed as deltas to the previous K block
    static int printed = 0;
    cudaStreamSynchronize(st);
    if (printed++ == 3) {
      static long long h[3 * 1024];
      cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "conv_tma trace Ci=%d Co=%d BN=%d stages=%d nsteps=%d\n", Ci, d->Co, a.BN, a.stages, a.nsteps);
      for (int k = 1; k < 1024 && h[1024 + k]; ++k)
        if (k < 100)
          fprintf(stderr, "k=%3d  prod +%5lld  mma_got +%5lld  commit +%5lld   (got-prod %6lld, commit-got %5lld)\n", k,
                  h[k] - h[k - 1], h[1024 + k] - h[1024 + k - 1], h[2048 + k] - h[2048 + k - 1], h[1024 + k] - h[k],
                  h[2048 + k] - h[1024 + k]);
    }
  }
  return CNB_OK;
}

}  // namespace cnb
