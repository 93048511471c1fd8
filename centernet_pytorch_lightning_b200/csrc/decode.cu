// Fused decode for CenterNet heads: 3x3 max-pool NMS + top-K peak extraction + offset gather.
//
// Replaces (reference file:line, all under CenterNet/):
//   utils/decode.py:5-10   _nms                 -> decode_scan_kernel (3x3 test in the smem tile, only where needed)
//   utils/decode.py:13-28  _topk                -> streaming selection with a per-image running threshold
//   utils/decode.py:31-40  _topk_channel        -> the same with one group per plane
//   utils/decode.py:59-63  _transpose_and_gather_feat -> direct NCHW gathers at the K winners
//   decode/ctdet.py:6-38   ctdet_decode         -> decode_scan_kernel (one launch; the CTA that completes an
//                                                  image merges its candidates and writes [K,6])
//   decode/multi_pose.py:7-96 multi_pose_decode -> decode_scan_kernel + multi_pose_assoc_kernel
//
// Data layout in HBM: heat maps NCHW fp32 exactly as the heads emit them.  They are streamed ONCE:
// persistent CTAs (2 per SM) each own a contiguous range of (group, plane, band-of-rows) chunks and pull them
// through a 4-stage ring of 1-D TMA bulk copies; algorithmic traffic = C*H*W*4 B per image (+ a few KB of
// candidates), see DESIGN.md.
//
// Selection.  A *group* is the set of planes that compete in one top-K (ctdet: the C class planes of an image;
// multi_pose: every plane on its own).  All selections order 64-bit keys
//     (score_bits << 32) | (0xFFFFFFFF - flat_index)        "larger key" == (higher score, then lower index)
// which are distinct, so every result is deterministic (the reference leaves ties to torch.topk).  Scores must
// be >= 0 (sigmoid outputs, as at every reference call site).  Per group the kernel keeps in global memory
//   thr   : a monotone lower bound of the group's K-th largest key (atomicMax),
//   hist  : counts of NMS survivors per score bin (7 mantissa bits, scores in [2^-8, 1]); whenever the counts
//           of the bins >= t reach K, (edge(t) << 32) is a sound new bound,
//   list  : appended candidate keys (only those >= thr at the time).
// A tile element is looked at closely only if its value reaches the bound (`max4 >= thr`): in steady state the
// scan is one LDS.128 + 3 FMNMX + 1 compare per 4 elements, so the kernel is HBM bound.  Chunks that hold more
// than ACAP live survivors (the first chunks of a group, plateaus of equal scores) take an exact radix select
// over the 64-bit keys, which also publishes an exact bound -- ties are then pruned by index.
#include "cnb_common.cuh"
#include <stdlib.h>

namespace cnb {
namespace {

constexpr int NT = 256;          // threads per CTA
constexpr int NW = NT / 32;
constexpr int MAX_K = 512;
constexpr int NST = 4;           // tile stages in flight per CTA
// Score histogram, three levels (all counts of NMS survivors, per group):
//   octave : exponent of the score, 2^-16 .. 1            (NOCT counters)
//   coarse : 7 mantissa bits, 128 bins per octave         (NCB counters)
//   fine   : 4 more mantissa bits, 16 sub-bins per coarse (NBF counters; large groups only)
constexpr int EXP0 = 111;                      // scores below 2^-16 are not counted (never bound anything)
constexpr int CB_BASE = EXP0 << 7;             // (bits >> 16) of 2^-16
constexpr int NOCT = 17;                       // octave 16 = scores >= 1
constexpr int NCB = 16 * 128 + 1;
constexpr int FSUB = 16;
constexpr int NBF = NCB * FSUB;
constexpr int OCT_PAD = 32;                    // ints reserved per group for the octave sums
constexpr int NCB_PAD = NCB + 3;               // keeps the per-group arrays 16-byte aligned
constexpr int MCAP = 2048;       // merge candidates staged in shared memory
constexpr int RANK_DIRECT = 384; // up to this many candidates are ranked by direct counting
constexpr int HCAP = 512;        // hit positions recorded per chunk before the dense path takes over
constexpr int SMALL_FLUSH = 128;   // flushes up to this size reserve their list space up front
constexpr int FLUSH_AT = 64;     // live survivors a CTA carries across chunks before it talks to global memory

struct ScanArgs {
  const float* t0;   // [B, C0, H, W]
  const float* t1;   // [B, C1, H, W] or nullptr
  int C0, C1;
  int PG;            // planes per group
  int G;             // groups
  int B, H, W, K;
  int R;             // rows per chunk
  int nbands;
  int rpt;           // rows per thread segment
  int acap;          // live survivors a chunk may append (>= K)
  int cpg;           // chunks per group
  int total_chunks;
  int gcap;          // candidate list capacity per group
  int tile_bytes;
  // per-group state (zeroed by the host wrapper before the launch)
  u64* thr;          // [G]
  int* gcount;       // [G]
  int* gdone;        // [G]
  int* goct;         // [G][OCT_PAD]
  int* ghist;        // [G][NCB_PAD]
  int* gfine;        // [G][NBF] or nullptr (small groups resolve well enough with the coarse bins)
  u64* lists;        // [G][gcap]
  // results
  u64* topk;         // multi_pose: [G][K] sorted keys
  int* have;         // multi_pose: [G]
  const float* wh;   // fused ctdet epilogue
  const float* reg;
  float* out;        // [B,K,6]
  int fuse_ctdet;
  int wflush;        // streaming scan: list fill level at which a warp flushes
  int debug;         // CNB_DECODE_DEBUG=1: per-CTA phase timestamps into g_dbg (tools/decode_timeline.py)
};

constexpr int DBG_SLOTS = 8;
__device__ unsigned long long g_dbg[1024 * DBG_SLOTS];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DBG_MARK(slot)                                                                      \
  do {                                                                                      \
    if (a.debug && threadIdx.x == 0 && blockIdx.x < 1024) g_dbg[blockIdx.x * DBG_SLOTS + (slot)] = gtimer(); \
  } while (0)
#define DBG_ADD(slot, v)                                                                    \
  do {                                                                                      \
    if (a.debug && threadIdx.x == 0 && blockIdx.x < 1024) g_dbg[blockIdx.x * DBG_SLOTS + (slot)] += (v); \
  } while (0)

__device__ __forceinline__ u64 make_key(float score, u32 flat) {
  return ((u64)__float_as_uint(score) << 32) | (u64)(0xFFFFFFFFu - flat);
}
__device__ __forceinline__ u32 key_hi(u64 k) { return (u32)(k >> 32); }
__device__ __forceinline__ u32 key_idx(u64 k) { return 0xFFFFFFFFu - (u32)(k & 0xFFFFFFFFull); }

// coarse bin of a score (bits of a positive float); < 0: below the histogram's range
__device__ __forceinline__ int coarse_bin(u32 bits) {
  const int b = (int)(bits >> 16) - CB_BASE;
  return min(b, NCB - 1);
}
__device__ __forceinline__ u32 bin_edge_bits(int cb, int sub) {
  return ((u32)(cb + CB_BASE) << 16) | ((u32)sub << 12);
}

__device__ __forceinline__ u64 ldcg_u64(const u64* p) {
  return __ldcg(reinterpret_cast<const unsigned long long*>(p));
}

// ---- shared-memory scratch shared by the selection routines -----------------------------------------------
struct Scratch {
  u32* hist;     // [256]
  int* red;      // [NW]
  int* misc;     // [8]
  u64* keyred;   // [2*NW]
};

// ---- candidate enumerators --------------------------------------------------------------------------------
template <int VEC>
struct FlagCands {       // survivors of the scanned chunk, kept as a per-thread bit mask
  const float* tile;     // tile row 0 == global row r0-1
  u64 flags;
  int rs, cg, W, r0;
  u32 flat_base;
  template <class F>
  __device__ __forceinline__ void for_each(F f) const {
    u64 fl = flags;
    while (fl) {
      const int bit = __ffsll((long long)fl) - 1;
      fl &= fl - 1;
      const int row = rs + bit / VEC;
      const int col = cg * VEC + bit % VEC;
      const float v = tile[(row - (r0 - 1)) * W + col];
      f(make_key(v, flat_base + (u32)(row * W + col)));
    }
  }
};
struct SmemCands {
  const u64* list;
  int n;
  template <class F>
  __device__ __forceinline__ void for_each(F f) const {
    for (int i = threadIdx.x; i < n; i += NT) f(list[i]);
  }
};
struct GlobalCands {
  const u64* list;
  int n;
  u64 floor;
  template <class F>
  __device__ __forceinline__ void for_each(F f) const {
    for (int i = threadIdx.x; i < n; i += NT) {
      const u64 k = ldcg_u64(list + i);
      if (k >= floor) f(k);
    }
  }
};

// Exact K-th largest of the CTA's (distinct) candidate keys; the caller guarantees at least K candidates.
// MSD radix select, 8-bit digits; digits on which all candidates agree are skipped (one AND/OR reduction).
template <class Cands>
__device__ u64 block_kth_key(const Cands& c, int K, const Scratch& sc) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  u64 kand = ~0ull, kor = 0ull;
  c.for_each([&](u64 k) {
    kand &= k;
    kor |= k;
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kand &= __shfl_xor_sync(0xffffffffu, kand, o);
    kor |= __shfl_xor_sync(0xffffffffu, kor, o);
  }
  if (lane == 0) {
    sc.keyred[warp] = kand;
    sc.keyred[NW + warp] = kor;
  }
  __syncthreads();
  kand = ~0ull;
  kor = 0ull;
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    kand &= sc.keyred[i];
    kor |= sc.keyred[NW + i];
  }
  const u64 varying = kand ^ kor;
  u64 prefix = 0;
  int rem = K;
  for (int shift = 56; shift >= 0; shift -= 8) {
    if (((varying >> shift) & 0xFFull) == 0) {   // block-uniform: every candidate has this digit
      prefix |= kand & (0xFFull << shift);
      continue;
    }
    sc.hist[tid] = 0;   // NT == 256 bins
    __syncthreads();
    const u64 himask = shift == 56 ? 0ull : (~0ull << (shift + 8));
    c.for_each([&](u64 k) {
      if ((k & himask) == prefix) atomicAdd(&sc.hist[(u32)(k >> shift) & 0xFFu], 1u);
    });
    __syncthreads();
    if (warp == 0) {
      u32 cnt[8];
      u32 s = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // lane 0 owns the 8 largest digits
        cnt[j] = sc.hist[255 - 8 * lane - j];
        s += cnt[j];
      }
      u32 incl = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const u32 excl = incl - s;
      if (excl < (u32)rem && (u32)rem <= incl) {
        u32 r = (u32)rem - excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (r <= cnt[j]) {
            sc.misc[0] = 255 - 8 * lane - j;
            sc.misc[1] = (int)r;
            break;
          }
          r -= cnt[j];
        }
      }
    }
    __syncthreads();
    prefix |= (u64)(u32)sc.misc[0] << shift;
    rem = sc.misc[1];
  }
  __syncthreads();
  return prefix;
}

// Is flat element `idx` of image-tensor `img` ([C,H,W]) a strictly positive NMS survivor?
__device__ __forceinline__ bool is_positive_peak(const float* img, int idx, int H, int W) {
  const int HW = H * W;
  const int pix = idx % HW;
  const int y = pix / W, x = pix % W;
  const float* pl = img + (size_t)(idx - pix);
  const float v = __ldg(pl + pix);
  if (!(v > 0.f)) return false;
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      if (__ldg(pl + yy * W + xx) > v) return false;
    }
  }
  return true;
}

// Reference semantics when a group has fewer than K positive peaks: torch.topk then returns zero-valued
// entries of heat*keep; we define their order as flat index ascending.
__device__ void block_zero_fill(const float* img, int n_elems, int H, int W, int have, int K, u64* out,
                                int* red) {
  int filled = have;
  for (int base = 0; base < n_elems && filled < K; base += NT) {
    const int idx = base + threadIdx.x;
    const bool z = idx < n_elems && !is_positive_peak(img, idx, H, W);
    const u32 bal = __ballot_sync(0xffffffffu, z);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = __popc(bal);
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int c = red[i];
      if (i < w) before += c;
      tot += c;
    }
    const int slot = filled + before + __popc(bal & ((1u << lane) - 1u));
    if (z && slot < K) out[slot] = make_key(0.f, (u32)idx);
    filled += tot;
    __syncthreads();
  }
}

// ---- ctdet epilogue: [K,6] rows from the sorted keys (decode/ctdet.py:15-38) ---------------------
__device__ void ctdet_write_rows(const ScanArgs& a, int b, const u64* keys) {
  const int HW = a.H * a.W;
  for (int k = threadIdx.x; k < a.K; k += NT) {
    const u64 key = keys[k];
    const float score = __uint_as_float(key_hi(key));
    const int idx = (int)key_idx(key);
    const int cls = idx / HW, pix = idx % HW;
    float xs = (float)(pix % a.W), ys = (float)(pix / a.W);
    if (a.reg) {
      xs = __fadd_rn(xs, __ldg(a.reg + ((size_t)b * 2 + 0) * HW + pix));
      ys = __fadd_rn(ys, __ldg(a.reg + ((size_t)b * 2 + 1) * HW + pix));
    } else {
      xs = __fadd_rn(xs, 0.5f);
      ys = __fadd_rn(ys, 0.5f);
    }
    const float hw = __fmul_rn(__ldg(a.wh + ((size_t)b * 2 + 0) * HW + pix), 0.5f);
    const float hh = __fmul_rn(__ldg(a.wh + ((size_t)b * 2 + 1) * HW + pix), 0.5f);
    float* o = a.out + ((size_t)b * a.K + k) * 6;
    o[0] = __fsub_rn(xs, hw);
    o[1] = __fsub_rn(ys, hh);
    o[2] = __fadd_rn(xs, hw);
    o[3] = __fadd_rn(ys, hh);
    o[4] = score;
    o[5] = (float)cls;
  }
}

// plane pointer of plane `pg` of group `g`
__device__ __forceinline__ const float* plane_ptr(const ScanArgs& a, int g, int pg, u32* flat_base) {
  const int HW = a.H * a.W;
  const int ppi = a.C0 + a.C1;
  const int P = g * a.PG + pg;
  const int b = P / ppi, p = P - b * ppi;
  *flat_base = (u32)(pg * HW);
  return p < a.C0 ? a.t0 + ((size_t)b * a.C0 + p) * HW : a.t1 + ((size_t)b * a.C1 + (p - a.C0)) * HW;
}

// Warp-collective: the largest sound lower bound of the group's K-th key that its survivor histogram supports
// (0 when fewer than K survivors are counted yet).
__device__ u64 warp_bound_from_hist(const int* oct, const int* hist, const int* fine, int K, u64 cur_thr) {
  const int lane = threadIdx.x & 31;
  // Largest threshold t with count(score >= t) >= K, resolved octave -> coarse bin -> fine sub-bin.  The
  // three levels are read in ONE round trip, speculating that the crossing octave / bin are the ones of
  // the bound we already hold; a level is re-read only when the crossing moved.
  const u32 tb = key_hi(cur_thr);
  const int cb_prev = cur_thr ? coarse_bin(tb) : -1;
  const int o_prev = cb_prev >= 0 ? cb_prev >> 7 : 15;
  int vo = lane < NOCT ? __ldcg(oct + (NOCT - 1 - lane)) : 0;           // lane 0 = top octave
  auto load_coarse = [&](int o, int (&cv)[4]) {                          // lane 0 = the octave's top bins
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cb = o * 128 + 127 - (lane * 4 + j);
      cv[j] = cb < NCB ? __ldcg(hist + cb) : 0;
    }
  };
  int cv[4];
  load_coarse(o_prev, cv);
  int fv = (fine && cb_prev >= 0 && lane < FSUB) ? __ldcg(fine + cb_prev * FSUB + (FSUB - 1 - lane)) : 0;
  // octave level
  int incl = vo;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  u32 ok = __ballot_sync(0xffffffffu, incl >= K);
  u64 nt = 0ull;
  if (ok) {
    const int lo = __ffs(ok) - 1;                       // lane of the crossing octave
    const int ostar = NOCT - 1 - lo;
    int above = __shfl_sync(0xffffffffu, incl - vo, lo);   // survivors counted in the octaves above it
    if (ostar != o_prev) load_coarse(ostar, cv);
    // coarse level (descending bins: lane 0 holds the octave's 4 largest)
    const int csum = cv[0] + cv[1] + cv[2] + cv[3];
    incl = csum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    ok = __ballot_sync(0xffffffffu, above + incl >= K);
    if (ok) {
      const int lc = __ffs(ok) - 1;
      int mycb = -1, myabove = 0;
      if (lane == lc) {
        int r = above + incl - csum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (r + cv[j] >= K && mycb < 0) {
            mycb = ostar * 128 + 127 - (lane * 4 + j);
            myabove = r;
          }
          r += cv[j];
        }
      }
      const int cbstar = __shfl_sync(0xffffffffu, mycb, lc);
      above = __shfl_sync(0xffffffffu, myabove, lc);
      int sub = 0;
      if (fine) {
        if (cbstar != cb_prev) fv = lane < FSUB ? __ldcg(fine + cbstar * FSUB + (FSUB - 1 - lane)) : 0;
        incl = fv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        ok = __ballot_sync(0xffffffffu, lane < FSUB && above + incl >= K);
        if (ok) sub = FSUB - 1 - (__ffs(ok) - 1);
      }
      nt = (u64)bin_edge_bits(cbstar, sub) << 32;
    } else {
      nt = (u64)bin_edge_bits(ostar * 128, 0) << 32;   // (stale coarse counts) the octave's lower edge
    }
  }
  return nt;
}

// Exact, sorted top-K of a finished group: candidates >= the final bound are staged in shared memory, cut with a
// radix select when there are many, ranked by counting; then the [K,6] rows (ctdet) or the sorted key list
// (multi_pose) are written.  Block-collective (NT threads).  s_sel holds >= K keys.
__device__ void merge_group(const ScanArgs& a, int g, u64* s_mlist, u64* s_list, u64* s_out, const Scratch& sc) {
  const int tid = threadIdx.x;
  const int HW = a.H * a.W;
    __threadfence();
    const int n_all = min(__ldcg(a.gcount + g), a.gcap);
    const u64 floor = ldcg_u64(a.thr + g);
    const u64* gl = a.lists + (size_t)g * a.gcap;
    if (tid == 0) sc.misc[4] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n_all; i0 += NT * 8) {   // 8 independent loads per thread in flight (the list sits in L2)
      u64 kk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + j * NT + tid;
        kk[j] = i < n_all ? ldcg_u64(gl + i) : 0ull;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (kk[j] != 0ull && kk[j] >= floor) {
          const int slot = atomicAdd(&sc.misc[4], 1);
          if (slot < MCAP) s_mlist[slot] = kk[j];
        }
      }
    }
    __syncthreads();
    int m = sc.misc[4];
    __syncthreads();
    const u64* L = s_mlist;
    if (m > MCAP) {   // more candidates than the staging area holds: exact K-th straight from global memory
      const GlobalCands gc{gl, n_all, floor};
      const u64 kth = block_kth_key(gc, a.K, sc);
      if (tid == 0) sc.misc[4] = 0;
      __syncthreads();
      gc.for_each([&](u64 key) {
        if (key >= kth) s_mlist[atomicAdd(&sc.misc[4], 1)] = key;
      });
      __syncthreads();
      m = sc.misc[4];
    } else if (m > RANK_DIRECT && m > a.K) {
      const SmemCands smc{s_mlist, m};
      const u64 kth = block_kth_key(smc, a.K, sc);
      if (tid == 0) sc.misc[4] = 0;
      __syncthreads();
      smc.for_each([&](u64 key) {
        if (key >= kth) s_list[atomicAdd(&sc.misc[4], 1)] = key;   // exactly K <= acap entries
      });
      __syncthreads();
      m = sc.misc[4];
      L = s_list;
    }
    for (int i = tid; i < m; i += NT) {   // exact rank by counting (keys distinct)
      const u64 key = L[i];
      int r = 0;
      for (int j = 0; j < m; ++j) r += (L[j] > key);
      if (r < a.K) s_out[r] = key;
    }
    __syncthreads();
    int have = min(m, a.K);
    if (a.fuse_ctdet) {
      if (have < a.K) {
        block_zero_fill(a.t0 + (size_t)g * a.C0 * HW, a.C0 * HW, a.H, a.W, have, a.K, s_out, sc.red);
        __syncthreads();
      }
      ctdet_write_rows(a, g, s_out);
    } else {
      const int ppi = a.C0 + a.C1;
      const int b = g / ppi, p = g - b * ppi;
      if (p < a.C0 && have < a.K) {   // detection heat map: zero-filled like torch.topk
        block_zero_fill(a.t0 + ((size_t)b * a.C0 + p) * HW, HW, a.H, a.W, have, a.K, s_out, sc.red);
        __syncthreads();
        have = a.K;
      }
      for (int i = tid; i < have; i += NT) a.topk[(size_t)g * a.K + i] = s_out[i];
      if (tid == 0) a.have[g] = have;
    }
}

// =================================================================================================
// decode_scan_kernel: persistent CTAs over contiguous ranges of (group, plane, band) chunks.
// =================================================================================================
template <int VEC>
__global__ void __launch_bounds__(NT, 2) decode_scan_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) u64 s_full[NST];
  __shared__ __align__(8) u64 s_keyred[2 * NW];
  __shared__ __align__(8) u64 s_thr;
  __shared__ u32 s_hist[256];
  __shared__ int s_red[NW];
  __shared__ int s_misc[8];
  __shared__ int s_cnt;        // live survivors carried in s_list
  __shared__ int s_nhit[3];    // hits of chunk k in s_hits[k % 3]
  __shared__ __align__(8) u64 s_ref_thr;   // periodic refresh of the group's bound from global memory
  __shared__ int s_ref_g;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NSTAGE = VEC == 4 ? NST : 1;
  u64* s_mlist = reinterpret_cast<u64*>(smem_raw + (size_t)NSTAGE * a.tile_bytes);   // [MCAP]
  const int LC = HCAP + FLUSH_AT;
  u64* s_list = s_mlist + MCAP;                                                      // [LC]
  u64* s_out = s_list + LC;                                                          // [MAX_K]
  u32* s_hits = reinterpret_cast<u32*>(s_out + MAX_K);                               // [3][HCAP]
  const Scratch sc{s_hist, s_red, s_misc, s_keyred};

  const int c_begin = (int)((long long)blockIdx.x * a.total_chunks / gridDim.x);
  const int c_end = (int)((long long)(blockIdx.x + 1) * a.total_chunks / gridDim.x);
  const int HW = a.H * a.W;

  auto issue_load = [&](int g, int pg, int band, int stage) {   // thread 0 only (VEC == 4)
    u32 fb;
    const float* plane = plane_ptr(a, g, pg, &fb);
    const int r0 = band * a.R, r1 = min(r0 + a.R, a.H);
    const int g0 = max(r0 - 1, 0), g1 = min(r1 + 1, a.H);
    const u32 bytes = (u32)((g1 - g0) * a.W * sizeof(float));
    float* tile = reinterpret_cast<float*>(smem_raw + (size_t)stage * a.tile_bytes);
    mbar_expect_tx(&s_full[stage], bytes);
    bulk_g2s(tile + (size_t)(g0 - (r0 - 1)) * a.W, plane + (size_t)g0 * a.W, bytes, &s_full[stage]);
  };
  auto advance = [&](int& g, int& pg, int& band) {
    if (++band == a.nbands) {
      band = 0;
      if (++pg == a.PG) {
        pg = 0;
        ++g;
      }
    }
  };
  // group of the chunk `d` (<= 3) positions after chunk (g, pg, band)
  auto group_after = [&](int g, int pg, int band, int d) {
    int rem = pg * a.nbands + band + d;
    while (rem >= a.cpg) {
      rem -= a.cpg;
      ++g;
    }
    return g;
  };

  if (a.debug && tid == 0 && blockIdx.x < 1024)
    for (int i = 0; i < DBG_SLOTS; ++i) g_dbg[blockIdx.x * DBG_SLOTS + i] = 0;
  DBG_MARK(0);
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&s_full[s], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
    s_cnt = 0;
    s_nhit[0] = s_nhit[1] = s_nhit[2] = 0;
  }
  __syncthreads();
  const int W4 = a.W / VEC;
  const int cg = tid % W4;
  const int seg = tid / W4;
  int done_in_group = 0;
  int carry = 0;           // live survivors carried in s_list across chunks (block-uniform)
  // chunk coordinates advance incrementally (no per-chunk integer divisions)
  int g = c_begin < c_end ? c_begin / a.cpg : 0;
  int pg, band;
  {
    const int rem = c_begin - g * a.cpg;
    pg = rem / a.nbands;
    band = rem - pg * a.nbands;
  }
  int lg = g, lpg = pg, lband = band;   // load cursor: the chunk NST ahead
  for (int s = 0; s < NST; ++s) {
    if (VEC == 4 && tid == 0 && c_begin + s < c_end) issue_load(lg, lpg, lband, s);
    advance(lg, lpg, lband);
  }
  // The group's bound lives in a block-uniform register: fetched once when the CTA enters a group, raised by the
  // CTA's own flushes (which read the group's histogram) and by a background refresh every 8 chunks whose L2
  // latency is given 4 chunks to hide.
  u64 tq = 0ull;           // thread 0: refresh in flight ...
  int tq_g = -1;           // ... and the group it was issued for
  u64 cur_thr = 0ull;
  if (tid == 0) {
    s_thr = c_begin < c_end ? ldcg_u64(a.thr + g) : 0ull;
    s_ref_thr = 0ull;
    s_ref_g = -1;
  }
  __syncthreads();
  cur_thr = s_thr;

  for (int c = c_begin; c < c_end; ++c) {
    const int k = c - c_begin;
    const int stage = VEC == 4 ? (k & (NST - 1)) : 0;
    u32 flat_base;
    const float* plane = plane_ptr(a, g, pg, &flat_base);
    const int r0 = band * a.R, r1 = min(r0 + a.R, a.H);
    float* tile = reinterpret_cast<float*>(smem_raw + (size_t)stage * a.tile_bytes);
    const bool group_ends = (c + 1 == c_end) || (pg == a.PG - 1 && band == a.nbands - 1);
    int* nhit = &s_nhit[k % 3];
    u32* hits = s_hits + (k % 3) * HCAP;
    if (tid == 0) s_nhit[(k + 1) % 3] = 0;

    if (s_ref_g == g && s_ref_thr > cur_thr) cur_thr = s_ref_thr;   // (written before the previous barrier)
    u64 thr = cur_thr;                            // a (possibly stale, hence sound) bound of the group
    if (tid == 0 && (k & 7) == 0) {
      tq = ldcg_u64(a.thr + g);
      tq_g = g;
    }
    if (VEC == 4) {
      const unsigned long long tw0 = a.debug ? gtimer() : 0ull;
      mbar_wait(&s_full[stage], (u32)(k / NST) & 1u);
      if (a.debug) DBG_ADD(7, gtimer() - tw0);
    } else {   // unaligned / odd widths: plain cooperative loads, no pipelining
      const int g0 = max(r0 - 1, 0), g1 = min(r1 + 1, a.H);
      const int n = (g1 - g0) * a.W;
      float* dst = tile + (size_t)(g0 - (r0 - 1)) * a.W;
      const float* src = plane + (size_t)g0 * a.W;
      for (int i = tid; i < n; i += NT) dst[i] = __ldg(src + i);
      __syncthreads();
    }

    // ---- scan: thread = (column group cg, row segment seg) --------------------------------------------
    // Elements that reach the bound are only *recorded* here (position in the tile); the 3x3 test runs
    // afterwards over the hit list with all threads busy.  With no bound yet (first chunk of a group) every
    // element would be a hit: that case goes straight to the dense path below.
    const float ts = __uint_as_float(key_hi(thr));
    const int rs = r0 + seg * a.rpt;
    const int re = min(rs + a.rpt, r1);
    const bool dense0 = thr == 0ull;              // block-uniform
    bool any_hit = false;
    if (!dense0) {
      // one look at the whole chunk first: the largest value of the thread's rows against the bound, one vote
      float mt = 0.f;
      if constexpr (VEC == 4) {
        float4 qq[4];
        for (int rb = 0; rb < a.rpt; rb += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int row = min(rs + rb + u, max(re - 1, r0));
            qq[u] = *reinterpret_cast<const float4*>(tile + (size_t)(row - (r0 - 1)) * a.W + cg * VEC);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mt = fmaxf(mt, fmaxf(fmaxf(qq[u].x, qq[u].y), fmaxf(qq[u].z, qq[u].w)));
        }
      } else {
        for (int row = rs; row < re; ++row) mt = fmaxf(mt, tile[(size_t)(row - (r0 - 1)) * a.W + cg]);
      }
      any_hit = __any_sync(0xffffffffu, rs < re && mt >= ts && mt > 0.f);
    }
    if (any_hit) {
      for (int rr = 0; rr < a.rpt; ++rr) {        // same trip count in every lane (votes inside)
        const int row = rs + rr;
        const bool live = row < re;
        const float* rowp = tile + (size_t)((live ? row : r0) - (r0 - 1)) * a.W + cg * VEC;
        float v[VEC];
        float m4;
        if constexpr (VEC == 4) {
          const float4 q = *reinterpret_cast<const float4*>(rowp);
          v[0] = q.x; v[1 % VEC] = q.y; v[2 % VEC] = q.z; v[3 % VEC] = q.w;
          m4 = fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w));
        } else {
          v[0] = rowp[0];
          m4 = v[0];
        }
        const bool maybe = live && m4 >= ts && m4 > 0.f;
        if (!__any_sync(0xffffffffu, maybe)) continue;   // the common case once the bound has tightened
        u32 hm = 0;
        if (maybe) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float x = v[i];
            if (x >= ts && x > 0.f &&
                make_key(x, flat_base + (u32)(row * a.W + cg * VEC + i)) >= thr)
              hm |= 1u << i;
          }
        }
        // warp-aggregated append of the hit positions: one shared-memory atomic per warp and row
        const int mine = __popc(hm);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
          int base = 0;
          if (lane == 31) base = atomicAdd(nhit, total);
          base = __shfl_sync(0xffffffffu, base, 31);
          int pos = base + incl - mine;
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            if (hm & (1u << i)) {
              if (pos < HCAP) hits[pos] = ((u32)(row - (r0 - 1)) << 16) | (u32)(cg * VEC + i);
              ++pos;
            }
          }
        }
      }
    }
    if (tid == 0 && (k & 7) == 4) {
      s_ref_thr = tq;
      s_ref_g = tq_g;   // tagged with the group it was read for
    }
    __syncthreads();
    const int nh = dense0 ? HCAP + 1 : *nhit;
    u64 flags = 0;
    if (nh > 0) {
      if (nh <= HCAP) {
        // ---- 3x3 max-pool NMS on the recorded hits (utils/decode.py:5-10: keep iff no neighbour is larger)
        for (int i = tid; i < nh; i += NT) {
          const u32 h = hits[i];
          const int trow = (int)(h >> 16), col = (int)(h & 0xFFFFu);
          const int row = trow + (r0 - 1);
          const float x = tile[(size_t)trow * a.W + col];
          bool peak = true;
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            const int gy = row + dy;
            if (gy < 0 || gy >= a.H) continue;
            const float* np = tile + (size_t)(trow + dy) * a.W;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int gx = col + dx;
              if (gx < 0 || gx >= a.W) continue;
              if (np[gx] > x) peak = false;
            }
          }
          if (peak) s_list[atomicAdd(&s_cnt, 1)] = make_key(x, flat_base + (u32)(row * a.W + col));
        }
      } else {
        // ---- dense path: separable sliding-window 3x3 max over the thread's rows -----------------------
        if (rs < re) {
          const float NEG = -INFINITY;
          float hp2[VEC], hp1[VEC], cp1[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) hp2[i] = hp1[i] = cp1[i] = NEG;
          for (int gy = rs - 1; gy <= re; ++gy) {
            float cv[VEC], hmx[VEC];
            if (gy >= 0 && gy < a.H) {
              const float* rowp = tile + (size_t)(gy - (r0 - 1)) * a.W + cg * VEC;
              if constexpr (VEC == 4) {
                const float4 q = *reinterpret_cast<const float4*>(rowp);
                cv[0] = q.x; cv[1 % VEC] = q.y; cv[2 % VEC] = q.z; cv[3 % VEC] = q.w;
              } else {
                cv[0] = rowp[0];
              }
              const float l = cg > 0 ? rowp[-1] : NEG;
              const float r = cg < W4 - 1 ? rowp[VEC] : NEG;
#pragma unroll
              for (int i = 0; i < VEC; ++i) {
                const float lft = i == 0 ? l : cv[(i - 1 + VEC) % VEC];
                const float rgt = i == VEC - 1 ? r : cv[(i + 1) % VEC];
                hmx[i] = max3f(lft, cv[i], rgt);
              }
            } else {
#pragma unroll
              for (int i = 0; i < VEC; ++i) cv[i] = hmx[i] = NEG;
            }
            if (gy >= rs + 1) {
              const int row = gy - 1;
#pragma unroll
              for (int i = 0; i < VEC; ++i) {
                const float m = max3f(hp2[i], hp1[i], hmx[i]);
                const float x = cp1[i];
                if (x == m && x > 0.f) {
                  const u64 key = make_key(x, flat_base + (u32)(row * a.W + cg * VEC + i));
                  if (key >= thr) {
                    flags |= 1ull << ((row - rs) * VEC + i);
                    const int slot = atomicAdd(&s_cnt, 1);
                    if (slot < LC) s_list[slot] = key;
                  }
                }
              }
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              hp2[i] = hp1[i];
              hp1[i] = hmx[i];
              cp1[i] = cv[i];
            }
          }
        }
      }
      __syncthreads();
    }
    int n = (nh > 0) ? s_cnt : carry;

    if (n > LC) {   // (dense path only) more live survivors than the list holds: exact local top-K
      const FlagCands<VEC> fc{tile, flags, rs, cg, a.W, r0, flat_base};
      const u64 kth = block_kth_key(fc, a.K, sc);
      if (tid == 0) s_misc[2] = carry;
      __syncthreads();
      fc.for_each([&](u64 key) {
        if (key >= kth) s_list[atomicAdd(&s_misc[2], 1)] = key;   // exactly K more entries
      });
      __syncthreads();
      n = s_misc[2];
      if (tid == 0) {
        s_cnt = n;
        atomicMax(reinterpret_cast<unsigned long long*>(a.thr + g), (unsigned long long)kth);
      }
      if (kth > cur_thr) cur_thr = kth;
      __syncthreads();
    }
    carry = n;

    // ---- flush: tighten the group's bound, append what can still matter --------------------------------
    if (carry > 0 && (group_ends || carry >= FLUSH_AT)) {   // block-uniform
      const unsigned long long tf0 = a.debug ? gtimer() : 0ull;
      DBG_ADD(5, 1);
      const int nf = carry;
      const bool small = nf <= SMALL_FLUSH;
      // small flush: room for all nf keys is reserved right away (the merge filters by the final bound anyway),
      // so the reservation's round trip overlaps the histogram's
      if (small && tid == 32) s_misc[5] = atomicAdd(a.gcount + g, nf);
      int* oct = a.goct + (size_t)g * OCT_PAD;
      int* hist = a.ghist + (size_t)g * NCB_PAD;
      int* fine = a.gfine ? a.gfine + (size_t)g * NBF : nullptr;
      if (tid < NOCT) s_hist[tid] = 0;   // octave counts are aggregated per flush: every CTA of a group would
      __syncthreads();                   // otherwise hammer the same few global counters
      for (int i = tid; i < nf; i += NT) {
        const u64 key = s_list[i];
        if (key >= cur_thr) {
          const u32 bits = key_hi(key);
          const int cb = coarse_bin(bits);
          if (cb >= 0) {
            atomicAdd(&s_hist[cb >> 7], 1u);
            atomicAdd(hist + cb, 1);
            if (fine) atomicAdd(fine + cb * FSUB + (int)((bits >> 12) & (FSUB - 1)), 1);
          }
        }
      }
      __syncthreads();
      if (tid >= 64 && tid < 64 + NOCT && s_hist[tid - 64]) atomicAdd(oct + (tid - 64), (int)s_hist[tid - 64]);
      // No fence: the histogram is read right away, possibly without this CTA's own increments -- counts only
      // grow, so a stale read gives a weaker but still sound bound; the next flush catches up.
      if (warp == 0) {
        const u64 nt = warp_bound_from_hist(oct, hist, fine, a.K, cur_thr);
        if (lane == 0) {
          u64 t = cur_thr;
          if (nt > t) {
            atomicMax(reinterpret_cast<unsigned long long*>(a.thr + g), (unsigned long long)nt);
            t = nt;
          }
          s_thr = t;
          s_cnt = 0;
        }
      }
      __syncthreads();
      cur_thr = s_thr;
      if (small) {
        const int base = s_misc[5];
        for (int i = tid; i < nf; i += NT)
          if (base + i < a.gcap) a.lists[(size_t)g * a.gcap + base + i] = s_list[i];
      } else {
        // large flush (the first chunks of a group): append only what the new bound lets through
        int kept = 0;
        for (int i = tid; i < nf; i += NT) kept += s_list[i] >= cur_thr;
        kept = warp_sum(kept);
        if (lane == 0) s_red[warp] = kept;
        __syncthreads();
        kept = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) kept += s_red[i];
        __syncthreads();
        u64 cut = cur_thr;
        if (kept > a.acap) {   // plateaus: never hand more than acap (>= K) keys per flush to the group's list
          const SmemCands smc{s_list, nf};
          cut = block_kth_key(smc, a.K, sc);
          if (tid == 0) atomicMax(reinterpret_cast<unsigned long long*>(a.thr + g), (unsigned long long)cut);
          if (cut > cur_thr) cur_thr = cut;
        }
        for (int i0 = 0; i0 < nf; i0 += NT) {   // warp-aggregated append
          const int i = i0 + tid;
          const u64 key = i < nf ? s_list[i] : 0ull;
          const bool keep = i < nf && key >= cut;
          const u32 bal = __ballot_sync(0xffffffffu, keep);
          if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(a.gcount + g, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            const int pos = base + __popc(bal & ((1u << lane) - 1u));
            if (keep && pos < a.gcap) a.lists[(size_t)g * a.gcap + pos] = key;
          }
        }
      }
      carry = 0;
      __syncthreads();
      if (a.debug) DBG_ADD(4, gtimer() - tf0);
    }

    if (VEC == 4 && tid == 0 && c + NST < c_end) issue_load(lg, lpg, lband, stage);
    advance(lg, lpg, lband);
    ++done_in_group;
    if (k == 0) DBG_MARK(1);
    if (nh > 0) DBG_ADD(6, 1);
    if (c + 1 == c_end) DBG_MARK(2);

    // ---- group boundary: account for the chunks done; the CTA that completes the group merges it -------
    if (group_ends) {
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const int prev = atomicAdd(a.gdone + g, done_in_group);
        s_misc[3] = (prev + done_in_group == a.cpg);
      }
      __syncthreads();
      done_in_group = 0;
      if (s_misc[3]) {
        merge_group(a, g, s_mlist, s_list, s_out, sc);
        __syncthreads();
      }
    }
    if (group_ends && c + 1 < c_end) {   // entering the next group: fetch its bound (once per group and CTA)
      if (tid == 0) s_thr = ldcg_u64(a.thr + g + 1);
      __syncthreads();
      cur_thr = s_thr;
      __syncthreads();
    }
    advance(g, pg, band);
  }
  DBG_MARK(3);
}

// =================================================================================================
// decode_stream_kernel: the same scan with autonomous warps -- no block barrier in the steady state.
// =================================================================================================
// The barrier-per-chunk structure above makes every chunk pay for its slowest warp (a warp that holds a hit runs
// a long single-lane path while 7 warps wait; tools/decode_timeline.py).  Here one producer warp keeps the tile ring
// full (TMA bulk copies, empty/full mbarriers) and each of the 8 consumer warps scans ITS rows of every chunk on
// its own: private survivor list, private copy of the group's bound, its own flushes (histogram increments, one
// speculative round trip for the new bound, append).  While one warp waits on L2 the others keep streaming.  The
// exact top-K of a group is taken afterwards by decode_merge_kernel (one CTA per group).  Used for W % 4 == 0 and
// K <= STREAM_MAX_K; everything else stays on decode_scan_kernel.
constexpr int SW = 8;                  // consumer warps
constexpr int SNT = (SW + 1) * 32;     // + the producer warp
constexpr int WLC = 544;               // keys a warp can carry: every element of its share of a chunk (<= 512) + what it carried
constexpr int WFLUSH = 12;             // ... and the fill level at which it flushes (small: the bound of a group is
                                       // only as good as what its warps have published)
constexpr int STREAM_MAX_K = 512;

__global__ void __launch_bounds__(SNT, 2) decode_stream_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) u64 s_full[NST];
  __shared__ __align__(8) u64 s_empty[NST];
  __shared__ int s_wcnt[SW];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  u64* s_wlist = reinterpret_cast<u64*>(smem_raw + (size_t)NST * a.tile_bytes);   // [SW][WLC]
  const int c_begin = (int)((long long)blockIdx.x * a.total_chunks / gridDim.x);
  const int c_end = (int)((long long)(blockIdx.x + 1) * a.total_chunks / gridDim.x);

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], SW);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (tid < SW) s_wcnt[tid] = 0;
  if (a.debug && tid == 0 && blockIdx.x < 1024)
    for (int i = 0; i < DBG_SLOTS; ++i) g_dbg[blockIdx.x * DBG_SLOTS + i] = 0;
  __syncthreads();
  if (c_begin >= c_end) return;
  DBG_MARK(0);

  int g = c_begin / a.cpg, pg, band;
  {
    const int rem = c_begin - g * a.cpg;
    pg = rem / a.nbands;
    band = rem - pg * a.nbands;
  }
  auto advance = [&](int& g_, int& pg_, int& band_) {
    if (++band_ == a.nbands) {
      band_ = 0;
      if (++pg_ == a.PG) {
        pg_ = 0;
        ++g_;
      }
    }
  };

  if (warp == SW) {
    // =============================== producer ================================================================
    if (lane == 0) {
      for (int c = c_begin; c < c_end; ++c) {
        const int k = c - c_begin, stage = k & (NST - 1);
        if (k >= NST) {
          const unsigned long long tw0 = a.debug ? gtimer() : 0ull;
          mbar_wait(&s_empty[stage], (u32)((k / NST) - 1) & 1u);
          if (a.debug && blockIdx.x < 1024) g_dbg[blockIdx.x * DBG_SLOTS + 6] += gtimer() - tw0;   // producer: ring full
        }
        u32 fb;
        const float* plane = plane_ptr(a, g, pg, &fb);
        const int r0 = band * a.R, r1 = min(r0 + a.R, a.H);
        const int g0 = max(r0 - 1, 0), g1 = min(r1 + 1, a.H);
        const u32 bytes = (u32)((g1 - g0) * a.W * sizeof(float));
        float* tile = reinterpret_cast<float*>(smem_raw + (size_t)stage * a.tile_bytes);
        mbar_expect_tx(&s_full[stage], bytes);
        bulk_g2s(tile + (size_t)(g0 - (r0 - 1)) * a.W, plane + (size_t)g0 * a.W, bytes, &s_full[stage]);
        advance(g, pg, band);
      }
    }
    return;
  }

  // ================================= consumers ================================================================
  int* wcnt = &s_wcnt[warp];
  u64* wl = s_wlist + warp * WLC;
  const int W4 = a.W / 4;
  const int cg = tid % W4;
  const int seg = tid / W4;
  int n = 0;               // entries in the warp's list (warp-uniform)
  int cur_g = -1;
  u64 cur_thr = 0ull;
  u64 tq = 0ull;           // refresh of the group's bound in flight
  int tq_g = -1;

  // hand the warp's list to the group: histogram increments, new bound, append
  auto warp_flush = [&](int fg) {
    const unsigned long long tf0 = a.debug ? gtimer() : 0ull;
    DBG_ADD(5, 1);
    int* oct = a.goct + (size_t)fg * OCT_PAD;
    int* hist = a.ghist + (size_t)fg * NCB_PAD;
    int* fine = a.gfine ? a.gfine + (size_t)fg * NBF : nullptr;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      int fb = -1;     // fine-bin index of this lane's entry (cb * FSUB + sub), -1: not counted
      if (i < n) {
        const u64 key = wl[i];
        if (key >= cur_thr) {
          const u32 bits = key_hi(key);
          const int cb = coarse_bin(bits);
          if (cb >= 0) fb = cb * FSUB + (int)((bits >> 12) & (FSUB - 1));
        }
      }
      // The counters are hot: every warp of a group hits the same octave, and on plateaus the same bin.  Lanes
      // that share a counter elect one of them to add the whole count (three match rounds: fine, coarse, octave).
      u32 peers = __match_any_sync(0xffffffffu, fb);
      if (fb >= 0 && fine && lane == __ffs(peers) - 1) atomicAdd(fine + fb, __popc(peers));
      const int cbv = fb >= 0 ? fb / FSUB : -1;
      peers = __match_any_sync(0xffffffffu, cbv);
      if (cbv >= 0 && lane == __ffs(peers) - 1) atomicAdd(hist + cbv, __popc(peers));
      const int o = cbv >= 0 ? cbv >> 7 : -1;
      peers = __match_any_sync(0xffffffffu, o);
      if (o >= 0 && lane == __ffs(peers) - 1) atomicAdd(oct + o, __popc(peers));
    }
    u64 nt = warp_bound_from_hist(oct, hist, fine, a.K, cur_thr);
    if (n >= a.K) {
      // The histogram bounds by score only.  A list that alone holds K survivors (plateaus of equal scores, or
      // a first chunk) also yields an exact key -- score AND index -- below which nothing of this group matters.
      u64 kand = ~0ull, kor = 0ull;     // bits on which all keys agree need no search step
      for (int i = lane; i < n; i += 32) {
        kand &= wl[i];
        kor |= wl[i];
      }
      kand = ((u64)__reduce_and_sync(0xffffffffu, (u32)(kand >> 32)) << 32) | __reduce_and_sync(0xffffffffu, (u32)kand);
      kor = ((u64)__reduce_or_sync(0xffffffffu, (u32)(kor >> 32)) << 32) | __reduce_or_sync(0xffffffffu, (u32)kor);
      const u64 varying = kand ^ kor;
      u64 t = kand & ~varying;
      for (int bit = 63 - __clzll((long long)(varying | 1ull)); bit >= 0; --bit) {
        if (!((varying >> bit) & 1ull)) continue;
        const u64 cand = t | (1ull << bit);
        int cge = 0;
        for (int i = lane; i < n; i += 32) cge += (wl[i] >= cand);
        if (warp_sum(cge) >= a.K) t = cand;
      }
      if (t > nt) nt = t;
    }
    if (nt > cur_thr) {
      if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(a.thr + fg), (unsigned long long)nt);
      cur_thr = nt;
    }
    // append what the new bound lets through (compacted: count, reserve, write)
    int kept = 0;
    for (int i = lane; i < n; i += 32) kept += (wl[i] >= cur_thr);
    int incl = kept;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 0 && total) base = atomicAdd(a.gcount + fg, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    int pos = base + incl - kept;
    for (int i = lane; i < n; i += 32) {
      const u64 key = wl[i];
      if (key >= cur_thr) {
        if (pos < a.gcap) a.lists[(size_t)fg * a.gcap + pos] = key;
        ++pos;
      }
    }
    n = 0;
    if (lane == 0) *wcnt = 0;
    __syncwarp();
    if (a.debug) DBG_ADD(4, gtimer() - tf0);
  };

  for (int c = c_begin; c < c_end; ++c) {
    const int k = c - c_begin, stage = k & (NST - 1);
    if (g != cur_g) {   // entering a group: hand over what belongs to the previous one, fetch the new bound
      if (n > 0) warp_flush(cur_g);
      cur_g = g;
      cur_thr = ldcg_u64(a.thr + g);
      tq_g = -1;
    }
    // background refresh of the bound: issued every 4 chunks, consumed 2 chunks later
    if ((k & 3) == 2 && tq_g == g && tq > cur_thr) cur_thr = tq;
    if ((k & 3) == 0) {
      tq = ldcg_u64(a.thr + g);
      tq_g = g;
    }
    u32 flat_base;
    plane_ptr(a, g, pg, &flat_base);
    const int r0 = band * a.R, r1 = min(r0 + a.R, a.H);
    const float* tile = reinterpret_cast<const float*>(smem_raw + (size_t)stage * a.tile_bytes);
    const u64 thr = cur_thr;
    const float ts = __uint_as_float(key_hi(thr));
    const int rs = r0 + seg * a.rpt;
    const int re = min(rs + a.rpt, r1);
    const int n_before = n;

    {
      const unsigned long long tw0 = a.debug ? gtimer() : 0ull;
      mbar_wait(&s_full[stage], (u32)(k / NST) & 1u);
      if (a.debug) DBG_ADD(7, gtimer() - tw0);
    }

    // one look at the thread's rows: largest value against the bound, one vote per warp
    float mt = 0.f;
    for (int row = rs; row < re; ++row) {
      const float4 q = *reinterpret_cast<const float4*>(tile + (size_t)(row - (r0 - 1)) * a.W + cg * 4);
      mt = fmaxf(mt, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
    }
    if (__any_sync(0xffffffffu, mt >= ts && mt > 0.f)) {
      for (int row = rs; row < re; ++row) {
        const float4 q = *reinterpret_cast<const float4*>(tile + (size_t)(row - (r0 - 1)) * a.W + cg * 4);
        const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x = v[i];
          if (!(x >= ts && x > 0.f)) continue;
          const int col = cg * 4 + i;
          const u64 key = make_key(x, flat_base + (u32)(row * a.W + col));
          if (key < thr) continue;
          bool peak = true;   // 3x3 max-pool NMS (utils/decode.py:5-10): keep iff no neighbour is larger
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            const int gy = row + dy;
            if (gy < 0 || gy >= a.H) continue;
            const float* np = tile + (size_t)(gy - (r0 - 1)) * a.W;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int gx = col + dx;
              if (gx < 0 || gx >= a.W) continue;
              if (np[gx] > x) peak = false;
            }
          }
          if (peak) {
            const int slot = atomicAdd(wcnt, 1);
            if (slot < WLC) wl[slot] = key;
          }
        }
      }
      __syncwarp();
      n = *reinterpret_cast<volatile int*>(wcnt);
      if (n > WLC) n = WLC;   // cannot happen: a warp owns at most 512 elements of a chunk (host-side plan)
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[stage]);   // this warp is done with the tile
    // flush when the list fills up, and right after a chunk scanned without any bound (the group needs one)
    if (n >= a.wflush || (n > 0 && thr == 0ull)) warp_flush(g);
    if (k == 0) DBG_MARK(1);
    if (c + 1 == c_end) DBG_MARK(2);
    advance(g, pg, band);
  }
  if (n > 0) warp_flush(cur_g);
  DBG_MARK(3);
}

// one CTA per group: the exact, sorted top-K of the candidates the streaming scan appended
__global__ void __launch_bounds__(NT) decode_merge_kernel(const ScanArgs a) {
  __shared__ __align__(8) u64 s_mlist[MCAP];
  __shared__ __align__(8) u64 s_sel[MAX_K];
  __shared__ __align__(8) u64 s_out[MAX_K];
  __shared__ __align__(8) u64 s_keyred[2 * NW];
  __shared__ u32 s_hist[256];
  __shared__ int s_red[NW];
  __shared__ int s_misc[8];
  const Scratch sc{s_hist, s_red, s_misc, s_keyred};
  merge_group(a, blockIdx.x, s_mlist, s_sel, s_out, sc);
}

// =================================================================================================
// multi_pose association: one CTA per (joint j, image b)  (decode/multi_pose.py:15-94)
// =================================================================================================
struct PoseArgs {
  const float *heat, *wh, *kps, *reg, *hm_hp, *hp_offset;
  float* out;
  int B, J, H, W, K;
  const u64* topk;   // [B*(1+J)][K] sorted keys per plane (plane 0 = detections, 1+j = joint j)
  const int* have;   // [B*(1+J)]
};

__global__ void __launch_bounds__(NT) multi_pose_assoc_kernel(const PoseArgs a) {
  __shared__ __align__(8) u64 s_det[MAX_K];
  __shared__ float s_hx[MAX_K], s_hy[MAX_K], s_hs[MAX_K];

  const int j = blockIdx.x, b = blockIdx.y;
  const int P = 1 + a.J, HW = a.H * a.W, K = a.K;
  const int tid = threadIdx.x;

  // (1) detections: exact top-K of the (single-class) heat plane, zero-filled like torch.topk
  for (int i = tid; i < K; i += NT) s_det[i] = a.topk[(size_t)(b * P) * K + i];
  // (2) joint candidates: exact top-K of hm_hp plane j (utils/decode.py:31-40) + offsets + threshold
  {
    const int gj = b * P + 1 + j;
    const int have = a.have[gj];
    for (int m = tid; m < K; m += NT) {
      float sc = 0.f, hx = 0.f, hy = 0.f;
      if (m < have) {
        const u64 key = a.topk[(size_t)gj * K + m];
        sc = __uint_as_float(key_hi(key));
        const int pix = (int)key_idx(key) % HW;
        hx = (float)(pix % a.W);
        hy = (float)(pix / a.W);
        if (a.hp_offset) {
          hx = __fadd_rn(hx, __ldg(a.hp_offset + ((size_t)b * 2 + 0) * HW + pix));
          hy = __fadd_rn(hy, __ldg(a.hp_offset + ((size_t)b * 2 + 1) * HW + pix));
        } else {
          hx = __fadd_rn(hx, 0.5f);
          hy = __fadd_rn(hy, 0.5f);
        }
      }
      // multi_pose.py:58-61 (entries beyond `have` are zero-score fill: mask = 0)
      const float mask = sc > 0.1f ? 1.f : 0.f;
      const float om = __fsub_rn(1.f, mask);
      s_hs[m] = __fadd_rn(__fmul_rn(om, -1.f), __fmul_rn(mask, sc));
      s_hy[m] = __fadd_rn(__fmul_rn(om, -10000.f), __fmul_rn(mask, hy));
      s_hx[m] = __fadd_rn(__fmul_rn(om, -10000.f), __fmul_rn(mask, hx));
    }
    __syncthreads();
  }
  // (3) per detection: regressed joint, nearest candidate, gating, blend
  const int ncol = 3 * a.J + 6;
  for (int k = tid; k < K; k += NT) {
    const u64 key = s_det[k];
    const float score = __uint_as_float(key_hi(key));
    const int pix = (int)key_idx(key) % HW;   // C == 1
    const float xi = (float)(pix % a.W), yi = (float)(pix / a.W);
    const float kx = __fadd_rn(__ldg(a.kps + ((size_t)b * 2 * a.J + 2 * j) * HW + pix), xi);
    const float ky = __fadd_rn(__ldg(a.kps + ((size_t)b * 2 * a.J + 2 * j + 1) * HW + pix), yi);
    float xs, ys;
    if (a.reg) {
      xs = __fadd_rn(xi, __ldg(a.reg + ((size_t)b * 2 + 0) * HW + pix));
      ys = __fadd_rn(yi, __ldg(a.reg + ((size_t)b * 2 + 1) * HW + pix));
    } else {
      xs = __fadd_rn(xi, 0.5f);
      ys = __fadd_rn(yi, 0.5f);
    }
    const float hw = __fmul_rn(__ldg(a.wh + ((size_t)b * 2 + 0) * HW + pix), 0.5f);
    const float hh = __fmul_rn(__ldg(a.wh + ((size_t)b * 2 + 1) * HW + pix), 0.5f);
    const float l = __fsub_rn(xs, hw), t = __fsub_rn(ys, hh);
    const float r = __fadd_rn(xs, hw), bt = __fadd_rn(ys, hh);
    float best = INFINITY;
    int bi = 0;
    for (int m = 0; m < K; ++m) {
      const float dx = __fsub_rn(kx, s_hx[m]);
      const float dy = __fsub_rn(ky, s_hy[m]);
      const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      if (d < best) {   // strict: first index wins ties (dist.min, multi_pose.py:68)
        best = d;
        bi = m;
      }
    }
    const float hx = s_hx[bi], hy = s_hy[bi], hs = s_hs[bi];
    const float lim = __fmul_rn(fmaxf(__fsub_rn(bt, t), __fsub_rn(r, l)), 0.3f);
    const bool gate = (hx < l) || (hx > r) || (hy < t) || (hy > bt) || (hs < 0.1f) || (best > lim);
    const float mk = gate ? 1.f : 0.f;
    const float om = __fsub_rn(1.f, mk);
    float* row = a.out + ((size_t)b * K + k) * ncol;
    row[5 + 2 * j] = __fadd_rn(__fmul_rn(om, hx), __fmul_rn(mk, kx));
    row[5 + 2 * j + 1] = __fadd_rn(__fmul_rn(om, hy), __fmul_rn(mk, ky));
    // hm_score is [B,J,K,1] .view(B,K,J) WITHOUT permute (multi_pose.py:90)
    const int f = j * K + k;
    a.out[((size_t)b * K + f / a.J) * ncol + (2 * a.J + 6) + f % a.J] = __fmul_rn(hs, om);
    if (j == 0) {
      row[0] = l; row[1] = t; row[2] = r; row[3] = bt;
      row[4] = score;
      row[5 + 2 * a.J] = 0.f;   // clses: (ind / K).int() with C == 1
    }
  }
}

// =================================================================================================
// host side
// =================================================================================================
struct ScanGeom {
  int vec, R, nbands, rpt, acap, tile_bytes;
  size_t smem;
};

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

static bool plan_scan(int H, int W, int K, bool aligned, ScanGeom* g) {
  const int vec = (aligned && W % 4 == 0) ? 4 : 1;
  const int W4 = W / vec;
  if (W4 > NT || W4 < 1) return false;
  const int S = NT / W4;                       // row segments scanned in parallel
  const int max_rpt = 64 / vec;                // survivor flags are one 64-bit mask per thread
  int rpt = env_int("CNB_DECODE_RPT", 0);
  if (rpt < 1) {                               // ~16 KB of rows per chunk
    rpt = (4096 / W + S / 2) / S;
    if (rpt < 1) rpt = 1;
  }
  if (rpt > max_rpt) rpt = max_rpt;
  int acap = K > 256 ? K : 256;
  const int nstage = vec == 4 ? NST : 1;
  const size_t fixed = (size_t)(MCAP + (HCAP + FLUSH_AT) + MAX_K) * sizeof(u64) + 3 * HCAP * sizeof(u32);
  int R, tile_bytes;
  size_t smem;
  for (;;) {
    R = S * rpt;
    if (R > H) R = H;
    tile_bytes = (int)((((size_t)(R + 2) * W * sizeof(float)) + 127) & ~(size_t)127);
    smem = (size_t)nstage * tile_bytes + fixed;
    if (smem <= 110 * 1024 || rpt == 1) break;
    rpt /= 2;
  }
  if (smem > 110 * 1024) return false;
  g->vec = vec;
  g->R = R;
  g->nbands = (H + R - 1) / R;
  g->rpt = (R + S - 1) / S;
  g->acap = acap;
  g->tile_bytes = tile_bytes;
  g->smem = smem;
  return true;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// workspace: [thr | gcount | gdone | octave sums | coarse bins | fine bins] (zeroed per call) | lists | topk | have
struct WsLayout {
  size_t thr_off, gcount_off, gdone_off, goct_off, ghist_off, gfine_off, zero_bytes, lists_off, topk_off, have_off, total;
  bool fine;
  int gcap;
};
// Fine sub-bins pay off when one coarse bin near the top can hold many more than K survivors: large groups
// (ctdet: C planes per image); single-plane groups (multi_pose) resolve well enough with the coarse bins.
static bool fine_for(int PG) { return PG >= 8; }

static WsLayout ws_layout(int G, int PG, int K, const ScanGeom& g, bool want_topk) {
  WsLayout w;
  // a warp flush appends at most K keys (an exact bound is taken whenever its list holds K or more)
  w.gcap = PG * g.nbands * (SW * K > g.acap ? SW * K : g.acap);
  w.thr_off = 0;
  w.gcount_off = align_up((size_t)G * sizeof(u64), 256);
  w.gdone_off = w.gcount_off + align_up((size_t)G * sizeof(int), 256);
  w.goct_off = w.gdone_off + align_up((size_t)G * sizeof(int), 256);
  w.ghist_off = w.goct_off + align_up((size_t)G * OCT_PAD * sizeof(int), 256);
  w.fine = fine_for(PG);
  w.gfine_off = w.ghist_off + align_up((size_t)G * NCB_PAD * sizeof(int), 256);
  w.zero_bytes = w.gfine_off + (w.fine ? align_up((size_t)G * NBF * sizeof(int), 256) : 0);
  w.lists_off = w.zero_bytes;
  w.topk_off = w.lists_off + align_up((size_t)G * w.gcap * sizeof(u64), 256);
  w.have_off = w.topk_off + (want_topk ? align_up((size_t)G * K * sizeof(u64), 256) : 0);
  w.total = w.have_off + (want_topk ? align_up((size_t)G * sizeof(int), 256) : 0);
  return w;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n < 1) n = 1;
  }
  return n;
}

// the warp-autonomous scan + the per-group merge (W % 4 == 0, K <= STREAM_MAX_K; CNB_DECODE_IMPL=v2 disables it)
static bool stream_ok(const ScanGeom& g, int K) {
  static const bool off = [] { const char* e = getenv("CNB_DECODE_IMPL"); return e && e[0] == 'v' && e[1] == '2'; }();
  return !off && g.vec == 4 && K <= STREAM_MAX_K && g.rpt * 4 <= 16;   // a warp owns <= 512 elements of a chunk
}

static cudaError_t launch_stream(const ScanArgs& a, cudaStream_t st) {
  static PerDeviceOnce once;
  const size_t smem = (size_t)NST * a.tile_bytes + (size_t)SW * WLC * sizeof(u64);
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(decode_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e != cudaSuccess) return e;
    once.mark();
  }
  int grid = env_int("CNB_DECODE_CTAS", 2 * num_sms());
  if (grid > a.total_chunks) grid = a.total_chunks;
  if (grid < 1) grid = 1;
  decode_stream_kernel<<<grid, SNT, smem, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  decode_merge_kernel<<<a.G, NT, 0, st>>>(a);
  return cudaGetLastError();
}

template <int VEC>
static cudaError_t launch_scan(const ScanArgs& a, size_t smem, cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(decode_scan_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         112 * 1024);
    if (e != cudaSuccess) return e;
    once.mark();
  }
  int grid = env_int("CNB_DECODE_CTAS", 2 * num_sms());
  if (grid > a.total_chunks) grid = a.total_chunks;
  if (grid < 1) grid = 1;
  decode_scan_kernel<VEC><<<grid, NT, smem, st>>>(a);
  return cudaGetLastError();
}

static int debug_on() {
  static const int on = env_int("CNB_DECODE_DEBUG", 0);
  return on;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static void fill_scan_args(ScanArgs& a, const ScanGeom& g, const WsLayout& w, unsigned char* ws) {
  a.R = g.R; a.nbands = g.nbands; a.rpt = g.rpt; a.acap = g.acap;
  a.cpg = a.PG * g.nbands;
  a.total_chunks = a.G * a.cpg;
  a.gcap = w.gcap;
  a.tile_bytes = g.tile_bytes;
  a.thr = (u64*)(ws + w.thr_off);
  a.gcount = (int*)(ws + w.gcount_off);
  a.gdone = (int*)(ws + w.gdone_off);
  a.goct = (int*)(ws + w.goct_off);
  a.ghist = (int*)(ws + w.ghist_off);
  a.gfine = w.fine ? (int*)(ws + w.gfine_off) : nullptr;
  a.lists = (u64*)(ws + w.lists_off);
}

}  // namespace
}  // namespace cnb

using namespace cnb;

// debug aid (not part of the public header): copies the per-CTA phase timestamps of the last decode launch
extern "C" int cnb_decode_debug_dump(unsigned long long* host, int n_words) {
  if (n_words > 1024 * DBG_SLOTS) n_words = 1024 * DBG_SLOTS;
  CNB_CUDA(cudaDeviceSynchronize());
  CNB_CUDA(cudaMemcpyFromSymbol(host, g_dbg, (size_t)n_words * sizeof(unsigned long long)));
  return CNB_OK;
}

extern "C" size_t cnb_ctdet_decode_workspace_bytes(int B, int C, int H, int W, int K) {
  ScanGeom g;
  if (B < 1 || C < 1 || K < 1 || K > MAX_K || !plan_scan(H, W, K, true, &g)) return 0;
  ScanGeom g1;   // the unaligned (scalar) fallback may band differently; take the larger of the two
  const size_t a = ws_layout(B, C, K, g, false).total;
  const size_t b = plan_scan(H, W, K, false, &g1) ? ws_layout(B, C, K, g1, false).total : 0;
  return a > b ? a : b;
}

extern "C" int cnb_ctdet_decode(const float* heat, const float* wh, const float* reg, float* out,
                                int B, int C, int H, int W, int K, void* workspace,
                                size_t workspace_bytes, cnb_stream_t stream) {
  CNB_CHECK_ARG(heat && wh && out && workspace, "ctdet_decode: null pointer");
  CNB_CHECK_ARG(B >= 1 && C >= 1 && H >= 1 && W >= 1, "ctdet_decode: bad shape");
  CNB_CHECK_ARG(K >= 1 && K <= MAX_K, "ctdet_decode: K=%d outside [1,%d]", K, MAX_K);
  CNB_CHECK_ARG((long long)C * H * W >= K, "ctdet_decode: K=%d exceeds C*H*W", K);
  CNB_CHECK_ARG((long long)C * H * W < (1ll << 31), "ctdet_decode: C*H*W too large");
  ScanGeom g;
  CNB_CHECK_ARG(plan_scan(H, W, K, aligned16(heat), &g), "ctdet_decode: unsupported map size %dx%d", H, W);
  CNB_CHECK_ARG((long long)B * C * g.nbands < (1ll << 31), "ctdet_decode: too many chunks");
  const WsLayout w = ws_layout(B, C, K, g, false);
  if (workspace_bytes < w.total) {
    set_error("ctdet_decode: workspace %zu < %zu bytes", workspace_bytes, w.total);
    return CNB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = (unsigned char*)workspace;
  ScanArgs a{};
  a.t0 = heat; a.t1 = nullptr; a.C0 = C; a.C1 = 0;
  a.PG = C; a.G = B;
  a.B = B; a.H = H; a.W = W; a.K = K;
  fill_scan_args(a, g, w, ws);
  a.topk = nullptr; a.have = nullptr;
  a.wh = wh; a.reg = reg; a.out = out;
  a.fuse_ctdet = 1;
  a.debug = debug_on();
  a.wflush = env_int("CNB_DECODE_WFLUSH", WFLUSH);
  CNB_CUDA(cudaMemsetAsync(ws, 0, w.zero_bytes, st));
  cudaError_t e;
  if (stream_ok(g, K)) {
    e = launch_stream(a, st);
    count_launch();
  } else {
    e = g.vec == 4 ? launch_scan<4>(a, g.smem, st) : launch_scan<1>(a, g.smem, st);
  }
  if (e != cudaSuccess) {
    set_error("ctdet_decode launch failed: %s", cudaGetErrorString(e));
    return CNB_ERR_CUDA;
  }
  count_launch();
  return CNB_OK;
}

extern "C" size_t cnb_multi_pose_decode_workspace_bytes(int B, int J, int H, int W, int K) {
  ScanGeom g;
  if (B < 1 || J < 1 || K < 1 || K > MAX_K || !plan_scan(H, W, K, true, &g)) return 0;
  ScanGeom g1;
  const size_t a = ws_layout(B * (1 + J), 1, K, g, true).total;
  const size_t b = plan_scan(H, W, K, false, &g1) ? ws_layout(B * (1 + J), 1, K, g1, true).total : 0;
  return a > b ? a : b;
}

extern "C" int cnb_multi_pose_decode(const float* heat, const float* wh, const float* kps,
                                     const float* reg, const float* hm_hp, const float* hp_offset,
                                     float* out, int B, int J, int H, int W, int K, void* workspace,
                                     size_t workspace_bytes, cnb_stream_t stream) {
  CNB_CHECK_ARG(heat && wh && kps && out && workspace, "multi_pose_decode: null pointer");
  CNB_CHECK_ARG(hm_hp, "multi_pose_decode: hm_hp is required (reference raises NameError, multi_pose.py:94)");
  CNB_CHECK_ARG(B >= 1 && J >= 1 && H >= 1 && W >= 1, "multi_pose_decode: bad shape");
  CNB_CHECK_ARG(K >= 1 && K <= MAX_K, "multi_pose_decode: K=%d outside [1,%d]", K, MAX_K);
  CNB_CHECK_ARG((long long)H * W >= K, "multi_pose_decode: K=%d exceeds H*W", K);
  CNB_CHECK_ARG((long long)J * H * W < (1ll << 31), "multi_pose_decode: J*H*W too large");
  CNB_CHECK_ARG(B <= 65535, "multi_pose_decode: B too large");
  ScanGeom g;
  CNB_CHECK_ARG(plan_scan(H, W, K, aligned16(heat) && aligned16(hm_hp), &g),
                "multi_pose_decode: unsupported map size %dx%d", H, W);
  const int P = 1 + J;
  CNB_CHECK_ARG((long long)B * P * g.nbands < (1ll << 31), "multi_pose_decode: too many chunks");
  const WsLayout w = ws_layout(B * P, 1, K, g, true);
  if (workspace_bytes < w.total) {
    set_error("multi_pose_decode: workspace %zu < %zu bytes", workspace_bytes, w.total);
    return CNB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = (unsigned char*)workspace;
  ScanArgs a{};
  a.t0 = heat; a.t1 = hm_hp; a.C0 = 1; a.C1 = J;
  a.PG = 1; a.G = B * P;
  a.B = B; a.H = H; a.W = W; a.K = K;
  fill_scan_args(a, g, w, ws);
  a.topk = (u64*)(ws + w.topk_off);
  a.have = (int*)(ws + w.have_off);
  a.wh = nullptr; a.reg = nullptr; a.out = nullptr;
  a.fuse_ctdet = 0;
  a.debug = debug_on();
  a.wflush = env_int("CNB_DECODE_WFLUSH", WFLUSH);
  CNB_CUDA(cudaMemsetAsync(ws, 0, w.zero_bytes, st));
  cudaError_t e;
  if (stream_ok(g, K)) {
    e = launch_stream(a, st);
    count_launch();
  } else {
    e = g.vec == 4 ? launch_scan<4>(a, g.smem, st) : launch_scan<1>(a, g.smem, st);
  }
  if (e != cudaSuccess) {
    set_error("multi_pose_decode scan launch failed: %s", cudaGetErrorString(e));
    return CNB_ERR_CUDA;
  }
  count_launch();
  PoseArgs pa{heat, wh, kps, reg, hm_hp, hp_offset, out, B, J, H, W, K, a.topk, a.have};
  multi_pose_assoc_kernel<<<dim3(J, B), NT, 0, st>>>(pa);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
