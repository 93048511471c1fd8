// Fused decode for CenterNet heads: 3x3 max-pool NMS + top-K peak extraction + offset gather.
//
// Replaces (reference file:line, all under CenterNet/):
//   utils/decode.py:5-10   _nms                 -> plane_scan_kernel (sliding 3x3 max in smem)
//   utils/decode.py:13-28  _topk                -> per-plane candidate lists + per-image merge
//   utils/decode.py:31-40  _topk_channel        -> exact per-plane lists
//   utils/decode.py:59-63  _transpose_and_gather_feat -> direct NCHW gathers at the K winners
//   decode/ctdet.py:6-38   ctdet_decode         -> plane_scan_kernel<.., FUSE_CTDET> (one launch;
//                                                  the last CTA of an image merges and writes [K,6])
//   decode/multi_pose.py:7-96 multi_pose_decode -> plane_scan_kernel + multi_pose_assoc_kernel
//
// Data layout in HBM: heat maps NCHW fp32 exactly as the heads emit them.  Every (image, class)
// plane (or a band of rows of it) is streamed ONCE from HBM into shared memory with a 1-D TMA bulk
// copy; algorithmic traffic = C*H*W*4 B per image (+ a few KB of candidates), see DESIGN.md.
//
// Ordering: all selections use 64-bit keys  (score_bits << 32) | (0xFFFFFFFF - flat_index)  so that
// "larger key" == (higher score, then lower flat index); keys are distinct, which makes every
// selection deterministic.  Scores must be >= 0 (sigmoid outputs, as at every reference call site).
#include "cnb_common.cuh"

namespace cnb {
namespace {

constexpr int NT = 256;          // threads per CTA
constexpr int NW = NT / 32;
constexpr int SURV_CAP = 1024;   // survivors kept in smem during a merge
constexpr int MAX_K = 512;

struct ScanArgs {
  const float* t0;   // [B, C0, H, W]
  const float* t1;   // [B, C1, H, W] or nullptr
  int C0, C1;
  int B, H, W, K;
  int R;        // rows per band
  int nbands;
  int rpt;      // rows per thread segment
  int cap;      // list capacity (keys)
  int exact0, exact1;   // emit exact sorted top-K lists for planes of t0 / t1
  u64* lists;   // per-list mode: [B*P*nbands][cap]; fused ctdet: [B][P*nbands*cap] append arrays
  int* counts;  // per-list mode: [B*P*nbands];      fused ctdet: [B] append cursors (zeroed)
  // fused ctdet epilogue
  int* done;    // [B] arrival counters (zeroed by the host wrapper)
  const float* wh;
  const float* reg;
  float* out;   // [B,K,6]
};

struct SelSmem {
  u64* list;     // [SURV_CAP]
  u64* out;      // [MAX_K]
  u32* vals;     // [NT]
  int* red;      // [NW]
  int* misc;     // [4]  (0: n, 1: flag, 2..3: bound lo/hi)
};

__device__ __forceinline__ u64 make_key(float score, u32 flat) {
  return ((u64)__float_as_uint(score) << 32) | (u64)(0xFFFFFFFFu - flat);
}
__device__ __forceinline__ u32 key_hi(u64 k) { return (u32)(k >> 32); }
__device__ __forceinline__ u32 key_idx(u64 k) { return 0xFFFFFFFFu - (u32)(k & 0xFFFFFFFFull); }

__device__ __forceinline__ int block_sum(int v, int* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int i = 0; i < NW; ++i) t += red[i];
  __syncthreads();
  return t;
}

// K-th largest (1-based) of the NT per-thread values; 0 when fewer than K values are non-zero or
// K > NT.  A sound lower bound for the K-th largest candidate of the block: the thread maxima are
// K distinct candidates >= the returned value.
__device__ __forceinline__ u32 block_kth_of_thread_max(u32 tmax, int K, const SelSmem& sm) {
  sm.vals[threadIdx.x] = tmax;
  __syncthreads();
  if (threadIdx.x < 32) {
    u32 v[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) v[i] = sm.vals[threadIdx.x + 32 * i];
    u32 t = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const u32 c = t | (1u << bit);
      int n = 0;
#pragma unroll
      for (int i = 0; i < NW; ++i) n += (v[i] >= c);
      n = warp_sum(n);
      if (n >= K) t = c;
    }
    if (threadIdx.x == 0) sm.misc[2] = (int)t;
  }
  __syncthreads();
  const u32 b = (u32)sm.misc[2];
  __syncthreads();
  return b;
}

// Exact ranking of the n (<= SURV_CAP) distinct keys in sm.list: key with rank r < K goes to sm.out[r].
__device__ __forceinline__ void block_rank_to_out(int n, int K, const SelSmem& sm) {
  for (int i = threadIdx.x; i < n; i += NT) {
    const u64 k = sm.list[i];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (sm.list[j] > k);
    if (r < K) sm.out[r] = k;
  }
  __syncthreads();
}

// ---- candidate iteration over a scanned band (flags -> scores in the smem tile) ------------------
template <int VEC>
struct BandCands {
  const float* tile;   // tile row 0 == global row r0-1
  u64 flags;
  int rs, cg, W, r0;
  u32 flat_base;       // plane_in_tensor*H*W
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

// ---- candidate iteration over global lists (merge) -----------------------------------------------
struct ListCands {
  const u64* lists;
  const int* counts;
  int nlists, cap;
  template <class F>
  __device__ __forceinline__ void for_each(F f) const {
    for (int l = 0; l < nlists; ++l) {
      const int n = __ldcg(counts + l);
      for (int s = threadIdx.x; s < n; s += NT) f(__ldcg(lists + (size_t)l * cap + s));
    }
  }
};

// Exact K-th largest 64-bit key over all candidates of the block (slow path; keys distinct).
template <class Cands>
__device__ u64 block_exact_kth(const Cands& c, int K, const SelSmem& sm) {
  u64 t = 0;
  for (int bit = 63; bit >= 0; --bit) {
    const u64 cand = t | (1ull << bit);
    int n = 0;
    c.for_each([&](u64 k) { n += (k >= cand); });
    n = block_sum(n, sm.red);
    if (n >= K) t = cand;
  }
  return t;
}

// Put a superset (<= limit keys) of the block's top-K candidates into sm.list; returns its size.
// `total` = number of candidates of the whole block, tmax = this thread's best score bits.
template <class Cands>
__device__ int block_collect(const Cands& c, int total, u32 tmax, int K, int limit, const SelSmem& sm) {
  u32 bound = 0;
  if (total > limit) bound = block_kth_of_thread_max(tmax, K, sm);
  if (threadIdx.x == 0) sm.misc[0] = 0;
  __syncthreads();
  if (tmax >= bound && tmax != 0) {
    c.for_each([&](u64 k) {
      if (key_hi(k) >= bound) {
        const int slot = atomicAdd(&sm.misc[0], 1);
        if (slot < limit) sm.list[slot] = k;
      }
    });
  }
  __syncthreads();
  int n = sm.misc[0];
  __syncthreads();
  if (n > limit) {  // adversarial distribution: fall back to an exact bitwise search
    const u64 t = block_exact_kth(c, K, sm);
    if (threadIdx.x == 0) sm.misc[0] = 0;
    __syncthreads();
    c.for_each([&](u64 k) {
      if (k >= t) {
        const int slot = atomicAdd(&sm.misc[0], 1);
        if (slot < limit) sm.list[slot] = k;
      }
    });
    __syncthreads();
    n = sm.misc[0];
    __syncthreads();
  }
  return n;
}

// Exact sorted top-K of several global candidate lists -> sm.out[0..ret)
__device__ int block_select_from_lists(const ListCands& c, int K, const SelSmem& sm) {
  u32 tmax = 0;
  int cnt = 0;
  c.for_each([&](u64 k) {
    tmax = max(tmax, key_hi(k));
    ++cnt;
  });
  const int total = block_sum(cnt, sm.red);
  const int n = block_collect(c, total, tmax, K, SURV_CAP, sm);
  block_rank_to_out(n, K, sm);
  return min(n, K);
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

// Reference semantics when an image has fewer than K positive peaks: torch.topk then returns
// zero-valued entries of heat*keep; we define their order as flat index ascending.
__device__ void block_zero_fill(const float* img, int n_elems, int H, int W, int have, int K,
                                const SelSmem& sm) {
  int filled = have;
  for (int base = 0; base < n_elems && filled < K; base += NT) {
    const int idx = base + threadIdx.x;
    const bool z = idx < n_elems && !is_positive_peak(img, idx, H, W);
    const u32 bal = __ballot_sync(0xffffffffu, z);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm.red[w] = __popc(bal);
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int c = sm.red[i];
      if (i < w) before += c;
      tot += c;
    }
    const int slot = filled + before + __popc(bal & ((1u << lane) - 1u));
    if (z && slot < K) sm.out[slot] = make_key(0.f, (u32)idx);
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

// =================================================================================================
// plane_scan_kernel: one CTA per (image, plane, band of rows).
// =================================================================================================
template <int VEC, bool FUSE_CTDET>
__global__ void __launch_bounds__(NT) plane_scan_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) u64 s_mbar;
  __shared__ u32 s_vals[NT];
  __shared__ int s_red[NW];
  __shared__ int s_misc[4];

  const int P = a.C0 + a.C1;
  const int band = blockIdx.x % a.nbands;
  const int bp = blockIdx.x / a.nbands;
  const int p = bp % P;
  const int b = bp / P;
  const int HW = a.H * a.W;
  const bool second = p >= a.C0;
  const int pl = second ? p - a.C0 : p;
  const float* plane = second ? a.t1 + ((size_t)b * a.C1 + pl) * HW : a.t0 + ((size_t)b * a.C0 + pl) * HW;
  const bool exact = second ? (a.exact1 != 0) : (a.exact0 != 0);

  const int r0 = band * a.R;
  const int r1 = min(r0 + a.R, a.H);
  const int g0 = max(r0 - 1, 0);
  const int g1 = min(r1 + 1, a.H);
  const int tile_rows = a.R + 2;
  float* tile = reinterpret_cast<float*>(smem_raw);
  const size_t tile_bytes = (((size_t)tile_rows * a.W * sizeof(float)) + 127) & ~(size_t)127;
  // band phase: list[cap] + out[K] live behind the tile; the merge phase (tile dead) re-carves from 0
  SelSmem sm;
  sm.list = reinterpret_cast<u64*>(smem_raw + tile_bytes);
  sm.out = sm.list + a.cap;
  sm.vals = s_vals;
  sm.red = s_red;
  sm.misc = s_misc;

  // ---- stage the band (+ halo rows) in shared memory -------------------------------------------
  const int tid = threadIdx.x;
  if (VEC == 4) {
    if (tid == 0) {
      mbar_init(&s_mbar, 1);
      fence_mbar_init();
      fence_proxy_async_smem();
    }
    __syncthreads();
    if (tid == 0) {
      const u32 bytes = (u32)((g1 - g0) * a.W * sizeof(float));
      mbar_expect_tx(&s_mbar, bytes);
      bulk_g2s(tile + (size_t)(g0 - (r0 - 1)) * a.W, plane + (size_t)g0 * a.W, bytes, &s_mbar);
    }
    mbar_wait(&s_mbar, 0);
  } else {
    const int n = (g1 - g0) * a.W;
    float* dst = tile + (size_t)(g0 - (r0 - 1)) * a.W;
    const float* src = plane + (size_t)g0 * a.W;
    for (int i = tid; i < n; i += NT) dst[i] = __ldg(src + i);
    __syncthreads();
  }

  // ---- sliding-window 3x3 max: thread = (column group cg, row segment seg) ----------------------
  const int W4 = a.W / VEC;
  const int cg = tid % W4;
  const int seg = tid / W4;
  const int rs = r0 + seg * a.rpt;
  const int re = min(rs + a.rpt, r1);
  u64 flags = 0;
  u32 tmax = 0;
  const float NEG = -INFINITY;
  if (rs < re) {
    float hp2[VEC], hp1[VEC], cp1[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) hp2[i] = hp1[i] = cp1[i] = NEG;
    for (int g = rs - 1; g <= re; ++g) {
      float c[VEC], hm[VEC];
      if (g >= 0 && g < a.H) {
        const float* rowp = tile + (size_t)(g - (r0 - 1)) * a.W + cg * VEC;
        if constexpr (VEC == 4) {
          const float4 q = *reinterpret_cast<const float4*>(rowp);
          c[0] = q.x; c[1 % VEC] = q.y; c[2 % VEC] = q.z; c[3 % VEC] = q.w;
        } else {
          c[0] = rowp[0];
        }
        const float l = cg > 0 ? rowp[-1] : NEG;
        const float r = cg < W4 - 1 ? rowp[VEC] : NEG;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float lft = i == 0 ? l : c[(i - 1 + VEC) % VEC];
          const float rgt = i == VEC - 1 ? r : c[(i + 1) % VEC];
          hm[i] = max3f(lft, c[i], rgt);
        }
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) c[i] = hm[i] = NEG;
      }
      if (g >= rs + 1) {
        const int rrel = g - 1 - rs;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float m = max3f(hp2[i], hp1[i], hm[i]);
          const float v = cp1[i];
          if (v == m && v > 0.f) {
            flags |= 1ull << (rrel * VEC + i);
            tmax = max(tmax, __float_as_uint(v));
          }
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        hp2[i] = hp1[i];
        hp1[i] = hm[i];
        cp1[i] = c[i];
      }
    }
  }

  // ---- candidate selection for this band ---------------------------------------------------------
  BandCands<VEC> cands{tile, flags, rs, cg, a.W, r0, (u32)(pl * HW)};
  const int total = block_sum(__popcll(flags), sm.red);
  int n = block_collect(cands, total, tmax, a.K, a.cap, sm);
  if (!FUSE_CTDET) {
    const int li = blockIdx.x;  // list index == (b*P + p)*nbands + band
    u64* gl = a.lists + (size_t)li * a.cap;
    if (exact) {
      block_rank_to_out(n, a.K, sm);
      n = min(n, a.K);
      for (int i = tid; i < n; i += NT) gl[i] = sm.out[i];
    } else {
      for (int i = tid; i < n; i += NT) gl[i] = sm.list[i];
    }
    if (tid == 0) a.counts[li] = n;
    return;
  }

  // ---- fused ctdet: append this band's survivors to the image's candidate array ------------------
  const size_t img_cap = (size_t)P * a.nbands * a.cap;
  if (tid == 0) s_misc[3] = n ? atomicAdd(a.counts + b, n) : 0;
  __syncthreads();
  {
    u64* gl = a.lists + (size_t)b * img_cap + s_misc[3];
    for (int i = tid; i < n; i += NT) gl[i] = sm.list[i];
  }

  // ---- last CTA of the image merges all lists and writes the detections -------------------------
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(a.done + b, 1);
    s_misc[1] = (prev == P * a.nbands - 1);
  }
  __syncthreads();
  if (!s_misc[1]) return;
  __threadfence();
  SelSmem msm = sm;   // tile is dead: SURV_CAP survivors + MAX_K outputs from the start of smem
  msm.list = reinterpret_cast<u64*>(smem_raw);
  msm.out = msm.list + SURV_CAP;
  ListCands lc{a.lists + (size_t)b * img_cap, a.counts + b, 1, 0};
  const int have = block_select_from_lists(lc, a.K, msm);
  if (have < a.K) {
    block_zero_fill(a.t0 + (size_t)b * a.C0 * HW, a.C0 * HW, a.H, a.W, have, a.K, msm);
    __syncthreads();
  }
  ctdet_write_rows(a, b, msm.out);
}

// =================================================================================================
// multi_pose association: one CTA per (joint j, image b)  (decode/multi_pose.py:15-94)
// =================================================================================================
struct PoseArgs {
  const float *heat, *wh, *kps, *reg, *hm_hp, *hp_offset;
  float* out;
  int B, J, H, W, K;
  int nbands, cap;
  const u64* lists;
  const int* counts;
};

__global__ void __launch_bounds__(NT) multi_pose_assoc_kernel(const PoseArgs a) {
  __shared__ __align__(8) u64 s_list[SURV_CAP];
  __shared__ __align__(8) u64 s_out[MAX_K];
  __shared__ __align__(8) u64 s_det[MAX_K];
  __shared__ float s_hx[MAX_K], s_hy[MAX_K], s_hs[MAX_K];
  __shared__ u32 s_vals[NT];
  __shared__ int s_red[NW];
  __shared__ int s_misc[4];
  SelSmem sm{s_list, s_out, s_vals, s_red, s_misc};

  const int j = blockIdx.x, b = blockIdx.y;
  const int P = 1 + a.J, HW = a.H * a.W, K = a.K;
  const int tid = threadIdx.x;

  // (1) detections: exact top-K of the (single-class) heat plane, zero-filled like torch.topk
  {
    ListCands lc{a.lists + (size_t)(b * P) * a.nbands * a.cap, a.counts + (size_t)(b * P) * a.nbands,
                 a.nbands, a.cap};
    const int have = block_select_from_lists(lc, K, sm);
    if (have < K) {
      block_zero_fill(a.heat + (size_t)b * HW, HW, a.H, a.W, have, K, sm);
      __syncthreads();
    }
    for (int i = tid; i < K; i += NT) s_det[i] = s_out[i];
    __syncthreads();
  }
  // (2) joint candidates: exact top-K of hm_hp plane j (utils/decode.py:31-40) + offsets + threshold
  {
    ListCands lc{a.lists + (size_t)(b * P + 1 + j) * a.nbands * a.cap,
                 a.counts + (size_t)(b * P + 1 + j) * a.nbands, a.nbands, a.cap};
    const int have = block_select_from_lists(lc, K, sm);
    for (int m = tid; m < K; m += NT) {
      float sc = 0.f, hx = 0.f, hy = 0.f;
      if (m < have) {
        const u64 key = s_out[m];
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
  int vec, R, nbands, rpt, cap;
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
  const int S = NT / W4;
  const int max_rpt = 64 / vec;
  int rpt = env_int("CNB_DECODE_RPT", 8);
  if (rpt < 1) rpt = 1;
  if (rpt > max_rpt) rpt = max_rpt;
  int R;
  size_t smem;
  int cap = 2 * K;
  if (cap < 64) cap = 64;
  cap = (cap + 31) & ~31;
  const size_t band_extra = (size_t)(cap + K) * sizeof(u64);
  const size_t merge_bytes = (size_t)(SURV_CAP + MAX_K) * sizeof(u64);
  for (;;) {
    R = S * rpt;
    if (R > H) R = H;
    smem = ((((size_t)(R + 2) * W * sizeof(float)) + 127) & ~(size_t)127) + band_extra;
    if (smem <= 100 * 1024 || rpt == 1) break;
    rpt /= 2;
  }
  if (smem > 200 * 1024) return false;
  if (smem < merge_bytes) smem = merge_bytes;
  g->vec = vec;
  g->R = R;
  g->nbands = (H + R - 1) / R;
  g->rpt = (R + S - 1) / S;
  g->cap = cap;
  g->smem = smem;
  return true;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WsLayout {
  size_t lists_off, counts_off, done_off, total;
};
static WsLayout ws_layout(int B, int P, const ScanGeom& g) {
  WsLayout w;
  const size_t nl = (size_t)B * P * g.nbands;
  w.lists_off = 0;
  w.counts_off = align_up(nl * g.cap * sizeof(u64), 256);
  w.done_off = w.counts_off + align_up((nl > (size_t)B ? nl : (size_t)B) * sizeof(int), 256);
  w.total = w.done_off + align_up((size_t)B * sizeof(int), 256);
  return w;
}

template <int VEC, bool FUSE>
static cudaError_t launch_scan(const ScanArgs& a, size_t smem, int grid, cudaStream_t st) {
  static bool configured = false;  // per template instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(plane_scan_kernel<VEC, FUSE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  plane_scan_kernel<VEC, FUSE><<<grid, NT, smem, st>>>(a);
  return cudaGetLastError();
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" size_t cnb_ctdet_decode_workspace_bytes(int B, int C, int H, int W, int K) {
  ScanGeom g;
  if (B < 1 || C < 1 || K < 1 || K > MAX_K || !plan_scan(H, W, K, true, &g)) return 0;
  ScanGeom g1;   // the unaligned (scalar) fallback may band differently; take the larger of the two
  const size_t a = ws_layout(B, C, g).total;
  const size_t b = plan_scan(H, W, K, false, &g1) ? ws_layout(B, C, g1).total : 0;
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
  const WsLayout w = ws_layout(B, C, g);
  if (workspace_bytes < w.total) {
    set_error("ctdet_decode: workspace %zu < %zu bytes", workspace_bytes, w.total);
    return CNB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = (unsigned char*)workspace;
  ScanArgs a{};
  a.t0 = heat; a.t1 = nullptr; a.C0 = C; a.C1 = 0;
  a.B = B; a.H = H; a.W = W; a.K = K;
  a.R = g.R; a.nbands = g.nbands; a.rpt = g.rpt; a.cap = g.cap;
  a.exact0 = 0; a.exact1 = 0;
  a.lists = (u64*)(ws + w.lists_off);
  a.counts = (int*)(ws + w.counts_off);
  a.done = (int*)(ws + w.done_off);
  a.wh = wh; a.reg = reg; a.out = out;
  // zero the per-image append cursors and arrival counters (contiguous: counts .. done)
  CNB_CUDA(cudaMemsetAsync(a.counts, 0, (w.done_off - w.counts_off) + (size_t)B * sizeof(int), st));
  const int grid = B * C * g.nbands;
  cudaError_t e = g.vec == 4 ? launch_scan<4, true>(a, g.smem, grid, st)
                             : launch_scan<1, true>(a, g.smem, grid, st);
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
  const size_t a = ws_layout(B, 1 + J, g).total;
  const size_t b = plan_scan(H, W, K, false, &g1) ? ws_layout(B, 1 + J, g1).total : 0;
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
  const WsLayout w = ws_layout(B, P, g);
  if (workspace_bytes < w.total) {
    set_error("multi_pose_decode: workspace %zu < %zu bytes", workspace_bytes, w.total);
    return CNB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = (unsigned char*)workspace;
  ScanArgs a{};
  a.t0 = heat; a.t1 = hm_hp; a.C0 = 1; a.C1 = J;
  a.B = B; a.H = H; a.W = W; a.K = K;
  a.R = g.R; a.nbands = g.nbands; a.rpt = g.rpt; a.cap = g.cap;
  a.exact0 = 1; a.exact1 = 1;
  a.lists = (u64*)(ws + w.lists_off);
  a.counts = (int*)(ws + w.counts_off);
  a.done = nullptr; a.wh = nullptr; a.reg = nullptr; a.out = nullptr;
  const int grid = B * P * g.nbands;
  cudaError_t e = g.vec == 4 ? launch_scan<4, false>(a, g.smem, grid, st)
                             : launch_scan<1, false>(a, g.smem, grid, st);
  if (e != cudaSuccess) {
    set_error("multi_pose_decode scan launch failed: %s", cudaGetErrorString(e));
    return CNB_ERR_CUDA;
  }
  count_launch();
  PoseArgs pa{heat, wh, kps, reg, hm_hp, hp_offset, out, B, J, H, W, K, g.nbands, g.cap, a.lists, a.counts};
  multi_pose_assoc_kernel<<<dim3(J, B), NT, 0, st>>>(pa);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
