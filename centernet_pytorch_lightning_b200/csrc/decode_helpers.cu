// Stand-alone versions of the decode primitives of CenterNet/utils/decode.py, for callers that import them by name:
//   _nms (:5-10)  _topk (:13-28)  _topk_channel (:31-40)  _gather_feat (:48-56)  _transpose_and_gather_feat (:59-63)
// ctdet_decode / multi_pose_decode do NOT go through these (decode.cu fuses all of them into one streaming pass);
// these exist so that the module swap of INTEGRATION.md leaves no dangling import.
#include "cnb_common.cuh"

namespace cnb {
namespace {

// keep = (maxpool3x3(heat) == heat); out = heat * keep   (padding never wins: max_pool2d pads with -inf)
__global__ void nms3x3_kernel(const float* __restrict__ heat, float* __restrict__ out, long long planes, int H, int W) {
  const long long total = planes * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const float* p = heat + (i - (long long)y * W - x);
    const float v = heat[i];
    float m = v;
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        m = fmaxf(m, p[(size_t)yy * W + xx]);
      }
    }
    out[i] = (m == v) ? v : v * 0.f;
  }
}

__device__ __forceinline__ u32 order_key(float f) {   // monotone float -> uint
  const u32 b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Exact top-K of every row (values descending, ties by ascending index): 4-pass radix select for the K-th key, an
// index-ordered compaction of {key > T} and the first needed {key == T}, then a rank sort of the K survivors.
constexpr int TK_THREADS = 256;
constexpr int TK_MAXK = 512;

__global__ void __launch_bounds__(TK_THREADS) topk_rows_kernel(const float* __restrict__ scores, int n, int K,
                                                               float* __restrict__ out_scores, long long* __restrict__ out_idx) {
  __shared__ u32 s_hist[256];
  __shared__ u32 s_prefix, s_need;
  __shared__ u32 s_key[TK_MAXK];
  __shared__ int s_idx[TK_MAXK];
  __shared__ int s_wsum[TK_THREADS / 32][2];
  __shared__ int s_base[2];
  const float* row = scores + (size_t)blockIdx.x * n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Keff = min(K, n);
  // ---- radix select: largest T such that count(key >= T) >= Keff
  if (tid == 0) {
    s_prefix = 0;
    s_need = (u32)Keff;
  }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    s_hist[tid] = 0;
    __syncthreads();
    const u32 prefix = s_prefix;
    const u32 himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = tid; i < n; i += TK_THREADS) {
      const u32 k = order_key(row[i]);
      if ((k & himask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      u32 need = s_need, d = 255;
      for (;; --d) {
        if (s_hist[d] >= need) break;
        need -= s_hist[d];
        if (d == 0) break;
      }
      s_prefix = prefix | (d << shift);
      s_need = need;   // still needed among keys equal to the prefix so far
    }
    __syncthreads();
  }
  const u32 T = s_prefix;
  const int need_eq = (int)s_need;
  // ---- index-ordered compaction
  if (tid == 0) s_base[0] = s_base[1] = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += TK_THREADS) {
    const int i = i0 + tid;
    const u32 k = i < n ? order_key(row[i]) : 0u;
    const int gt = (i < n && k > T) ? 1 : 0, eq = (i < n && k == T) ? 1 : 0;
    const u32 bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) {
      s_wsum[warp][0] = __popc(bg);
      s_wsum[warp][1] = __popc(be);
    }
    __syncthreads();
    int og = s_base[0], oe = s_base[1];
    for (int w = 0; w < warp; ++w) {
      og += s_wsum[w][0];
      oe += s_wsum[w][1];
    }
    og += __popc(bg & ((1u << lane) - 1u));
    oe += __popc(be & ((1u << lane) - 1u));
    // slots: [0, n_gt) for key > T (n_gt = Keff - need_eq), then the first need_eq equal keys
    if (gt) {
      s_key[og] = k;
      s_idx[og] = i;
    } else if (eq && oe < need_eq) {
      s_key[Keff - need_eq + oe] = k;
      s_idx[Keff - need_eq + oe] = i;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 0; w < TK_THREADS / 32; ++w) {
        s_base[0] += s_wsum[w][0];
        s_base[1] += s_wsum[w][1];
      }
    }
    __syncthreads();
  }
  // ---- rank sort
  for (int a = tid; a < Keff; a += TK_THREADS) {
    const u32 ka = s_key[a];
    const int ia = s_idx[a];
    int rank = 0;
    for (int b = 0; b < Keff; ++b) {
      const u32 kb = s_key[b];
      rank += (kb > ka || (kb == ka && s_idx[b] < ia)) ? 1 : 0;
    }
    out_scores[(size_t)blockIdx.x * K + rank] = row[ia];
    out_idx[(size_t)blockIdx.x * K + rank] = ia;
  }
}

// out[b][k][c] = feat[b][c][ind[b][k]]   (feat NCHW with HW flattened)
__global__ void gather_nchw_kernel(const float* __restrict__ feat, const long long* __restrict__ ind, float* __restrict__ out,
                                   int B, int C, int HW, int K) {
  const long long total = (long long)B * K * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bk = i / C;
    const int b = (int)(bk / K);
    const long long id = ind[bk];
    out[i] = (id >= 0 && id < HW) ? feat[((size_t)b * C + c) * HW + id] : 0.f;
  }
}

// out[b][k][c] = feat[b][ind[b][k]][c]   (feat [B,N,C])
__global__ void gather_rows_kernel(const float* __restrict__ feat, const long long* __restrict__ ind, float* __restrict__ out,
                                   int B, int N, int C, int K) {
  const long long total = (long long)B * K * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bk = i / C;
    const int b = (int)(bk / K);
    const long long id = ind[bk];
    out[i] = (id >= 0 && id < N) ? feat[((size_t)b * N + id) * C + c] : 0.f;
  }
}

inline int grid_for(long long total) {
  long long g = (total + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" int cnb_nms3x3(const float* heat, float* out, long long planes, int H, int W, cnb_stream_t stream) {
  CNB_CHECK_ARG(heat && out && planes >= 1 && H >= 1 && W >= 1, "nms3x3: bad argument");
  nms3x3_kernel<<<grid_for(planes * H * W), 256, 0, (cudaStream_t)stream>>>(heat, out, planes, H, W);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_topk_rows(const float* scores, int rows, int n, int K, float* out_scores, long long* out_idx,
                             cnb_stream_t stream) {
  CNB_CHECK_ARG(scores && out_scores && out_idx && rows >= 1 && n >= 1, "topk_rows: bad argument");
  CNB_CHECK_ARG(K >= 1 && K <= TK_MAXK && K <= n, "topk_rows: K=%d must be in [1, min(%d, n=%d)]", K, TK_MAXK, n);
  topk_rows_kernel<<<rows, TK_THREADS, 0, (cudaStream_t)stream>>>(scores, n, K, out_scores, out_idx);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_gather_feat(const float* feat, const long long* ind, float* out, int B, int C, int N, int K,
                               int feat_is_nchw, cnb_stream_t stream) {
  CNB_CHECK_ARG(feat && ind && out && B >= 1 && C >= 1 && N >= 1 && K >= 1, "gather_feat: bad argument");
  if (feat_is_nchw)
    gather_nchw_kernel<<<grid_for((long long)B * K * C), 256, 0, (cudaStream_t)stream>>>(feat, ind, out, B, C, N, K);
  else
    gather_rows_kernel<<<grid_for((long long)B * K * C), 256, 0, (cudaStream_t)stream>>>(feat, ind, out, B, N, C, K);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
