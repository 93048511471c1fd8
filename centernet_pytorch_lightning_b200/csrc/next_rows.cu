// The callers and data formats either side of the hot path ("next" rows), on the GPU:
//   (f)2 target encoding   CenterNet/sample/ctdet.py:39-90 + utils/gaussian.py:6-58   -> ctdet_encode_kernel
//   (f)3 soft-NMS          CenterNet/utils/nms.py:5-106                               -> soft_nms_kernel
//   (f)1 test-time augmentation around the engine
//        prologue  centernet_detection.py:143-156 (pad to (h|31)+1, normalise, hflip copy) -> tta_prologue_kernel
//        flip merge centernet_detection.py:167-171                                          -> tta_flip_merge_kernel
//        epilogue  centernet_detection.py:188-204 (x4, -pad, /scale, group by class)        -> ctdet_post_kernel
#include "cnb_common.cuh"
#include <math.h>

namespace cnb {
namespace {

// ---- (f)2 ------------------------------------------------------------------------------------------------------
// utils/gaussian.py:6-27, evaluated in double like the Python original
__device__ double gaussian_radius_d(double height, double width, double min_overlap) {
  const double b1 = height + width;
  const double c1 = width * height * (1 - min_overlap) / (1 + min_overlap);
  const double r1 = (b1 + sqrt(b1 * b1 - 4 * c1)) / 2;
  const double b2 = 2 * (height + width);
  const double c2 = (1 - min_overlap) * width * height;
  const double r2 = (b2 + sqrt(b2 * b2 - 16 * c2)) / 2;
  const double a3 = 4 * min_overlap;
  const double b3 = -2 * min_overlap * (height + width);
  const double c3 = (min_overlap - 1) * width * height;
  const double r3 = (b3 + sqrt(b3 * b3 - 4 * a3 * c3)) / 2;
  return fmin(r1, fmin(r2, r3));
}

// One CTA per (object slot k, image b).  boxes: [B,M,4] float64 COCO (x, y, w, h) in input pixels (double so that
// x + w is formed like `_coco_box_to_bbox` does, in Python floats, before the cast to float32).
__global__ void __launch_bounds__(128) ctdet_encode_kernel(const double* __restrict__ boxes, const int* __restrict__ cls,
                                                           const int* __restrict__ nobj, float* __restrict__ heat,
                                                           long long* __restrict__ ind, unsigned char* __restrict__ mask,
                                                           float* __restrict__ wh, float* __restrict__ reg, int C, int H,
                                                           int W, int M, int down_ratio) {
  const int k = blockIdx.x, b = blockIdx.y;
  __shared__ int s_r, s_cx, s_cy, s_c;
  __shared__ float s_den;
  if (threadIdx.x == 0) {
    s_r = -1;
    const size_t o = (size_t)b * M + k;
    if (k < nobj[b]) {
      const double* bx = boxes + o * 4;
      float x1 = (float)bx[0], y1 = (float)bx[1], x2 = (float)(bx[0] + bx[2]), y2 = (float)(bx[1] + bx[3]);
      const float dr = (float)down_ratio;
      x1 = fminf(fmaxf(x1 / dr, 0.f), (float)(W - 1));
      y1 = fminf(fmaxf(y1 / dr, 0.f), (float)(H - 1));
      x2 = fminf(fmaxf(x2 / dr, 0.f), (float)(W - 1));
      y2 = fminf(fmaxf(y2 / dr, 0.f), (float)(H - 1));
      const float h = y2 - y1, w = x2 - x1;
      if (h > 0.f && w > 0.f) {
        const double rad = gaussian_radius_d(ceil((double)h), ceil((double)w), 0.7);
        const int r = max(0, (int)rad);
        const float ctx = (x1 + x2) / 2.f, cty = (y1 + y2) / 2.f;
        const int cx = (int)ctx, cy = (int)cty;
        s_r = r;
        s_cx = cx;
        s_cy = cy;
        s_c = cls[o];
        const double sigma = (2 * r + 1) / 6.0;
        s_den = (float)(2 * sigma * sigma);
        wh[o * 2] = w;
        wh[o * 2 + 1] = h;
        ind[o] = (long long)cy * W + cx;
        reg[o * 2] = ctx - (float)cx;
        reg[o * 2 + 1] = cty - (float)cy;
        mask[o] = 1;
      }
    }
  }
  __syncthreads();
  const int r = s_r;
  if (r < 0 || s_c < 0 || s_c >= C) return;
  const int cx = s_cx, cy = s_cy, d = 2 * r + 1;
  int* plane = reinterpret_cast<int*>(heat + ((size_t)b * C + s_c) * H * W);
  for (int i = threadIdx.x; i < d * d; i += blockDim.x) {
    const int dy = i / d - r, dx = i % d - r;
    const int y = cy + dy, x = cx + dx;
    if (y < 0 || y >= H || x < 0 || x >= W) continue;
    float g = expf(-(float)(dx * dx + dy * dy) / s_den);
    if (g < 1.1920929e-07f) g = 0.f;                       // h[h < eps * h.max()] = 0 with h.max() == 1
    atomicMax(plane + (size_t)y * W + x, __float_as_int(g));   // max-composition; values >= 0: int order == float order
  }
}

// ---- (f)3 ------------------------------------------------------------------------------------------------------
// One CTA per box list.  Sequential semantics of the numba original kept exactly: selection order, the
// swap-with-last compaction (done by one thread after the parallel decay), float64 overlap arithmetic.
__global__ void __launch_bounds__(256) soft_nms_kernel(float* __restrict__ boxes_all, const int* __restrict__ counts,
                                                       int* __restrict__ kept, int stride_boxes, int ncol, double sigma,
                                                       double Nt, double threshold, int method) {
  extern __shared__ float s_box[];   // [N][ncol]
  __shared__ float s_best[256];
  __shared__ int s_bidx[256];
  __shared__ int s_N;
  float* boxes = boxes_all + (size_t)blockIdx.x * stride_boxes * ncol;
  const int N0 = counts[blockIdx.x];
  const int tid = threadIdx.x;
  for (int i = tid; i < N0 * ncol; i += 256) s_box[i] = boxes[i];
  if (tid == 0) s_N = N0;
  __syncthreads();
  for (int i = 0; i < N0; ++i) {
    const int N = s_N;
    if (i < N) {
      // first maximum of the scores in [i, N)
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int p = i + tid; p < N; p += 256) {
        const float sc = s_box[p * ncol + 4];
        if (sc > best) {
          best = sc;
          bi = p;
        }
      }
      s_best[tid] = best;
      s_bidx[tid] = bi;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
          const float ob = s_best[tid + o];
          const int oi = s_bidx[tid + o];
          if (ob > s_best[tid] || (ob == s_best[tid] && oi < s_bidx[tid])) {
            s_best[tid] = ob;
            s_bidx[tid] = oi;
          }
        }
        __syncthreads();
      }
      const int maxpos = s_bidx[0];
      if (tid < ncol && maxpos != i) {   // swap the whole row (the 39-column variant carries key points along)
        const float t = s_box[i * ncol + tid];
        s_box[i * ncol + tid] = s_box[maxpos * ncol + tid];
        s_box[maxpos * ncol + tid] = t;
      }
      __syncthreads();
      // numba's typing of the original: box coordinates are float32, differences / min / max are formed in float32 and
      // only the "+ 1" promotes to float64 -- reproduced literally so that threshold decisions cannot differ
      const float tx1 = s_box[i * ncol], ty1 = s_box[i * ncol + 1], tx2 = s_box[i * ncol + 2], ty2 = s_box[i * ncol + 3];
      for (int p = i + 1 + tid; p < N; p += 256) {
        const float x1 = s_box[p * ncol], y1 = s_box[p * ncol + 1], x2 = s_box[p * ncol + 2], y2 = s_box[p * ncol + 3];
        const double area = ((double)(x2 - x1) + 1) * ((double)(y2 - y1) + 1);
        const double iw = (double)(fminf(tx2, x2) - fmaxf(tx1, x1)) + 1;
        if (iw > 0) {
          const double ih = (double)(fminf(ty2, y2) - fmaxf(ty1, y1)) + 1;
          if (ih > 0) {
            const double ua = ((double)(tx2 - tx1) + 1) * ((double)(ty2 - ty1) + 1) + area - iw * ih;
            const double ov = iw * ih / ua;
            double weight;
            if (method == 1) weight = ov > Nt ? 1 - ov : 1;
            else if (method == 2) weight = exp(-(ov * ov) / sigma);
            else weight = ov > Nt ? 0 : 1;
            s_box[p * ncol + 4] = (float)(weight * (double)s_box[p * ncol + 4]);
          }
        }
      }
      __syncthreads();
      if (tid == 0) {   // discard boxes whose score fell below the threshold: swap with the last box, in scan order
        int n = N, pos = i + 1;
        while (pos < n) {
          if ((double)s_box[pos * ncol + 4] < threshold) {
            for (int c = 0; c < ncol; ++c) s_box[pos * ncol + c] = s_box[(n - 1) * ncol + c];
            --n;
          } else {
            ++pos;
          }
        }
        s_N = n;
      }
      __syncthreads();
    }
  }
  const int N = s_N;
  for (int i = tid; i < N * ncol; i += 256) boxes[i] = s_box[i];
  if (tid == 0) kept[blockIdx.x] = N;
}

// ---- (f)1 ------------------------------------------------------------------------------------------------------
// out[f][c][y][x] = (pad(img)[c][y][x'] - mean[c]) / std[c],  x' = x (f = 0) or Wp-1-x (f = 1: the hflip copy)
__global__ void tta_prologue_kernel(const float* __restrict__ img, float* __restrict__ out, int C, int H, int W, int pad_lr,
                                    int pad_tb, float m0, float m1, float m2, float s0, float s1, float s2, int copies) {
  const int Hp = H + 2 * pad_tb, Wp = W + 2 * pad_lr;
  const long long total = (long long)copies * C * Hp * Wp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wp);
    long long p = i / Wp;
    const int y = (int)(p % Hp);
    p /= Hp;
    const int c = (int)(p % C);
    const int f = (int)(p / C);
    const int xs = (f ? Wp - 1 - x : x) - pad_lr, ys = y - pad_tb;
    const float v = (xs >= 0 && xs < W && ys >= 0 && ys < H) ? img[((size_t)c * H + ys) * W + xs] : 0.f;
    const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), s = c == 0 ? s0 : (c == 1 ? s1 : s2);
    out[i] = (v - m) / s;
  }
}

// out[0][c][y][x] = (a[0][c][y][x] + a[1][c][y][W-1-x]) / 2
__global__ void tta_flip_merge_kernel(const float* __restrict__ a, float* __restrict__ out, int C, int H, int W) {
  const long long total = (long long)C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    out[i] = (a[i] + a[total + i - x + (W - 1 - x)]) / 2.f;
  }
}

// det [K,6] (x1,y1,x2,y2,score,cls) in output-stride units -> rows (x1,y1,x2,y2,score) in image pixels grouped by class:
// out[offset[c] + j], j = rank of the row among its class in input order; counts[c] rows per class.
__global__ void __launch_bounds__(256) ctdet_post_kernel(const float* __restrict__ det, float* __restrict__ out,
                                                         int* __restrict__ counts, int* __restrict__ offsets, int K, int C,
                                                         float down_ratio, float pad_x, float pad_y, float scale_x, float scale_y) {
  extern __shared__ int s_cnt[];   // [C] counts, then [C] offsets
  int* s_off = s_cnt + C;
  for (int c = threadIdx.x; c < C; c += 256) s_cnt[c] = 0;
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += 256) {
    const int c = (int)det[k * 6 + 5];
    if (c >= 0 && c < C) atomicAdd(&s_cnt[c], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int c = 0; c < C; ++c) {
      s_off[c] = run;
      run += s_cnt[c];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    counts[c] = s_cnt[c];
    offsets[c] = s_off[c];
  }
  for (int k = threadIdx.x; k < K; k += 256) {
    const int c = (int)det[k * 6 + 5];
    if (c < 0 || c >= C) continue;
    int rank = 0;
    for (int j = 0; j < k; ++j) rank += ((int)det[j * 6 + 5] == c);
    float* o = out + (size_t)(s_off[c] + rank) * 5;
    o[0] = (det[k * 6 + 0] * down_ratio - pad_x) / scale_x;
    o[1] = (det[k * 6 + 1] * down_ratio - pad_y) / scale_y;
    o[2] = (det[k * 6 + 2] * down_ratio - pad_x) / scale_x;
    o[3] = (det[k * 6 + 3] * down_ratio - pad_y) / scale_y;
    o[4] = det[k * 6 + 4];
  }
}

inline int grid_for(long long total) {
  long long g = (total + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" int cnb_ctdet_encode(const double* boxes, const int* cls, const int* nobj, float* heatmap, long long* indices,
                                unsigned char* mask, float* wh, float* reg, int B, int C, int H, int W, int M,
                                int down_ratio, cnb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CNB_CHECK_ARG(boxes && cls && nobj && heatmap && indices && mask && wh && reg, "ctdet_encode: null pointer");
  CNB_CHECK_ARG(B >= 1 && C >= 1 && H >= 1 && W >= 1 && M >= 1 && down_ratio >= 1, "ctdet_encode: bad shape");
  CNB_CUDA(cudaMemsetAsync(heatmap, 0, (size_t)B * C * H * W * 4, st));
  CNB_CUDA(cudaMemsetAsync(indices, 0, (size_t)B * M * 8, st));
  CNB_CUDA(cudaMemsetAsync(mask, 0, (size_t)B * M, st));
  CNB_CUDA(cudaMemsetAsync(wh, 0, (size_t)B * M * 8, st));
  CNB_CUDA(cudaMemsetAsync(reg, 0, (size_t)B * M * 8, st));
  ctdet_encode_kernel<<<dim3((unsigned)M, (unsigned)B), 128, 0, st>>>(boxes, cls, nobj, heatmap, indices, mask, wh, reg, C, H, W,
                                                                       M, down_ratio);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_soft_nms(float* boxes, const int* counts, int* kept, int nlists, int max_boxes, int ncol, double sigma,
                            double Nt, double threshold, int method, cnb_stream_t stream) {
  CNB_CHECK_ARG(boxes && counts && kept && nlists >= 1 && max_boxes >= 1 && ncol >= 5 && ncol <= 64, "soft_nms: bad argument");
  const size_t smem = (size_t)max_boxes * ncol * sizeof(float);
  CNB_CHECK_ARG(smem <= 200 * 1024, "soft_nms: list too long for shared memory");
  static PerDeviceOnce once;
  if (once.need()) {
    CNB_CUDA(cudaFuncSetAttribute(soft_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    once.mark();
  }
  soft_nms_kernel<<<nlists, 256, smem, (cudaStream_t)stream>>>(boxes, counts, kept, max_boxes, ncol, sigma, Nt, threshold, method);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_tta_prologue(const float* img, float* out, int C, int H, int W, int pad_lr, int pad_tb, const float* mean3,
                                const float* std3, int flip, cnb_stream_t stream) {
  CNB_CHECK_ARG(img && out && mean3 && std3 && C == 3 && pad_lr >= 0 && pad_tb >= 0, "tta_prologue: bad argument (C must be 3; mean/std are HOST arrays)");
  const int copies = flip ? 2 : 1;
  const long long total = (long long)copies * C * (H + 2 * pad_tb) * (W + 2 * pad_lr);
  tta_prologue_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(img, out, C, H, W, pad_lr, pad_tb, mean3[0], mean3[1],
                                                                         mean3[2], std3[0], std3[1], std3[2], copies);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_tta_flip_merge(const float* pair, float* out, int C, int H, int W, cnb_stream_t stream) {
  CNB_CHECK_ARG(pair && out && C >= 1 && H >= 1 && W >= 1, "tta_flip_merge: bad argument");
  tta_flip_merge_kernel<<<grid_for((long long)C * H * W), 256, 0, (cudaStream_t)stream>>>(pair, out, C, H, W);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_ctdet_post(const float* det, float* out, int* counts, int* offsets, int K, int C, float down_ratio,
                              float pad_x, float pad_y, float scale_x, float scale_y, cnb_stream_t stream) {
  CNB_CHECK_ARG(det && out && counts && offsets && K >= 1 && C >= 1 && C <= 4096, "ctdet_post: bad argument");
  ctdet_post_kernel<<<1, 256, (size_t)2 * C * sizeof(int), (cudaStream_t)stream>>>(det, out, counts, offsets, K, C, down_ratio,
                                                                                  pad_x, pad_y, scale_x, scale_y);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
