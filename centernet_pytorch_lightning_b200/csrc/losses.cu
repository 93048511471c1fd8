// Fused CenterNet losses (forward + backward in one pass over the maps).
//
// Replaces (reference file:line under CenterNet/):
//   utils/decode.py:43-45  sigmoid_clamped
//   utils/losses.py:14-39  _neg_loss  (FocalLoss :42-50)   -- ~10 elementwise passes + 3 sums + a
//                                                             host sync at :35 in the reference
//   utils/losses.py:53-63  RegL1Loss, :81-91 RegWeightedL1Loss (+ utils/decode.py:59-63 gather copy)
//
// HBM-bound: focal = read pred/logits + gt, write grad = 12 B/element.
// Reductions are deterministic: fixed grid, per-CTA partials in the workspace, one finalising CTA.
#include "cnb_common.cuh"

namespace cnb {
namespace {

constexpr int FT = 256;                // threads per CTA
constexpr int FOCAL_GRID = 148 * 8;    // persistent grid-stride: 8 CTAs per SM

struct FocalPartial {
  float pos, neg, npos, pad;
};

// One element of _neg_loss (losses.py:21-29) and d(-L)/dp.
__device__ __forceinline__ void focal_elem(float p, float gt, float& pos, float& neg, float& npos,
                                           float& g) {
  g = 0.f;
  if (gt == 1.f) {
    const float om = 1.f - p;
    const float lg = logf(p);
    pos += lg * om * om;
    npos += 1.f;
    g = -(om * om / p - 2.f * om * lg);
  } else if (gt < 1.f) {
    const float om = 1.f - p;
    const float w = (1.f - gt) * (1.f - gt);
    const float w4 = w * w;
    const float lg = logf(om);
    neg += lg * p * p * w4;
    g = -(2.f * p * lg - p * p / om) * w4;
  }
}

template <bool FROM_LOGITS>
__global__ void __launch_bounds__(FT) focal_kernel(const float* __restrict__ x,
                                                   const float* __restrict__ gt,
                                                   float* __restrict__ prob_out,
                                                   float* __restrict__ grad, FocalPartial* partial,
                                                   long long n) {
  float pos = 0.f, neg = 0.f, npos = 0.f;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * FT;
  for (long long i = (long long)blockIdx.x * FT + threadIdx.x; i < n4; i += stride) {
    const float4 xv = __ldcs(reinterpret_cast<const float4*>(x) + i);
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(gt) + i);
    float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
    float gr[4], pr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float p = xs[k], dpdx = 1.f;
      if (FROM_LOGITS) {
        const float s = 1.f / (1.f + expf(-xs[k]));
        p = fminf(fmaxf(s, 1e-4f), 1.f - 1e-4f);
        dpdx = (s >= 1e-4f && s <= 1.f - 1e-4f) ? s * (1.f - s) : 0.f;
      }
      pr[k] = p;
      float g;
      focal_elem(p, gs[k], pos, neg, npos, g);
      gr[k] = g * dpdx;
    }
    if (grad) __stcs(reinterpret_cast<float4*>(grad) + i, make_float4(gr[0], gr[1], gr[2], gr[3]));
    if (FROM_LOGITS && prob_out)
      __stcs(reinterpret_cast<float4*>(prob_out) + i, make_float4(pr[0], pr[1], pr[2], pr[3]));
  }
  // tail (n % 4) handled by CTA 0
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += FT) {
      float p = x[i], dpdx = 1.f;
      if (FROM_LOGITS) {
        const float s = 1.f / (1.f + expf(-x[i]));
        p = fminf(fmaxf(s, 1e-4f), 1.f - 1e-4f);
        dpdx = (s >= 1e-4f && s <= 1.f - 1e-4f) ? s * (1.f - s) : 0.f;
        if (prob_out) prob_out[i] = p;
      }
      float g;
      focal_elem(p, gt[i], pos, neg, npos, g);
      if (grad) grad[i] = g * dpdx;
    }
  }
  __shared__ float s_red[3][FT / 32];
  pos = warp_sum_f(pos);
  neg = warp_sum_f(neg);
  npos = warp_sum_f(npos);
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = pos;
    s_red[1][threadIdx.x >> 5] = neg;
    s_red[2][threadIdx.x >> 5] = npos;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int w = 0; w < FT / 32; ++w) {
      a += s_red[0][w];
      b += s_red[1][w];
      c += s_red[2][w];
    }
    partial[blockIdx.x] = FocalPartial{a, b, c, 0.f};
  }
}

// loss_out[0] = loss, [1] = gradient normaliser (1/num_pos, or 1 when num_pos == 0), [2] = num_pos
__global__ void focal_finalize_kernel(const FocalPartial* partial, int nparts, float* loss_out) {
  __shared__ double s[3][32];
  double a = 0, b = 0, c = 0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    a += partial[i].pos;
    b += partial[i].neg;
    c += partial[i].npos;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s[0][threadIdx.x >> 5] = a;
    s[1][threadIdx.x >> 5] = b;
    s[2][threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = b = c = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += s[0][w];
      b += s[1][w];
      c += s[2][w];
    }
    // losses.py:35-38
    if (c == 0) {
      loss_out[0] = (float)(-b);
      loss_out[1] = 1.f;
    } else {
      loss_out[0] = (float)(-(a + b) / c);
      loss_out[1] = (float)(1.0 / c);
    }
    loss_out[2] = (float)c;
  }
}

// ---- sigmoid_clamped (utils/decode.py:43-45) forward / backward --------------------------------
__global__ void sigmoid_clamp_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                         float lo, float hi) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float s = 1.f / (1.f + expf(-x[i]));
    y[i] = fminf(fmaxf(s, lo), hi);
  }
}
// dx = dy * y(1-y) inside the clamp range, 0 where the clamp was active
__global__ void sigmoid_clamp_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                         float* __restrict__ dx, long long n, float lo, float hi) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = y[i];
    dx[i] = (v > lo && v < hi) ? dy[i] * v * (1.f - v) : 0.f;
  }
}

// ---- gather-L1 ---------------------------------------------------------------------------------
struct RegArgs {
  const float* output;
  const void* mask;
  const long long* ind;
  const float* target;
  float* doutput;
  float* loss_out;
  const float* grad_scale;   // device scalar or nullptr (== 1)
  int B, C, H, W, M, mask_per_channel;
};

__global__ void __launch_bounds__(1024) reg_l1_kernel(const RegArgs a) {
  __shared__ float s_a[32], s_b[32];
  __shared__ float s_den;
  const int HW = a.H * a.W;
  const long long n = (long long)a.B * a.M * a.C;
  float num = 0.f, den = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = (int)(i % a.C);
    const long long bm = i / a.C;
    const int b = (int)(bm / a.M);
    const float m = a.mask_per_channel ? ((const float*)a.mask)[i]
                                       : (((const unsigned char*)a.mask)[bm] ? 1.f : 0.f);
    const long long idx = a.ind[bm];
    if (idx < 0 || idx >= HW) continue;   // out-of-range index: the entry is dropped (the reference's gather asserts)
    const float pred = a.output[((size_t)b * a.C + c) * HW + idx];
    num += fabsf(pred * m - a.target[i] * m);
    den += m;
  }
  num = warp_sum_f(num);
  den = warp_sum_f(den);
  if ((threadIdx.x & 31) == 0) {
    s_a[threadIdx.x >> 5] = num;
    s_b[threadIdx.x >> 5] = den;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0.f, y = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      x += s_a[w];
      y += s_b[w];
    }
    y += 1e-4f;   // losses.py:62
    a.loss_out[0] = x / y;
    s_den = y;
  }
  __syncthreads();
  if (!a.doutput) return;
  const float gs = (a.grad_scale ? *a.grad_scale : 1.f) / s_den;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = (int)(i % a.C);
    const long long bm = i / a.C;
    const int b = (int)(bm / a.M);
    const float m = a.mask_per_channel ? ((const float*)a.mask)[i]
                                       : (((const unsigned char*)a.mask)[bm] ? 1.f : 0.f);
    const long long idx = a.ind[bm];
    if (idx < 0 || idx >= HW) continue;
    const size_t o = ((size_t)b * a.C + c) * HW + idx;
    const float d = a.output[o] * m - a.target[i] * m;
    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    if (sg != 0.f && m != 0.f) atomicAdd(a.doutput + o, sg * m * gs);
  }
}

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" size_t cnb_focal_loss_workspace_bytes(long long n) {
  (void)n;
  return sizeof(FocalPartial) * FOCAL_GRID;
}

template <bool LOGITS>
static int focal_run(const float* x, const float* gt, float* prob_out, float* grad, float* loss_out,
                     long long n, void* ws, size_t ws_bytes, cnb_stream_t stream) {
  CNB_CHECK_ARG(x && gt && loss_out && ws, "focal_loss: null pointer");
  CNB_CHECK_ARG(n > 0, "focal_loss: empty input");
  CNB_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)gt & 15) == 0 && ((uintptr_t)grad & 15) == 0 &&
                    ((uintptr_t)prob_out & 15) == 0,
                "focal_loss: pointers must be 16-byte aligned");
  if (ws_bytes < sizeof(FocalPartial) * FOCAL_GRID) {
    set_error("focal_loss: workspace too small");
    return CNB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  long long want = ((n >> 2) + FT - 1) / FT;
  int grid = (int)(want < 1 ? 1 : (want > FOCAL_GRID ? FOCAL_GRID : want));
  focal_kernel<LOGITS><<<grid, FT, 0, st>>>(x, gt, prob_out, grad, (FocalPartial*)ws, n);
  CNB_LAUNCH_CHECK();
  focal_finalize_kernel<<<1, 256, 0, st>>>((const FocalPartial*)ws, grid, loss_out);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_focal_loss_fwd_bwd(const float* logits, const float* gt, float* prob_out,
                                      float* dlogits, float* loss_out, long long n, void* workspace,
                                      size_t workspace_bytes, cnb_stream_t stream) {
  return focal_run<true>(logits, gt, prob_out, dlogits, loss_out, n, workspace, workspace_bytes, stream);
}

extern "C" int cnb_focal_loss_prob_fwd_bwd(const float* pred, const float* gt, float* dpred,
                                           float* loss_out, long long n, void* workspace,
                                           size_t workspace_bytes, cnb_stream_t stream) {
  return focal_run<false>(pred, gt, nullptr, dpred, loss_out, n, workspace, workspace_bytes, stream);
}

extern "C" int cnb_reg_l1_fwd_bwd(const float* output, const void* mask, const long long* ind,
                                  const float* target, float* doutput, float* loss_out, int B, int C,
                                  int H, int W, int M, int mask_per_channel,
                                  const float* grad_scale_dev, cnb_stream_t stream) {
  CNB_CHECK_ARG(output && mask && ind && target && loss_out, "reg_l1: null pointer");
  CNB_CHECK_ARG(B >= 1 && C >= 1 && H >= 1 && W >= 1 && M >= 1, "reg_l1: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (doutput) CNB_CUDA(cudaMemsetAsync(doutput, 0, (size_t)B * C * H * W * sizeof(float), st));
  RegArgs a{output, mask, ind, target, doutput, loss_out, grad_scale_dev, B, C, H, W, M, mask_per_channel};
  reg_l1_kernel<<<1, 1024, 0, st>>>(a);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_sigmoid_clamped_fwd(const float* x, float* y, long long n, float lo, float hi,
                                       cnb_stream_t stream) {
  CNB_CHECK_ARG(x && y && n > 0, "sigmoid_clamped_fwd: bad argument");
  long long want = (n + 255) / 256;
  const int grid = (int)(want > 148 * 16 ? 148 * 16 : want);
  sigmoid_clamp_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, n, lo, hi);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_sigmoid_clamped_bwd(const float* y, const float* dy, float* dx, long long n, float lo,
                                       float hi, cnb_stream_t stream) {
  CNB_CHECK_ARG(y && dy && dx && n > 0, "sigmoid_clamped_bwd: bad argument");
  long long want = (n + 255) / 256;
  const int grid = (int)(want > 148 * 16 ? 148 * 16 : want);
  sigmoid_clamp_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, dy, dx, n, lo, hi);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
