// fp32-strict execution mode: the same network operators as the tensor-core engine, computed entirely in fp32 on the
// CUDA cores (NCHW fp32, the reference's own layout and precision; SURVEY.md section 7, hard part 2).
//
// Purpose: north_star's END-TO-END tolerance ("bit-exact top-k indices, coords within 1e-4" against the reference's
// fp32 CPU path) cannot be demonstrated through a 60-layer bf16 network; these kernels carry the whole forward in
// fp32 (sequential fmaf accumulation per output element) so that small-shape end-to-end checks compare like with
// like.  They are deliberately simple -- one thread per output element -- and are NOT the measured fast path.
// Operators (reference file:line under CenterNet/models/): nn.Conv2d + folded eval BatchNorm + residual + ReLU
// (backbones/pose_dla_dcn.py:28-68, :165-188, :351-370; heads.py:4-25), DCNv2 (pose_dla_dcn.py:441-449) as
// sampled columns + 1x1 conv, ConvTranspose2d depthwise / dense (pose_dla_dcn.py:466-475, resnet_dcn.py:212-220),
// MaxPool2d (pose_dla_dcn.py:243, resnet_dcn.py:139).
#include "cnb_common.cuh"

namespace cnb {
namespace {

inline int grid_for(long long total, int threads = 256) {
  long long g = (total + threads - 1) / threads;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__global__ void conv2d_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const float* __restrict__ res, float* __restrict__ y,
                                  int B, int Ci, int Hi, int Wi, int Co, int KH, int KW, int stride, int pad, int Ho,
                                  int Wo, int act) {
  const long long total = (long long)B * Co * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    long long p = i / Wo;
    const int oy = (int)(p % Ho);
    p /= Ho;
    const int co = (int)(p % Co);
    const int n = (int)(p / Co);
    float acc = 0.f;
    const float* wp = w + (size_t)co * Ci * KH * KW;
    for (int ci = 0; ci < Ci; ++ci) {
      const float* xp = x + ((size_t)n * Ci + ci) * Hi * Wi;
      for (int kh = 0; kh < KH; ++kh) {
        const int iy = oy * stride - pad + kh;
        if (iy < 0 || iy >= Hi) continue;
        for (int kw = 0; kw < KW; ++kw) {
          const int ix = ox * stride - pad + kw;
          if (ix < 0 || ix >= Wi) continue;
          acc = fmaf(__ldg(xp + (size_t)iy * Wi + ix), __ldg(wp + (ci * KH + kh) * KW + kw), acc);
        }
      }
    }
    float v = acc;
    if (scale) v *= scale[co];
    if (shift) v += shift[co];
    if (res) v += res[i];
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = 1.f / (1.f + expf(-v));
    y[i] = v;
  }
}

// col[n][ci*9 + k][oy][ox] = sigmoid(om[n][18+k]) * bilinear(x[n][ci], (oy-1+kh+om[n][2k], ox-1+kw+om[n][2k+1]))
__global__ void dcn_im2col_f32_kernel(const float* __restrict__ x, const float* __restrict__ om, float* __restrict__ col,
                                      int B, int C, int H, int W) {
  const long long total = (long long)B * C * 9 * H * W;
  const int HW = H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W);
    long long p = i / W;
    const int oy = (int)(p % H);
    p /= H;
    const int k = (int)(p % 9);
    p /= 9;
    const int ci = (int)(p % C);
    const int n = (int)(p / C);
    const float* omp = om + (size_t)n * 27 * HW + (size_t)oy * W + ox;
    const float dy = omp[(size_t)(2 * k) * HW], dx = omp[(size_t)(2 * k + 1) * HW];
    const float mk = 1.f / (1.f + expf(-omp[(size_t)(18 + k) * HW]));
    const float h = (float)(oy - 1 + k / 3) + dy, wq = (float)(ox - 1 + k % 3) + dx;
    float val = 0.f;
    if (h > -1.f && wq > -1.f && h < (float)H && wq < (float)W) {
      const int hl = (int)floorf(h), wl = (int)floorf(wq);
      const int hh = hl + 1, wh = wl + 1;
      const float lh = h - (float)hl, lw = wq - (float)wl;
      const float uh = 1.f - lh, uw = 1.f - lw;
      const float* xp = x + ((size_t)n * C + ci) * HW;
      const float v1 = (hl >= 0 && wl >= 0) ? xp[hl * W + wl] : 0.f;
      const float v2 = (hl >= 0 && wh <= W - 1) ? xp[hl * W + wh] : 0.f;
      const float v3 = (hh <= H - 1 && wl >= 0) ? xp[hh * W + wl] : 0.f;
      const float v4 = (hh <= H - 1 && wh <= W - 1) ? xp[hh * W + wh] : 0.f;
      val = uh * uw * v1 + uh * lw * v2 + lh * uw * v3 + lh * lw * v4;
    }
    col[i] = val * mk;
  }
}

// ConvTranspose2d(Ci, Co, K, stride, pad, groups in {1, Ci == Co}) (+ add): gather form
__global__ void conv_transpose2d_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                            const float* __restrict__ add, float* __restrict__ y, int B, int Ci, int Hi,
                                            int Wi, int Co, int K, int stride, int pad, int depthwise, int act) {
  const int Ho = (Hi - 1) * stride - 2 * pad + K, Wo = (Wi - 1) * stride - 2 * pad + K;
  const long long total = (long long)B * Co * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    long long p = i / Wo;
    const int oy = (int)(p % Ho);
    p /= Ho;
    const int co = (int)(p % Co);
    const int n = (int)(p / Co);
    float acc = 0.f;
    const int ci0 = depthwise ? co : 0, ci1 = depthwise ? co + 1 : Ci;
    for (int ci = ci0; ci < ci1; ++ci) {
      const float* xp = x + ((size_t)n * Ci + ci) * Hi * Wi;
      // weight layout [Ci][Co/groups][K][K]
      const float* wp = depthwise ? w + (size_t)ci * K * K : w + ((size_t)ci * Co + co) * K * K;
      for (int ky = 0; ky < K; ++ky) {
        const int ty = oy + pad - ky;
        if (ty < 0 || ty % stride) continue;
        const int iy = ty / stride;
        if (iy >= Hi) continue;
        for (int kx = 0; kx < K; ++kx) {
          const int tx = ox + pad - kx;
          if (tx < 0 || tx % stride) continue;
          const int ix = tx / stride;
          if (ix >= Wi) continue;
          acc = fmaf(xp[(size_t)iy * Wi + ix], wp[ky * K + kx], acc);
        }
      }
    }
    float v = acc;
    if (scale) v *= scale[co];
    if (shift) v += shift[co];
    if (add) v += add[i];
    if (act == 1) v = fmaxf(v, 0.f);
    y[i] = v;
  }
}

__global__ void maxpool2d_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int BC, int H, int W, int k,
                                     int stride, int pad, int Ho, int Wo) {
  const long long total = (long long)BC * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    long long p = i / Wo;
    const int oy = (int)(p % Ho);
    const long long bc = p / Ho;
    float m = -INFINITY;
    for (int a = 0; a < k; ++a) {
      const int iy = oy * stride - pad + a;
      if (iy < 0 || iy >= H) continue;
      for (int b = 0; b < k; ++b) {
        const int ix = ox * stride - pad + b;
        if (ix < 0 || ix >= W) continue;
        m = fmaxf(m, x[((size_t)bc * H + iy) * W + ix]);
      }
    }
    y[i] = m;
  }
}

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" int cnb_strict_conv2d_f32(const float* x, const float* w, const float* scale, const float* shift,
                                     const float* res, float* y, int B, int Ci, int Hi, int Wi, int Co, int KH, int KW,
                                     int stride, int pad, int act, cnb_stream_t stream) {
  CNB_CHECK_ARG(x && w && y && B >= 1 && Ci >= 1 && Co >= 1 && stride >= 1, "strict_conv2d_f32: bad argument");
  const int Ho = (Hi + 2 * pad - KH) / stride + 1, Wo = (Wi + 2 * pad - KW) / stride + 1;
  conv2d_f32_kernel<<<grid_for((long long)B * Co * Ho * Wo), 256, 0, (cudaStream_t)stream>>>(
      x, w, scale, shift, res, y, B, Ci, Hi, Wi, Co, KH, KW, stride, pad, Ho, Wo, act);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_strict_dcn_im2col_f32(const float* x, const float* om, float* col, int B, int C, int H, int W,
                                         cnb_stream_t stream) {
  CNB_CHECK_ARG(x && om && col, "strict_dcn_im2col_f32: null pointer");
  dcn_im2col_f32_kernel<<<grid_for((long long)B * C * 9 * H * W), 256, 0, (cudaStream_t)stream>>>(x, om, col, B, C, H, W);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_strict_conv_transpose2d_f32(const float* x, const float* w, const float* scale, const float* shift,
                                               const float* add, float* y, int B, int Ci, int Hi, int Wi, int Co, int K,
                                               int stride, int pad, int depthwise, int act, cnb_stream_t stream) {
  CNB_CHECK_ARG(x && w && y && (!depthwise || Ci == Co), "strict_conv_transpose2d_f32: bad argument");
  const int Ho = (Hi - 1) * stride - 2 * pad + K, Wo = (Wi - 1) * stride - 2 * pad + K;
  conv_transpose2d_f32_kernel<<<grid_for((long long)B * Co * Ho * Wo), 256, 0, (cudaStream_t)stream>>>(
      x, w, scale, shift, add, y, B, Ci, Hi, Wi, Co, K, stride, pad, depthwise, act);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_strict_maxpool2d_f32(const float* x, float* y, int BC, int H, int W, int k, int stride, int pad,
                                        cnb_stream_t stream) {
  CNB_CHECK_ARG(x && y && k >= 1 && stride >= 1, "strict_maxpool2d_f32: bad argument");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  maxpool2d_f32_kernel<<<grid_for((long long)BC * Ho * Wo), 256, 0, (cudaStream_t)stream>>>(x, y, BC, H, W, k, stride, pad, Ho, Wo);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
