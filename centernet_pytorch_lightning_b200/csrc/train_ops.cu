// Training-path kernels that are bandwidth-bound (everything around the tensor-core GEMMs of the backward pass).
//
// What autograd + ATen/cuDNN give the reference for free in CenterNet/centernet.py:70-80 (training_step ->
// loss.backward()) for the modules of the hot path, here as hand-written sm_100a kernels over NHWC bf16:
//   * train-mode nn.BatchNorm2d (pose_dla_dcn.py:40, momentum 0.1; batch statistics, running-stat update) forward
//     and backward, fused with the ReLU and the residual add that follow it in BasicBlock / Root / DeformConv
//     (pose_dla_dcn.py:28-68, :165-188, :435-454);
//   * per-channel sums (conv bias gradients: heads.py:9-16, DCN bias, conv_offset_mask bias);
//   * nn.MaxPool2d(2, 2) backward (pose_dla_dcn.py:243);
//   * the depthwise bilinear ConvTranspose2d of IDAUp (pose_dla_dcn.py:466-475) backward: dX and dW (the reference
//     never freezes these weights, so they train);
//   * DCNv2 (external tteepe/DCNv2; call site pose_dla_dcn.py:441-449) for training: the sampled column matrix is
//     materialised once (modulated_deformable_im2col) so that forward, dW and d(columns) are plain tensor-core GEMMs
//     (cnb_conv2d_fprop / cnb_conv2d_wgrad on a 1x1 geometry), and one gather/scatter kernel turns d(columns) into
//     dX (fp32 reductions), d(offset) and d(mask) -- the published DCNv2 backward (col2im + col2im_coord);
//   * Adam (centernet.py:94-95: torch.optim.Adam defaults) on flat fp32 parameter / gradient buffers.
#include "cnb_common.cuh"

namespace cnb {
namespace {

__device__ __forceinline__ float2 bf2f(u32 v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ u32 f2bf(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<u32*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const float2 a = bf2f(v.x), b = bf2f(v.y), c = bf2f(v.z), d = bf2f(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(f2bf(f[0], f[1]), f2bf(f[2], f[3]), f2bf(f[4], f[5]), f2bf(f[6], f[7]));
}

inline int grid_for(long long total, int threads = 256) {
  long long g = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---------------------------------------------------------------------------------------------------------
// Per-channel reductions over the pixels of an NHWC bf16 map [M, C] (C % 8 == 0, C <= 2048).
// A thread owns one 8-channel chunk and strides over rows; threads with the same chunk are reduced through shared
// memory, one atomicAdd per (CTA, channel, quantity) into the zero-filled fp32 output.
//   MODE 0: out[0][c] = sum z,        out[1][c] = sum z^2                      (BatchNorm statistics)
//   MODE 1: out[0][c] = sum g,        out[1][c] = sum g * xhat                 (BatchNorm backward; g = dy * relu mask)
//   MODE 2: out[0][c] = sum g                                                  (bias gradient; g = dy * relu mask)
// ---------------------------------------------------------------------------------------------------------
struct RedArgs {
  const __nv_bfloat16* a;      // z (MODE 0) or dy (MODE 1, 2)
  const __nv_bfloat16* y;      // post-activation output (ReLU mask: y > 0) or nullptr
  const __nv_bfloat16* z;      // MODE 1: pre-BN conv output
  const float* mean;           // MODE 1
  const float* invstd;         // MODE 1
  float* out;                  // [2][C] (MODE 2: [C])
  long long M;
  int C, a_cstride, a_coffset;
};

template <int MODE>
__global__ void __launch_bounds__(256) channel_reduce_kernel(const RedArgs r) {
  __shared__ float s_red[2][256][9];   // padded: 8 values per thread
  const int tpr = (r.C + 7) / 8;       // threads per row (the last chunk may hold padding channels: never written out)
  const int tid = threadIdx.x;
  const int chunk = tid % tpr, rlane = tid / tpr;
  const int rows_per_pass = 256 / tpr;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
  float mu[8], is[8];
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu[j] = r.mean[chunk * 8 + j];
      is[j] = r.invstd[chunk * 8 + j];
    }
  }
  if (rlane < rows_per_pass) {
    for (long long m = (long long)blockIdx.x * rows_per_pass + rlane; m < r.M; m += (long long)gridDim.x * rows_per_pass) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(r.a + (size_t)m * r.a_cstride + r.a_coffset + chunk * 8)), v);
      if (MODE != 0 && r.y) {   // dense map (C % 8 == 0 checked by the callers that pass a mask)
        float yv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(r.y + (size_t)m * r.C + chunk * 8)), yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = yv[j] > 0.f ? v[j] : 0.f;
      }
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s0[j] += v[j];
          s1[j] = fmaf(v[j], v[j], s1[j]);
        }
      } else if (MODE == 1) {
        float zv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(r.z + (size_t)m * r.C + chunk * 8)), zv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s0[j] += v[j];
          s1[j] = fmaf(v[j], (zv[j] - mu[j]) * is[j], s1[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s0[j] += v[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_red[0][tid][j] = s0[j];
    s_red[1][tid][j] = s1[j];
  }
  __syncthreads();
  // thread t < C sums channel t over the row lanes
  for (int c = tid; c < r.C; c += 256) {
    const int ch = c / 8, j = c % 8;
    float t0 = 0.f, t1 = 0.f;
    for (int rl = 0; rl < rows_per_pass; ++rl) {
      t0 += s_red[0][rl * tpr + ch][j];
      t1 += s_red[1][rl * tpr + ch][j];
    }
    atomicAdd(r.out + c, t0);
    if (MODE != 2) atomicAdd(r.out + r.C + c, t1);
  }
}

// mean / invstd / folded (scale, shift) from the sums; running statistics like nn.BatchNorm2d (unbiased variance)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, float count, float eps, float momentum,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ stats /* [4][C]: mean, invstd, scale, shift */, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = sums[c] / count;
  const float var = fmaxf(sums[C + c] / count - mean * mean, 0.f);
  const float invstd = rsqrtf(var + eps);
  const float sc = gamma[c] * invstd;
  stats[c] = mean;
  stats[C + c] = invstd;
  stats[2 * C + c] = sc;
  stats[3 * C + c] = beta[c] - mean * sc;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    const float unb = count > 1.f ? var * count / (count - 1.f) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
  }
}

// y = act(z * scale + shift (+ res)), NHWC bf16, per-channel fp32 scale/shift
__global__ void scale_shift_act_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
                                       const float* __restrict__ shift, const __nv_bfloat16* __restrict__ res,
                                       __nv_bfloat16* __restrict__ y, long long M, int C, int act) {
  const int groups = C / 8;
  const long long total = M * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(z) + i), v);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + g * 8)), s1 = __ldg(reinterpret_cast<const float4*>(scale + g * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + g * 8)), b1 = __ldg(reinterpret_cast<const float4*>(shift + g * 8 + 4));
    v[0] = fmaf(v[0], s0.x, b0.x); v[1] = fmaf(v[1], s0.y, b0.y); v[2] = fmaf(v[2], s0.z, b0.z); v[3] = fmaf(v[3], s0.w, b0.w);
    v[4] = fmaf(v[4], s1.x, b1.x); v[5] = fmaf(v[5], s1.y, b1.y); v[6] = fmaf(v[6], s1.z, b1.z); v[7] = fmaf(v[7], s1.w, b1.w);
    if (res) {
      float rv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(res) + i), rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += rv[j];
    }
    if (act == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    reinterpret_cast<uint4*>(y)[i] = pack8(v);
  }
}

// BatchNorm backward, second pass: dz = scale * (g - sum_g / N - xhat * sum_gx / N), g = dy * relu mask;
// optionally also writes g itself (the gradient of the residual branch).
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                    const __nv_bfloat16* __restrict__ z, const float* __restrict__ stats,
                                    const float* __restrict__ sums, float inv_count, __nv_bfloat16* __restrict__ dz,
                                    __nv_bfloat16* __restrict__ dres, long long M, int C) {
  const int groups = C / 8;
  const long long total = M * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % groups) * 8;
    float g[8], zv[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + i), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(z) + i), zv);
    if (y) {
      float yv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
    }
    if (dres) reinterpret_cast<uint4*>(dres)[i] = pack8(g);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float xhat = (zv[j] - stats[c]) * stats[C + c];
      o[j] = stats[2 * C + c] * (g[j] - sums[c] * inv_count - xhat * sums[C + c] * inv_count);
    }
    reinterpret_cast<uint4*>(dz)[i] = pack8(o);
  }
}

// dgamma = sum_gx, dbeta = sum_g (accumulated into the fp32 gradient buffers)
__global__ void bn_param_grad_kernel(const float* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     int C, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] = (accumulate ? dbeta[c] : 0.f) + sums[c];
  dgamma[c] = (accumulate ? dgamma[c] : 0.f) + sums[C + c];
}

// dx = (y > 0) ? dy : 0
__global__ void relu_bwd_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy,
                                __nv_bfloat16* __restrict__ dx, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float g[8], yv[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + i), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), yv);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
    reinterpret_cast<uint4*>(dx)[i] = pack8(g);
  }
}

// out = bf16(acc_f32 (+ addend_bf16))
__global__ void f32_to_bf16_add_kernel(const float* __restrict__ acc, const __nv_bfloat16* __restrict__ addend,
                                       __nv_bfloat16* __restrict__ out, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(acc) + 2 * i), a1 = __ldg(reinterpret_cast<const float4*>(acc) + 2 * i + 1);
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    if (addend) {
      float b[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(addend) + i), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += b[j];
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(v);
  }
}

// out = a + b (bf16, fp32 add)
__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ out, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(a) + i), x);
    unpack8(__ldg(reinterpret_cast<const uint4*>(b) + i), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    reinterpret_cast<uint4*>(out)[i] = pack8(x);
  }
}

// ---- MaxPool2d(2, 2) backward: the gradient goes to the FIRST maximum of the window in scan order (ATen's rule) --
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                    __nv_bfloat16* __restrict__ dx, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, groups = C / 8;
  const long long total = (long long)B * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long p = i / groups;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float v[4][8], gy[8];
    size_t off[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      off[t] = ((size_t)(b * H + 2 * oy + (t >> 1)) * W + 2 * ox + (t & 1)) * C + g * 8;
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + off[t])), v[t]);
    }
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + i), gy);
    float o[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int best = 0;
      float mx = v[0][j];
#pragma unroll
      for (int t = 1; t < 4; ++t)
        if (v[t][j] > mx) {
          mx = v[t][j];
          best = t;
        }
#pragma unroll
      for (int t = 0; t < 4; ++t) o[t][j] = t == best ? gy[j] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) *reinterpret_cast<uint4*>(dx + off[t]) = pack8(o[t]);
  }
}

// ---- depthwise ConvTranspose2d(C, C, 2f, stride f, pad f/2) backward ---------------------------------------------
// forward: y[oy,ox,c] = sum x[iy,ix,c] * w[c,ky,kx], oy = iy*f - f/2 + ky.
// dx[iy,ix,c] = sum_{ky,kx} dy[iy*f - f/2 + ky, ix*f - f/2 + kx, c] * w[c,ky,kx]     (one thread per (pixel, 8 channels))
// dw[c,ky,kx] = sum_{b,iy,ix} x[b,iy,ix,c] * dy[b, iy*f - f/2 + ky, ...]              (shared-memory accumulation per
//               CTA over its pixels, then one atomicAdd per (CTA, tap, channel) into dwt [(2f)^2][C] fp32)
template <bool REG16>   // REG16: 2f = 4 -> the 16 x 8 weight-gradient partials of a thread live in registers (its 8-channel
                        // group is fixed: the grid stride is a multiple of C/8) and reach shared memory once at the end
__global__ void __launch_bounds__(256) dw_deconv_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt,
                                                            const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                                            float* __restrict__ dwt, int B, int H, int W, int C, int f) {
  extern __shared__ float s_dw[];   // [(2f)^2][C]
  const int ks = 2 * f, pad = f / 2, Ho = H * f, Wo = W * f, groups = C / 8, ntap = ks * ks;
  for (int i = threadIdx.x; i < ntap * C; i += blockDim.x) s_dw[i] = 0.f;
  __syncthreads();
  float racc[REG16 ? 16 : 1][8];
  if (REG16) {
#pragma unroll
    for (int t = 0; t < 16; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) racc[t][j] = 0.f;
  }
  const long long total = (long long)B * H * W * groups;
  const int g = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % groups);   // constant over the loop
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long p = i / groups;
    const int ix = (int)(p % W);
    p /= W;
    const int iy = (int)(p % H);
    const int b = (int)(p / H);
    float xv[8], acc[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (REG16) {   // 2f == 4: fully unrolled, every racc index is a compile-time constant (registers, not local memory)
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int oy = iy * 2 - 1 + ky;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          const int ox = ix * 2 - 1 + kx;
          if (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) {
            float gy[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(dy + ((size_t)(b * Ho + oy) * Wo + ox) * C + g * 8)), gy);
            const float* wp = wt + (size_t)(ky * 4 + kx) * C + g * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[j] = fmaf(gy[j], __ldg(wp + j), acc[j]);
              racc[ky * 4 + kx][j] = fmaf(gy[j], xv[j], racc[ky * 4 + kx][j]);
            }
          }
        }
      }
    } else {
      for (int ky = 0; ky < ks; ++ky) {
        const int oy = iy * f - pad + ky;
        if (oy < 0 || oy >= Ho) continue;
        for (int kx = 0; kx < ks; ++kx) {
          const int ox = ix * f - pad + kx;
          if (ox < 0 || ox >= Wo) continue;
          float gy[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(dy + ((size_t)(b * Ho + oy) * Wo + ox) * C + g * 8)), gy);
          const float* wp = wt + (size_t)(ky * ks + kx) * C + g * 8;
          float* sp = s_dw + (size_t)(ky * ks + kx) * C + g * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[j] = fmaf(gy[j], __ldg(wp + j), acc[j]);
            atomicAdd(sp + j, gy[j] * xv[j]);
          }
        }
      }
    }
    reinterpret_cast<uint4*>(dx)[i] = pack8(acc);
  }
  if (REG16) {
#pragma unroll
    for (int t = 0; t < 16; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(s_dw + (size_t)t * C + g * 8 + j, racc[t][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ntap * C; i += blockDim.x) {
    const float v = s_dw[i];
    if (v != 0.f) atomicAdd(dwt + i, v);
  }
}

// dwt [(2f)^2][C] fp32 -> dw [C,1,2f,2f] (+=)
__global__ void dw_deconv_unpack_wgrad_kernel(const float* __restrict__ dwt, float* __restrict__ dw, int C, int kk, int accumulate) {
  const int total = C * kk;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i / kk, t = i % kk;
    const float v = dwt[(size_t)t * C + c];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

// ---- DCNv2 for training: sampled columns ---------------------------------------------------------------------
// One (pixel, tap) sampling position from the raw offset/mask channels (DCNv2 layout: offsets (dy, dx) at channels
// 2k, 2k+1, mask logit at 18 + k), identical arithmetic to the inference sampler (dcn_ws.cu setup warps) but with an
// accurate sigmoid.
struct Samp {
  float w[4];        // bilinear weights of the 4 corners with validity folded in (mask NOT folded in)
  float mk;          // sigmoid(mask logit)
  float ly, lx;      // fractional position (for the coordinate gradients)
  int idx[4];        // element offsets (pixel * C) of the clamped corners
  bool valid[4];
  bool inside;
};

__device__ __forceinline__ Samp make_samp(const float* __restrict__ omp, int tap, int n, int oy, int ox, int H, int W, int C) {
  Samp s;
  const int kh = tap / 3, kw = tap - 3 * kh;
  const float dy = omp[2 * tap], dx = omp[2 * tap + 1];
  s.mk = 1.f / (1.f + expf(-omp[18 + tap]));
  const float py = (float)(oy - 1 + kh) + dy;
  const float px = (float)(ox - 1 + kw) + dx;
  s.inside = py > -1.f && px > -1.f && py < (float)H && px < (float)W;
  const int y0 = (int)floorf(py), x0 = (int)floorf(px);
  s.ly = py - (float)y0;
  s.lx = px - (float)x0;
  const float hy = 1.f - s.ly, hx = 1.f - s.lx;
  const bool vy0 = y0 >= 0 && y0 <= H - 1, vy1 = y0 + 1 >= 0 && y0 + 1 <= H - 1;
  const bool vx0 = x0 >= 0 && x0 <= W - 1, vx1 = x0 + 1 >= 0 && x0 + 1 <= W - 1;
  s.valid[0] = s.inside && vy0 && vx0;
  s.valid[1] = s.inside && vy0 && vx1;
  s.valid[2] = s.inside && vy1 && vx0;
  s.valid[3] = s.inside && vy1 && vx1;
  s.w[0] = s.valid[0] ? hy * hx : 0.f;
  s.w[1] = s.valid[1] ? hy * s.lx : 0.f;
  s.w[2] = s.valid[2] ? s.ly * hx : 0.f;
  s.w[3] = s.valid[3] ? s.ly * s.lx : 0.f;
  const int y0c = min(max(y0, 0), H - 1), y1c = min(max(y0 + 1, 0), H - 1);
  const int x0c = min(max(x0, 0), W - 1), x1c = min(max(x0 + 1, 0), W - 1);
  const int base = n * H;
  s.idx[0] = ((base + y0c) * W + x0c) * C;
  s.idx[1] = ((base + y0c) * W + x1c) * C;
  s.idx[2] = ((base + y1c) * W + x0c) * C;
  s.idx[3] = ((base + y1c) * W + x1c) * C;
  return s;
}

// col[m][tap*C + c] = mask * bilinear(x[n, :, :, c], p + tap + offset): one thread per (pixel, tap, 8 channels)
__global__ void __launch_bounds__(256) dcn_im2col_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ om,
                                                         int om_cstride, __nv_bfloat16* __restrict__ col, int B, int H,
                                                         int W, int C) {
  const int groups = C / 8;
  const long long total = (long long)B * H * W * 9 * groups;
  const int HW = H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long p = i / groups;
    const int tap = (int)(p % 9);
    const long long m = p / 9;
    const int n = (int)(m / HW);
    const int rem = (int)(m - (long long)n * HW);
    const int oy = rem / W, ox = rem - oy * W;
    const Samp s = make_samp(om + (size_t)m * om_cstride, tap, n, oy, ox, H, W, C);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (!s.valid[c]) continue;
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + s.idx[c] + g * 8)), v);
      const float w = s.w[c] * s.mk;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, v[j], acc[j]);
    }
    reinterpret_cast<uint4*>(col)[i] = pack8(acc);   // i == (m*9 + tap)*groups + g
  }
}
// (A warp-per-item variant of this kernel with the sampling state broadcast by shuffles, like dcn_col2im below, was
// measured 1.4x SLOWER -- 345 vs ~250 us at 64 channels, 128x128, batch 16: the forward has no reduction to amortise the
// shuffles over, and one thread per 16-byte chunk keeps far more loads in flight.)

// d(columns) -> dX (fp32 scatter-add), d(offset), d(mask logit).
// Work item = (pixel, tap) of one 64-channel slab.  Each lane derives the sampling position of one item; the warp then
// walks its 32 items two at a time (one per half-warp) broadcasting that state -- no redundant coordinate math -- with
// the 16 lanes of a half-warp splitting the 64 channels four apiece: one 8-byte load per corner and ONE
// red.global.add.v4.f32 per (corner, lane).  The L2 reduction units are the limit of this kernel (measured: ~300 G
// reduction instructions/s whatever their width), so the widest fp32 vector form is used; shared-memory staging does not
// help because sm_100 has no native shared-memory fp32 add (red.shared.add.f32 compiles to a CAS loop: measured 1.7x
// SLOWER than going to L2 directly).  The three per-item sums (d offset_y, d offset_x, d mask) are half-warp reductions.
__global__ void __launch_bounds__(256) dcn_col2im_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ om,
                                                         int om_cstride, const __nv_bfloat16* __restrict__ dcol,
                                                         float* __restrict__ dx_acc, float* __restrict__ dom, int B,
                                                         int H, int W, int C) {
  const int lane = threadIdx.x & 31;
  const int hsel = lane >> 4, q = lane & 15;
  const int slabs = C / 64;
  const long long HW = (long long)H * W;
  const long long items = (long long)B * HW * 9 * slabs;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long base = warp0 * 32; base < items; base += nwarps * 32) {
    // ---- lane-parallel: the sampling state of item base + lane  (item = ((m * 9 + tap) * slabs + slab))
    const long long item = base + lane;
    int flag = 0, y0 = 0, x0 = 0, tap = 0, slab = 0, n = 0;
    float ly = 0.f, lx = 0.f, mk = 0.f;
    long long m = 0;
    if (item < items) {
      slab = (int)(item % slabs);
      const long long r = item / slabs;
      tap = (int)(r % 9);
      m = r / 9;
      n = (int)(m / HW);
      const int rem = (int)(m - (long long)n * HW);
      const int oy = rem / W, ox = rem - oy * W;
      const float* omp = om + (size_t)m * om_cstride;
      const int kh = tap / 3, kw = tap - 3 * kh;
      const float py = (float)(oy - 1 + kh) + omp[2 * tap];
      const float px = (float)(ox - 1 + kw) + omp[2 * tap + 1];
      mk = 1.f / (1.f + expf(-omp[18 + tap]));
      if (py > -1.f && px > -1.f && py < (float)H && px < (float)W) {
        y0 = (int)floorf(py);
        x0 = (int)floorf(px);
        ly = py - (float)y0;
        lx = px - (float)x0;
        flag = 1;
      }
    }
    float my_sy = 0.f, my_sx = 0.f, my_sm = 0.f;
    const u32 any = __ballot_sync(0xffffffffu, flag);
    for (int j = 0; j < 16; ++j) {
      if (!((any >> (2 * j)) & 3u)) continue;   // warp-uniform: neither item of the pair samples inside the image
      const int src = 2 * j + hsel;
      const int jf = __shfl_sync(0xffffffffu, flag, src);
      const int jy0 = __shfl_sync(0xffffffffu, y0, src), jx0 = __shfl_sync(0xffffffffu, x0, src);
      const float jly = __shfl_sync(0xffffffffu, ly, src), jlx = __shfl_sync(0xffffffffu, lx, src);
      const float jmk = __shfl_sync(0xffffffffu, mk, src);
      const long long jm = __shfl_sync(0xffffffffu, m, src);
      const int jtap = __shfl_sync(0xffffffffu, tap, src), jslab = __shfl_sync(0xffffffffu, slab, src);
      const int jn = __shfl_sync(0xffffffffu, n, src);
      float sy = 0.f, sx = 0.f, sm = 0.f;
      if (jf) {
        const int cbase = jslab * 64 + 4 * q;
        const float hy = 1.f - jly, hx = 1.f - jlx;
        const bool vy0 = jy0 >= 0, vy1 = jy0 + 1 <= H - 1, vx0 = jx0 >= 0, vx1 = jx0 + 1 <= W - 1;
        const bool valid[4] = {vy0 && vx0, vy0 && vx1, vy1 && vx0, vy1 && vx1};
        const float wgt[4] = {hy * hx, hy * jlx, jly * hx, jly * jlx};
        const uint2 gr = __ldg(reinterpret_cast<const uint2*>(dcol + ((size_t)jm * 9 + jtap) * C + cbase));
        const float2 g01 = bf2f(gr.x), g23 = bf2f(gr.y);
        const float g[4] = {g01.x, g01.y, g23.x, g23.y};
        float v[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cy = jy0 + (c >> 1), cx = jx0 + (c & 1);
#pragma unroll
          for (int e = 0; e < 4; ++e) v[c][e] = 0.f;
          if (valid[c]) {
            const size_t off = (((size_t)jn * H + cy) * W + cx) * C + cbase;
            const uint2 xr = __ldg(reinterpret_cast<const uint2*>(x + off));
            const float2 a01 = bf2f(xr.x), a23 = bf2f(xr.y);
            v[c][0] = a01.x; v[c][1] = a01.y; v[c][2] = a23.x; v[c][3] = a23.y;
            const float w = wgt[c] * jmk;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dx_acc + off), "f"(w * g[0]), "f"(w * g[1]),
                         "f"(w * g[2]), "f"(w * g[3])
                         : "memory");
          }
        }
        // value before modulation and its coordinate derivatives (torchvision get_coordinate_weight)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sm += g[e] * (wgt[0] * v[0][e] + wgt[1] * v[1][e] + wgt[2] * v[2][e] + wgt[3] * v[3][e]);
          sy += g[e] * (jlx * (v[3][e] - v[1][e]) + hx * (v[2][e] - v[0][e]));
          sx += g[e] * (jly * (v[3][e] - v[2][e]) + hy * (v[1][e] - v[0][e]));
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {   // half-warp sums
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
      }
      // hand the sums to the lane that owns the item: lane 2j takes half-warp 0's, lane 2j+1 half-warp 1's
      const float oy_ = __shfl_sync(0xffffffffu, sy, 16), ox_ = __shfl_sync(0xffffffffu, sx, 16), om_ = __shfl_sync(0xffffffffu, sm, 16);
      const float ey_ = __shfl_sync(0xffffffffu, sy, 0), ex_ = __shfl_sync(0xffffffffu, sx, 0), em_ = __shfl_sync(0xffffffffu, sm, 0);
      if (lane == 2 * j) {
        my_sy = ey_ * mk; my_sx = ex_ * mk; my_sm = em_ * mk * (1.f - mk);
      } else if (lane == 2 * j + 1) {
        my_sy = oy_ * mk; my_sx = ox_ * mk; my_sm = om_ * mk * (1.f - mk);
      }
    }
    if (flag) {
      float* dp = dom + (size_t)m * om_cstride;
      if (slabs == 1) {
        dp[2 * tap] = my_sy;
        dp[2 * tap + 1] = my_sx;
        dp[18 + tap] = my_sm;
      } else {   // several channel slabs add into the same three numbers (dom is zero-filled by the host wrapper)
        atomicAdd(dp + 2 * tap, my_sy);
        atomicAdd(dp + 2 * tap + 1, my_sx);
        atomicAdd(dp + 18 + tap, my_sm);
      }
    }
  }
}

// fp32 NHWC [M, cs] -> bf16 NHWC [M, cs] (offset/mask gradients feeding the offset conv's backward GEMMs)
__global__ void f32_to_bf16_kernel(const float* __restrict__ a, __nv_bfloat16* __restrict__ out, long long n8, int cs, int cvalid) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(a) + 2 * i), a1 = __ldg(reinterpret_cast<const float4*>(a) + 2 * i + 1);
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const int c0 = (int)((i * 8) % cs);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j >= cvalid) v[j] = 0.f;
    reinterpret_cast<uint4*>(out)[i] = pack8(v);
  }
}

// ---- Adam (torch.optim.Adam defaults: no weight decay, no amsgrad) on flat fp32 buffers ---------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, int step, const int* __restrict__ step_dev,
                            const float* __restrict__ lr_dev, float gscale) {
  // step / learning rate may live in device memory so that a captured CUDA graph of the training step sees them change
  const int t = step_dev ? *step_dev : step;
  const float lrv = lr_dev ? *lr_dev : lr;
  const float bc1 = 1.f - powf(b1, (float)t), bc2_sqrt = sqrtf(1.f - powf(b2, (float)t));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lrv / bc1) * (mi / denom);
  }
}

}  // namespace
}  // namespace cnb

using namespace cnb;
typedef const __nv_bfloat16* cbf;
typedef __nv_bfloat16* bf;

static int launch_reduce(int mode, const RedArgs& r, cudaStream_t st) {
  CNB_CHECK_ARG(r.C >= 1 && r.C <= 2048, "channel reduce: C=%d must be in [1, 2048]", r.C);
  CNB_CHECK_ARG(mode == 2 || r.C % 8 == 0, "channel reduce: BatchNorm maps need C %% 8 == 0");
  CNB_CHECK_ARG(r.a_coffset + (r.C + 7) / 8 * 8 <= r.a_cstride && r.a_cstride % 8 == 0 && r.a_coffset % 8 == 0,
                "channel reduce: the 8-channel chunks of [coffset, coffset + C) must lie inside the channel stride");
  CNB_CHECK_ARG(r.M >= 1, "channel reduce: empty map");
  const int tpr = (r.C + 7) / 8, rpp = 256 / tpr;
  long long want = (r.M + rpp - 1) / rpp;
  const int grid = (int)(want > 148 * 4 ? 148 * 4 : want);
  if (mode == 0) channel_reduce_kernel<0><<<grid, 256, 0, st>>>(r);
  else if (mode == 1) channel_reduce_kernel<1><<<grid, 256, 0, st>>>(r);
  else channel_reduce_kernel<2><<<grid, 256, 0, st>>>(r);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_bn_train_fwd(const void* z, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, float momentum, float eps, const void* res, int act, void* y,
                                float* stats, float* sums_ws, long long M, int C, cnb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CNB_CHECK_ARG(z && gamma && beta && y && stats && sums_ws, "bn_train_fwd: null pointer");
  CNB_CUDA(cudaMemsetAsync(sums_ws, 0, (size_t)2 * C * sizeof(float), st));
  RedArgs r{(cbf)z, nullptr, nullptr, nullptr, nullptr, sums_ws, M, C, C, 0};
  int rc = launch_reduce(0, r, st);
  if (rc) return rc;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums_ws, (float)M, eps, momentum, gamma, beta, running_mean,
                                                       running_var, stats, C);
  CNB_LAUNCH_CHECK();
  scale_shift_act_kernel<<<grid_for(M * (C / 8)), 256, 0, st>>>((cbf)z, stats + 2 * C, stats + 3 * C, (cbf)res, (bf)y, M, C, act);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_scale_shift_act(const void* z, const float* scale, const float* shift, const void* res, int act,
                                   void* y, long long M, int C, cnb_stream_t stream) {
  CNB_CHECK_ARG(z && scale && shift && y && C % 8 == 0, "scale_shift_act: bad argument");
  scale_shift_act_kernel<<<grid_for(M * (C / 8)), 256, 0, (cudaStream_t)stream>>>((cbf)z, scale, shift, (cbf)res, (bf)y, M, C, act);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_bn_train_bwd(const void* dy, const void* y_or_null, const void* z, const float* stats, void* dz,
                                void* dres_or_null, float* dgamma, float* dbeta, int accumulate, float* sums_ws,
                                long long M, int C, cnb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CNB_CHECK_ARG(dy && z && stats && dz && dgamma && dbeta && sums_ws, "bn_train_bwd: null pointer");
  CNB_CUDA(cudaMemsetAsync(sums_ws, 0, (size_t)2 * C * sizeof(float), st));
  RedArgs r{(cbf)dy, (cbf)y_or_null, (cbf)z, stats, stats + C, sums_ws, M, C, C, 0};
  int rc = launch_reduce(1, r, st);
  if (rc) return rc;
  bn_bwd_apply_kernel<<<grid_for(M * (C / 8)), 256, 0, st>>>((cbf)dy, (cbf)y_or_null, (cbf)z, stats, sums_ws,
                                                             1.f / (float)M, (bf)dz, (bf)dres_or_null, M, C);
  CNB_LAUNCH_CHECK();
  bn_param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums_ws, dgamma, dbeta, C, accumulate);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_channel_sum(const void* dy, int cstride, int coffset, const void* y_mask_or_null, float* out,
                               int accumulate, long long M, int C, cnb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CNB_CHECK_ARG(dy && out, "channel_sum: null pointer");
  CNB_CHECK_ARG(!y_mask_or_null || (cstride == C && coffset == 0 && C % 8 == 0), "channel_sum: the ReLU mask needs a dense map");
  if (!accumulate) CNB_CUDA(cudaMemsetAsync(out, 0, (size_t)C * sizeof(float), st));
  RedArgs r{(cbf)dy, (cbf)y_mask_or_null, nullptr, nullptr, nullptr, out, M, C, cstride, coffset};
  return launch_reduce(2, r, st);
}

extern "C" int cnb_relu_bwd(const void* y, const void* dy, void* dx, long long n, cnb_stream_t stream) {
  CNB_CHECK_ARG(y && dy && dx && n % 8 == 0, "relu_bwd: bad argument");
  relu_bwd_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>((cbf)y, (cbf)dy, (bf)dx, n / 8);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_add_bf16(const void* a, const void* b, void* out, long long n, cnb_stream_t stream) {
  CNB_CHECK_ARG(a && b && out && n % 8 == 0, "add_bf16: bad argument");
  add_bf16_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>((cbf)a, (cbf)b, (bf)out, n / 8);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_f32_to_bf16(const float* acc, const void* addend_or_null, void* out, long long n, int cstride,
                               int cvalid, cnb_stream_t stream) {
  CNB_CHECK_ARG(acc && out && n % 8 == 0, "f32_to_bf16: bad argument");
  if (cvalid > 0 && cvalid < cstride) {
    CNB_CHECK_ARG(!addend_or_null && cstride % 8 == 0, "f32_to_bf16: channel masking excludes the addend");
    f32_to_bf16_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>(acc, (bf)out, n / 8, cstride, cvalid);
  } else {
    f32_to_bf16_add_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>(acc, (cbf)addend_or_null, (bf)out, n / 8);
  }
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int B, int H, int W, int C, cnb_stream_t stream) {
  CNB_CHECK_ARG(x && dy && dx && C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2x2_bwd: bad argument");
  maxpool2_bwd_kernel<<<grid_for((long long)B * (H / 2) * (W / 2) * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
      (cbf)x, (cbf)dy, (bf)dx, B, H, W, C);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_dw_deconv_bwd(const void* x, const float* wt, const void* dy, void* dx, float* dwt_acc, int B, int H,
                                 int W, int C, int f, cnb_stream_t stream) {
  CNB_CHECK_ARG(x && wt && dy && dx && dwt_acc && C % 8 == 0 && f >= 2 && f % 2 == 0, "dw_deconv_bwd: bad argument");
  const size_t smem = (size_t)4 * f * f * C * sizeof(float);
  CNB_CHECK_ARG(smem <= 160 * 1024, "dw_deconv_bwd: (2f)^2 * C too large for shared memory");
  static PerDeviceOnce once;
  if (once.need()) {
    CNB_CUDA(cudaFuncSetAttribute(dw_deconv_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CNB_CUDA(cudaFuncSetAttribute(dw_deconv_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    once.mark();
  }
  const long long total = (long long)B * H * W * (C / 8);
  long long want = (total + 255) / 256;
  const int grid = (int)(want > 148 * 2 ? 148 * 2 : want);
  // the register path needs a thread's channel group to stay fixed: grid stride (grid * 256) % (C / 8) == 0
  if (f == 2 && 256 % (C / 8) == 0)
    dw_deconv_bwd_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>((cbf)x, wt, (cbf)dy, (bf)dx, dwt_acc, B, H, W, C, f);
  else
    dw_deconv_bwd_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>((cbf)x, wt, (cbf)dy, (bf)dx, dwt_acc, B, H, W, C, f);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_dw_deconv_unpack_wgrad(const float* dwt_acc, float* dw, int C, int f, int accumulate, cnb_stream_t stream) {
  CNB_CHECK_ARG(dwt_acc && dw, "dw_deconv_unpack_wgrad: null pointer");
  dw_deconv_unpack_wgrad_kernel<<<grid_for((long long)C * 4 * f * f), 256, 0, (cudaStream_t)stream>>>(dwt_acc, dw, C, 4 * f * f, accumulate);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_dcnv2_im2col(const void* x, const float* om, int om_cstride, void* col, int B, int H, int W, int C,
                                cnb_stream_t stream) {
  CNB_CHECK_ARG(x && om && col && C % 8 == 0 && om_cstride >= 27, "dcnv2_im2col: bad argument");
  CNB_CHECK_ARG((long long)B * H * W * C < (1ll << 31), "dcnv2_im2col: tensor too large for 32-bit element offsets");
  dcn_im2col_kernel<<<grid_for((long long)B * H * W * 9 * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
      (cbf)x, om, om_cstride, (bf)col, B, H, W, C);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_dcnv2_col2im(const void* x, const float* om, int om_cstride, const void* dcol, float* dx_acc,
                                float* dom, int B, int H, int W, int C, cnb_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CNB_CHECK_ARG(x && om && dcol && dx_acc && dom && C % 64 == 0 && om_cstride >= 27, "dcnv2_col2im: bad argument (C %% 64 == 0)");
  CNB_CHECK_ARG((long long)B * H * W * C < (1ll << 31), "dcnv2_col2im: tensor too large for 32-bit element offsets");
  CNB_CUDA(cudaMemsetAsync(dx_acc, 0, (size_t)B * H * W * C * sizeof(float), st));
  CNB_CUDA(cudaMemsetAsync(dom, 0, (size_t)B * H * W * om_cstride * sizeof(float), st));
  const long long items = (long long)B * H * W * 9 * (C / 64);
  dcn_col2im_kernel<<<grid_for(items), 256, 0, st>>>((cbf)x, om, om_cstride, (cbf)dcol, dx_acc, dom, B, H, W, C);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                             float beta2, float eps, int step, const int* step_dev, const float* lr_dev, float grad_scale,
                             cnb_stream_t stream) {
  CNB_CHECK_ARG(p && g && m && v && n >= 1 && (step >= 1 || step_dev), "adam_step: bad argument");
  adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, step, step_dev, lr_dev, grad_scale);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
