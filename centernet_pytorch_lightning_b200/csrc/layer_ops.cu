// HBM-bound layer ops of the DLA-34 / ResNet-DCN forward, NHWC bf16, 16-byte (8-channel) vectors.
//
// Replaces (reference file:line under CenterNet/models/backbones/):
//   pose_dla_dcn.py:243      nn.MaxPool2d(stride, stride)            -> maxpool_kernel
//   pose_dla_dcn.py:466-488  depthwise bilinear ConvTranspose2d + `layers[i] + layers[i-1]` -> dw_deconv_kernel
//   input / output layout changes around the NHWC bf16 engine        -> nchw_to_nhwc / nhwc_to_nchw kernels
#include "cnb_common.cuh"

namespace cnb {
namespace {

__device__ __forceinline__ float2 bf2_to_f2(u32 v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ u32 f2_to_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<u32*>(&h);
}
__device__ __forceinline__ u32 bf2_max(u32 a, u32 b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<u32*>(&r);
}

// ---- NCHW fp32 -> NHWC bf16 with zero channel padding (C -> Cp, Cp % 8 == 0) -------------------------
// one thread per (pixel, 8-channel group); reads are coalesced along W for each channel plane
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int C,
                                    int HW, int Cp) {
  const int groups = Cp / 8;
  const long long total = (long long)B * HW * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix_all = i % ((long long)B * HW);   // pixel fastest -> coalesced plane reads
    const int g = (int)(i / ((long long)B * HW));
    const int b = (int)(pix_all / HW);
    const int pix = (int)(pix_all % HW);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      v[j] = c < C ? __ldg(x + ((size_t)b * C + c) * HW + pix) : 0.f;
    }
    uint4 o = make_uint4(f2_to_bf2(v[0], v[1]), f2_to_bf2(v[2], v[3]), f2_to_bf2(v[4], v[5]), f2_to_bf2(v[6], v[7]));
    *reinterpret_cast<uint4*>(y + ((size_t)b * HW + pix) * Cp + g * 8) = o;
  }
}

// The same for Cp == 4 (C <= 4): 8 bytes per pixel.  Two neighbouring pixels of this tensor are one 8-channel
// "super-pixel" of the space-to-depth stem (ops.pack_stem_s2d_weights).
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int C,
                                     int HW) {
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int pix = (int)(i - (long long)b * HW);
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = c < C ? __ldg(x + ((size_t)b * C + c) * HW + pix) : 0.f;
    *reinterpret_cast<uint2*>(y + (size_t)i * 4) = make_uint2(f2_to_bf2(v[0], v[1]), f2_to_bf2(v[2], v[3]));
  }
}

// ---- NHWC bf16 (channel slice) -> NCHW fp32 -------------------------------------------------------------
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int B, int C,
                                    int HW, int cstride, int coffset) {
  const long long total = (long long)B * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const long long bc = i / HW;
    const int c = (int)(bc % C), b = (int)(bc / C);
    y[i] = __bfloat162float(x[((size_t)b * HW + pix) * cstride + coffset + c]);
  }
}

// ---- MaxPool2d(k, stride, pad) with -inf padding (resnet stem: 3x3, stride 2, pad 1) ------------------------
__global__ void maxpool_pad_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H,
                                   int W, int C, int Ho, int Wo, int k, int stride, int pad) {
  const int groups = C / 8;
  const long long total = (long long)B * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long p = i / groups;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    uint4 m = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);   // bf16 -inf pairs
    for (int dy = 0; dy < k; ++dy) {
      const int iy = oy * stride - pad + dy;
      if (iy < 0 || iy >= H) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int ix = ox * stride - pad + dx;
        if (ix < 0 || ix >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * H + iy) * W + ix) * C + g * 8));
        m.x = bf2_max(m.x, v.x); m.y = bf2_max(m.y, v.y); m.z = bf2_max(m.z, v.z); m.w = bf2_max(m.w, v.w);
      }
    }
    *reinterpret_cast<uint4*>(y + ((size_t)(b * Ho + oy) * Wo + ox) * C + g * 8) = m;
  }
}

// ---- depth-to-space (pixel shuffle) by 2: x [B,H,W,4*C] (phase-major channel blocks, phase = py*2+px) ->
//      y [B,2H,2W,C];  the second half of a dense ConvTranspose2d(4, stride 2, pad 1) run as one 3x3 conv ------
__global__ void depth_to_space2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B,
                                       int H, int W, int C) {
  const int groups = C / 8;
  const long long total = (long long)B * H * W * 4 * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long p = i / groups;
    const int ph = (int)(p % 4);
    p /= 4;
    const int ix = (int)(p % W);
    p /= W;
    const int iy = (int)(p % H);
    const int b = (int)(p / H);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)(b * H + iy) * W + ix) * 4 + ph) * C + g * 8));
    const int oy = 2 * iy + (ph >> 1), ox = 2 * ix + (ph & 1);
    *reinterpret_cast<uint4*>(y + ((size_t)(b * 2 * H + oy) * (2 * W) + ox) * C + g * 8) = v;
  }
}

// ---- MaxPool2d(k, stride=k), floor mode -----------------------------------------------------------------
__global__ void maxpool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H,
                               int W, int C, int xcs, int xco, int ycs, int yco, int k) {
  const int Ho = H / k, Wo = W / k, groups = C / 8;
  const long long total = (long long)B * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long p = i / groups;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    uint4 m;
    bool first = true;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(
            x + ((size_t)(b * H + oy * k + dy) * W + ox * k + dx) * xcs + xco + g * 8));
        if (first) {
          m = v;
          first = false;
        } else {
          m.x = bf2_max(m.x, v.x); m.y = bf2_max(m.y, v.y); m.z = bf2_max(m.z, v.z); m.w = bf2_max(m.w, v.w);
        }
      }
    *reinterpret_cast<uint4*>(y + ((size_t)(b * Ho + oy) * Wo + ox) * ycs + yco + g * 8) = m;
  }
}

// ---- depthwise ConvTranspose2d(C, C, 2f, stride=f, padding=f/2, groups=C, bias=False) [+ add] -----------
// out[oy,ox,c] = sum_{ky,kx} w[c,ky,kx] * in[iy,ix,c]  with  oy = iy*f - f/2 + ky  (<= 2x2 contributing inputs)
// w is passed re-laid-out as [2f*2f][C] fp32 so that 8 channels of one tap are contiguous.
__global__ void dw_deconv_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt,
                                 const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ y, int B,
                                 int H, int W, int C, int f) {
  const int Ho = H * f, Wo = W * f, groups = C / 8, ks = 2 * f, pad = f / 2;
  const long long total = (long long)B * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long p = i / groups;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // contributing input rows: iy with 0 <= oy + pad - iy*f < 2f
    const int ty = oy + pad, tx = ox + pad;
    const int iy_hi = ty / f, ix_hi = tx / f;
    for (int a = 0; a < 2; ++a) {
      const int iy = iy_hi - a;
      const int ky = ty - iy * f;
      if (iy < 0 || iy >= H || ky < 0 || ky >= ks) continue;
      for (int c2 = 0; c2 < 2; ++c2) {
        const int ix = ix_hi - c2;
        const int kx = tx - ix * f;
        if (ix < 0 || ix >= W || kx < 0 || kx >= ks) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * H + iy) * W + ix) * C + g * 8));
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(ky * ks + kx) * C + g * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(ky * ks + kx) * C + g * 8 + 4));
        const float2 a0 = bf2_to_f2(v.x), a1 = bf2_to_f2(v.y), a2 = bf2_to_f2(v.z), a3 = bf2_to_f2(v.w);
        acc[0] += w0.x * a0.x; acc[1] += w0.y * a0.y; acc[2] += w0.z * a1.x; acc[3] += w0.w * a1.y;
        acc[4] += w1.x * a2.x; acc[5] += w1.y * a2.y; acc[6] += w1.z * a3.x; acc[7] += w1.w * a3.y;
      }
    }
    const size_t o = ((size_t)(b * Ho + oy) * Wo + ox) * C + g * 8;
    if (add) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(add + o));
      const float2 a0 = bf2_to_f2(v.x), a1 = bf2_to_f2(v.y), a2 = bf2_to_f2(v.z), a3 = bf2_to_f2(v.w);
      acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
      acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
    }
    *reinterpret_cast<uint4*>(y + o) = make_uint4(f2_to_bf2(acc[0], acc[1]), f2_to_bf2(acc[2], acc[3]),
                                                  f2_to_bf2(acc[4], acc[5]), f2_to_bf2(acc[6], acc[7]));
  }
}

// The same with the taps held in registers.  All outputs of one phase (py, px) = ((oy + f/2) % f, (ox + f/2) % f) use
// the same four filter taps {py, py+f} x {px, px+f}; blockIdx.y picks the phase, a thread keeps its 8-channel group
// fixed, so its 4 x 8 weights are loaded once and every output costs 4 input loads + (add) + 1 store instead of
// 13 loads (the weight loads were most of the L1 wavefronts of the generic kernel above).
__global__ void __launch_bounds__(256)
dw_deconv_phase_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt,
                       const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ y, int B, int H, int W,
                       int C, int f) {
  const int Ho = H * f, Wo = W * f, groups = C / 8, ks = 2 * f, pad = f / 2;
  const int py = blockIdx.y / f, px = blockIdx.y - py * f;
  const int g = threadIdx.x % groups;               // blockDim.x and the grid stride are multiples of `groups`
  float w[4][8];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int ky = py + (t >> 1) * f, kx = px + (t & 1) * f;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(ky * ks + kx) * C + g * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(ky * ks + kx) * C + g * 8 + 4));
    w[t][0] = w0.x; w[t][1] = w0.y; w[t][2] = w0.z; w[t][3] = w0.w;
    w[t][4] = w1.x; w[t][5] = w1.y; w[t][6] = w1.z; w[t][7] = w1.w;
  }
  const int QW = W + 1, QH = H + 1;
  const long long total = (long long)B * QH * QW * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long p = i / groups;
    const int qx = (int)(p % QW);
    p /= QW;
    const int qy = (int)(p % QH);
    const int b = (int)(p / QH);
    const int oy = qy * f + py - pad, ox = qx * f + px - pad;
    if (oy < 0 || oy >= Ho || ox < 0 || ox >= Wo) continue;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {                    // tap (ky, kx) = (py + a*f, px + c*f) reads input (qy - a, qx - c)
      const int iy = qy - (t >> 1), ix = qx - (t & 1);
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * H + iy) * W + ix) * C + g * 8));
      const float2 a0 = bf2_to_f2(v.x), a1 = bf2_to_f2(v.y), a2 = bf2_to_f2(v.z), a3 = bf2_to_f2(v.w);
      acc[0] += w[t][0] * a0.x; acc[1] += w[t][1] * a0.y; acc[2] += w[t][2] * a1.x; acc[3] += w[t][3] * a1.y;
      acc[4] += w[t][4] * a2.x; acc[5] += w[t][5] * a2.y; acc[6] += w[t][6] * a3.x; acc[7] += w[t][7] * a3.y;
    }
    const size_t o = ((size_t)(b * Ho + oy) * Wo + ox) * C + g * 8;
    if (add) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(add + o));
      const float2 a0 = bf2_to_f2(v.x), a1 = bf2_to_f2(v.y), a2 = bf2_to_f2(v.z), a3 = bf2_to_f2(v.w);
      acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
      acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
    }
    *reinterpret_cast<uint4*>(y + o) = make_uint4(f2_to_bf2(acc[0], acc[1]), f2_to_bf2(acc[2], acc[3]),
                                                  f2_to_bf2(acc[4], acc[5]), f2_to_bf2(acc[6], acc[7]));
  }
}

// The phase kernel without per-element index division and with the taps amortised: 3-D grid (x: input column x
// channel group, y: block of RPT cell rows, z: image x phase); a thread keeps its phase's 4 x 8 taps in registers,
// walks RPT rows down its column and rolls the two input pixels of the previous row through registers: per output
// 2 input loads + 1 (add) load + 1 store.  (The grid-stride version above spends more issue slots on 64-bit div/mod
// than on the blend, and reloading the 8 tap vectors per output made the L1 wavefront count the bound.)
constexpr int UP_RPT = 8;
__global__ void __launch_bounds__(256)
dw_deconv_phase3d_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt,
                         const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ y, int B, int H, int W,
                         int C, int f, int gshift) {
  const int Ho = H * f, Wo = W * f, groups = C / 8, ks = 2 * f, pad = f / 2;
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int qx = t >> gshift, g = t & (groups - 1);
  const int ff = f * f;
  const int b = blockIdx.z / ff, phase = blockIdx.z - b * ff;
  const int py = phase / f, px = phase - py * f;
  const int ox = qx * f + px - pad;
  // taps of this block's phase in shared memory ([4 taps][C] fp32, <= 4 KB): registers go to loads in flight instead
  __shared__ __align__(16) float s_w[4 * 256];
  const bool active = qx <= W && ox >= 0 && ox < Wo;
  for (int i = threadIdx.x; i < 4 * C; i += 256) {
    const int tp = i / C, c = i - tp * C;
    const int ky = py + (tp >> 1) * f, kx = px + (tp & 1) * f;     // tap (ky, kx) reads input (qy - a, qx - c)
    s_w[i] = __ldg(wt + (size_t)(ky * ks + kx) * C + c);
  }
  __syncthreads();
  if (!active) return;
  auto load_q = [&](int iy, int ix) -> uint4 {
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) return make_uint4(0u, 0u, 0u, 0u);
    return __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * H + iy) * W + ix) * C + g * 8));
  };
  auto out_off = [&](int qy, bool* ok) -> size_t {
    const int oy = qy * f + py - pad;
    *ok = oy >= 0 && oy < Ho;
    return ((size_t)(b * Ho + (*ok ? oy : 0)) * Wo + ox) * C + g * 8;
  };
  const int qy0 = blockIdx.y * UP_RPT;
  const int qy1 = min(qy0 + UP_RPT, H + 1);
  uint4 v[4];                                          // v[tp]: input (qy - (tp>>1), qx - (tp&1))
  v[2] = load_q(qy0 - 1, qx);
  v[3] = load_q(qy0 - 1, qx - 1);
  v[0] = load_q(qy0, qx);
  v[1] = load_q(qy0, qx - 1);
  bool ok;
  size_t o = out_off(qy0, &ok);
  uint4 av = (add && ok) ? __ldg(reinterpret_cast<const uint4*>(add + o)) : make_uint4(0u, 0u, 0u, 0u);
  for (int qy = qy0; qy < qy1; ++qy) {
    // next row's loads first: they are in flight while this row is blended
    const bool more = qy + 1 < qy1;
    const uint4 n0 = more ? load_q(qy + 1, qx) : make_uint4(0u, 0u, 0u, 0u);
    const uint4 n1 = more ? load_q(qy + 1, qx - 1) : make_uint4(0u, 0u, 0u, 0u);
    bool nok = false;
    const size_t no = more ? out_off(qy + 1, &nok) : 0;
    const uint4 nav = (add && nok) ? __ldg(reinterpret_cast<const uint4*>(add + no)) : make_uint4(0u, 0u, 0u, 0u);
    if (ok) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) {
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[tp * C + g * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s_w[tp * C + g * 8 + 4]);
        const float2 a0 = bf2_to_f2(v[tp].x), a1 = bf2_to_f2(v[tp].y), a2 = bf2_to_f2(v[tp].z), a3 = bf2_to_f2(v[tp].w);
        acc[0] += w0.x * a0.x; acc[1] += w0.y * a0.y; acc[2] += w0.z * a1.x; acc[3] += w0.w * a1.y;
        acc[4] += w1.x * a2.x; acc[5] += w1.y * a2.y; acc[6] += w1.z * a3.x; acc[7] += w1.w * a3.y;
      }
      if (add) {
        const float2 a0 = bf2_to_f2(av.x), a1 = bf2_to_f2(av.y), a2 = bf2_to_f2(av.z), a3 = bf2_to_f2(av.w);
        acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
        acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
      }
      *reinterpret_cast<uint4*>(y + o) = make_uint4(f2_to_bf2(acc[0], acc[1]), f2_to_bf2(acc[2], acc[3]),
                                                    f2_to_bf2(acc[4], acc[5]), f2_to_bf2(acc[6], acc[7]));
    }
    v[2] = v[0];
    v[3] = v[1];
    v[0] = n0;
    v[1] = n1;
    av = nav;
    o = no;
    ok = nok;
  }
}

// Tile kernel.  What bounds the kernels above is memory-level parallelism, not bandwidth: a thread has 3 x 16 bytes in
// flight and waits a full DRAM latency per output row (768 threads x 48 B = 37 KB per SM, i.e. ~2.5 TB/s at ~2 us loaded
// latency -- where every variant sat; torch's copy_ of the same bytes runs at 5.4 TB/s).  Here a CTA requests everything
// its tile needs up front with cp.async -- the (tqh + 1) x 17 input pixels and the tqh*f x 16*f pixels of the `add` map,
// ~43 KB -- five such CTAs per SM keep > 200 KB in flight, and the blend then runs out of shared memory.  A thread keeps
// (8-channel group, phase) fixed (its 4 x 8 taps in registers); same tap order as the kernels above: bit-identical results.
constexpr int UPT_W = 16;
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__global__ void __launch_bounds__(256)
dw_deconv_tile_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt,
                      const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ y, int B, int H, int W, int C,
                      int f, int tqh) {
  extern __shared__ __align__(16) unsigned char up_smem[];
  const int Ho = H * f, Wo = W * f, groups = C / 8, ks = 2 * f, pad = f / 2;
  const int npix = (tqh + 1) * (UPT_W + 1);                // input pixels of the tile
  const int arows = tqh * f, acols = UPT_W * f;            // output / add pixels of the tile
  uint4* s_in = reinterpret_cast<uint4*>(up_smem);         // [npix][groups]
  uint4* s_add = s_in + npix * groups;                     // [arows][acols][groups]
  const int b = blockIdx.z, qy0 = blockIdx.y * tqh, qx0 = blockIdx.x * UPT_W;
  const int oy0 = qy0 * f - pad, ox0 = qx0 * f - pad;      // first output pixel of the tile (may be -pad)
  // ---- everything the tile reads, requested at once ------------------------------------------------------
  for (int i = threadIdx.x; i < npix * groups; i += 256) {
    const int p = i / groups, gg = i - p * groups;
    const int ry = p / (UPT_W + 1), rx = p - ry * (UPT_W + 1);
    const int iy = qy0 - 1 + ry, ix = qx0 - 1 + rx;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) cp_async16(s_in + i, x + ((size_t)(b * H + iy) * W + ix) * C + gg * 8);
    else s_in[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (add) {
    for (int i = threadIdx.x; i < arows * acols * groups; i += 256) {
      const int p = i / groups, gg = i - p * groups;
      const int ry = p / acols, rx = p - ry * acols;
      const int oy = oy0 + ry, ox = ox0 + rx;
      if (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) cp_async16(s_add + i, add + ((size_t)(b * Ho + oy) * Wo + ox) * C + gg * 8);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // ---- this thread's taps while the copies fly ---------------------------------------------------------------
  const int per = groups * f * f;                          // threads per cell: (phase, channel group)
  const int idx = threadIdx.x % per, slot = threadIdx.x / per, nslot = 256 / per;
  const int phase = idx / groups, g = idx - phase * groups;
  const int py = phase / f, px = phase - py * f;
  float w[4][8];
#pragma unroll
  for (int t = 0; t < 4; ++t) {                            // tap (ky, kx) = (py + a f, px + c f) reads input (qy - a, qx - c)
    const int ky = py + (t >> 1) * f, kx = px + (t & 1) * f;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(ky * ks + kx) * C + g * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(ky * ks + kx) * C + g * 8 + 4));
    w[t][0] = w0.x; w[t][1] = w0.y; w[t][2] = w0.z; w[t][3] = w0.w;
    w[t][4] = w1.x; w[t][5] = w1.y; w[t][6] = w1.z; w[t][7] = w1.w;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  for (int c = slot; c < tqh * UPT_W; c += nslot) {
    const int cy = c / UPT_W, cx = c - cy * UPT_W;
    const int qy = qy0 + cy, qx = qx0 + cx;
    const int ry = cy * f + py, rx = cx * f + px;          // position inside the tile's output block
    const int oy = oy0 + ry, ox = ox0 + rx;
    if (qy > H || qx > W || oy < 0 || oy >= Ho || ox < 0 || ox >= Wo) continue;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const uint4 v = s_in[((cy + 1 - (t >> 1)) * (UPT_W + 1) + (cx + 1 - (t & 1))) * groups + g];
      const float2 a0 = bf2_to_f2(v.x), a1 = bf2_to_f2(v.y), a2 = bf2_to_f2(v.z), a3 = bf2_to_f2(v.w);
      acc[0] += w[t][0] * a0.x; acc[1] += w[t][1] * a0.y; acc[2] += w[t][2] * a1.x; acc[3] += w[t][3] * a1.y;
      acc[4] += w[t][4] * a2.x; acc[5] += w[t][5] * a2.y; acc[6] += w[t][6] * a3.x; acc[7] += w[t][7] * a3.y;
    }
    if (add) {
      const uint4 av = s_add[(ry * acols + rx) * groups + g];
      const float2 a0 = bf2_to_f2(av.x), a1 = bf2_to_f2(av.y), a2 = bf2_to_f2(av.z), a3 = bf2_to_f2(av.w);
      acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
      acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
    }
    *reinterpret_cast<uint4*>(y + ((size_t)(b * Ho + oy) * Wo + ox) * C + g * 8) =
        make_uint4(f2_to_bf2(acc[0], acc[1]), f2_to_bf2(acc[2], acc[3]), f2_to_bf2(acc[4], acc[5]), f2_to_bf2(acc[6], acc[7]));
  }
}

// f == 2 (every large up-sampling of the DLA-34 / ResNet nets): a thread owns (8-channel group, input column qx, output
// row phase py) and walks down the rows.  Its 2 x 4 x 8 taps sit in registers, the two input rows of a cell roll
// through registers, and each cell yields the two horizontally adjacent outputs: 2 input loads + 2 (add) loads +
// 2 stores per 2 outputs (the phase kernel: 5 loads + 1 store per output, plus 64-bit index arithmetic per output;
// block / grid indices replace all of that here).  Same tap order as above -> bit-identical results.
__global__ void __launch_bounds__(128, 4)
dw_deconv_f2_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt,
                    const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ y, int B, int H, int W, int C,
                    int rows_per_block) {
  constexpr int F = 2, KS = 4;
  const int Ho = H * F, Wo = W * F, groups = C / 8;
  const int t = blockIdx.x * 128 + threadIdx.x;
  const int qx = t / groups, g = t - qx * groups;
  if (qx > W) return;
  const int b = blockIdx.z >> 1, py = blockIdx.z & 1;
  float w[2][KS][8];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) {
      const float* wp = wt + (size_t)((py + a * F) * KS + kx) * C + g * 8;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
      w[a][kx][0] = w0.x; w[a][kx][1] = w0.y; w[a][kx][2] = w0.z; w[a][kx][3] = w0.w;
      w[a][kx][4] = w1.x; w[a][kx][5] = w1.y; w[a][kx][6] = w1.z; w[a][kx][7] = w1.w;
    }
  // raw 16-byte loads (zero outside the image); converted to fp32 only where they are used, so that the loads of
  // the NEXT row and of this row's `add` values are in flight while the current row is blended
  auto load_q = [&](int iy, int ix) -> uint4 {
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) return make_uint4(0u, 0u, 0u, 0u);
    return __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * H + iy) * W + ix) * C + g * 8));
  };
  auto unpack8 = [](const uint4& q, float (&v)[8]) {
    const float2 a0 = bf2_to_f2(q.x), a1 = bf2_to_f2(q.y), a2 = bf2_to_f2(q.z), a3 = bf2_to_f2(q.w);
    v[0] = a0.x; v[1] = a0.y; v[2] = a1.x; v[3] = a1.y; v[4] = a2.x; v[5] = a2.y; v[6] = a3.x; v[7] = a3.y;
  };
  const int qy0 = blockIdx.y * rows_per_block;
  const int qy1 = min(qy0 + rows_per_block, H + 1);
  uint4 up0 = load_q(qy0 - 1, qx), up1 = load_q(qy0 - 1, qx - 1);      // input row qy-1 at columns qx / qx-1
  uint4 cur0 = load_q(qy0, qx), cur1 = load_q(qy0, qx - 1);            // input row qy
  for (int qy = qy0; qy < qy1; ++qy) {
    const int oy = qy * F + py - 1;
    const bool row_ok = oy >= 0 && oy < Ho;
    const int ox0 = qx * F - 1;
    const bool ok0 = row_ok && ox0 >= 0, ok1 = row_ok && ox0 + 1 < Wo;
    const size_t o0 = ((size_t)(b * Ho + (row_ok ? oy : 0)) * Wo + (ox0 >= 0 ? ox0 : 0)) * C + g * 8;
    const size_t o1 = ((size_t)(b * Ho + (row_ok ? oy : 0)) * Wo + (ox0 + 1 < Wo ? ox0 + 1 : 0)) * C + g * 8;
    uint4 ad0 = make_uint4(0u, 0u, 0u, 0u), ad1 = ad0;
    if (add && ok0) ad0 = __ldg(reinterpret_cast<const uint4*>(add + o0));
    if (add && ok1) ad1 = __ldg(reinterpret_cast<const uint4*>(add + o1));
    const uint4 nxt0 = qy + 1 < qy1 ? load_q(qy + 1, qx) : make_uint4(0u, 0u, 0u, 0u);
    const uint4 nxt1 = qy + 1 < qy1 ? load_q(qy + 1, qx - 1) : make_uint4(0u, 0u, 0u, 0u);
    if (row_ok) {
      float c0[8], c1[8], u0[8], u1[8];
      unpack8(cur0, c0);
      unpack8(cur1, c1);
      unpack8(up0, u0);
      unpack8(up1, u1);
#pragma unroll
      for (int px = 0; px < F; ++px) {
        if (!(px ? ok1 : ok0)) continue;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // taps (ky, kx) = (py, px), (py, px+2), (py+2, px), (py+2, px+2)
          acc[j] = w[0][px][j] * c0[j];
          acc[j] += w[0][px + F][j] * c1[j];
          acc[j] += w[1][px][j] * u0[j];
          acc[j] += w[1][px + F][j] * u1[j];
        }
        if (add) {
          float av[8];
          unpack8(px ? ad1 : ad0, av);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += av[j];
        }
        *reinterpret_cast<uint4*>(y + (px ? o1 : o0)) =
            make_uint4(f2_to_bf2(acc[0], acc[1]), f2_to_bf2(acc[2], acc[3]), f2_to_bf2(acc[4], acc[5]),
                       f2_to_bf2(acc[6], acc[7]));
      }
    }
    up0 = cur0;
    up1 = cur1;
    cur0 = nxt0;
    cur1 = nxt1;
  }
}

// [C,1,ks,ks] fp32 -> [ks*ks][C] fp32
__global__ void dw_weight_relayout_kernel(const float* __restrict__ w, float* __restrict__ wt, int C, int kk) {
  const int total = C * kk;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, t = i / C;
    wt[i] = w[(size_t)c * kk + t];
  }
}

inline int grid_for(long long total, int threads = 256) {
  long long g = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" int cnb_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int C, int H, int W, int C_pad,
                                         cnb_stream_t s) {
  CNB_CHECK_ARG(x && y && B >= 1 && C >= 1 && H >= 1 && W >= 1, "nchw_to_nhwc: bad argument");
  CNB_CHECK_ARG(C_pad >= C && (C_pad % 8 == 0 || C_pad == 4),
                "nchw_to_nhwc: C_pad must be >= C and a multiple of 8 (or 4)");
  if (C_pad == 4) {
    nchw_to_nhwc4_kernel<<<grid_for((long long)B * H * W), 256, 0, (cudaStream_t)s>>>(x, (__nv_bfloat16*)y, B, C,
                                                                                      H * W);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
  }
  const long long total = (long long)B * H * W * (C_pad / 8);
  nchw_to_nhwc_kernel<<<grid_for(total), 256, 0, (cudaStream_t)s>>>(x, (__nv_bfloat16*)y, B, C, H * W, C_pad);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int C, int H, int W, int x_cstride,
                                         int x_coffset, cnb_stream_t s) {
  CNB_CHECK_ARG(x && y && B >= 1 && C >= 1 && H >= 1 && W >= 1 && x_cstride >= C, "nhwc_to_nchw: bad argument");
  const long long total = (long long)B * C * H * W;
  nhwc_to_nchw_kernel<<<grid_for(total), 256, 0, (cudaStream_t)s>>>((const __nv_bfloat16*)x, y, B, C, H * W,
                                                                    x_cstride, x_coffset);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_maxpool2d(const void* x, void* y, int B, int H, int W, int C, int x_cstride, int x_coffset,
                             int y_cstride, int y_coffset, int k, cnb_stream_t s) {
  CNB_CHECK_ARG(x && y && B >= 1 && H >= k && W >= k && k >= 1, "maxpool2d: bad argument");
  CNB_CHECK_ARG(C % 8 == 0 && x_cstride % 8 == 0 && x_coffset % 8 == 0 && y_cstride % 8 == 0 && y_coffset % 8 == 0,
                "maxpool2d: channel counts/strides/offsets must be multiples of 8");
  const long long total = (long long)B * (H / k) * (W / k) * (C / 8);
  maxpool_kernel<<<grid_for(total), 256, 0, (cudaStream_t)s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, H, W,
                                                               C, x_cstride, x_coffset, y_cstride, y_coffset, k);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_dw_deconv_relayout_weights(const float* w, float* wt, int C, int f, cnb_stream_t s) {
  CNB_CHECK_ARG(w && wt && C >= 1 && f >= 1, "dw_deconv_relayout_weights: bad argument");
  const int kk = 4 * f * f;
  dw_weight_relayout_kernel<<<grid_for((long long)C * kk), 256, 0, (cudaStream_t)s>>>(w, wt, C, kk);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_dw_deconv_up(const void* x, const float* wt, const void* add, void* y, int B, int H, int W,
                                int C, int f, cnb_stream_t s) {
  CNB_CHECK_ARG(x && wt && y && B >= 1 && H >= 1 && W >= 1, "dw_deconv_up: bad argument");
  CNB_CHECK_ARG(C % 8 == 0 && f >= 1 && f % 2 == 0, "dw_deconv_up: C %% 8 == 0 and even upsampling factor required");
  const int groups = C / 8;
  static const int up_tile = [] { const char* e = getenv("CNB_DW_DECONV_IMPL"); return e ? atoi(e) : 4; }();
  {
    const int per = groups * f * f;
    const int tqh = 128 / (f * f * groups);   // add tile (tqh f) x (16 f) pixels x C channels <= 32 KB
    // cp.async tile kernel: default for the large outputs, where it measured faster (B=32: 64x64 -> 128x128 59.7 -> 53.4 us,
    // 32x32 -> 128x128 (f=4) 65.2 -> 54.0 us; 33 vs 33 us and 24 vs 22 us at 64x64 / 32x32 outputs); CNB_DW_DECONV_IMPL=5
    // forces it everywhere, =3 / =2 select the phase kernels
    if ((up_tile == 5 || (up_tile == 4 && (long long)H * f >= 128)) && per <= 256 && 256 % per == 0 && tqh >= 1 && B <= 65535) {
      const size_t smem = ((size_t)(tqh + 1) * (UPT_W + 1) + (size_t)tqh * f * UPT_W * f) * groups * 16;
      static PerDeviceOnce once;
      if (once.need()) {
        CNB_CUDA(cudaFuncSetAttribute(dw_deconv_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        once.mark();
      }
      if (smem <= 64 * 1024) {
        dim3 grid((unsigned)((W + 1 + UPT_W - 1) / UPT_W), (unsigned)((H + 1 + tqh - 1) / tqh), (unsigned)B);
        dw_deconv_tile_kernel<<<grid, 256, smem, (cudaStream_t)s>>>((const __nv_bfloat16*)x, wt, (const __nv_bfloat16*)add,
                                                                    (__nv_bfloat16*)y, B, H, W, C, f, tqh);
        CNB_LAUNCH_CHECK();
        return CNB_OK;
      }
    }
  }
  static const bool no_f2 = [] { const char* e = getenv("CNB_DW_DECONV_F2"); return e && e[0] == '0'; }();
  static const int up_impl0 = [] { const char* e = getenv("CNB_DW_DECONV_IMPL"); return e ? atoi(e) : 3; }();
  if (f == 2 && !no_f2 && B <= 32767 && H >= 64 && up_impl0 == 2) {   // measured: 71 vs 78 us at 64x64 -> 128x128, no gain on small maps
    const int rows_per_block = 8;
    dim3 grid((unsigned)(((long long)(W + 1) * groups + 127) / 128), (unsigned)((H + 1 + rows_per_block - 1) / rows_per_block),
              (unsigned)(B * 2));
    dw_deconv_f2_kernel<<<grid, 128, 0, (cudaStream_t)s>>>((const __nv_bfloat16*)x, wt, (const __nv_bfloat16*)add,
                                                            (__nv_bfloat16*)y, B, H, W, C, rows_per_block);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
  }
  static const int up_impl = [] { const char* e = getenv("CNB_DW_DECONV_IMPL"); return e ? atoi(e) : 3; }();
  if (up_impl == 3 && (groups & (groups - 1)) == 0 && f <= 8 && C <= 256 && (long long)B * f * f <= 65535 && H + 1 <= 65535) {
    int gshift = 0;
    while ((1 << gshift) < groups) ++gshift;
    dim3 grid((unsigned)(((long long)(W + 1) * groups + 255) / 256), (unsigned)((H + UP_RPT) / UP_RPT), (unsigned)(B * f * f));
    dw_deconv_phase3d_kernel<<<grid, 256, 0, (cudaStream_t)s>>>((const __nv_bfloat16*)x, wt, (const __nv_bfloat16*)add,
                                                                 (__nv_bfloat16*)y, B, H, W, C, f, gshift);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
  }
  if (256 % groups == 0 && f <= 8) {   // phase kernel: taps in registers (every geometry of the DLA-34 / ResNet nets)
    const long long per_phase = (long long)B * (H + 1) * (W + 1) * groups;
    long long gx = (per_phase + 255) / 256;
    const long long cap = 148LL * 16 / (f * f) > 148 ? 148LL * 16 / (f * f) : 148;
    if (gx > cap) gx = cap;
    dw_deconv_phase_kernel<<<dim3((unsigned)gx, (unsigned)(f * f)), 256, 0, (cudaStream_t)s>>>(
        (const __nv_bfloat16*)x, wt, (const __nv_bfloat16*)add, (__nv_bfloat16*)y, B, H, W, C, f);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
  }
  const long long total = (long long)B * H * f * W * f * groups;
  dw_deconv_kernel<<<grid_for(total), 256, 0, (cudaStream_t)s>>>((const __nv_bfloat16*)x, wt,
                                                                 (const __nv_bfloat16*)add, (__nv_bfloat16*)y, B, H,
                                                                 W, C, f);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_maxpool2d_pad(const void* x, void* y, int B, int H, int W, int C, int k, int stride, int pad,
                                 cnb_stream_t stream) {
  CNB_CHECK_ARG(x && y && B >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "maxpool2d_pad: bad argument");
  CNB_CHECK_ARG(k >= 1 && stride >= 1 && pad >= 0 && 2 * pad <= k, "maxpool2d_pad: bad window");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  CNB_CHECK_ARG(Ho >= 1 && Wo >= 1, "maxpool2d_pad: empty output");
  const long long total = (long long)B * Ho * Wo * (C / 8);
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  maxpool_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, H, W, C,
                                                             Ho, Wo, k, stride, pad);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

extern "C" int cnb_depth_to_space2(const void* x, void* y, int B, int H, int W, int C, cnb_stream_t stream) {
  CNB_CHECK_ARG(x && y && B >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "depth_to_space2: bad argument");
  const long long total = (long long)B * H * W * 4 * (C / 8);
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  depth_to_space2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, H, W, C);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}
