/*
 * centernet_b200 -- C ABI of the B200-native CenterNet hot path (libcenternet_b200.so).
 *
 * The reference (tteepe/CenterNet-pytorch-lightning) is pure Python and has no FFI of its own;
 * its plugin surface is Python duck typing.  Each entry point below is what a
 * ctypes binding for the corresponding reference function binds to (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - every pointer is a raw CUDA *device* pointer unless the function name ends in `_host`;
 *   - the caller owns all memory (nothing is allocated inside, except by the `_host` variants which
 *     keep a cached device staging arena per process);
 *   - calls are stream-ordered and asynchronous on `stream` (a cudaStream_t; NULL = legacy default);
 *   - stateless / re-entrant per stream; return 0 on success, a negative cnb_status otherwise, and
 *     cnb_last_error() returns a thread-local message for the last failure; never exit()/abort();
 *   - activations between conv layers are NHWC bf16; network input is NCHW fp32 and the head maps
 *     handed to decode are NCHW fp32 (the contract of decode/ctdet.py:6).
 */
#ifndef CENTERNET_B200_H_
#define CENTERNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cnb_stream_t; /* cudaStream_t */

enum cnb_status {
  CNB_OK = 0,
  CNB_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  CNB_ERR_WORKSPACE = -2, /* workspace too small */
  CNB_ERR_CUDA = -3,      /* CUDA runtime error (message in cnb_last_error) */
  CNB_ERR_NO_DEVICE = -4
};

int cnb_version(void);
const char* cnb_last_error(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
unsigned long long cnb_launch_count(void);

/* ---------------------------------------------------------------- decode ------------------- */
/* Replaces CenterNet/decode/ctdet.py:6-38 `ctdet_decode(heat, wh, reg=None, K=100)` including
 * utils/decode.py:5-10 `_nms`, :13-28 `_topk`, :59-63 `_transpose_and_gather_feat`.
 * heat [B,C,H,W] fp32 (already sigmoided, >= 0), wh/reg [B,2,H,W] fp32 (reg may be NULL -> +0.5),
 * out [B,K,6] fp32 = (x1,y1,x2,y2,score,cls).  Ties: (score desc, flat index asc). */
size_t cnb_ctdet_decode_workspace_bytes(int B, int C, int H, int W, int K);
int cnb_ctdet_decode(const float* heat, const float* wh, const float* reg, float* out,
                     int B, int C, int H, int W, int K,
                     void* workspace, size_t workspace_bytes, cnb_stream_t stream);
/* same, host buffers: H2D of the three maps, decode, D2H of out, synchronous on return */
int cnb_ctdet_decode_host(const float* heat, const float* wh, const float* reg, float* out,
                          int B, int C, int H, int W, int K);

/* Replaces CenterNet/decode/multi_pose.py:7-96 `multi_pose_decode` (+ utils/decode.py:31-40
 * `_topk_channel`).  heat [B,1,H,W], wh/reg/hp_offset [B,2,H,W], kps [B,2J,H,W], hm_hp [B,J,H,W];
 * reg and hp_offset may be NULL (-> +0.5); hm_hp is required (the reference raises NameError without
 * it, multi_pose.py:94).  out [B,K,3J+6] = bbox(4) score(1) kps(2J) cls(1) hm_score(J). */
size_t cnb_multi_pose_decode_workspace_bytes(int B, int J, int H, int W, int K);
int cnb_multi_pose_decode(const float* heat, const float* wh, const float* kps, const float* reg,
                          const float* hm_hp, const float* hp_offset, float* out,
                          int B, int J, int H, int W, int K,
                          void* workspace, size_t workspace_bytes, cnb_stream_t stream);

/* ---------------------------------------------------------------- losses ------------------- */
/* Replaces utils/decode.py:43-45 `sigmoid_clamped` + utils/losses.py:14-39 `_neg_loss` (FocalLoss)
 * forward AND backward in one pass over (logits, gt):
 *   p = clamp(sigmoid(x), 1e-4, 1-1e-4);  loss = -(sum pos + sum neg)/num_pos  (or -sum neg).
 * workspace: cnb_focal_loss_workspace_bytes(n) bytes (per-CTA partials; deterministic reduction).
 * loss_out: 3 device floats = { loss, gnorm, num_pos } with gnorm = 1/num_pos (1 if num_pos == 0).
 * If dlogits != NULL it receives the UNNORMALISED gradient g:  dloss/dlogits = g * gnorm  (num_pos is
 * only known after the reduction; the caller folds gnorm into its upstream gradient, no host sync;
 * the clamp passes zero gradient outside [1e-4, 1-1e-4], as torch.clamp does).  If prob_out != NULL
 * the clamped sigmoid is written there (what the reference stores in output["heatmap"]).
 * All pointers 16-byte aligned. */
size_t cnb_focal_loss_workspace_bytes(long long n);
int cnb_focal_loss_fwd_bwd(const float* logits, const float* gt, float* prob_out, float* dlogits,
                           float* loss_out, long long n,
                           void* workspace, size_t workspace_bytes, cnb_stream_t stream);
/* pred already sigmoid-clamped: the literal FocalLoss()(pred, gt) call (losses.py:42-50);
 * dpred (nullable) receives the unnormalised gradient as above. */
int cnb_focal_loss_prob_fwd_bwd(const float* pred, const float* gt, float* dpred,
                                float* loss_out, long long n,
                                void* workspace, size_t workspace_bytes, cnb_stream_t stream);

/* utils/decode.py:43-45 `sigmoid_clamped` on its own (the drop-in call made by
 * centernet_detection.py:104): y = clamp(sigmoid(x), lo, hi); x may alias y (in place, like
 * `x.sigmoid_()`); backward: dx = dy * y(1-y) where the clamp was inactive, else 0. */
int cnb_sigmoid_clamped_fwd(const float* x, float* y, long long n, float lo, float hi,
                            cnb_stream_t stream);
int cnb_sigmoid_clamped_bwd(const float* y, const float* dy, float* dx, long long n, float lo,
                            float hi, cnb_stream_t stream);

/* Replaces utils/losses.py:53-63 `RegL1Loss` and :81-91 `RegWeightedL1Loss` forward+backward without
 * the NCHW->NHWC copy of utils/decode.py:59-63.
 * output [B,C,H,W] fp32; ind [B,M] int64 (each in [0,H*W)); target [B,M,C] fp32;
 * mask: mask_per_channel==0 -> [B,M] uint8 (bool), expanded over C (denominator counts C*n);
 *       mask_per_channel==1 -> [B,M,C] fp32 weights (RegWeightedL1Loss).
 * loss_out: 1 device float.  If doutput != NULL it is zero-filled by this call and receives
 * (*grad_scale_dev) * dloss/doutput (scatter-add; duplicate indices add); grad_scale_dev is a device
 * scalar (NULL = 1) so that autograd's upstream gradient never needs a host sync. */
int cnb_reg_l1_fwd_bwd(const float* output, const void* mask, const long long* ind,
                       const float* target, float* doutput, float* loss_out,
                       int B, int C, int H, int W, int M, int mask_per_channel,
                       const float* grad_scale_dev, cnb_stream_t stream);

/* ---------------------------------------------------------------- convolution engine ------- */
/* Implicit-GEMM convolution on tcgen05 (bf16 x bf16 -> fp32 in TMEM), NHWC bf16 activations.
 * Replaces nn.Conv2d + nn.BatchNorm2d(eval) + residual add + nn.ReLU chains of
 * models/backbones/pose_dla_dcn.py:28-68 (BasicBlock), :165-188 (Root), :351-370, heads.py:4-25.
 *   x    [B,Hi,Wi,Ci]  bf16 (Ci % 8 == 0), optionally a channel slice of a wider tensor (x_cstride)
 *   wpk  packed weights from cnb_conv_pack_weights: [Co_pad][Kpad] bf16, K order (kh,kw,ci)
 *   scale/shift [Co] fp32 applied to the fp32 accumulator (folded BN, or scale=1/shift=bias)
 *   res  optional residual, same geometry as y (added before ReLU)
 *   y    [B,Ho,Wo,Co] bf16 NHWC (y_cstride/y_coffset allow writing into a concat buffer), or
 *        fp32 NCHW (out_nchw_f32 == 1, head maps) or fp32 NHWC (== 2); act: 0 none, 1 relu, 2 sigmoid.
 * For Ci not a multiple of 8 (the 3-channel stem) pad the activation's channels with
 * cnb_nchw_f32_to_nhwc_bf16(..., C_pad=8); cnb_conv_pack_weights pads the filter the same way. */
typedef struct cnb_conv_desc {
  int B, Hi, Wi, Ci;      /* input geometry */
  int Co, KH, KW;         /* filter */
  int stride, pad, dil;
  int Ho, Wo;
  int x_cstride;          /* channel stride (elements) between pixels of x; >= Ci */
  int x_coffset;
  int y_cstride;          /* channel stride of y / res (NHWC mode) */
  int y_coffset;
  int res_cstride;
  int res_coffset;
  int act;                /* 0 none, 1 relu, 2 sigmoid */
  int out_nchw_f32;       /* output format: 0 NHWC bf16, 1 NCHW fp32 (head maps), 2 NHWC fp32 (DCN offsets) */
  int w_kw;               /* filter width the weights were PACKED with (0 = KW).  The 8-channel stem packs its 7x7
                             filter as 7x8 (zero column) so that pairs of taps form one K=16 tensor-core step. */
  int pad_w1;             /* horizontal padding + 1 when it differs from `pad` (0 = same as pad).  With KH != KW this
                             describes the rectangular 7x5 filter of the space-to-depth stem (cnb_stem_s2d_pack_weights);
                             rectangular filters run on the row-window kernel only (stride 1, Wo % 128 == 0). */
} cnb_conv_desc;

/* Data gradient of a convolution = cnb_conv2d_fprop on a transformed filter (no separate kernel):
 *   stride 1: dX = conv(dY, W') with W'[ci][co][kh][kw] = W[co][ci][KH-1-kh][KW-1-kw], pad' = K-1-pad
 *             (host mirror: ops.pack_conv_weights_dgrad / ops.conv2d_dgrad);
 *   stride 2 (3x3, pad 1, even input): the four output phases (dX[2i+a, 2j+b]) are the channel blocks of ONE 3x3
 *             stride-1 convolution over dY with 4*Ci output channels, followed by cnb_depth_to_space2
 *             (ops.pack_conv_weights_dgrad_s2 / ops.conv2d_dgrad_s2).
 * Weight gradient: cnb_conv2d_wgrad below. */
size_t cnb_conv_packed_weight_bytes(int Co, int Ci, int KH, int KW);
/* w: [Co,Ci,KH,KW] fp32 (PyTorch layout, device) -> wpk bf16 device */
int cnb_conv_pack_weights(const float* w, void* wpk, int Co, int Ci, int KH, int KW,
                          cnb_stream_t stream);
int cnb_conv2d_fprop(const cnb_conv_desc* d, const void* x, const void* wpk, const float* scale,
                     const float* shift, const void* res, void* y, cnb_stream_t stream);

/* Weight gradient (training path; what autograd gives the reference in centernet.py:70-80):
 *   dW[co][kh][kw][ci] += sum over output pixels of dY[.., co] * X[.., shifted by (kh,kw), ci]
 * tcgen05 GEMM with the pixel index as the contraction (both operands MN-major, the forward kernel's TMA im2col
 * tiles reused as they lie).  `d` is the FORWARD convolution's descriptor (x geometry, Co, filter, stride, pad, w_kw);
 * dy [B,Ho,Wo,dy_cstride] bf16 NHWC, channels [dy_coffset, dy_coffset+Co);  dw_acc fp32 [Co][KH*KWp*Ci] (KWp = w_kw
 * or KW; K order (kh,kw,ci) = the packed-weight order) is ACCUMULATED into with vector reductions: zero it first.
 * cnb_conv_unpack_wgrad turns it into the PyTorch layout [Co][Ci][KH][KW] fp32 (dw = (accumulate ? dw : 0) +
 * scale * unpacked). */
size_t cnb_conv_wgrad_acc_elems(int Co, int Ci_pad, int KH, int KWp);
int cnb_conv2d_wgrad(const cnb_conv_desc* d, const void* x, const void* dy, int dy_cstride, int dy_coffset,
                     float* dw_acc, cnb_stream_t stream);
int cnb_conv_unpack_wgrad(const float* dw_acc, float* dw, int Co, int Ci, int Ci_pad, int KH, int KW, int KWp,
                          int accumulate, float scale, cnb_stream_t stream);

/* CenterHead in one kernel (models/heads.py:4-50): for every head, conv3x3(Ci=64 -> 256, bias) -> ReLU ->
 * conv1x1(256 -> c_out, bias) [-> act], with the 256-channel intermediate kept in tensor memory (never written to HBM).
 *   x     [B,H,W,x_cstride] bf16 NHWC feature map (64 channels at x_coffset)
 *   w3pk  cnb_conv_pack_weights of the heads' 3x3 filters concatenated along Co: [nheads*256][576]
 *   bias3 [nheads*256] fp32;  w1pk: the heads' packed 1x1 filters concatenated ([sum round16(c_out)][256]);
 *   bias1[h] [c_out[h]] fp32 (nullable);  y[h] [B,c_out[h],H,W] fp32 NCHW;  act[h]: 0 none, 2 sigmoid.
 * bias1 / y are HOST arrays of device pointers (read during the call). */
typedef struct cnb_head_desc {
  int B, H, W, Ci, x_cstride, x_coffset, head_conv, nheads;
  int c_out[8];
  int act[8];
} cnb_head_desc;
int cnb_head_fused_supported(const cnb_head_desc* d);
int cnb_head_fused_fprop(const cnb_head_desc* d, const void* x, const void* w3pk, const float* bias3,
                         const void* w1pk, const float* const* bias1, float* const* y, cnb_stream_t stream);

/* Modulated deformable convolution v2 (DCN.dcn_v2.DCN forward; call sites pose_dla_dcn.py:441-449,
 * resnet_dcn.py:202-210): 3x3, stride 1, pad 1, dil 1, deformable_groups 1.
 *   om [B,H,W,om_cstride] fp32 NHWC holds the 27 raw channels of the conv_offset_mask conv (run it
 *   through cnb_conv2d_fprop with out_nchw_f32 == 2).  DCNv2 does o1,o2,mask = chunk(om,3);
 *   offset = cat(o1,o2) = channels 0..17, so tap k samples at (dy,dx) = (ch 2k, ch 2k+1) and is
 *   modulated by sigmoid(ch 18+k); samples outside (-1,H)x(-1,W) are 0.
 * The bilinear sampler writes the A operand tile directly into shared memory; same epilogue as conv. */
int cnb_dcnv2_fprop(const cnb_conv_desc* d, const void* x, const float* om, int om_cstride,
                    const void* wpk, const float* scale, const float* shift, void* y,
                    cnb_stream_t stream);

/* ---------------------------------------------------------------- memory-bound layer ops --- */
/* MaxPool2d(k, stride=k) NHWC bf16 (pose_dla_dcn.py:243); channel-slice addressing on both sides so that
 * the pooled map can be written straight into a Root concat buffer (pose_dla_dcn.py:182). */
int cnb_maxpool2d(const void* x, void* y, int B, int H, int W, int C, int x_cstride, int x_coffset,
                  int y_cstride, int y_coffset, int k, cnb_stream_t stream);
/* MaxPool2d(k, stride, padding) with -inf padding, dense NHWC bf16 (the ResNet stem's 3x3 / stride 2 / pad 1 pool,
 * resnet_dcn.py:139, msra_resnet.py:112). */
int cnb_maxpool2d_pad(const void* x, void* y, int B, int H, int W, int C, int k, int stride, int pad,
                      cnb_stream_t stream);
/* Pixel shuffle by 2: x [B,H,W,4C] (channel blocks ordered by output phase py*2+px) -> y [B,2H,2W,C] NHWC bf16.
 * Second half of the dense ConvTranspose2d(4, stride 2, pad 1) of resnet_dcn.py:212-220 / msra_resnet.py:165-175,
 * whose four output phases are computed together by ONE 3x3 tensor-core conv with 4C output channels. */
int cnb_depth_to_space2(const void* x, void* y, int B, int H, int W, int C, cnb_stream_t stream);
/* depthwise ConvTranspose2d(C,C,2f,stride=f,padding=f/2,groups=C,bias=False) (pose_dla_dcn.py:466-475).
 * w [C,1,2f,2f] fp32 is first re-laid-out to wt [(2f)^2][C] fp32; then y = up(x) (+ add if add != NULL:
 * fuses `layers[i] + layers[i-1]`, pose_dla_dcn.py:488).  x [B,H,W,C], y/add [B,H*f,W*f,C] NHWC bf16. */
int cnb_dw_deconv_relayout_weights(const float* w, float* wt, int C, int f, cnb_stream_t stream);
int cnb_dw_deconv_up(const void* x, const float* wt, const void* add, void* y, int B, int H, int W,
                     int C, int f, cnb_stream_t stream);
/* layout changes at the two ends of the NHWC bf16 engine: network input [B,C,H,W] fp32 -> NHWC bf16 with
 * channels zero-padded to C_pad (3 -> 8 for the 7x7 stem, pose_dla_dcn.py:281-285); a channel slice of an
 * NHWC bf16 map -> NCHW fp32 (the backbone output contract). */
int cnb_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int C, int H, int W, int C_pad,
                              cnb_stream_t stream);
int cnb_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int C, int H, int W, int x_cstride,
                              int x_coffset, cnb_stream_t stream);

/* ---------------------------------------------------------------- training path (bandwidth-bound pieces) ---- */
/* Train-mode nn.BatchNorm2d (pose_dla_dcn.py:40, momentum 0.1) fused with the residual add and ReLU that follow it.
 * z [M,C] bf16 NHWC = conv output; y = act(bn(z) (+ res)).  stats [4][C] fp32 receives (mean, invstd, scale, shift)
 * for the backward; running_mean/var are updated like PyTorch (unbiased variance; pass NULL to skip);
 * sums_ws: 2*C floats of scratch. */
int cnb_bn_train_fwd(const void* z, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, const void* res, int act, void* y,
                     float* stats, float* sums_ws, long long M, int C, cnb_stream_t stream);
/* y = act(z * scale + shift (+ res)), per-channel fp32 scale/shift, NHWC bf16 */
int cnb_scale_shift_act(const void* z, const float* scale, const float* shift, const void* res, int act,
                        void* y, long long M, int C, cnb_stream_t stream);
/* backward of the above: g = dy * (y > 0) (y_or_null == NULL: no ReLU); dz = scale*(g - mean(g) - xhat*mean(g*xhat));
 * dres_or_null receives g (gradient of the residual branch); dgamma/dbeta fp32 [C] (+= if accumulate). */
int cnb_bn_train_bwd(const void* dy, const void* y_or_null, const void* z, const float* stats, void* dz,
                     void* dres_or_null, float* dgamma, float* dbeta, int accumulate, float* sums_ws,
                     long long M, int C, cnb_stream_t stream);
/* out[c] (+)= sum over pixels of dy[m][coffset + c] (* (y > 0) when y_mask_or_null is given): conv bias gradients */
int cnb_channel_sum(const void* dy, int cstride, int coffset, const void* y_mask_or_null, float* out,
                    int accumulate, long long M, int C, cnb_stream_t stream);
int cnb_relu_bwd(const void* y, const void* dy, void* dx, long long n, cnb_stream_t stream);
int cnb_add_bf16(const void* a, const void* b, void* out, long long n, cnb_stream_t stream);
/* out = bf16(acc (+ addend)); with 0 < cvalid < cstride: channels >= cvalid of every [.., cstride] pixel are zeroed */
int cnb_f32_to_bf16(const float* acc, const void* addend_or_null, void* out, long long n, int cstride,
                    int cvalid, cnb_stream_t stream);
/* nn.MaxPool2d(2, 2) backward (pose_dla_dcn.py:243): gradient to the first maximum of each window */
int cnb_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int B, int H, int W, int C, cnb_stream_t stream);
/* depthwise ConvTranspose2d(C,C,2f,stride f,pad f/2) backward (pose_dla_dcn.py:466-475): dx, and dW accumulated into
 * dwt_acc [(2f)^2][C] fp32 (zero it first); cnb_dw_deconv_unpack_wgrad -> [C,1,2f,2f]. */
int cnb_dw_deconv_bwd(const void* x, const float* wt, const void* dy, void* dx, float* dwt_acc, int B, int H,
                      int W, int C, int f, cnb_stream_t stream);
int cnb_dw_deconv_unpack_wgrad(const float* dwt_acc, float* dw, int C, int f, int accumulate, cnb_stream_t stream);
/* DCNv2 for training (DCN.dcn_v2.DCN backward; pose_dla_dcn.py:441-449): the sampled, modulated column matrix
 * col [B*H*W][9*C] bf16 (K order (tap, ci)) is materialised so that y = conv1x1(col; W), dW and dcol are tensor-core
 * GEMMs (cnb_conv2d_fprop / cnb_conv2d_wgrad); cnb_dcnv2_col2im turns dcol into dX (fp32, zero-filled here,
 * scatter-added), and the gradients of the 27 raw offset/mask channels dom [B,H,W,om_cstride] fp32. */
int cnb_dcnv2_im2col(const void* x, const float* om, int om_cstride, void* col, int B, int H, int W, int C,
                     cnb_stream_t stream);
int cnb_dcnv2_col2im(const void* x, const float* om, int om_cstride, const void* dcol, float* dx_acc,
                     float* dom, int B, int H, int W, int C, cnb_stream_t stream);
/* torch.optim.Adam step (centernet.py:94-95 defaults) on flat fp32 buffers; g is multiplied by grad_scale first
 * (1/world_size after a sum all-reduce).  step_dev / lr_dev (nullable device scalars) override step / lr: a captured
 * CUDA graph of the training step reads the values current at replay time. */
int cnb_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                  float beta2, float eps, int step, const int* step_dev, const float* lr_dev, float grad_scale,
                  cnb_stream_t stream);

/* Gradient all-reduce over NVLink peer memory (no NCCL on the data path): every rank's flat gradient buffer is a
 * symmetric-memory allocation mapped into all peers; one kernel per rank and bucket publishes "written", waits for the
 * peers, reduces its 1/world shard over all peers' buffers (peer loads; multimem.ld_reduce in the switch when a
 * multicast mapping is given) and stores the sum back to every peer.  bufs / sigs: HOST arrays of `world` device
 * pointers (buffer and signal pad of each rank, cnb_p2p_signal_bytes() bytes each, zero-initialised); counters: >= 64
 * zeroed u32 (local); epoch_dev: device u32 advanced once per step by cnb_p2p_next_step; [off, off+n) floats with
 * off % 4 == 0 and n % (4*world) == 0; slot < 64 identifies the bucket.  cnb_p2p_wait blocks the stream until the
 * shards of slots [0, nslots) have arrived from every peer. */
size_t cnb_p2p_signal_bytes(void);
int cnb_p2p_next_step(unsigned int* epoch_dev, cnb_stream_t stream);
int cnb_p2p_allreduce(float* const* bufs, unsigned int* const* sigs, float* multicast_or_null,
                      unsigned int* counters, const unsigned int* epoch_dev, long long off, long long n,
                      int rank, int world, int slot, int ctas, cnb_stream_t stream);
int cnb_p2p_wait(const unsigned int* sig_local, const unsigned int* epoch_dev, int nslots, int world,
                 cnb_stream_t stream);

/* ---------------------------------------------------------------- fp32-strict mode (NCHW fp32, CUDA cores) ---- */
/* The same operators in the reference's own precision, for small-shape END-TO-END parity checks against the fp32
 * CPU path (hard part 2).  Not the measured fast path. */
int cnb_strict_conv2d_f32(const float* x, const float* w, const float* scale, const float* shift,
                          const float* res, float* y, int B, int Ci, int Hi, int Wi, int Co, int KH, int KW,
                          int stride, int pad, int act, cnb_stream_t stream);
int cnb_strict_dcn_im2col_f32(const float* x, const float* om, float* col, int B, int C, int H, int W,
                              cnb_stream_t stream);
int cnb_strict_conv_transpose2d_f32(const float* x, const float* w, const float* scale, const float* shift,
                                    const float* add, float* y, int B, int Ci, int Hi, int Wi, int Co, int K,
                                    int stride, int pad, int depthwise, int act, cnb_stream_t stream);
int cnb_strict_maxpool2d_f32(const float* x, float* y, int BC, int H, int W, int k, int stride, int pad,
                             cnb_stream_t stream);

/* ---------------------------------------------------------------- decode primitives by name ---- */
/* utils/decode.py:5-10 `_nms`; :13-40 the torch.topk inside `_topk` / `_topk_channel` (exact, sorted, ties by
 * ascending index); :48-63 `_gather_feat` / `_transpose_and_gather_feat`.  ctdet_decode / multi_pose_decode fuse all
 * of these and do not call them. */
int cnb_nms3x3(const float* heat, float* out, long long planes, int H, int W, cnb_stream_t stream);
int cnb_topk_rows(const float* scores, int rows, int n, int K, float* out_scores, long long* out_idx,
                  cnb_stream_t stream);
int cnb_gather_feat(const float* feat, const long long* ind, float* out, int B, int C, int N, int K,
                    int feat_is_nchw, cnb_stream_t stream);

/* ---------------------------------------------------------------- callers either side of the path ---- */
/* (f)2 target encoding, CenterNet/sample/ctdet.py:39-90 + utils/gaussian.py:6-58 (umich Gaussian), batched on the device:
 * boxes [B,M,4] float64 COCO (x,y,w,h) in input pixels, cls [B,M] int32, nobj [B] int32 ->
 * heatmap [B,C,H,W] fp32 (zero-filled here), indices [B,M] int64, mask [B,M] uint8 (bool), wh / reg [B,M,2] fp32. */
int cnb_ctdet_encode(const double* boxes, const int* cls, const int* nobj, float* heatmap, long long* indices,
                     unsigned char* mask, float* wh, float* reg, int B, int C, int H, int W, int M,
                     int down_ratio, cnb_stream_t stream);
/* (f)3 soft-NMS, CenterNet/utils/nms.py:5-106 (ncol = 5) / :109-206 (ncol = 39): nlists independent lists of
 * counts[l] <= max_boxes rows [x1,y1,x2,y2,score,...], processed in place (kept rows first, selection order,
 * decayed scores); kept[l] = number of rows kept.  method 0 hard, 1 linear, 2 gaussian. */
int cnb_soft_nms(float* boxes, const int* counts, int* kept, int nlists, int max_boxes, int ncol, double sigma,
                 double Nt, double threshold, int method, cnb_stream_t stream);
/* (f)1 test-time augmentation around the engine (centernet_detection.py:143-171, 188-204):
 * prologue: img [3,H,W] fp32 -> out [(flip ? 2 : 1),3,H+2*pad_tb,W+2*pad_lr] = normalize(zero-pad(img)) (+ its hflip);
 *           mean3 / std3 are HOST arrays of 3 floats;
 * flip merge: pair [2,C,H,W] -> out [C,H,W] = (pair[0] + hflip(pair[1])) / 2;
 * post: det [K,6] -> out [K,5] rows (x1,y1,x2,y2,score) in image pixels ((v * down_ratio - pad) / scale) grouped by class
 *       (class c occupies rows [offsets[c], offsets[c] + counts[c]), input order kept inside a class). */
int cnb_tta_prologue(const float* img, float* out, int C, int H, int W, int pad_lr, int pad_tb, const float* mean3,
                     const float* std3, int flip, cnb_stream_t stream);
int cnb_tta_flip_merge(const float* pair, float* out, int C, int H, int W, cnb_stream_t stream);
int cnb_ctdet_post(const float* det, float* out, int* counts, int* offsets, int K, int C, float down_ratio,
                   float pad_x, float pad_y, float scale_x, float scale_y, cnb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CENTERNET_B200_H_ */
